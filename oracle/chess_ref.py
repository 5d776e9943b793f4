"""Chess side of the oracle.  TEST INFRASTRUCTURE ONLY.

The reference's chess environment (/root/reference/custom_alphazero/chess/board.py) is a thin subclass of
python-chess's Board; python-chess (chess 1.9.4 / python-chess 1.999 in the reference's poetry.lock) is neither
vendored nor installed, so chess parity is UNPINNED against the reference itself.  What this file pins instead:

* the rules: oracle/c/chess_oracle.c (mailbox, pseudo-legal + king test) against the published perft counts;
* the reference's own layer on top of python-chess, restated here line by line: the int8 `array`
  (board.py:100-139), `state` / `full_state` (:50-73), `play` with mirror (:162-173), `get_result` (:178-190),
  the action list (utils.py:11-32) built by the same procedure (lone queen / knight on every square, pawns on the
  seventh rank) over the oracle's move generator, `Move` ordering (move.py:28-32).

python-chess behaviour assumed where the reference leans on it (from its documentation / source, unverifiable here):
Board.mirror() returns a stack-less copy made through type(self)(None) - for the reference's subclass that runs
its __init__ with the initial position, so after every keep_same_player move `state_history` is a fresh deque
[0, 0, 0, 0, 0, 0, 0, state(initial position)] to which update_array() appends the new state; the fullmove number
only advances after black moves, and on that path black never moves; is_repetition() needs the move stack and
is always False there.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "_build", "libchess_oracle.so")
HOSTCHECK_LIB = os.path.join(HERE, "_build", "libchess_hostcheck.so")

PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]  # ConfigChess.piece_symbols (config.py:27)
PROMO_LETTERS = ["", "b", "n", "q", "r"]              # sorted UCI suffixes = the engine's promo codes 0..4
PROMO_TYPE = {"": 0, "n": 2, "b": 3, "r": 4, "q": 5}  # python-chess piece types
START_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"

PERFT = {  # published perft node counts (chessprogramming wiki "Perft Results")
    START_FEN: [20, 400, 8902, 197281, 4865609],
    "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1": [48, 2039, 97862, 4085603],
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1": [14, 191, 2812, 43238, 674624],
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1": [6, 264, 9467, 422333],
    "r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1": [6, 264, 9467, 422333],
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8": [44, 1486, 62379, 2103487],
    "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10": [46, 2079, 89890, 3894594],
}


class CoState(ctypes.Structure):
    _fields_ = [("sq", ctypes.c_int8 * 64), ("turn", ctypes.c_int32), ("castling", ctypes.c_int32),
                ("ep", ctypes.c_int32), ("halfmove", ctypes.c_int32), ("fullmove", ctypes.c_int32)]

    def copy(self):
        c = CoState()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(CoState))
        return c


class CoMove(ctypes.Structure):
    _fields_ = [("from_", ctypes.c_int8), ("to", ctypes.c_int8), ("promo", ctypes.c_int8), ("pad", ctypes.c_int8)]


def build(force=False):
    srcs = [os.path.join(HERE, "c", "chess_oracle.c"), os.path.join(HERE, "c", "chess_hostcheck.cpp"),
            os.path.join(HERE, "..", "custom-alphazero_b200", "csrc", "az_chess.cuh"),
            os.path.join(HERE, "..", "custom-alphazero_b200", "csrc", "az_chess_tables.inc")]
    newest = max(os.path.getmtime(s) for s in srcs)
    for lib in (ORACLE_LIB, HOSTCHECK_LIB):
        if force or not os.path.exists(lib) or os.path.getmtime(lib) < newest:
            subprocess.check_call(["make", "-s", "-C", HERE, "_build/" + os.path.basename(lib)])


_olib = None
_hlib = None


def olib():
    global _olib
    if _olib is None:
        build()
        L = ctypes.CDLL(ORACLE_LIB)
        P = ctypes.POINTER(CoState)
        L.co_legal.restype = ctypes.c_int
        L.co_legal.argtypes = [P, ctypes.POINTER(CoMove)]
        L.co_push.restype = None
        L.co_push.argtypes = [P, CoMove]
        L.co_mirror.restype = None
        L.co_mirror.argtypes = [P]
        L.co_status.restype = ctypes.c_int
        L.co_status.argtypes = [P]
        L.co_in_check.restype = ctypes.c_int
        L.co_in_check.argtypes = [P]
        for name in ("co_perft", "co_perft_mirrored"):
            getattr(L, name).restype = ctypes.c_uint64
            getattr(L, name).argtypes = [P, ctypes.c_int]
        _olib = L
    return _olib


def hlib():
    """The device rules header compiled for the host (oracle/c/chess_hostcheck.cpp)."""
    global _hlib
    if _hlib is None:
        build()
        L = ctypes.CDLL(HOSTCHECK_LIB)
        U = ctypes.POINTER(ctypes.c_uint64)
        L.hc_legal.restype = ctypes.c_int
        L.hc_legal.argtypes = [U, U, ctypes.POINTER(ctypes.c_int)]
        L.hc_play.restype = None
        L.hc_play.argtypes = [U, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, U]
        L.hc_mirror.restype = None
        L.hc_mirror.argtypes = [U, U]
        L.hc_status.restype = ctypes.c_int
        L.hc_status.argtypes = [U]
        L.hc_perft_mirrored.restype = ctypes.c_uint64
        L.hc_perft_mirrored.argtypes = [U, ctypes.c_int]
        _hlib = L
    return _hlib


# ---------------------------------------------------------------- states
def from_fen(fen):
    parts = fen.split()
    s = CoState()
    rank, file = 7, 0
    for ch in parts[0]:
        if ch == "/":
            rank, file = rank - 1, 0
        elif ch.isdigit():
            file += int(ch)
        else:
            v = PIECE_SYMBOLS.index(ch.lower())
            s.sq[rank * 8 + file] = v if ch.isupper() else -v
            file += 1
    s.turn = 1 if (len(parts) < 2 or parts[1] == "w") else 0
    rights = parts[2] if len(parts) > 2 else "-"
    s.castling = sum(b for c, b in zip("KQkq", (1, 2, 4, 8)) if c in rights)
    ep = parts[3] if len(parts) > 3 else "-"
    s.ep = -1 if ep == "-" else (ord(ep[0]) - 97) + 8 * (int(ep[1]) - 1)
    s.halfmove = int(parts[4]) if len(parts) > 4 else 0
    s.fullmove = int(parts[5]) if len(parts) > 5 else 1
    return s


def start_state():
    return from_fen(START_FEN)


def legal(state):
    """[(from, to, promo letter)] of the side to move, in the oracle's generation order."""
    buf = (CoMove * 256)()
    n = olib().co_legal(ctypes.byref(state), buf)
    inv = {v: k for k, v in PROMO_TYPE.items()}
    return [(buf[i].from_, buf[i].to, inv[buf[i].promo]) for i in range(n)]


def push(state, move, keep_same_player=False):
    """chess/board.py:162-173 on a copy."""
    s = state.copy()
    olib().co_push(ctypes.byref(s), CoMove(move[0], move[1], PROMO_TYPE[move[2]], 0))
    if keep_same_player:
        olib().co_mirror(ctypes.byref(s))
        s.turn = 1
    return s


def status(state):
    return olib().co_status(ctypes.byref(state))


def result(state):
    """chess/board.py:178-190: None while the game goes on, 0 draw, +1 white won, -1 black won."""
    st = status(state)
    if st == 0:
        return None
    if st == 2:
        return 0
    return -1 if state.turn else 1


def perft(state, depth, mirrored=False):
    fn = olib().co_perft_mirrored if mirrored else olib().co_perft
    return int(fn(ctypes.byref(state), depth))


# ---------------------------------------------------------------- the reference's layer
def uci(move):
    f, t, p = move
    return "abcdefgh"[f & 7] + str((f >> 3) + 1) + "abcdefgh"[t & 7] + str((t >> 3) + 1) + p


def move_key(move):
    """chess/move.py:28-32: ordered by (pos_from, pos_to) = ((file, rank), (file, rank, promotion letter))."""
    f, t, p = move
    return ((f & 7, f >> 3), (t & 7, t >> 3, p))


def all_possible_moves():
    """chess/utils.py:11-32 with the oracle's move generator in python-chess's place."""
    moves = set()
    for sq in range(64):
        for piece in (5, 2):  # "Q", "N"
            s = CoState()
            s.turn, s.ep, s.fullmove = 1, -1, 1
            s.sq[sq] = piece
            moves.update(legal(s))
    s = CoState()
    s.turn, s.ep, s.fullmove = 1, -1, 1
    for f in range(8):
        s.sq[48 + f] = 1  # array[1, :] = "P": white pawns on the seventh rank
    moves.update(legal(s))
    for f in range(8):
        s.sq[56 + f] = -1  # array[0, :] = "p"
    moves.update(legal(s))
    return sorted(moves, key=move_key)


def array_of(state):
    """Board.array (chess/board.py:119-131): int8 [8][8], row 0 = rank 8, white positive."""
    a = np.zeros((8, 8), dtype=np.int8)
    for sq in range(64):
        a[7 - (sq >> 3), sq & 7] = state.sq[sq]
    return a


def state_planes(state, repetition=False):
    """Board.state (chess/board.py:50-56): one-hot over 13 values (negative values wrap) + repetition plane."""
    return np.dstack([np.eye(13)[array_of(state)], np.full((8, 8), repetition)])


def full_state(state, history):
    """Board.full_state (chess/board.py:58-73).  history: the 7 older entries of the deque, oldest first, each a
    state_planes() array or None for the zero padding; the current state is appended like update_array does."""
    hist = [np.zeros((8, 8, 14)) if h is None else h for h in history] + [state_planes(state)]
    assert len(hist) == 8
    turn = 1 if state.turn else -1
    own_q, own_k = (2, 1) if turn > 0 else (8, 4)
    opp_q, opp_k = (8, 4) if turn > 0 else (2, 1)
    feats = [bool(state.castling & own_q), bool(state.castling & own_k), bool(state.castling & opp_q),
             bool(state.castling & opp_k), state.fullmove, state.halfmove]
    return np.dstack([np.dstack(hist)] + [np.full((8, 8), f) for f in feats]).astype(np.float64)


def selfplay_history():
    """What state_history holds before update_array() on the keep_same_player path (see the module docstring)."""
    return [None] * 6 + [state_planes(start_state())]


def history_of(state):
    """The deque a self-play position is evaluated with: a Board() that has not moved yet still has the constructor's
    seven empty entries (chess/board.py:37-40); after play() + mirror() it is selfplay_history().  On that path the
    un-moved board is the start position with a zero halfmove clock (its recurrences have a non-zero clock)."""
    s0 = start_state()
    fresh = list(state.sq) == list(s0.sq) and state.halfmove == 0 and bool(state.turn) and state.castling == s0.castling
    return [None] * 7 if fresh else selfplay_history()


# ---------------------------------------------------------------- bridge to the engine's bitboard position
def to_pos(state, repetition=False, valid=True):
    """co_state -> the engine's Pos as 8 uint64 (pawns, knights, bishops, rooks, queens, kings, white, meta)."""
    bb = [0] * 7
    for sq in range(64):
        v = state.sq[sq]
        if v:
            bb[abs(v) - 1] |= 1 << sq
            if v > 0:
                bb[6] |= 1 << sq
    meta = (state.castling & 15) | ((state.ep + 1) << 4) | ((0 if state.turn else 1) << 11) | (state.halfmove << 16) | \
        (state.fullmove << 32) | (int(repetition) << 48) | (int(valid) << 49)
    return np.array(bb + [meta], dtype=np.uint64)


def from_pos(pos):
    s = CoState()
    pos = [int(x) for x in pos]
    for sq in range(64):
        for t in range(6):
            if (pos[t] >> sq) & 1:
                s.sq[sq] = (t + 1) if (pos[6] >> sq) & 1 else -(t + 1)
    meta = pos[7]
    s.castling = meta & 15
    s.ep = ((meta >> 4) & 127) - 1
    s.turn = 0 if (meta >> 11) & 1 else 1
    s.halfmove = (meta >> 16) & 0xFFFF
    s.fullmove = (meta >> 32) & 0xFFFF
    return s


def states_equal(a, b):
    return (bytes(a.sq) == bytes(b.sq) and a.turn == b.turn and a.castling == b.castling and a.ep == b.ep
            and a.halfmove == b.halfmove and a.fullmove == b.fullmove)


def mask_to_actions(mask_words):
    out = []
    for w, word in enumerate(mask_words):
        word = int(word)
        while word:
            b = word & -word
            out.append(w * 64 + b.bit_length() - 1)
            word ^= b
    return out


def _u64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def host_legal(pos):
    """Legal actions of `pos` by the device header compiled for the host -> (sorted action indices, in_check, unlisted)."""
    mask = np.zeros(30, dtype=np.uint64)
    flags = (ctypes.c_int * 2)()
    n = hlib().hc_legal(_u64p(pos), _u64p(mask), flags)
    acts = mask_to_actions(mask)
    assert len(acts) == n - flags[1]  # the count includes black promotions, the mask cannot hold them
    return acts, bool(flags[0]), flags[1]


def host_play(pos, move_code, keep_same_player):
    out = np.zeros(8, dtype=np.uint64)
    hlib().hc_play(_u64p(pos), move_code & 63, (move_code >> 6) & 63, move_code >> 12, int(keep_same_player), _u64p(out))
    return out


def host_status(pos):
    return hlib().hc_status(_u64p(pos))


def host_perft_mirrored(pos, depth):
    return int(hlib().hc_perft_mirrored(_u64p(pos), depth))


# ---------------------------------------------------------------- MCTS over chess (oracle/c/chess_oracle.c: co_mcts_game)
class CoMctsCfg(ctypes.Structure):
    _fields_ = [("eval_kind", ctypes.c_int32), ("prior_mode", ctypes.c_int32), ("sims", ctypes.c_int32),
                ("greedy_idx", ctypes.c_int32), ("max_plies", ctypes.c_int32), ("c_puct", ctypes.c_double)]


_act_index = None


def act_index_table():
    """int16 [4096]: index of the promotion-less move joining two squares in all_possible_moves(), -1 if none."""
    global _act_index
    if _act_index is None:
        t = np.full(4096, -1, dtype=np.int16)
        for i, (f, to, p) in enumerate(all_possible_moves()):
            if p == "":
                t[f * 64 + to] = i
        _act_index = t
    return _act_index


def mcts_game(start=None, sims=100, evaluator="uniform", prior_mode="f64", greedy_idx=8, max_plies=512, c_puct=1.5,
              uniforms=None):
    """One self-play game by the C oracle.  Returns dict(k, act, n, choice [plies...], result, sims, evals)."""
    L = olib()
    L.co_mcts_game.restype = ctypes.c_int
    start = start_state() if start is None else start
    cfg = CoMctsCfg({"uniform": 0, "hash": 1}[evaluator], {"f64": 0, "f32": 1}[prior_mode], sims, greedy_idx, max_plies, c_puct)
    P = max_plies
    out_k = np.zeros(P, dtype=np.int32)
    out_act = np.zeros((P, 224), dtype=np.uint16)
    out_n = np.zeros((P, 224), dtype=np.int32)
    out_choice = np.zeros(P, dtype=np.int32)
    result = ctypes.c_int32(0)
    counters = (ctypes.c_longlong * 2)()
    table = act_index_table()
    u = None
    if uniforms is not None:
        u = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64))
        assert len(u) >= P
    vp = ctypes.c_void_p
    plies = L.co_mcts_game(ctypes.byref(cfg), ctypes.byref(start), vp(table.ctypes.data),
                           vp(u.ctypes.data) if u is not None else vp(None), vp(out_k.ctypes.data), vp(out_act.ctypes.data),
                           vp(out_n.ctypes.data), vp(out_choice.ctypes.data), ctypes.byref(result), counters)
    return dict(plies=plies, k=out_k[:plies], act=out_act[:plies], n=out_n[:plies], choice=out_choice[:plies],
                result=result.value, sims=counters[0], evals=counters[1])
