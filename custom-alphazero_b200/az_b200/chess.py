"""Batched chess environment calls (az_chess_* kernels through the C ABI).

A board is 8 uint64 words (include/az_b200.h, az_chess_pos): pawns, knights, bishops, rooks, queens, kings (both
colours), white, meta.  Batches are numpy uint64 [n, 8] (or torch int64 [n, 8] on the GPU, same bits); numpy in ->
numpy out.  The rules of chess live in the kernels only: without libaz_b200 or without a CUDA device these raise.
The FEN / array conversions below are data-format glue (chess/board.py:119-153), not rules.
"""
import ctypes

import numpy as np
import torch

from . import native
from .engine import _ptr, _stream
from .native import NativeError, check, lib

N_ACTIONS, MASK_WORDS, PLANES = native.AZ_CHESS_ACTIONS, native.AZ_CHESS_MASK_WORDS, native.AZ_CHESS_PLANES
PROMO_LETTERS = ("", "b", "n", "q", "r")  # promo codes 0..4 of the action table = sorted UCI suffixes
PIECE_SYMBOLS = (None, "p", "n", "b", "r", "q", "k")  # ConfigChess.piece_symbols (config.py:27)
START_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
META_TURN, META_REP, META_VALID = 1 << 11, 1 << 48, 1 << 49

_table = None


def action_table():
    """uint16 [1880]: from | to << 6 | promo << 12 in the order of get_all_possible_moves() (chess/utils.py:11-32)."""
    global _table
    if _table is None:
        buf = (ctypes.c_uint16 * N_ACTIONS)()
        check(lib().az_chess_action_table(buf))
        _table = np.frombuffer(buf, dtype=np.uint16).copy()
    return _table


def action_uci(a):
    code = int(action_table()[a])
    f, t, p = code & 63, (code >> 6) & 63, code >> 12
    return "abcdefgh"[f & 7] + str((f >> 3) + 1) + "abcdefgh"[t & 7] + str((t >> 3) + 1) + PROMO_LETTERS[p]


_uci_index = None


def uci_action(uci):
    global _uci_index
    if _uci_index is None:
        _uci_index = {action_uci(a): a for a in range(N_ACTIONS)}
    return _uci_index.get(uci, -1)


# ------------------------------------------------------------------ data-format glue (host)
def pack_position(array, turn=True, castling=15, ep_square=None, halfmove_clock=0, fullmove_number=1,
                  repetition=False, valid=True):
    """Board.array (int8 [8, 8], row 0 = rank 8, white positive; chess/board.py:119-131) + state -> uint64 [8]."""
    array = np.asarray(array)
    bb = [0] * 7
    for row in range(8):
        for col in range(8):
            v = int(array[row, col])
            if v:
                sq = (7 - row) * 8 + col
                bb[abs(v) - 1] |= 1 << sq
                if v > 0:
                    bb[6] |= 1 << sq
    meta = (castling & 15) | ((0 if ep_square is None else ep_square + 1) << 4) | (0 if turn else META_TURN) | \
        (halfmove_clock << 16) | (fullmove_number << 32) | (META_REP if repetition else 0) | (META_VALID if valid else 0)
    return np.array(bb + [meta], dtype=np.uint64)


def unpack_position(pos):
    """uint64 [8] -> dict(array, turn, castling, ep_square, halfmove_clock, fullmove_number)."""
    w = [int(x) for x in np.asarray(pos).view(np.uint64)]
    array = np.zeros((8, 8), dtype=np.int8)
    for t in range(6):
        b = w[t]
        while b:
            low = b & -b
            sq = low.bit_length() - 1
            array[7 - (sq >> 3), sq & 7] = (t + 1) if w[6] & low else -(t + 1)
            b ^= low
    meta = w[7]
    ep = ((meta >> 4) & 127) - 1
    return dict(array=array, turn=not (meta & META_TURN), castling=meta & 15, ep_square=None if ep < 0 else ep,
                halfmove_clock=(meta >> 16) & 0xFFFF, fullmove_number=(meta >> 32) & 0xFFFF)


def position_from_fen(fen=START_FEN):
    parts = fen.split()
    array = np.zeros((8, 8), dtype=np.int8)
    row = col = 0
    for ch in parts[0]:
        if ch == "/":
            row, col = row + 1, 0
        elif ch.isdigit():
            col += int(ch)
        else:
            v = PIECE_SYMBOLS.index(ch.lower())
            array[row, col] = v if ch.isupper() else -v
            col += 1
    rights = parts[2] if len(parts) > 2 else "-"
    ep = parts[3] if len(parts) > 3 else "-"
    return pack_position(array, turn=(len(parts) < 2 or parts[1] == "w"),
                         castling=sum(b for c, b in zip("KQkq", (1, 2, 4, 8)) if c in rights),
                         ep_square=None if ep == "-" else (ord(ep[0]) - 97) + 8 * (int(ep[1]) - 1),
                         halfmove_clock=int(parts[4]) if len(parts) > 4 else 0,
                         fullmove_number=int(parts[5]) if len(parts) > 5 else 1)


# ------------------------------------------------------------------ device calls
def _dev():
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the chess kernels have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _pos_dev(pos):
    if torch.is_tensor(pos):
        t = pos.to(_dev()).contiguous()
    else:
        a = np.ascontiguousarray(np.asarray(pos, dtype=np.uint64))
        t = torch.from_numpy(a.view(np.int64)).to(_dev())
    assert t.dtype == torch.int64 and t.shape[-1] == 8
    return t


def _back(t, like, view=None):
    if torch.is_tensor(like) and like.is_cuda:
        return t
    a = t.cpu().numpy()
    return a.view(view) if view is not None else a


def chess_legal(pos):
    """-> (mask bool [n, 1880], count int32 [n], status int32 [n]); status bits as in include/az_b200.h."""
    p = _pos_dev(pos).reshape(-1, 8)
    n = p.shape[0]
    mask = torch.empty((n, MASK_WORDS), dtype=torch.int64, device=p.device)
    count = torch.empty(n, dtype=torch.int32, device=p.device)
    status = torch.empty(n, dtype=torch.int32, device=p.device)
    check(lib().az_chess_legal(_ptr(p), n, _ptr(mask), _ptr(count), _ptr(status), _stream()))
    bits = (mask.unsqueeze(-1) >> torch.arange(64, device=p.device)) & 1
    legal = bits.reshape(n, MASK_WORDS * 64)[:, :N_ACTIONS].bool()
    return _back(legal, pos), _back(count, pos), _back(status, pos)


def chess_play(pos, actions, keep_same_player=True):
    """Board.play for n boards -> (pos_out [n, 8], status [n]): 0 ongoing, 1 checkmate, 2 draw, -1 illegal action."""
    p = _pos_dev(pos).reshape(-1, 8)
    n = p.shape[0]
    a = torch.as_tensor(np.asarray(actions) if not torch.is_tensor(actions) else actions, dtype=torch.int32).to(p.device)
    assert a.numel() == n
    out = torch.empty_like(p)
    status = torch.empty(n, dtype=torch.int32, device=p.device)
    check(lib().az_chess_play(_ptr(p), _ptr(a.contiguous()), n, int(bool(keep_same_player)), _ptr(out), _ptr(status), _stream()))
    return _back(out, pos, np.uint64), _back(status, pos)


def chess_encode(pos, history=None, dtype=torch.float32):
    """Board.full_state for n boards -> [n, 8, 8, 118].  history: [n, 7, 8] older deque entries (oldest first; entries
    without the valid flag are zero padding) or None for the keep_same_player deque."""
    p = _pos_dev(pos).reshape(-1, 8)
    n = p.shape[0]
    h = None
    if history is not None:
        h = _pos_dev(history).reshape(n, 7, 8)
    out = torch.empty((n, 8, 8, PLANES), dtype=dtype, device=p.device)
    code = {torch.float32: native.AZ_F32, torch.bfloat16: native.AZ_BF16}[dtype]
    check(lib().az_chess_encode(_ptr(p), _ptr(h), n, code, _ptr(out), _stream()))
    return _back(out, pos) if dtype == torch.float32 else out


def chess_perft(pos, depth):
    """Move paths of length `depth` from each board along the self-play path -> uint64 [n]."""
    p = _pos_dev(pos).reshape(-1, 8)
    n = p.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=p.device)
    check(lib().az_chess_perft(_ptr(p), n, int(depth), _ptr(out), _stream()))
    return _back(out, pos, np.uint64)
