"""GPU parity tests of the chess kernels through the C ABI: published perft counts, the mailbox
oracle on random playouts (legal masks, play + mirror, game end, clocks), the 118-plane encoder against the numpy
restatement of Board.full_state, and the drop-in custom_alphazero.chess classes."""
import ctypes

import numpy as np
import pytest

from oracle import chess_ref as cr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _chess():
    from az_b200 import chess

    return chess


def _lcg(seed):
    s = (seed * 0x9E3779B97F4A7C15 + 1) % 2 ** 64
    while True:
        s = (s * 6364136223846793005 + 1442695040888963407) % 2 ** 64
        yield s >> 33


def test_action_table_matches_the_oracle_list():
    chess = _chess()
    acts = cr.all_possible_moves()
    assert [chess.action_uci(a) for a in range(chess.N_ACTIONS)] == [cr.uci(m) for m in acts]


def test_perft_known_answers_on_the_device():
    chess = _chess()
    fens = [f for f in cr.PERFT if f.split()[1] == "w"]
    pos = np.stack([chess.position_from_fen(f) for f in fens])
    for depth in range(1, 5):
        got = chess.chess_perft(pos, depth)
        assert [int(g) for g in got] == [cr.PERFT[f][depth - 1] for f in fens], depth
    # black to move: the kernel mirrors first
    black = "r2q1rk1/pP1p2pp/Q4n2/bbp1p3/Np6/1B3NBn/pPPP1PPP/R3K2R b KQ - 0 1"
    assert int(chess.chess_perft(chess.position_from_fen(black)[None], 4)[0]) == 422333
    assert int(chess.chess_perft(chess.position_from_fen()[None], 5)[0]) == 4865609


def test_perft_split_over_many_roots():
    """Depth-6 start position count (119 060 324) as 8 902 independent depth-3 roots: the bench workload."""
    chess = _chess()
    front = chess.position_from_fen()[None]
    for _ in range(3):
        mask, count, status = chess.chess_legal(front)
        idx, act = np.nonzero(mask)
        front, st = chess.chess_play(front[idx], act.astype(np.int32), keep_same_player=True)
        assert (st >= 0).all()
    assert front.shape[0] == 8902
    assert int(chess.chess_perft(front, 3).sum()) == 119060324


@pytest.mark.parametrize("keep", [True, False])
def test_random_playouts_against_the_mailbox_oracle(keep):
    chess = _chess()
    acts = cr.all_possible_moves()
    index = {m: i for i, m in enumerate(acts)}
    G = 96
    states = [cr.start_state() for _ in range(G)]
    rngs = [_lcg(g + (500 if keep else 0)) for g in range(G)]
    pos = np.stack([cr.to_pos(s) for s in states])
    alive = list(range(G))
    n_checked = n_over = 0
    for ply in range(220):
        if not alive:
            break
        mask, count, status = chess.chess_legal(pos[alive])
        chosen, nxt = [], []
        for row, g in enumerate(alive):
            s = states[g]
            moves = cr.legal(s)
            listed = sorted(index[m] for m in moves if m in index)
            assert list(np.nonzero(mask[row])[0]) == listed, (g, ply)
            assert int(count[row]) == len(moves)
            st = cr.status(s)
            assert int(status[row]) & 3 == st
            assert bool(int(status[row]) & 4) == bool(cr.olib().co_in_check(ctypes.byref(s)))
            assert int(status[row]) >> 8 == len(moves) - len(listed)
            n_checked += 1
            if st != 0 or not listed:
                n_over += st != 0
                continue
            a = listed[next(rngs[g]) % len(listed)]
            chosen.append(a)
            nxt.append(g)
        if not nxt:
            break
        out, st_new = chess.chess_play(pos[nxt], np.array(chosen, dtype=np.int32), keep_same_player=keep)
        for row, (g, a) in enumerate(zip(nxt, chosen)):
            states[g] = cr.push(states[g], acts[a], keep_same_player=keep)
            assert cr.states_equal(cr.from_pos(out[row]), states[g]), (g, ply)
            assert int(st_new[row]) == cr.status(states[g])
            pos[g] = out[row]
        alive = nxt
    assert n_checked > 8000 and n_over >= 2


def test_illegal_actions_are_flagged_not_played():
    chess = _chess()
    start = chess.position_from_fen()
    bad = chess.uci_action("e2e5")   # in the action list (queen line) but illegal here
    good = chess.uci_action("e2e4")
    out, st = chess.chess_play(np.stack([start, start, start]), np.array([bad, good, -1], dtype=np.int32))
    assert list(st) == [-1, 0, -1]
    assert (out[0] == start).all() and (out[2] == start).all() and not (out[1] == start).all()
    assert chess.chess_legal(np.zeros((0, 8), dtype=np.uint64))[0].shape == (0, 1880)


def test_encoder_matches_full_state():
    chess = _chess()
    acts = cr.all_possible_moves()
    index = {m: i for i, m in enumerate(acts)}
    rng = _lcg(7)
    s = cr.start_state()
    samples = []
    for ply in range(60):
        moves = [m for m in cr.legal(s) if m in index]
        if cr.status(s) != 0 or not moves:
            break
        samples.append(s.copy())
        s = cr.push(s, moves[next(rng) % len(moves)], keep_same_player=True)
    pos = np.stack([cr.to_pos(x) for x in samples])
    # (i) the self-play deque (history=None)
    got = chess.chess_encode(pos)
    assert got.shape == (len(samples), 8, 8, 118) and got.dtype == np.float32
    for i, x in enumerate(samples):
        want = cr.full_state(x, cr.history_of(x))  # ply 0 = Board() itself: seven empty entries
        assert np.array_equal(got[i].astype(np.float64), want), i
    assert not got[0][:, :, :98].any() and got[1][:, :, 84:98].any() and not got[1][:, :, :84].any()
    # (ii) an explicit 7-entry history: sliding window over the same game, with a repetition flag and padding
    hist = np.zeros((len(samples), 7, 8), dtype=np.uint64)
    wants = []
    for i, x in enumerate(samples):
        older = []
        for k in range(7):
            j = i - 7 + k
            if j < 0:
                older.append(None)
            else:
                rep = (j % 5 == 0)
                hist[i, k] = cr.to_pos(samples[j], repetition=rep, valid=True)
                older.append(cr.state_planes(samples[j], repetition=rep))
        wants.append(cr.full_state(x, older))
    got = chess.chess_encode(pos, hist)
    for i in range(len(samples)):
        assert np.array_equal(got[i].astype(np.float64), wants[i]), i
    # bf16 output holds the same values (0/1 planes and small integers are exact in bf16)
    got16 = chess.chess_encode(pos, hist, dtype=torch.bfloat16).float().cpu().numpy()
    assert np.array_equal(got16, got)


def test_drop_in_board_follows_the_reference_flow():
    from custom_alphazero.chess.board import Board
    from custom_alphazero.chess.move import Move
    from custom_alphazero.chess.utils import get_all_possible_moves

    all_moves = get_all_possible_moves()
    assert len(all_moves) == 1880 and all_moves == sorted(all_moves)
    b = Board()
    assert b.turn is True and b.fullmove_number == 1 and len(b.moves) == 20
    assert b.full_state.shape == (8, 8, 118)
    mask = b.legal_moves_mask(all_moves)
    assert mask.sum() == 20 and all_moves[int(np.nonzero(mask)[0][0])] in b.moves
    s = cr.start_state()
    hist = [None] * 7
    assert np.array_equal(b.full_state, cr.full_state(s, hist))
    # fool's mate on the keep_same_player path (every move is "white's")
    for u in ("f2f3", "e2e4", "g2g4", "d1h5"):
        assert b.get_result() is None
        mv = Move(uci=u)
        assert mv in b.moves
        s = cr.push(s, (mv.pos_from[0] + 8 * mv.pos_from[1], mv.pos_to[0] + 8 * mv.pos_to[1], mv.pos_to[2]), True)
        b.play(mv, keep_same_player=True)
        assert np.array_equal(b.array, cr.array_of(s)) and b.turn is True and b.fullmove_number == 1
        assert np.array_equal(b.full_state, cr.full_state(s, cr.selfplay_history()))
    assert b.is_game_over() and b.result() == "0-1" and b.get_result() == -1 and b.get_result(keep_same_player=True) == 1
    # without keep_same_player the turn alternates and the deque slides
    c = Board()
    c.play(Move(uci="e2e4"))
    assert c.turn is False and c.ep_square == 20 and len(c.moves) == 20
    c.play(Move(uci="e7e5"))
    assert c.turn is True and c.fullmove_number == 2 and c.array[3, 4] == -1 and c.array[4, 4] == 1
    d = c.play(Move(uci="g1f3"), on_copy=True)
    assert d is not c and c.array[5, 5] == 0 and d.array[5, 5] == 2
    with pytest.raises(ValueError):
        c.play(Move(uci="e1e3"))


def test_integration_md_chess_stub_runs():
    """The ctypes stub printed in INTEGRATION.md for chess is executed as written (only the library path is made
    absolute) on an object with python-chess's Board attributes."""
    import os
    import re
    import types

    chess = _chess()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    stub = next(b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "def pack(board)" in b)
    stub = stub.replace('ctypes.CDLL("libaz_b200.so")', 'ctypes.CDLL("{}")'.format(os.path.join(root, "custom-alphazero_b200", "libaz_b200.so")))
    ns = {}
    exec(stub, ns)

    def fake_board(fen):  # what python-chess exposes: piece bitboards, occupied_co, castling_rights as a rook-square mask
        w = [int(x) for x in chess.position_from_fen(fen)]
        f = chess.unpack_position(np.array(w, dtype=np.uint64))
        occ = w[0] | w[1] | w[2] | w[3] | w[4] | w[5]
        rights = sum(1 << sq for bit, sq in ((1, 7), (2, 0), (4, 63), (8, 56)) if f["castling"] & bit)
        return types.SimpleNamespace(pawns=w[0], knights=w[1], bishops=w[2], rooks=w[3], queens=w[4], kings=w[5],
                                     occupied_co={True: w[6], False: occ & ~w[6]}, castling_rights=rights,
                                     ep_square=f["ep_square"], turn=bool(f["turn"]), halfmove_clock=f["halfmove_clock"],
                                     fullmove_number=f["fullmove_number"])

    fens = [chess.START_FEN, "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
            "rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 1 3"]
    boards = [fake_board(f) for f in fens]
    for b, f in zip(boards, fens):
        want = chess.position_from_fen(f)
        want[7] &= np.uint64(~(3 << 48) & (2 ** 64 - 1))  # the valid / repetition flags only matter for history entries
        assert np.array_equal(ns["pack"](b), want)
    mask, count, status = ns["legal_moves_mask"](boards)
    torch.cuda.synchronize()
    assert count.tolist() == [20, 48, 0] and [s & 3 for s in status.tolist()] == [0, 0, 1]
    planes = ns["full_state"](boards)
    assert np.array_equal(planes.cpu().numpy(), chess.chess_encode(np.stack([chess.position_from_fen(f) for f in fens])))


def test_published_perft_divide_on_the_device():
    """The published per-root-move perft counts (tests/golden/chess_perft_divide.json, typed in from the published tables)
    through az_chess_legal / az_chess_play / az_chess_perft."""
    import json
    import os

    chess = _chess()
    with open(os.path.join(os.path.dirname(__file__), "golden", "chess_perft_divide.json")) as fp:
        cases = json.load(fp)["cases"]
    for case in cases:
        root = chess.position_from_fen(case["fen"])[None]
        mask, count, status = chess.chess_legal(root)
        _, act = np.nonzero(mask)
        children, st = chess.chess_play(np.repeat(root, len(act), axis=0), act.astype(np.int32), keep_same_player=True)
        assert (st >= 0).all()
        counts = chess.chess_perft(children, case["depth"] - 1)
        got = {chess.action_uci(int(a)): int(c) for a, c in zip(act, counts)}
        assert got == case["divide"], case["fen"]
