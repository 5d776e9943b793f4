"""CPU suite for the chess row: the mailbox oracle against the published perft counts, and the
DEVICE rules header (az_chess.cuh, compiled for the host by oracle/c/chess_hostcheck.cpp) against the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import chess_ref as cr


@pytest.mark.parametrize("fen", list(cr.PERFT))
def test_oracle_and_device_header_reproduce_published_perft(fen):
    s = cr.from_fen(fen)
    pos = cr.to_pos(s)
    if not s.turn:  # the self-play path always has white to move: mirror black-to-move positions first
        out = np.zeros(8, dtype=np.uint64)
        cr.hlib().hc_mirror(cr._u64p(pos), cr._u64p(out))
        pos = out
    for depth, want in enumerate(cr.PERFT[fen][:3], 1):
        assert cr.perft(s, depth) == want
        assert cr.host_perft_mirrored(pos, depth) == want
        if s.turn:
            assert cr.perft(s, depth, mirrored=True) == want


def test_deep_perft_startpos_and_kiwipete():
    assert cr.perft(cr.start_state(), 4) == 197281
    assert cr.host_perft_mirrored(cr.to_pos(cr.start_state()), 4) == 197281
    kiwi = cr.from_fen("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    assert cr.host_perft_mirrored(cr.to_pos(kiwi), 4) == 4085603


def test_action_list_is_the_references_procedure():
    acts = cr.all_possible_moves()
    assert len(acts) == 1880
    assert acts == sorted(set(acts), key=cr.move_key)
    h = cr.hlib()
    h.hc_act_move.restype = ctypes.c_int
    h.hc_act_index.restype = ctypes.c_int
    for i, (f, t, p) in enumerate(acts):
        code = h.hc_act_move(i)
        assert (code & 63, (code >> 6) & 63, cr.PROMO_LETTERS[code >> 12]) == (f, t, p)
        assert h.hc_act_index(f, t) + cr.PROMO_LETTERS.index(p) == i
    # first and last entries in the reference's order: a1 -> a2 ... h8 -> h7
    assert cr.uci(acts[0]) == "a1a2" and cr.uci(acts[-1]) == "h8h7"


def _lcg(seed):
    s = (seed * 0x9E3779B97F4A7C15 + 1) % 2 ** 64
    while True:
        s = (s * 6364136223846793005 + 1442695040888963407) % 2 ** 64
        yield s >> 33


@pytest.mark.parametrize("keep", [True, False])
def test_random_playouts_device_header_equals_oracle(keep):
    acts = cr.all_possible_moves()
    index = {m: i for i, m in enumerate(acts)}
    n_pos = n_end = 0
    for game in range(60):
        rng = _lcg(game + (1000 if keep else 0))
        s = cr.start_state()
        pos = cr.to_pos(s)
        for ply in range(300):
            moves = cr.legal(s)
            listed = sorted(index[m] for m in moves if m in index)
            got, in_check, unlisted = cr.host_legal(pos)
            assert got == listed, (game, ply)
            assert unlisted == len(moves) - len(listed)
            assert in_check == bool(cr.olib().co_in_check(ctypes.byref(s)))
            assert cr.host_status(pos) == cr.status(s)
            n_pos += 1
            if cr.status(s) != 0:
                n_end += 1
                break
            if not listed:  # only black promotions left (not in the action list): stop this playout
                break
            a = listed[next(rng) % len(listed)]
            s = cr.push(s, acts[a], keep_same_player=keep)
            code = acts[a][0] | acts[a][1] << 6 | cr.PROMO_LETTERS.index(acts[a][2]) << 12
            pos = cr.host_play(pos, code, keep)
            assert cr.states_equal(cr.from_pos(pos), s), (game, ply)
            if keep:
                assert s.turn == 1 and s.fullmove == 1
    assert n_pos > 5000 and n_end >= 3


def test_full_state_layout():
    s = cr.start_state()
    fs = cr.full_state(s, cr.selfplay_history())
    assert fs.shape == (8, 8, 118)
    cur = fs[:, :, 98:112]
    assert cur[7, 4, 6] == 1 and cur[0, 4, 7] == 1      # white king e1 -> plane 6, black king e8 -> plane 13 - 6
    assert cur[6, :, 1].sum() == 8 and cur[1, :, 12].sum() == 8  # pawns: white plane 1, black plane 12
    assert cur[2:6, :, 0].sum() == 32 and cur[:, :, 13].sum() == 0
    assert fs[:, :, :84].sum() == 0 and np.array_equal(fs[:, :, 84:98], cur)
    assert [fs[0, 0, 112 + i] for i in range(6)] == [1, 1, 1, 1, 1, 0]
    # Board() before any move: the constructor's deque, seven empty entries + the state (chess/board.py:37-40)
    assert cr.history_of(s) == [None] * 7
    fresh = cr.full_state(s, cr.history_of(s))
    assert fresh[:, :, :98].sum() == 0 and np.array_equal(fresh[:, :, 98:], fs[:, :, 98:])
    s2 = cr.push(s, (12, 28, ""), keep_same_player=True)  # e2e4 then mirror
    assert len(cr.history_of(s2)) == 7 and cr.history_of(s2)[6] is not None and cr.history_of(s2)[5] is None
    assert s2.ep == (20 ^ 56) and s2.halfmove == 0 and cr.array_of(s2)[3, 4] == -1
    assert cr.result(s2) is None
    # fool's mate on the mirrored path: the side to move (always "white") is mated -> -1 (chess/board.py:183-187)
    t = cr.start_state()
    for u in ("f2f3", "e2e4", "g2g4", "d1h5"):
        mv = (ord(u[0]) - 97 + 8 * (int(u[1]) - 1), ord(u[2]) - 97 + 8 * (int(u[3]) - 1), "")
        assert mv in cr.legal(t), u
        t = cr.push(t, mv, keep_same_player=True)
    assert cr.status(t) == 1 and cr.result(t) == -1


@pytest.mark.parametrize("fen,want", [
    ("8/8/8/4k3/8/8/8/4K3 w - - 0 1", 2),                       # K v K: insufficient material
    ("8/8/8/4k3/8/8/8/4KN2 w - - 0 1", 2),                      # K+N v K
    ("8/8/8/4k3/8/8/8/4KB2 w - - 0 1", 2),                      # K+B v K
    ("8/8/8/2b1k3/8/8/8/4KB2 w - - 0 1", 0),                    # bishops on opposite colours (c5 dark, f1 light): mate is possible
    ("8/8/8/3bk3/8/8/8/4KB2 w - - 0 1", 2),                     # all bishops on the same colour (d5 and f1 are light)
    ("8/8/8/4k3/8/8/8/3NKN2 w - - 0 1", 0),                     # K+N+N v K: python-chess does not call it insufficient
    ("8/8/8/4k3/8/8/4P3/4K3 w - - 0 1", 0),                     # a pawn is enough
    ("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1", 2),                      # stalemate (black to move, no moves, not in check)
    ("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1", 1),                      # checkmate
    ("4k3/8/8/8/8/8/4R3/4K3 w - - 149 80", 0),                  # halfmove clock 149: still running
    ("4k3/8/8/8/8/8/4R3/4K3 w - - 150 80", 2),                  # 75-move rule
    ("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 1 3", 1),  # fool's mate
])
def test_game_end_rules_oracle_and_device_header_agree(fen, want):
    s = cr.from_fen(fen)
    assert cr.status(s) == want
    assert cr.host_status(cr.to_pos(s)) == want


def test_castling_and_en_passant_bookkeeping():
    acts = cr.all_possible_moves()
    code = lambda u: next((m[0] | m[1] << 6 | cr.PROMO_LETTERS.index(m[2]) << 12, m) for m in acts if cr.uci(m) == u)  # noqa: E731
    # capturing the rook on h8 removes black's king-side right; moving the a1 rook removes white's queen-side right
    s = cr.from_fen("r3k2r/8/8/8/8/8/6B1/R3K2R w KQkq - 0 1")
    c, m = code("a1a2")
    assert cr.push(s, m).castling == 1 | 4 | 8 and cr.from_pos(cr.host_play(cr.to_pos(s), c, False)).castling == 1 | 4 | 8
    s2 = cr.from_fen("r3k2r/8/8/8/8/8/8/R3K2B w Qkq - 0 1")
    c, m = code("h1a8")
    after = cr.push(s2, m)
    assert after.castling == 2 | 4 and cr.states_equal(cr.from_pos(cr.host_play(cr.to_pos(s2), c, False)), after)
    # castling through an attacked square is illegal, castling out of check too; the rook may pass an attacked square
    s3 = cr.from_fen("4k3/8/8/8/8/5r2/8/R3K2R w KQ - 0 1")  # f1 attacked: no O-O, O-O-O fine
    legal = {cr.uci(m) for m in cr.legal(s3)}
    assert "e1c1" in legal and "e1g1" not in legal
    got, _, _ = cr.host_legal(cr.to_pos(s3))
    assert {cr.uci(acts[a]) for a in got} == legal
    s4 = cr.from_fen("4k3/8/8/8/8/1r6/8/R3K2R w KQ - 0 1")  # b1 attacked: the rook passes it, O-O-O is legal
    assert "e1c1" in {cr.uci(m) for m in cr.legal(s4)}
    assert "e1c1" in {cr.uci(acts[a]) for a in cr.host_legal(cr.to_pos(s4))[0]}
    # en passant that would expose the king along the rank is illegal (the classic perft trap)
    s5 = cr.from_fen("8/8/8/K2pP2r/8/8/8/4k3 w - d6 0 1")
    assert "e5d6" not in {cr.uci(m) for m in cr.legal(s5)}
    assert "e5d6" not in {cr.uci(acts[a]) for a in cr.host_legal(cr.to_pos(s5))[0]}
    s6 = cr.from_fen("8/8/8/3pP3/8/8/8/K3k3 w - d6 0 1")
    assert "e5d6" in {cr.uci(m) for m in cr.legal(s6)} and "e5d6" in {cr.uci(acts[a]) for a in cr.host_legal(cr.to_pos(s6))[0]}
    c, m = code("e5d6")
    after = cr.push(s6, m)
    assert after.sq[35] == 0 and after.sq[43] == 1 and cr.states_equal(cr.from_pos(cr.host_play(cr.to_pos(s6), c, False)), after)


def _golden(name):
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".json")) as fp:
        return json.load(fp)


def test_oracle_reproduces_the_committed_chess_fixtures():
    """tests/golden/chess_*.json (tests/golden/make_chess_golden.py) freeze the oracle's behaviour: playout fingerprint
    and whole MCTS games."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "make_chess_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_chess_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    want = _golden("chess_playouts")
    got = mk.playouts(want["games"], want["max_plies"])
    assert got == want
    for name in ("chess_mcts_hash_f64_100", "chess_mcts_uniform_f32_64", "chess_mcts_hash_f32_200"):
        w = _golden(name)
        assert mk.mcts(w["evaluator"], w["prior_mode"], w["sims"], w["max_plies"], w["greedy_idx"]) == w


def test_device_header_reproduces_the_playout_fixture():
    """The same 300 playouts through the DEVICE rules header compiled for the host: same moves, same final positions."""
    import hashlib

    want = _golden("chess_playouts")
    acts = cr.all_possible_moves()
    lines = []
    for g in range(want["games"]):
        rng = _lcg(g)
        pos = cr.to_pos(cr.start_state())
        picked = []
        for _ in range(want["max_plies"]):
            if cr.host_status(pos):
                break
            legal, _, _ = cr.host_legal(pos)
            a = legal[next(rng) % len(legal)]
            picked.append(a)
            m = acts[a]
            pos = cr.host_play(pos, m[0] | m[1] << 6 | cr.PROMO_LETTERS.index(m[2]) << 12, True)
        lines.append("{}|{}|{}\n".format(",".join(map(str, picked)), cr.host_status(pos), pos.tolist()))
    assert hashlib.sha256("".join(lines).encode()).hexdigest()[:16] == want["sha"]


def _divide_cases():
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "chess_perft_divide.json")) as fp:
        return json.load(fp)["cases"]


@pytest.mark.parametrize("case", _divide_cases(), ids=lambda c: f"{c['fen'].split()[0][:12]}-d{c['depth']}")
def test_published_perft_divide_per_root_move(case):
    """the published per-root-move counts (typed in from the published tables, see the fixture), for
    the mailbox oracle and for the DEVICE rules header compiled for the host.  A wrong count under one root move names
    the piece of the generator that is wrong, which a total cannot."""
    acts = cr.all_possible_moves()
    index = {m: i for i, m in enumerate(acts)}
    s = cr.from_fen(case["fen"])
    pos = cr.to_pos(s)
    d = case["depth"]
    assert sum(case["divide"].values()) == case["total"]
    got_oracle, got_header = {}, {}
    for m in cr.legal(s):
        got_oracle[cr.uci(m)] = cr.perft(cr.push(s, m, keep_same_player=False), d - 1)
        code = m[0] | m[1] << 6 | cr.PROMO_LETTERS.index(m[2]) << 12
        got_header[cr.uci(m)] = cr.host_perft_mirrored(cr.host_play(pos, code, True), d - 1)
    assert got_oracle == case["divide"]
    assert got_header == case["divide"]
    listed, _, unlisted = cr.host_legal(pos)
    assert unlisted == 0 and sorted(index[m] for m in cr.legal(s)) == listed
