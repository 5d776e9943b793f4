"""Pins oracle/ref_port.py (Python restatement) to the golden vectors that
tests/golden/make_golden.py produced from the unmodified reference."""
import hashlib

import numpy as np
import pytest

from oracle import evaluators, ref_port
from tests.helpers import golden_names, lcg_next, lcg_start, load_golden, trace_sha16


def rules_of(case):
    return ref_port.Rules(width=case["W"], height=case["H"], n=case["n"], gravity=case["gravity"])


@pytest.mark.parametrize("name", golden_names("env_"))
def test_port_env_fingerprint(name):
    case = load_golden(name)
    rules = rules_of(case)
    games = case["games"] if case["W"] * case["H"] <= 42 else min(case["games"], 60)
    sha = hashlib.sha256()
    wins = draws = plies = 0
    for g in range(games):
        s = lcg_start(g)
        b = ref_port.RefBoard(rules)
        picked = []
        while not b.over:
            moves = b.legal_moves()
            s = lcg_next(s)
            idx = (s >> 33) % len(moves)
            picked.append(idx)
            b.play(moves[idx], keep_same_player=True)
        res = b.result(keep_same_player=True)
        wins += res == 1
        draws += res == 0
        plies += len(picked)
        sha.update("{}|{}|{}\n".format(",".join(map(str, picked)), res, b.text()).encode())
        if g < len(case["detail"]):
            d = case["detail"][g]
            assert picked == d["picked"] and res == d["result"] and b.text() == d["repr"]
            assert [float(x) for x in b.full_state().sum(axis=(0, 1))] == d["full_state_channel_sums"]
    if games == case["games"]:
        assert (wins, draws, plies) == (case["wins"], case["draws"], case["plies"])
        assert sha.hexdigest()[:16] == case["sha16"]


@pytest.mark.parametrize("name", golden_names("search_"))
def test_port_single_search(name):
    case = load_golden(name)
    rules = rules_of(case)
    b = ref_port.RefBoard(rules)
    actions = rules.all_actions()
    for a in case.get("prefix", []):
        b.play(actions[a], keep_same_player=True)
    assert b.text() == case["repr"] and b.plies == case["fullmove_number"] and b.to_move == case["turn"]
    s = ref_port.RefSearch(b, evaluators.make(case["evaluator"], rules.n_actions))
    s.search(case["sims"])
    edges = s.current.edges
    assert [e.n for e in edges] == case["edge_N"]
    assert [float(e.w) for e in edges] == case["edge_W"]
    assert [float(e.prior) for e in edges] == case["edge_P"]
    assert [e.child.board.over for e in edges] == case["children_terminal"]
    assert s.evals == case["evaluator_calls"]


def _check_game(case, out):
    got = out["trace"]
    want = case["plies"]
    for i, (g, w) in enumerate(zip(got, want)):
        assert g["N"] == w["N"], f"ply {i}"
        assert g["move"] == w["move"], f"ply {i}"
    assert len(got) == case["n_plies"]
    assert out["result"] == case["result"]
    assert trace_sha16(got) == case["sha16"]
    assert hashlib.sha256(out["states"].tobytes()).hexdigest()[:16] == case["states_sha16"]
    np.testing.assert_array_equal(out["policies"], np.asarray([p["policy"] for p in want]))


FAST_GAMES = [n for n in golden_names("game_") if "9x9" not in n and "800" not in n]
SLOW_GAMES = [n for n in golden_names("game_") if n not in FAST_GAMES]


@pytest.mark.parametrize("name", FAST_GAMES)
def test_port_full_game(name):
    case = load_golden(name)
    rules = rules_of(case)
    out = ref_port.play_game(rules, evaluators.make(case["evaluator"], rules.n_actions), case["sims"],
                             uniforms=case.get("uniforms"))
    _check_game(case, out)


@pytest.mark.slow
@pytest.mark.parametrize("name", ["game_6x7_800_hash"])
def test_port_full_game_800(name):
    case = load_golden(name)
    rules = rules_of(case)
    out = ref_port.play_game(rules, evaluators.make(case["evaluator"], rules.n_actions), case["sims"])
    _check_game(case, out)


def test_port_rewards_alternate():
    case = load_golden("game_6x7_250_hash")
    rules = rules_of(case)
    out = ref_port.play_game(rules, evaluators.make("hash", 7), 250)
    r = out["rewards"]
    assert r[-1] == 1 and all(r[-1 - i] == (1 if i % 2 == 0 else -1) for i in range(len(r)))


# ------------------------------------------------------------------ C restatement
from oracle import c_oracle  # noqa: E402


def c_rules_of(case):
    return c_oracle.make_rules(case["W"], case["H"], case["n"], case["gravity"])


@pytest.mark.parametrize("name", golden_names("env_"))
def test_c_env_fingerprint(name):
    case = load_golden(name)
    rules = c_rules_of(case)
    symbols = {-1: "O", 0: ".", 1: "X"}
    sha = hashlib.sha256()
    wins = draws = plies = 0
    for g in range(case["games"]):
        picked, res, cells = c_oracle.env_playout(rules, lcg_start(g))
        text = "\n".join("".join(symbols[int(v)] for v in row) for row in cells)
        wins += res == 1
        draws += res == 0
        plies += len(picked)
        sha.update("{}|{}|{}\n".format(",".join(map(str, picked)), res, text).encode())
        if g < len(case["detail"]):
            d = case["detail"][g]
            assert picked == d["picked"] and res == d["result"] and text == d["repr"]
    assert (wins, draws, plies) == (case["wins"], case["draws"], case["plies"])
    assert sha.hexdigest()[:16] == case["sha16"]


@pytest.mark.parametrize("name", golden_names("search_"))
def test_c_single_search(name):
    case = load_golden(name)
    out = c_oracle.search_once(c_rules_of(case), case.get("prefix", []), case["sims"], case["evaluator"])
    assert out["actions"] == case["edge_actions"]
    assert out["N"] == case["edge_N"]
    assert out["W"] == case["edge_W"]
    assert out["P"] == case["edge_P"]
    # the reference memoises evaluations by position text (mcts.py:123-124); the C port counts expansions
    assert out["evals"] >= case["evaluator_calls"]


@pytest.mark.parametrize("name", golden_names("game_"))
def test_c_full_game(name):
    case = load_golden(name)
    rules = c_rules_of(case)
    out = c_oracle.play_game(rules, case["sims"], case["evaluator"], uniforms=case.get("uniforms"))
    want = case["plies"]
    assert len(out["moves"]) == case["n_plies"] and out["result"] == case["result"]
    for t, w in enumerate(want):
        assert out["moves"][t] == w["move"], f"ply {t}"
        got = [int(out["visits"][t][a]) for a in w["actions"]]
        assert got == w["N"], f"ply {t}"
        assert int((out["visits"][t] >= 0).sum()) == len(w["actions"])
        np.testing.assert_array_equal(out["policies"][t], np.asarray(w["policy"]))
    assert out["evals"] >= case["evaluator_calls"]  # reference memoises evaluations, the C port does not


def test_c_normalise_matches_numpy():
    rng = np.random.default_rng(5)
    for k in list(range(1, 82)) * 3:
        p = rng.random(k)
        if k % 7 == 0:
            p[rng.integers(0, k)] = 0.0
        want = p / p.sum()
        np.testing.assert_array_equal(c_oracle.normalise(p), want)
        p32 = p.astype(np.float32)
        want32 = (p32 / p32.sum()).astype(np.float64)
        np.testing.assert_array_equal(c_oracle.normalise(p32.astype(np.float64), c_oracle.PRIOR_F32), want32)
    np.testing.assert_array_equal(c_oracle.normalise(np.zeros(5)), np.full(5, 1 / 5))


def test_c_pow_half_is_python_pow():
    lib = c_oracle.lib()
    for n in list(range(0, 70000)) + [10**6 + i for i in range(1000)]:
        assert lib.azo_pow_half(n) == n**0.5
    assert lib.azo_pow_half(2921) != float(np.sqrt(2921.0))  # Q3: pow, not sqrt


def test_c_callback_evaluator_equals_builtin():
    rules = c_oracle.make_rules(7, 6, 4, True)
    a = c_oracle.play_game(rules, 60, "hash")
    b = c_oracle.play_game(rules, 60, "callback", callback=evaluators.hash_evaluator(7))
    np.testing.assert_array_equal(a["visits"], b["visits"])
    np.testing.assert_array_equal(a["moves"], b["moves"])
