"""GPU edge cases through the C ABI: empty batches, a full finished-game ring, exhausted pools and tables
(loud, never silent), drawn games in the sample decoder."""
import numpy as np
import pytest

from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _mods():
    from az_b200 import engine, env, native, selfplay

    return engine, env, native, selfplay


def test_empty_batches_are_accepted():
    engine, env, native, selfplay = _mods()
    rules = engine.Rules(7, 6, 4, True)
    out, status = env.env_play(rules, np.zeros((0, 6, 7), np.int8), np.zeros(0, np.int32))
    assert out.shape == (0, 6, 7) and status.shape == (0,)
    assert env.env_legal(rules, np.zeros((0, 6, 7), np.int8)).shape == (0, 7)
    assert env.env_encode(rules, np.zeros((0, 6, 7), np.int8)).shape == (0, 6, 7, 4)
    eng = engine.TreeEngine(rules, n_trees=2, sims_per_move=8, eval_mode="uniform", prior_mode="f64")
    s, p, v = selfplay.decode_samples(rules, {k: t[:0] for k, t in {
        "game_id": eng.view("fin_game_id"), "len": eng.view("fin_len"), "result": eng.view("fin_result"),
        "visits": eng.view("fin_visits"), "action": eng.view("fin_action"), "board": eng.view("fin_board")}.items()})
    assert s.shape == (0, 6, 7, 4) and p.shape == (0, 7) and v.shape == (0,)


def test_full_ring_stalls_and_recovers():
    """Finished games wait (AZ_PHASE_STALLED) when the ring is full and go on once the host drained it."""
    engine, env, native, selfplay = _mods()
    case = load_golden("game_3x3ng_n3_40_hash")
    rules = engine.Rules(3, 3, 3, False)
    eng = engine.TreeEngine(rules, n_trees=4, sims_per_move=40, eval_mode="hash", prior_mode="f64", games_target=10,
                            auto_restart=True, fin_capacity=2)
    got = []
    for _ in range(400):
        eng.search()
        eng.play()
        ph = eng.phases().cpu().numpy()
        if (ph == native.AZ_PHASE_STALLED).any() or int(eng.view("fin_count")[0]) == 2:
            fin = eng.drain_finished()
            got.extend(fin["game_id"].tolist())
        if (ph == native.AZ_PHASE_IDLE).all():
            break
    fin = eng.drain_finished()
    got.extend(fin["game_id"].tolist())
    eng.check_status()
    assert sorted(got) == list(range(10))
    assert eng.totals()["games"] == 10 and eng.totals()["moves"] == 10 * case["n_plies"]


def test_pool_and_table_exhaustion_are_loud():
    engine, env, native, selfplay = _mods()
    rules = engine.Rules(7, 6, 4, True)
    eng = engine.TreeEngine(rules, n_trees=2, sims_per_move=200, eval_mode="uniform", prior_mode="f64", node_capacity=256)
    eng.search()
    with pytest.raises(native.NativeError, match="node pool exhausted"):
        eng.check_status()
    eng = engine.TreeEngine(rules, n_trees=2, sims_per_move=200, eval_mode="uniform", prior_mode="f64", pow_lut_len=50)
    eng.search()
    with pytest.raises(native.NativeError, match="pow-half table"):
        eng.check_status()
    eng = engine.TreeEngine(rules, n_trees=2, sims_per_move=200, eval_mode="uniform", prior_mode="f64")
    eng.play()  # nothing searched yet: trees are not READY, nothing happens
    eng.check_status()
    assert int(eng.view("rec_len").sum()) == 0


def test_decoder_rewards_draws_and_greedy_targets():
    """Sample decoder on a mix of decisive and drawn 3x3 games: rewards alternate backwards from the last ply
    (self_play.py:71-78), drawn games give zeros and are dropped with exclude_null_games (:155-162), greedy plies
    give one-hot targets at the first maximum in EDGE order (mcts.py:189-192)."""
    engine, env, native, selfplay = _mods()
    rules = engine.Rules(3, 3, 3, False)
    G = 64
    rng = np.random.RandomState(3)
    eng = engine.TreeEngine(rules, n_trees=G, sims_per_move=24, eval_mode="hash", prior_mode="f64",
                            move_mode="host_uniforms", index_move_greedy=4)
    eng.set_uniforms(rng.random_sample((G, 9)))
    for _ in range(10):
        eng.search()
        eng.play()
    torch.cuda.synchronize()
    fin_host = eng.drain_finished()
    assert len(fin_host["len"]) == G
    results = fin_host["result"]
    assert (results == 0).any() and (results == 1).any()  # both draws and wins occur on 3x3
    n = G
    e = eng
    fin = {"game_id": e.view("fin_game_id")[:n], "len": e.view("fin_len")[:n], "result": e.view("fin_result")[:n],
           "visits": e.view("fin_visits")[:n], "action": e.view("fin_action")[:n], "board": e.view("fin_board")[:n]}
    states, policies, values = selfplay.decode_samples(rules, fin)
    order = np.argsort(fin_host["game_id"])
    k = 0
    for g in order:
        L, res = int(fin_host["len"][g]), int(fin_host["result"][g])
        for i in range(L):
            assert values[k] == (res if (L - 1 - i) % 2 == 0 else -res)
            vis = fin_host["visits"][g][i]
            legal = vis >= 0
            # legal <=> empty cell of the parent state (action = x * H + y)
            empty = np.transpose(states[k][..., 0], (1, 0)).reshape(-1) == 1
            assert (legal == empty).all()
            if i >= 4:  # greedy: one-hot at the first maximum in row-major (board) order
                cells = sorted(np.nonzero(legal)[0], key=lambda a: (a % 3, a // 3))
                best = max(cells, key=lambda a: (vis[a], -cells.index(a)))
                want = np.zeros(9)
                want[best] = 1.0
            else:
                want = np.where(legal, vis, 0).astype(np.float64)
                want = want / want.sum() if want.sum() > 0 else legal / legal.sum()
            np.testing.assert_array_equal(policies[k], want)
            k += 1
    assert k == len(values)
    s2, p2, v2 = selfplay.decode_samples(rules, fin, exclude_null_games=True)
    assert len(v2) == int(fin_host["len"][results != 0].sum()) and (v2 != 0).all()
