"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libaz_b200.so; expected values are the golden vectors from the unmodified reference and the C oracle."""
import hashlib

import numpy as np
import pytest

from tests.helpers import golden_names, lcg_next, lcg_start, load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _engine_mod():
    from az_b200 import engine, env

    return engine, env


def rules_of(case):
    engine, _ = _engine_mod()
    return engine.Rules(case["W"], case["H"], case["n"], case["gravity"])


SYMBOLS = {-1: "O", 0: ".", 1: "X"}


def text_of(cells):
    return "\n".join("".join(SYMBOLS[int(v)] for v in row) for row in cells)


# ------------------------------------------------------------------ K2 / K3 environment kernels
@pytest.mark.parametrize("name", golden_names("env_"))
def test_env_kernels_reproduce_reference_playouts(name):
    engine, env = _engine_mod()
    case = load_golden(name)
    rules = rules_of(case)
    G = case["games"]
    cells = np.zeros((G, rules.height, rules.width), dtype=np.int8)
    lcg = [lcg_start(g) for g in range(G)]
    picked = [[] for _ in range(G)]
    result = [None] * G
    final = [None] * G
    live = list(range(G))
    while live:
        legal = env.env_legal(rules, cells[live])
        actions = []
        for row, g in enumerate(live):
            order = env.board_order_actions(rules, legal[row])
            lcg[g] = lcg_next(lcg[g])
            idx = (lcg[g] >> 33) % len(order)
            picked[g].append(int(idx))
            actions.append(order[idx])
        out, status = env.env_play(rules, cells[live], np.asarray(actions, dtype=np.int32))
        assert (status >= 0).all()
        nxt = []
        for row, g in enumerate(live):
            cells[g] = out[row]
            if status[row] == 0:
                nxt.append(g)
            else:
                result[g] = 1 if status[row] == 1 else 0
                final[g] = out[row].copy()
        live = nxt
    sha = hashlib.sha256()
    for g in range(G):
        sha.update("{}|{}|{}\n".format(",".join(map(str, picked[g])), result[g], text_of(final[g])).encode())
    for g, d in enumerate(case["detail"]):
        assert picked[g] == d["picked"] and result[g] == d["result"] and text_of(final[g]) == d["repr"]
    assert sum(r == 1 for r in result) == case["wins"] and sum(r == 0 for r in result) == case["draws"]
    assert sum(len(p) for p in picked) == case["plies"]
    assert sha.hexdigest()[:16] == case["sha16"]
    # K3: channel sums of the final positions (full_state, board.py:83-98)
    states = env.env_encode(rules, np.stack(final[: len(case["detail"])]))
    for g, d in enumerate(case["detail"]):
        assert [float(x) for x in states[g].sum(axis=(0, 1))] == d["full_state_channel_sums"]


def test_env_illegal_action_is_reported():
    engine, env = _engine_mod()
    rules = engine.Rules(7, 6, 4, True)
    cells = np.zeros((2, 6, 7), dtype=np.int8)
    cells[0, :, 3] = [1, -1, 1, -1, 1, -1]  # column 3 full
    out, status = env.env_play(rules, cells, np.asarray([3, 3], dtype=np.int32))
    assert status.tolist() == [-1, 0]
    np.testing.assert_array_equal(out[0], cells[0])
    assert out[1][5, 3] == -1  # the stone just played belongs to the opponent after mirroring


# ------------------------------------------------------------------ single searches (root N, W, P)
def _root_after_prefix(env, rules, prefix):
    cells = np.zeros((1, rules.height, rules.width), dtype=np.int8)
    for a in prefix:
        cells, status = env.env_play(rules, cells, np.asarray([a], dtype=np.int32))
        assert status[0] == 0
    return cells


@pytest.mark.parametrize("name", golden_names("search_"))
def test_single_search_root_statistics(name):
    engine, env = _engine_mod()
    case = load_golden(name)
    rules = rules_of(case)
    eng = engine.TreeEngine(rules, n_trees=3, sims_per_move=case["sims"], eval_mode=case["evaluator"], prior_mode="f64")
    prefix = case.get("prefix", [])
    cells = _root_after_prefix(env, rules, prefix)
    eng.set_roots([0, 1, 2], np.repeat(cells, 3, axis=0), [len(prefix)] * 3)
    eng.begin_search(case["sims"])
    eng.search()
    torch.cuda.synchronize()
    eng.check_status()
    for t in range(3):
        n, w, p = eng.root_stats(t)
        assert n == case["edge_N"]
        assert w == case["edge_W"]
        assert p == case["edge_P"]
    assert eng.totals()["sims"] == 3 * case["sims"]


# ------------------------------------------------------------------ full games
def _play_games(eng, max_plies):
    """search + play until every tree is idle; returns the finished games sorted by game id."""
    for _ in range(max_plies + 1):
        eng.search()
        eng.play()
        if int((eng.phases() != 0).sum()) == 0:
            break
    torch.cuda.synchronize()
    eng.check_status()
    fin = eng.drain_finished()
    order = np.argsort(fin["game_id"])
    return {k: v[order] for k, v in fin.items()}


def _check_against_golden(case, fin, g):
    T = fin["len"][g]
    assert T == case["n_plies"] and fin["result"][g] == case["result"]
    for t, w in enumerate(case["plies"]):
        assert (fin["action"][g][t] & 0xFFFF) == w["move"], f"ply {t}"
        assert bool(fin["action"][g][t] >> 16) == w["greedy"]
        assert [int(fin["visits"][g][t][a]) for a in w["actions"]] == w["N"], f"ply {t}"
        assert int((fin["visits"][g][t] >= 0).sum()) == len(w["actions"])


@pytest.mark.parametrize("name", golden_names("game_"))
def test_full_game_visit_counts_bit_exact(name):
    engine, _ = _engine_mod()
    case = load_golden(name)
    rules = rules_of(case)
    seeded = case.get("seed") is not None
    n_trees = 5
    eng = engine.TreeEngine(rules, n_trees=n_trees, sims_per_move=case["sims"], eval_mode=case["evaluator"],
                            prior_mode="f64", move_mode="host_uniforms" if seeded else "argmax")
    if seeded:
        eng.set_uniforms(np.tile(np.asarray(case["uniforms"][: rules.max_plies]), (n_trees, 1)))
    fin = _play_games(eng, rules.max_plies)
    assert len(fin["len"]) == n_trees
    for g in range(n_trees):
        _check_against_golden(case, fin, g)


# ------------------------------------------------------------------ external evaluator path (k_step)
@pytest.mark.parametrize("name", ["game_6x7_120_hash_seed99", "game_5x5ng_n3_60_hash", "game_3x3ng_n3_40_hash"])
def test_external_evaluator_path_bit_exact(name):
    """Drives az_step with the hash evaluator computed on the HOST from the states the kernel encodes
    (float64 priors, exactly what serving/factory.py:55 hands the reference): same golden trace."""
    from oracle import evaluators

    engine, _ = _engine_mod()
    case = load_golden(name)
    rules = rules_of(case)
    seeded = case.get("seed") is not None
    T, A = 2, rules.n_actions
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=case["sims"], eval_mode="external", prior_mode="f64",
                            move_mode="host_uniforms" if seeded else "argmax", max_free_sims=3)
    if seeded:
        eng.set_uniforms(np.tile(np.asarray(case["uniforms"][: rules.max_plies]), (T, 1)))
    states = torch.zeros((T, rules.height, rules.width, 4), dtype=torch.float32, device="cuda")
    valid = torch.zeros(T, dtype=torch.int32, device="cuda")
    priors = torch.zeros((T, A), dtype=torch.float64, device="cuda")
    values = torch.zeros(T, dtype=torch.float64, device="cuda")
    f = evaluators.hash_evaluator(A)
    first = True
    for _ in range(200000):
        eng.step(None if first else priors, None if first else values, states, valid)
        first = False
        v = valid.cpu().numpy()
        if v.any():
            st = states.cpu().numpy()
            p = np.zeros((T, A))
            val = np.zeros(T)
            for t in range(T):
                if v[t]:
                    p[t], val[t] = f(st[t])
            priors.copy_(torch.from_numpy(p))
            values.copy_(torch.from_numpy(val))
        ph = eng.phases().cpu().numpy()
        if (ph == 2).any():
            eng.play()
        if (ph == 0).all():
            break
    eng.check_status()
    fin = eng.drain_finished()
    assert len(fin["len"]) == T
    for g in range(T):
        _check_against_golden(case, fin, g)


# ------------------------------------------------------------------ many different games vs the C oracle
@pytest.mark.parametrize("capacity", [30000, 45000])
def test_on_demand_compaction_keeps_results(capacity):
    """A pool half too small for in-place re-rooting forces the breadth-first compaction path on most
    moves; the golden 800-simulation game must still be reproduced exactly."""
    engine, _ = _engine_mod()
    case = load_golden("game_6x7_800_hash")
    rules = rules_of(case)
    eng = engine.TreeEngine(rules, n_trees=3, sims_per_move=800, eval_mode="hash", prior_mode="f64",
                            node_capacity=capacity)
    fin = _play_games(eng, rules.max_plies)
    for g in range(3):
        _check_against_golden(case, fin, g)
    assert eng.totals()["reroot_nodes"] > 0  # compaction really ran


def test_batch_of_different_games_matches_c_oracle():
    from oracle import c_oracle

    engine, _ = _engine_mod()
    rules = engine.Rules(7, 6, 4, True)
    T, sims = 64, 200
    rng = np.random.RandomState(2024)
    uniforms = rng.random_sample((T, rules.max_plies))
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=sims, eval_mode="hash", prior_mode="f64",
                            move_mode="host_uniforms")
    eng.set_uniforms(uniforms)
    fin = _play_games(eng, rules.max_plies)
    assert len(fin["len"]) == T
    crules = c_oracle.make_rules(7, 6, 4, True)
    lengths = set()
    for g in range(T):
        want = c_oracle.play_game(crules, sims, "hash", uniforms=uniforms[g])
        n = len(want["moves"])
        lengths.add(n)
        assert fin["len"][g] == n and fin["result"][g] == want["result"]
        np.testing.assert_array_equal(fin["action"][g][:n] & 0xFFFF, want["moves"])
        np.testing.assert_array_equal(fin["visits"][g][:n], want["visits"])
    assert len(lengths) > 3  # the games really differ


def test_full_size_batch_reproduces_golden_in_every_tree():
    """BASELINE config C2 size: 4096 concurrent games x 800 simulations per move, uniform evaluator,
    deterministic play: every one of the 4096 trees must reproduce the reference's golden game."""
    engine, _ = _engine_mod()
    case = load_golden("game_6x7_800_uniform")
    rules = rules_of(case)
    T = 4096
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=800, eval_mode="uniform", prior_mode="f64")
    fin = _play_games(eng, rules.max_plies)
    assert len(fin["len"]) == T
    assert (fin["len"] == case["n_plies"]).all() and (fin["result"] == case["result"]).all()
    assert (fin["visits"] == fin["visits"][0]).all() and (fin["action"] == fin["action"][0]).all()
    _check_against_golden(case, fin, 0)
    _check_against_golden(case, fin, T - 1)
    tot = eng.totals()
    assert tot["sims"] == T * 800 * case["n_plies"] and tot["games"] == T and tot["moves"] == T * case["n_plies"]


def _philox_uniform(seed, game, ply):
    """Philox4x32-10 exactly as csrc/az_tree.cuh:philox_uniform (counter = game id lo/hi, ply, 0)."""
    M0, M1, W0, W1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c = [game & mask, (game >> 32) & mask, ply & mask, 0]
    k0, k1 = seed & mask, (seed >> 32) & mask
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> 32) ^ c[3] ^ k1) & mask, p0 & mask]
        k0, k1 = (k0 + W0) & mask, (k1 + W1) & mask
    return ((c[0] >> 5) * 67108864.0 + (c[1] >> 6)) / 9007199254740992.0


@pytest.mark.parametrize("inline_play", [False, True])
def test_async_refill_philox_games_match_c_oracle(inline_play):
    """The production configuration of the tree engine: more games than trees (refill), device Philox
    move sampling, moves played inside az_step (inline_play) or by az_play, external evaluator through
    az_step (hash evaluator computed on the host in float64).  Every finished game, whichever tree slot
    played it, must equal the C oracle's game for the same uniforms."""
    from oracle import c_oracle, evaluators

    engine, _ = _engine_mod()
    rules = engine.Rules(7, 6, 4, True)
    T, G, sims, seed, base = 6, 14, 48, 99, 1000
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=sims, eval_mode="external", move_mode="philox", seed=seed,
                            games_target=G, game_id_base=base, auto_restart=True, fin_capacity=G, max_free_sims=3,
                            inline_play=inline_play)
    A = rules.n_actions
    states = torch.zeros((T, 6, 7, 4), dtype=torch.float32, device="cuda")
    valid = torch.zeros(T, dtype=torch.int32, device="cuda")
    priors = torch.zeros((T, A), dtype=torch.float64, device="cuda")
    values = torch.zeros(T, dtype=torch.float64, device="cuda")
    f = evaluators.hash_evaluator(A)
    for it in range(400000):
        eng.step(priors, values, states, valid)
        v = valid.cpu().numpy()
        st = states.cpu().numpy()
        p = np.zeros((T, A))
        val = np.zeros(T)
        for t in range(T):
            if v[t]:
                p[t], val[t] = f(st[t])
        priors.copy_(torch.from_numpy(p))
        values.copy_(torch.from_numpy(val))
        if not inline_play or it % 16 == 0:
            eng.play()
        if it % 64 == 0 and int((eng.phases() != 0).sum()) == 0:
            break
    eng.check_status()
    fin = eng.drain_finished()
    assert sorted(fin["game_id"].tolist()) == list(range(base, base + G))
    crules = c_oracle.make_rules(7, 6, 4, True)
    for i, g in enumerate(fin["game_id"]):
        u = [_philox_uniform(seed, int(g), ply) for ply in range(rules.max_plies)]
        want = c_oracle.play_game(crules, sims, "hash", uniforms=u)
        n = len(want["moves"])
        assert fin["len"][i] == n and fin["result"][i] == want["result"], int(g)
        np.testing.assert_array_equal(fin["action"][i][:n] & 0xFFFF, want["moves"])
        np.testing.assert_array_equal(fin["visits"][i][:n], want["visits"])
    tot = eng.totals()
    assert tot["games"] == G and tot["moves"] == int(fin["len"].sum()) and tot["sims"] == tot["moves"] * sims


@pytest.mark.parametrize("W,H,n,gravity", [(11, 10, 5, False), (11, 10, 4, True), (2, 2, 2, False), (8, 7, 4, True)])
def test_extreme_board_sizes_match_c_oracle(W, H, n, gravity):
    """Largest board the 128-bit layout admits (H*(W+1) = 120 bits, 110 actions), the smallest legal one, and the
    largest one-word board (7 rows x 9 = 63 bits): full seeded games against the C oracle."""
    from oracle import c_oracle

    engine, _ = _engine_mod()
    rules = engine.Rules(W, H, n, gravity)
    T, sims = 3, 24
    rng = np.random.RandomState(W * 100 + H)
    uniforms = rng.random_sample((T, rules.max_plies))
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=sims, eval_mode="hash", prior_mode="f64",
                            move_mode="host_uniforms")
    eng.set_uniforms(uniforms)
    fin = _play_games(eng, rules.max_plies)
    assert len(fin["len"]) == T
    crules = c_oracle.make_rules(W, H, n, gravity)
    for g in range(T):
        want = c_oracle.play_game(crules, sims, "hash", uniforms=uniforms[g])
        k = len(want["moves"])
        assert fin["len"][g] == k and fin["result"][g] == want["result"]
        np.testing.assert_array_equal(fin["action"][g][:k] & 0xFFFF, want["moves"])
        np.testing.assert_array_equal(fin["visits"][g][:k], want["visits"])


# ------------------------------------------------------------------ Dirichlet root noise
def test_dirichlet_sampler_has_the_right_distribution():
    """The device sampler behind the root noise against the closed-form moments of Dirichlet(alpha * 1_k) and
    against np.random.dirichlet itself (statistical parity: same distribution, different stream)."""
    import ctypes

    from az_b200 import engine, native

    for alpha, k, n in ((0.03, 7, 40000), (1.0, 3, 20000), (0.3, 81, 4000)):
        out = torch.zeros((n, k), dtype=torch.float64, device="cuda")
        native.check(native.lib().az_debug_dirichlet(ctypes.c_uint64(123), alpha, k, n, engine._ptr(out), engine._stream()))
        x = out.cpu().numpy()
        assert np.isfinite(x).all() and (x >= 0).all() and np.allclose(x.sum(1), 1.0)
        mean, var = 1.0 / k, (1.0 / k) * (1 - 1.0 / k) / (k * alpha + 1)
        se_mean = np.sqrt(var / n)
        assert np.abs(x.mean(0) - mean).max() < 5 * se_mean
        assert np.abs(x.var(0) - var).max() < 0.08 * var + 5e-5
        ref = np.random.RandomState(0).dirichlet(np.full(k, alpha), n)
        # same shape of the distribution: quantiles of the first component agree with numpy's sampler
        q = [0.5, 0.9, 0.99]
        assert np.allclose(np.quantile(x[:, 0], q), np.quantile(ref[:, 0], q), atol=0.03)
        assert abs((x.max(1) > 0.99).mean() - (ref.max(1) > 0.99).mean()) < 0.02  # mass on the corners (alpha << 1)


def test_root_noise_plumbing():
    """ratio = 0 goes through the noise kernels but must reproduce the golden game exactly; ratio = 0.25 changes
    the search, keeps its invariants, depends on the seed and is reproducible for a seed."""
    engine, _ = _engine_mod()
    case = load_golden("game_6x7_250_hash")
    rules = rules_of(case)
    eng = engine.TreeEngine(rules, n_trees=2, sims_per_move=250, eval_mode="hash", prior_mode="f64", dirichlet_noise=True,
                            dirichlet_ratio=0.0)
    fin = _play_games(eng, rules.max_plies)
    _check_against_golden(case, fin, 0)
    runs = []
    for seed in (1, 1, 2):
        eng = engine.TreeEngine(rules, n_trees=2, sims_per_move=250, eval_mode="hash", prior_mode="f64", seed=seed,
                                dirichlet_noise=True, dirichlet_alpha=0.03, dirichlet_ratio=0.25)
        eng.search()
        torch.cuda.synchronize()
        eng.check_status()
        n0, w0, p0 = eng.root_stats(0)
        n1, _, _ = eng.root_stats(1)
        assert sum(n0) == 249 and sum(n1) == 249 and p0 == case["plies"][0]["P"]  # stored priors are untouched
        runs.append((n0, n1))
    assert runs[0] == runs[1] and runs[0] != runs[2]
    assert runs[0][0] != runs[0][1]  # different game ids draw different noise
    assert runs[0][0] != case["plies"][0]["N"]
