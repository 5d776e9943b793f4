"""SURVEY 8a row a22 (the net on the search path), tightened as VERDICT r1 item 6 asks:
 (i)   prior-level bound of the bf16 whole-net kernel against the float32 module on 4096 reachable positions;
 (ii)  800-simulation root policy targets, bf16 kernel vs float32 evaluator, with a sharpened net, counted in flipped visits;
 (iii) plumbing exactness independent of bf16: the GPU tree engine fed by the float32 module reproduces, visit for visit,
       the C oracle's MCTS fed by the same module.
The architecture itself is pinned on the CPU by tests/test_net_keras_restatement.py (item iv).  All through the C ABI."""
import contextlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    from az_b200 import engine, env, net

    return engine, env, net


@contextlib.contextmanager
def _exact_fp32():
    """float32 means float32: no TF32 in cuDNN / cuBLAS while the checker runs."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _reachable_positions(env, rules, n, seed, max_plies=14):
    """n non-terminal positions after 0..max_plies uniformly random legal moves (K2 / K3 kernels, side to move = +1)."""
    rng = np.random.default_rng(seed)
    cells = np.zeros((n, rules.height, rules.width), dtype=np.int8)
    plies = np.zeros(n, dtype=np.int32)
    target = rng.integers(0, max_plies + 1, n)
    for _ in range(max_plies):
        legal = env.env_legal(rules, cells)
        u = rng.random(legal.shape) * legal
        a = u.argmax(-1).astype(np.int32)
        nxt, status = env.env_play(rules, cells, a)
        go = (plies < target) & (status == 0)
        cells[go] = nxt[go]
        plies[go] += 1
    return cells, plies


def _sharpen(m, scale):
    """Random-initialised heads give near-uniform priors and values near 0, i.e. PUCT scores tied to within the bf16 error
    everywhere; a trained net is decisive.  Scaling the last dense layers makes the random net decisive too."""
    with torch.no_grad():
        m.policy_fc.weight.mul_(scale)
        m.value_fc2.weight.mul_(scale)
    return m


def _record(name, payload):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "a22_measured.jsonl"), "a") as fp:
        fp.write(json.dumps({"test": name, **payload}) + "\n")


# ------------------------------------------------------------------ (i)
def test_prior_level_bound_on_4096_reachable_positions():
    """max |dp| <= 4e-3 and max |dv| <= 4e-3 over 4096 positions reached by random play, bf16 az_net_forward against the
    float32 PolicyValueNet (randomised BN statistics, random-initialised weights).  Measured on B200: see
    profiles/a22_measured_r2.jsonl."""
    engine, env, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    torch.manual_seed(0)
    fp32 = net.randomise_bn(net.PolicyValueNet(6, 7, 7)).eval()
    inf = net.InferenceNet(fp32)
    assert inf.fused_net
    cells, _ = _reachable_positions(env, rules, 4096, seed=11)
    x = torch.from_numpy(env.env_encode(rules, cells))
    p, v = inf(x.cuda().to(torch.bfloat16))
    with _exact_fp32(), torch.no_grad():
        p32, v32 = fp32.cuda()(x.cuda())
    dp = float((p - p32).abs().max())
    dv = float((v - v32.reshape(-1)).abs().max())
    _record("prior_level", {"n": 4096, "max_dp": dp, "max_dv": dv, "mean_dp": float((p - p32).abs().mean())})
    assert dp <= 4e-3 and dv <= 4e-3, (dp, dv)


# ------------------------------------------------------------------ (ii)
def _root_visits(engine, rules, cells, plies, sims, evaluate, state_dtype):
    """Runs one `sims`-simulation search from every position (one tree each) with an external evaluator and returns the
    root visit counts [T, A] (0 for illegal actions)."""
    T, A = len(plies), rules.n_actions
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=sims, eval_mode="external", prior_mode="f32", max_free_sims=4)
    eng.set_roots(np.arange(T), cells, plies)
    eng.begin_search(sims)
    states = torch.zeros((T, rules.height, rules.width, 4), dtype=state_dtype, device="cuda")
    valid = torch.zeros(T, dtype=torch.int32, device="cuda")
    priors = values = None
    for it in range(4 * sims):
        eng.step(priors, values, states, valid)
        if it % 16 == 15 and int((eng.phases() == 1).sum()) == 0:
            break
        priors, values = evaluate(states)
    eng.check_status()
    assert int((eng.phases() == 2).sum()) == T
    out = np.zeros((T, A), dtype=np.int64)
    legal = np.zeros((T, A), dtype=bool)
    for t in range(T):
        n, _, _ = eng.root_stats(t)
        cols = [x for x in range(rules.width) if cells[t, 0, x] == 0]  # gravity: legal columns in board order
        assert len(cols) == len(n)
        out[t, cols] = n
        legal[t, cols] = True
    return out, legal


def _pi_stats(a, b):
    moved = np.abs(a - b).sum(-1) // 2
    dpi = np.abs(a - b).max(-1) / 799.0
    return {"share_identical": float((moved == 0).mean()), "share_within_one_visit": float((moved <= 1).mean()),
            "share_within_5e-3": float((dpi <= 5e-3).mean()), "median_moved_visits": float(np.median(moved)),
            "mean_dpi": float(dpi.mean()), "max_dpi": float(dpi.max())}


@pytest.mark.parametrize("scale", [1.0, 6.0])
def test_root_policy_targets_800_simulations_in_moved_visits(scale):
    """north_star's example bound is max |d pi| <= 1e-3 under bf16 vs fp32.  One visit out of 799 is 1.25e-3, so that bound
    means 'no visit moves'.  Measured on B200 (profiles/a22_measured_r2.jsonl), 128 reachable roots x 800 simulations:
    random-initialised net 15 % of the roots identical, 56 % within one moved visit, 95 % within 5e-3, mean 3.6e-3; with
    sharpened heads (x6) 9 % / 23 % / 47 %, mean 7.4e-3, median 4 moved visits.  Two control arms put that in context: the
    float32 evaluator against itself with its outputs perturbed by 1e-6 relative moves NO visit (the search is not chaotic,
    so these differences are the bf16 evaluator's), and perturbed by 2e-3 relative - about the bf16 path's prior error on
    the random-initialised net, test (i) - it moves visits in 45 % of the roots, mean 2.4e-3 (bf16: 3.6e-3), and tips the
    same bistable root by 0.23 that bf16 tips (the max of both arms).  With sharpened heads the same control gives 1.3e-3:
    there the bf16 arm is 6x the control because the x6 dense layer also multiplies the bf16 error of its input.  A search
    is 800 chained argmax decisions over PUCT scores that differ by less than the prior error wherever two moves are
    close, so an evaluator error of 2.4e-3 cannot give root targets within 1.25e-3; asserted is what was measured, with a
    head-room of 2x: mean max |d pi| <= 8e-3 (1.5e-2 sharpened), >= 85 % of the roots within 5e-3 and the bf16 arm within
    3x of the 2e-3 control arm (random-initialised), median moved visits <= 2 (8 sharpened)."""
    engine, env, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    torch.manual_seed(1)
    fp32 = _sharpen(net.randomise_bn(net.PolicyValueNet(6, 7, 7)), scale).eval()
    inf = net.InferenceNet(fp32)
    gpu32 = fp32.cuda()
    cells, plies = _reachable_positions(env, rules, 128, seed=21, max_plies=10)

    def eval_bf16(states):
        p, v = inf(states)
        return p.contiguous(), v.contiguous()

    def eval_fp32(states, eps=0.0):
        with _exact_fp32(), torch.no_grad():
            p, v = gpu32(states)
        if eps:  # deterministic relative perturbation of the evaluator's outputs
            wob = 1.0 + eps * torch.cos(1e4 * p)
            p, v = p * wob, v * (1.0 + eps)
        return p.contiguous(), v.reshape(-1).contiguous()

    n16, _ = _root_visits(engine, rules, cells, plies, 800, eval_bf16, torch.bfloat16)
    n32, _ = _root_visits(engine, rules, cells, plies, 800, eval_fp32, torch.float32)
    n32e, _ = _root_visits(engine, rules, cells, plies, 800, lambda s: eval_fp32(s, 1e-6), torch.float32)
    n32c, _ = _root_visits(engine, rules, cells, plies, 800, lambda s: eval_fp32(s, 2e-3), torch.float32)
    n32b, _ = _root_visits(engine, rules, cells, plies, 800, eval_fp32, torch.float32)
    assert (n16.sum(-1) == 799).all() and (n32.sum(-1) == 799).all()  # quirk Q2: a fresh root counts sims - 1
    assert (n32 == n32b).all()  # the float32 arm is reproducible: the control measures the perturbation, not noise
    bf16, tiny, control = _pi_stats(n16, n32), _pi_stats(n32e, n32), _pi_stats(n32c, n32)
    _record("root_pi_800", {"roots": 128, "head_scale": scale, "bf16_vs_fp32": bf16, "fp32_perturbed_1e-6_vs_fp32": tiny,
                            "fp32_perturbed_2e-3_vs_fp32": control})
    assert tiny["share_identical"] == 1.0
    if scale == 1.0:
        assert bf16["mean_dpi"] <= 8e-3 and bf16["share_within_5e-3"] >= 0.85 and bf16["median_moved_visits"] <= 2, bf16
        assert bf16["mean_dpi"] <= 3.0 * max(control["mean_dpi"], 1e-3), (bf16, control)
    else:
        assert bf16["mean_dpi"] <= 1.5e-2 and bf16["median_moved_visits"] <= 8, bf16


# ------------------------------------------------------------------ (iii)
def test_engine_with_the_float32_module_equals_the_oracle_with_the_float32_module():
    """Same evaluator on both sides (the float32 PolicyValueNet on the CPU, one position per call, so both sides get the
    same bits): whole games, argmax moves, 120 simulations per move - moves and every root visit count identical between
    the GPU tree engine (az_step, float32 priors) and the C oracle's MCTS (float32 mode).  Nothing here depends on bf16:
    any difference would be plumbing (state encoding, prior / move pairing, normalisation dtype, value sign)."""
    from oracle import c_oracle

    engine, env, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    torch.manual_seed(2)
    fp32 = _sharpen(net.randomise_bn(net.PolicyValueNet(6, 7, 7)), 3.0).eval()
    sims = 120

    def one(state):
        with torch.no_grad():
            p, v = fp32(torch.from_numpy(np.ascontiguousarray(state, dtype=np.float32))[None])
        return p[0].numpy(), float(v[0, 0])

    def cb(state):
        p, v = one(state)
        return p.astype(np.float64), v

    want = c_oracle.play_game(c_oracle.make_rules(7, 6, 4, True), sims, "callback", prior_mode=c_oracle.PRIOR_F32, callback=cb)
    T, A = 2, 7
    eng = engine.TreeEngine(rules, n_trees=T, sims_per_move=sims, eval_mode="external", prior_mode="f32", move_mode="argmax",
                            max_free_sims=2)
    states = torch.zeros((T, 6, 7, 4), dtype=torch.float32, device="cuda")
    valid = torch.zeros(T, dtype=torch.int32, device="cuda")
    priors = torch.zeros((T, A), dtype=torch.float32, device="cuda")
    values = torch.zeros(T, dtype=torch.float32, device="cuda")
    first = True
    for _ in range(100000):
        eng.step(None if first else priors, None if first else values, states, valid)
        first = False
        v = valid.cpu().numpy()
        if v.any():
            st = states.cpu().numpy()
            p, val = np.zeros((T, A), dtype=np.float32), np.zeros(T, dtype=np.float32)
            for t in range(T):
                if v[t]:
                    p[t], val[t] = one(st[t])
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
    L = len(want["moves"])
    for g in range(T):
        assert int(fin["len"][g]) == L and int(fin["result"][g]) == want["result"]
        np.testing.assert_array_equal(fin["action"][g][:L] & 0xFFFF, want["moves"])
        np.testing.assert_array_equal(fin["visits"][g][:L], want["visits"])
    assert L >= 7 and want["evals"] > 0


# ------------------------------------------------------------------ the headline route end to end
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


@pytest.mark.parametrize("memo", [0, 12])
def test_headline_route_equals_the_oracle_search_fed_by_the_same_net_kernel(memo):
    """The route bench.py times - az_step_gather + az_net_forward_trees under CUDA graphs: gathered leaf batch, tree warps
    inside the net kernel, moves played in line, Philox sampling, refill, finished-game ring (and, second case, the device
    evaluation memo of the product entry point) - against the C oracle's MCTS (float32 mode) whose evaluator is the SAME net
    kernel called on one position at a time (the kernel is batch independent, tests/test_gpu_tower.py).  Every game, whichever
    tree slot played it: moves, per-ply visit counts and result identical.  The evaluator is a real, position-dependent
    bf16 net, so a wrong prior / move pairing, value sign, leaf list entry or parked leaf would show."""
    from az_b200 import selfplay
    from oracle import c_oracle

    engine, env, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    torch.manual_seed(4)
    fp32 = _sharpen(net.randomise_bn(net.PolicyValueNet(6, 7, 7)), 3.0).eval()
    T, G, sims, seed, base = 24, 36, 64, 77, 5000
    r = selfplay.SelfPlayRunner(rules, n_trees=T, sims_per_move=sims, net=fp32, games_target=G, game_id_base=base, seed=seed,
                                move_mode="philox", auto_restart=True, unroll=4, fin_capacity=G, eval_cache_log2=memo)
    assert r.whole_net and r.net_tree_sims > 0 and r.max_free_sims == 2
    r.run_until_done(poll_every=64, max_advances=400000)
    fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
    assert sorted(fin["game_id"].tolist()) == list(range(base, base + G))
    inf = r.net
    x1 = torch.zeros((1, 6, 7, 4), dtype=torch.bfloat16, device="cuda")

    def cb(state):
        x1.copy_(torch.from_numpy(np.ascontiguousarray(state, dtype=np.float32))[None])
        p, v = inf(x1)
        return p[0].cpu().numpy().astype(np.float64), float(v[0])

    crules = c_oracle.make_rules(7, 6, 4, True)
    lengths = set()
    for i, g in enumerate(fin["game_id"]):
        u = [_philox_uniform(seed, int(g), ply) for ply in range(rules.max_plies)]
        want = c_oracle.play_game(crules, sims, "callback", uniforms=u, prior_mode=c_oracle.PRIOR_F32, callback=cb)
        n = len(want["moves"])
        lengths.add(n)
        assert fin["len"][i] == n and fin["result"][i] == want["result"], int(g)
        np.testing.assert_array_equal(fin["action"][i][:n] & 0xFFFF, want["moves"])
        np.testing.assert_array_equal(fin["visits"][i][:n], want["visits"])
    assert len(lengths) > 3  # the games really differ
    tot = r.totals()
    assert tot["games"] == G and tot["sims"] == sims * tot["moves"]
    if memo:
        assert tot["memo_hits"] > 0
