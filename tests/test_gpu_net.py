"""GPU tests of the policy/value net path: hand-written stem / heads kernels and the bf16 tower
against the fp32 PyTorch module (floating point: tolerances stated per test)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _mods():
    from az_b200 import engine, native, net

    return engine, native, net


def _random_states(n, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    code = torch.randint(0, 3, (n, H, W), generator=g)
    x = torch.zeros(n, H, W, 4)
    x.scatter_(3, code[..., None], 1.0)
    x[..., 3] = 1.0
    return x


@pytest.mark.parametrize("H,W", [(6, 7), (9, 9), (3, 3)])
def test_stem_kernel_matches_fp32_conv(H, W):
    engine, native, net = _mods()
    torch.manual_seed(1)
    n = 257
    x = _random_states(n, H, W)
    w = torch.randn(128, 4, 3, 3) * 0.3
    b = torch.randn(128) * 0.1
    # the kernel rounds the weights to bf16 (tensor-core operands) and accumulates in fp32
    wq = w.to(torch.bfloat16).float()
    want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), wq, b, padding=1)).permute(0, 2, 3, 1)
    xd = x.to("cuda", torch.bfloat16).contiguous()
    out = torch.empty((n, H, W, 128), dtype=torch.bfloat16, device="cuda")
    wd, bd = w.cuda().contiguous(), b.cuda()
    native.check(native.lib().az_net_stem(engine._ptr(xd), engine._ptr(wd), engine._ptr(bd), n, H, W, 128,
                                          engine._ptr(out), engine._stream()))
    got = out.float().cpu()
    # fp32 accumulation of exact 0/1 inputs; only the bf16 rounding of the output differs: <= 2^-8 relative
    assert torch.allclose(got, want, rtol=2 ** -8, atol=1e-5)


@pytest.mark.parametrize("H,W,A", [(6, 7, 7), (9, 9, 9), (9, 9, 81), (3, 3, 9)])
def test_heads_kernel_matches_fp32(H, W, A):
    engine, native, net = _mods()
    torch.manual_seed(2)
    n, cells = 301, H * W
    x = (torch.randn(n, cells, 128).relu() * 0.7).to(torch.bfloat16)
    cw, cb = torch.randn(3, 128) * 0.1, torch.randn(3) * 0.1
    pw, pb = torch.randn(A, 2 * cells) * 0.2, torch.randn(A) * 0.1
    v1w, v1b = torch.randn(256, cells) * 0.2, torch.randn(256) * 0.1
    v2w, v2b = torch.randn(256) * 0.1, torch.randn(1) * 0.1
    xf = x.float()
    h = torch.relu(xf @ cw.to(torch.bfloat16).float().T + cb)  # [n, cells, 3]; conv weights are bf16 operands
    want_p = torch.softmax(h[..., :2].reshape(n, -1) @ pw.T + pb, -1)
    want_v = torch.tanh(torch.relu(h[..., 2] @ v1w.T + v1b) @ v2w + v2b)
    pw_pad = torch.nn.functional.pad(pw, (0, 1))  # rows padded to the odd stride az_net_heads expects
    v1w_pad = v1w.t().contiguous()  # Dense(256) weights transposed [cells][256]
    dev = [t.cuda().contiguous() for t in (cw, cb, pw_pad, pb, v1w_pad, v1b, v2w, v2b)]
    hw = native.AzHeadWeights(*[t.data_ptr() for t in dev])
    priors = torch.empty((n, A), device="cuda")
    values = torch.empty(n, device="cuda")
    xd = x.cuda().contiguous()
    native.check(native.lib().az_net_heads(engine._ptr(xd), ctypes.byref(hw), n, cells, 128, A, engine._ptr(priors),
                                           engine._ptr(values), engine._stream()))
    # same bf16 inputs, fp32 arithmetic on both sides: only summation order differs
    assert (priors.cpu() - want_p).abs().max() < 2e-5
    assert (values.cpu() - want_v).abs().max() < 2e-5
    assert torch.allclose(priors.sum(-1).cpu(), torch.ones(n), atol=1e-5)


@pytest.mark.parametrize("H,W,A", [(6, 7, 7), (9, 9, 9)])
def test_bf16_inference_net_close_to_fp32_module(H, W, A):
    engine, native, net = _mods()
    torch.manual_seed(0)
    ref = net.randomise_bn(net.PolicyValueNet(H, W, A)).eval()
    x = _random_states(512, H, W, seed=3)
    with torch.no_grad():
        want_p, want_v = ref(x)
    inf = net.InferenceNet(ref, dtype=torch.bfloat16, device="cuda")
    assert inf.fast
    p, v = inf(x.cuda().to(torch.bfloat16))
    dp = (p.cpu() - want_p).abs().max().item()
    dv = (v.cpu() - want_v.reshape(-1)).abs().max().item()
    # bf16 weights and activations through 9 sequential 3x3 convolutions vs fp32: tolerance 8e-3 on the softmax outputs
    # and on the tanh values (measured 2.4e-3 / 1.6e-3 on 4096 reachable positions: tests/test_gpu_a22.py asserts 4e-3 there)
    print("bf16 vs fp32", H, W, dp, dv)
    assert dp < 8e-3 and dv < 8e-3, (dp, dv)
    # the slow (library-only) GPU path computes the same function
    inf.fast = False
    p2, v2 = inf(x.cuda().to(torch.bfloat16))
    assert (p2 - p).abs().max().item() < 2e-2 and (v2 - v).abs().max().item() < 5e-2


@pytest.mark.parametrize("W,H,n,gravity", [(9, 9, 5, True), (9, 9, 5, False), (5, 4, 3, True)])
def test_selfplay_runner_other_boards(W, H, n, gravity):
    """BASELINE config C4 (9x9 connect-5, with and without gravity) through the whole path: generic-rules
    tree kernels (two-word bitboards, up to 81 children), hand-written stem / heads at that size, bf16 tower."""
    from az_b200 import selfplay

    engine, native, net = _mods()
    rules = engine.Rules(W, H, n, gravity)
    torch.manual_seed(0)
    r = selfplay.SelfPlayRunner(rules, n_trees=48, sims_per_move=20, games_target=64, unroll=4, seed=5)
    r.run_until_done(poll_every=64, max_advances=400000)
    tot = r.totals()
    states, policies, values = r.collect()
    assert tot["games"] == 64 and len(values) == tot["moves"] and tot["sims"] == tot["moves"] * 20
    assert states.shape[1:] == (H, W, 4) and policies.shape[1] == rules.n_actions
    assert np.allclose(policies.sum(-1), 1.0) and (policies >= 0).all()
    assert (states[..., :3].sum(-1) == 1).all() and (states[..., 3] == 1).all()
    # a policy target never puts mass on an occupied cell / full column
    if gravity:
        legal = states[:, 0, :, 0] == 1                       # top cell of the column empty
        assert (policies[~legal] == 0).all()
    else:
        legal = np.transpose(states[..., 0], (0, 2, 1)).reshape(len(values), -1) == 1   # action = x * H + y
        assert (policies[~legal] == 0).all()


def test_fused_advance_equals_the_three_kernel_route():
    """az_advance_fused (heads + tree step + stem in one launch) against az_net_heads / az_step / az_net_stem
    in sequence: same weights, same Philox seeds -> the finished games must be identical, bit for bit."""
    from az_b200 import selfplay

    engine, native, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    out = []
    for fused, extra, mf in ((False, 0, 8), (True, 0, 8), (True, 12, 1)):
        torch.manual_seed(0)
        fp32 = net.randomise_bn(net.PolicyValueNet())
        r = selfplay.SelfPlayRunner(rules, n_trees=96, sims_per_move=40, net=fp32, games_target=160, unroll=4, seed=11,
                                    fused=fused, extra_sims=extra, max_free_sims=mf, whole_net=False)
        assert r.fused == fused and not r.whole_net
        r.run_until_done(poll_every=64, max_advances=400000)
        fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
        order = np.argsort(fin["game_id"])
        out.append(({k: v[order] for k, v in fin.items()}, r.totals()))
    (a, ta) = out[0]
    for b, tb in out[1:]:  # the fused launch, and the fused launch with evaluator-free simulations beside the tower
        assert ta["games"] == tb["games"] == 160 and ta["sims"] == tb["sims"] and ta["evals"] == tb["evals"]
        for k in ("game_id", "len", "result"):
            np.testing.assert_array_equal(a[k], b[k])
        for g in range(160):
            n = a["len"][g]
            np.testing.assert_array_equal(a["visits"][g][:n], b["visits"][g][:n])
            np.testing.assert_array_equal(a["action"][g][:n], b["action"][g][:n])


def test_games_do_not_depend_on_how_they_are_sharded():
    """SURVEY 4 test (5): 1 rank playing games [0, G) and 2 ranks playing [0, G/2) and [G/2, G) (az_b200.dist.shard_games)
    produce the same games id by id - move sampling is keyed by (seed, game id, ply), never by rank or tree slot."""
    from az_b200 import dist as azdist
    from az_b200 import selfplay

    engine, native, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    G = 120

    def play(base, count, trees):
        torch.manual_seed(0)
        r = selfplay.SelfPlayRunner(rules, n_trees=trees, sims_per_move=32, net=net.PolicyValueNet(), games_target=count,
                                    game_id_base=base, unroll=4, seed=21)
        r.run_until_done(poll_every=64, max_advances=400000)
        fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
        return {int(g): (fin["len"][i], fin["result"][i], fin["action"][i][: fin["len"][i]].tolist(),
                         fin["visits"][i][: fin["len"][i]].tolist()) for i, g in enumerate(fin["game_id"])}

    whole = play(0, G, 64)
    parts = {}
    for rank in range(2):
        base, count = azdist.shard_games(G, rank, 2)
        parts.update(play(base, count, 48))
    assert sorted(whole) == sorted(parts) == list(range(G))
    for g in range(G):
        assert whole[g][0] == parts[g][0] and whole[g][1] == parts[g][1], g
        assert whole[g][2] == parts[g][2] and whole[g][3] == parts[g][3], g


def test_evaluation_memo_changes_nothing_but_the_number_of_evaluations():
    """The reference memoises evaluations by position (plays_inferences, mcts.py:122-143).  With the device memo on,
    every finished game must be identical to the run without it - the evaluator is a pure function of the position -
    while a good share of the leaves no longer needs the net."""
    from az_b200 import selfplay

    engine, native, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    out = []
    # the whole-net route (az_step_gather + az_net_forward_gathered, the default) with and without the memo, then the two
    # older routes against their own memo-free run
    for log2, fused, whole in ((0, True, True), (18, True, True), (10, True, True)):
        torch.manual_seed(0)
        fp32 = net.randomise_bn(net.PolicyValueNet())
        r = selfplay.SelfPlayRunner(rules, n_trees=128, sims_per_move=64, net=fp32, games_target=256, unroll=4, seed=5,
                                    fused=fused, eval_cache_log2=log2, whole_net=whole)
        assert r.whole_net == whole
        r.run_until_done(poll_every=64, max_advances=400000)
        fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
        order = np.argsort(fin["game_id"])
        out.append(({k: v[order] for k, v in fin.items()}, r.totals()))
    (a, ta) = out[0]
    assert ta["memo_hits"] == 0
    for b, tb in out[1:]:
        assert tb["games"] == 256 and tb["sims"] == ta["sims"] and tb["moves"] == ta["moves"]
        assert tb["memo_hits"] > 0 and tb["evals"] + tb["memo_hits"] == ta["evals"]
        for k in ("game_id", "len", "result"):
            np.testing.assert_array_equal(a[k], b[k])
        for g in range(256):
            n = a["len"][g]
            np.testing.assert_array_equal(a["visits"][g][:n], b["visits"][g][:n])
            np.testing.assert_array_equal(a["action"][g][:n], b["action"][g][:n])
    # a big table hits more often than a tiny, collision-ridden one
    assert out[1][1]["memo_hits"] > out[2][1]["memo_hits"] > 0
    assert out[1][1]["memo_hits"] > 0.15 * ta["evals"]
    # az_advance_fused and the three-kernel route: same property
    for fused in (True, False):
        pair = []
        for log2 in (0, 18):
            torch.manual_seed(0)
            fp32 = net.randomise_bn(net.PolicyValueNet())
            r = selfplay.SelfPlayRunner(rules, n_trees=128, sims_per_move=64, net=fp32, games_target=256, unroll=4, seed=5,
                                        fused=fused, eval_cache_log2=log2, whole_net=False)
            r.run_until_done(poll_every=64, max_advances=400000)
            fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
            order = np.argsort(fin["game_id"])
            pair.append(({k: v[order] for k, v in fin.items()}, r.totals()))
        (a2, t2), (b2, u2) = pair
        assert u2["memo_hits"] > 0 and u2["evals"] + u2["memo_hits"] == t2["evals"] and u2["sims"] == t2["sims"]
        for k in ("game_id", "len", "result", "action", "visits"):
            np.testing.assert_array_equal(a2[k], b2[k])


def test_tcgen05_shortcut_gemm_matches_the_library():
    """az_net_conv1x1 (hand-written tcgen05 GEMM) against the same product in float32 and against cuDNN's bf16 result."""
    import ctypes

    from az_b200 import native

    torch.manual_seed(3)
    lib = native.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    w = (torch.randn(128, 128, device="cuda") * 0.1).to(torch.bfloat16).contiguous()
    for rows in (128, 1, 127, 129, 4096 * 42, 300 * 128 + 77):
        x = torch.randn(rows, 128, device="cuda").to(torch.bfloat16).contiguous()
        y = torch.full((rows + 1, 128), 7.0, device="cuda", dtype=torch.bfloat16)  # one guard row
        native.check(lib.az_net_conv1x1(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()), rows, 128,
                                        ctypes.c_void_p(y.data_ptr()), stream))
        torch.cuda.synchronize()
        ref = x.float() @ w.float().t()
        err = (y[:rows].float() - ref).abs().max().item()
        assert err <= 2.0 ** -7 * max(1.0, ref.abs().max().item()), (rows, err)  # bf16 output rounding only
        assert bool((y[rows] == 7.0).all()), rows  # nothing written past the last row
    # every row and column really lands where it should: a one-hot probe
    x = torch.zeros(256, 128, device="cuda", dtype=torch.bfloat16)
    x[torch.arange(256), torch.arange(256) % 128] = 1.0
    y = torch.empty_like(x)
    native.check(lib.az_net_conv1x1(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()), 256, 128,
                                    ctypes.c_void_p(y.data_ptr()), stream))
    assert torch.equal(y, w.t()[torch.arange(256) % 128])


def test_tower_with_tcgen05_shortcut_equals_cudnn_tower():
    from az_b200.net import InferenceNet, PolicyValueNet, randomise_bn

    torch.manual_seed(5)
    net = randomise_bn(PolicyValueNet()).eval()
    inf = InferenceNet(net, dtype=torch.bfloat16, device="cuda")
    h0 = torch.randn(777, 6, 7, 128, device="cuda").to(torch.bfloat16)
    inf.tc_shortcut = True
    a = inf.tower(h0).float()
    inf.tc_shortcut = False
    b = inf.tower(h0).float()
    scale = b.abs().max().item()
    assert (a - b).abs().max().item() <= 0.02 * scale  # two bf16 roundings of the shortcut per block, different order


def test_tcgen05_stem_equals_the_float32_convolution():
    """az_net_stem_tc against conv3x3 in float32 on the bf16-rounded weights, for 6x7, 9x9 and 3x3 boards and batch sizes
    that leave the last tile partly empty."""
    import ctypes

    from az_b200 import native
    from az_b200.net import InferenceNet, PolicyValueNet, randomise_bn

    lib = native.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    for (H, W, A) in ((6, 7, 7), (9, 9, 81), (3, 3, 9)):
        torch.manual_seed(H)
        inf = InferenceNet(randomise_bn(PolicyValueNet(H, W, A)).eval(), dtype=torch.bfloat16, device="cuda")
        assert inf.tc_stem
        for n in (4096, 1, 2, 3, 4, 1001):
            code = torch.randint(0, 3, (n, H, W), device="cuda")
            x = torch.cat([torch.nn.functional.one_hot(code, 3).float(), torch.ones(n, H, W, 1, device="cuda")], -1).to(torch.bfloat16)
            out = torch.full((n + 1, H, W, 128), 7.0, dtype=torch.bfloat16, device="cuda")
            native.check(lib.az_net_stem_tc(P(x), P(inf.stem_w16_k), P(inf.stem_b32), n, H, W, 128, P(out),
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            w = inf.stem_w.detach().float()  # the same bf16-rounded weights
            ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, inf.stem_b32, padding=1)).permute(0, 2, 3, 1)
            err = (out[:n].float() - ref).abs().max().item()
            assert err <= 2.0 ** -7 * max(1.0, ref.abs().max().item()), (H, W, n, err)
            assert bool((out[n] == 7.0).all())
        # and the mma.sync kernel it replaces gives the same tensor up to the output rounding
        x = torch.zeros(64, H, W, 4, device="cuda", dtype=torch.bfloat16)
        x[..., 0] = 1
        x[..., 3] = 1
        inf.tc_stem = True
        a = inf(x)[0].clone()
        inf.tc_stem = False
        b = inf(x)[0]
        assert (a - b).abs().max().item() <= 2e-3


def test_split_heads_equal_the_single_kernel():
    """az_net_head_convs + az_net_heads_dense against az_net_heads on the same tower output."""
    from az_b200.net import InferenceNet, PolicyValueNet, randomise_bn

    for (H, W, A) in ((6, 7, 7), (9, 9, 81)):
        torch.manual_seed(H + 1)
        inf = InferenceNet(randomise_bn(PolicyValueNet(H, W, A)).eval(), dtype=torch.bfloat16, device="cuda")
        code = torch.randint(0, 3, (777, H, W), device="cuda")
        x = torch.cat([torch.nn.functional.one_hot(code, 3).float(), torch.ones(777, H, W, 1, device="cuda")], -1).to(torch.bfloat16)
        inf.split_heads = False
        p0, v0 = inf(x)
        p0, v0 = p0.clone(), v0.clone()
        inf.split_heads = True
        p1, v1 = inf(x)
        # az_net_heads rounds the 1x1 convolution weights to bf16 for mma.sync; az_net_head_convs keeps them in float32
        dp, dv = (p1 - p0).abs().max().item(), (v1 - v0).abs().max().item()
        assert dp <= 4e-3 and dv <= 4e-3, (H, W, dp, dv)

