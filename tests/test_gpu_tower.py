"""GPU tests of az_net_tower (csrc/az_tower.cu): the whole residual tower (model/tensorflow/base_layers.py:85-125 x
depth, model.py:48-66) as one tcgen05 kernel, through the C ABI, against float32 convolutions on the same bf16-rounded
operands and against the cuDNN route it replaces.  Floating point: tolerances stated per test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
F = torch.nn.functional


def _mods():
    from az_b200 import engine, native, net

    return engine, native, net


def _blocks(depth, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(depth):
        w1 = (torch.randn(128, 128, 3, 3, generator=g) * 0.03 * scale).to(torch.bfloat16).float()
        w2 = (torch.randn(128, 128, 3, 3, generator=g) * 0.03 * scale).to(torch.bfloat16).float()
        wp = (torch.randn(128, 128, 1, 1, generator=g) * 0.08 * scale).to(torch.bfloat16).float()
        out.append((w1, torch.randn(128, generator=g) * 0.1, w2, wp, torch.randn(128, generator=g) * 0.1))
    return out


def _reference(x, blocks):
    """float32 convolutions, activations rounded to bf16 where the kernel rounds them (h and every block output)."""
    t = x.float().permute(0, 3, 1, 2)
    for w1, b1, w2, wp, b2p in blocks:
        h = F.relu(F.conv2d(t, w1, b1, padding=1)).to(torch.bfloat16).float()
        t = F.relu(F.conv2d(h, w2, b2p, padding=1) + F.conv2d(t, wp)).to(torch.bfloat16).float()
    return t.permute(0, 2, 3, 1)


LAYOUT = {"layout": 1}  # 0 = one CTA per tile, 1 = CTA pairs (cta_group::2); switched by the `layout` fixture


@pytest.fixture(params=[0, 1], ids=["single-cta", "cta-pair"])
def layout(request, monkeypatch):
    """Every test below runs for both kernel variants: az_net_tower / az_net_forward layout 0 and 1."""
    LAYOUT["layout"] = request.param
    monkeypatch.setenv("AZ_TOWER_PAIR", str(request.param))
    yield request.param
    LAYOUT["layout"] = 1


def _run(x, blocks, H, W):
    engine, native, net = _mods()
    img, bias = net.pack_tower_weights(blocks, pair=bool(LAYOUT["layout"]))
    img, bias = img.cuda(), bias.cuda()
    xd = x.to("cuda", torch.bfloat16).contiguous()
    out = torch.full_like(xd, float("nan"))
    native.check(native.lib().az_net_tower(engine._ptr(xd), engine._ptr(img), engine._ptr(bias), xd.shape[0], H, W, 128,
                                           len(blocks), LAYOUT["layout"], engine._ptr(out), engine._stream()))
    torch.cuda.synchronize()
    return out.float().cpu()


@pytest.mark.parametrize("H,W,n,depth", [(6, 7, 3, 1), (6, 7, 1, 1), (6, 7, 500, 4), (8, 8, 131, 4), (7, 7, 9, 2), (9, 9, 5, 1)])
def test_fused_tower_equals_the_float32_convolutions(H, W, n, depth, layout):
    blocks = _blocks(depth, seed=H * 10 + depth)
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, H, W, 128, generator=g).to(torch.bfloat16)
    got = _run(x, blocks, H, W)
    want = _reference(x, blocks)
    assert torch.isfinite(got).all()
    # same operands, float32 accumulation in another order; an activation that lands on a bf16 rounding boundary may
    # differ by one ulp (2^-8 relative) per layer and is then carried through the later layers
    err = (got - want).abs()
    assert float(err.max()) <= 2 ** -5 * max(1.0, float(want.abs().max())), float(err.max())
    assert float((err > 2 ** -7 * want.abs().clamp(min=1.0)).float().mean()) < 0.01


def test_every_tap_lands_on_the_right_cell_and_edges_are_zero_padded(layout):
    """One-hot probes: input 1 at one cell / channel, conv1 = a single tap copying channel 0 -> 0, conv2 = centre-tap
    identity, no shortcut: the output must be the input shifted by the tap, and nothing may wrap around a board edge or
    leak into a neighbouring position of the same tile."""
    H, W, n = 6, 7, 4
    eye = torch.zeros(128, 128, 3, 3)
    eye[:, :, 1, 1] = torch.eye(128)
    zero_p, zero_b = torch.zeros(128, 128, 1, 1), torch.zeros(128)
    x = torch.zeros(n, H, W, 128)
    cells = [(0, 0), (0, 6), (5, 0), (5, 6), (2, 3), (0, 3), (3, 0), (3, 6)]
    for i, (y, xx) in enumerate(cells):
        x[i % n, y, xx, i] = 1.0 + i  # a different channel per probe so they can share a launch
    for ky in range(3):
        for kx in range(3):
            w1 = torch.zeros(128, 128, 3, 3)
            w1[:, :, ky, kx] = torch.eye(128)
            got = _run(x, [(w1, zero_b, eye, zero_p, zero_b)], H, W)
            want = torch.zeros_like(x)
            for i, (y, xx) in enumerate(cells):
                oy, ox = y - (ky - 1), xx - (kx - 1)  # out[oy, ox] = in[oy + ky - 1, ox + kx - 1]
                if 0 <= oy < H and 0 <= ox < W:
                    want[i % n, oy, ox, i] = 1.0 + i
            assert torch.equal(got, want), (ky, kx, (got - want).abs().nonzero()[:8].tolist())


def test_shortcut_and_biases_reach_the_output(layout):
    H, W, n = 6, 7, 5
    g = torch.Generator().manual_seed(3)
    x = torch.rand(n, H, W, 128, generator=g).to(torch.bfloat16)
    wp = (torch.randn(128, 128, 1, 1, generator=g) * 0.1).to(torch.bfloat16).float()
    z3 = torch.zeros(128, 128, 3, 3)
    b1, b2 = torch.randn(128, generator=g), torch.randn(128, generator=g)
    got = _run(x, [(z3, b1, z3, wp, b2)], H, W)
    want = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wp) + b2[None, :, None, None]).permute(0, 2, 3, 1)
    assert torch.allclose(got, want.to(torch.bfloat16).float(), rtol=2 ** -7, atol=1e-6)


def test_inference_net_routes_through_the_fused_tower_and_matches_the_library_route(layout):
    engine, native, net = _mods()
    torch.manual_seed(5)
    fp32 = net.randomise_bn(net.PolicyValueNet(6, 7, 7))
    inf = net.InferenceNet(fp32)
    assert inf.fused_tower and inf.tower_layout == layout
    h0 = torch.rand(1000, 6, 7, 128, device="cuda").to(torch.bfloat16)
    a = inf.tower(h0).float()
    b = inf.tower_library(h0).float()
    # both are bf16 pipelines with float32 accumulation; the library rounds the shortcut to bf16 before the add
    assert float((a - b).abs().max()) <= 2 ** -5 * max(1.0, float(b.abs().max()))
    assert float(((a - b).abs() > 2 ** -6 * b.abs().clamp(min=1.0)).float().mean()) < 0.01


def test_fused_tower_is_deterministic_and_independent_of_the_batch_split(layout):
    """A position's output may not depend on which tile / CTA it lands in."""
    blocks = _blocks(2, seed=11)
    g = torch.Generator().manual_seed(12)
    x = torch.rand(1000, 6, 7, 128, generator=g).to(torch.bfloat16)
    full = _run(x, blocks, 6, 7)
    again = _run(x, blocks, 6, 7)
    assert torch.equal(full, again)
    part = _run(x[301:555], blocks, 6, 7)
    assert torch.equal(full[301:555], part)


def _states(n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    code = torch.randint(0, 3, (n, H, W), generator=g)
    x = torch.zeros(n, H, W, 4)
    x.scatter_(3, code[..., None], 1.0)
    x[..., 3] = torch.where(torch.rand(n, generator=g) < 0.5, 1.0, -1.0)[:, None, None]  # side to move (board.py:83-98)
    return x


@pytest.mark.parametrize("n", [1, 3, 500, 4097])
def test_whole_net_kernel_equals_the_float32_module(n, layout):
    """az_net_forward (stem + tower + heads in one kernel) against the fp32 PolicyValueNet and against the
    multi-kernel bf16 route it replaces."""
    import os

    engine, native, net = _mods()
    torch.manual_seed(7)
    fp32 = net.randomise_bn(net.PolicyValueNet(6, 7, 7)).eval()
    inf = net.InferenceNet(fp32)
    assert inf.fused_net
    x = _states(n, 6, 7, seed=n)
    p, v = inf(x.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        p32, v32 = fp32(x)
    p, v = p.cpu(), v.cpu()
    assert torch.isfinite(p).all() and torch.allclose(p.sum(-1), torch.ones(n), atol=1e-5)
    # bf16 activations through 13 layers against float32: same bounds as the multi-kernel route (tests/test_gpu_net.py)
    print("whole net vs fp32", n, float((p - p32).abs().max()), float((v - v32.reshape(-1)).abs().max()))
    assert float((p - p32).abs().max()) <= 8e-3 and float((v - v32.reshape(-1)).abs().max()) <= 8e-3
    os.environ["AZ_FUSED_NET"] = "0"
    try:
        old = net.InferenceNet(fp32)
    finally:
        del os.environ["AZ_FUSED_NET"]
    assert inf.tower_layout == layout
    assert not old.fused_net and old.fused_tower
    p2, v2 = old(x.cuda())
    # identical bf16 pipeline up to summation order inside the heads
    assert float((p - p2.cpu()).abs().max()) <= 1e-3 and float((v - v2.cpu()).abs().max()) <= 2e-3


def test_whole_net_kernel_is_batch_independent(layout):
    engine, native, net = _mods()
    torch.manual_seed(8)
    inf = net.InferenceNet(net.randomise_bn(net.PolicyValueNet(6, 7, 7)))
    x = _states(700, 6, 7, seed=1).cuda()
    p, v = inf(x)
    p, v = p.clone(), v.clone()
    p2, v2 = inf(x[100:461])
    assert torch.equal(p[100:461], p2) and torch.equal(v[100:461], v2)


def test_gathered_batch_evaluates_exactly_the_listed_trees(layout):
    """az_net_forward_gathered: rows index[:count] get the same priors / values as a plain forward, every other row keeps
    what it held; count = 0 is a no-op."""
    engine, native, net = _mods()
    torch.manual_seed(9)
    inf = net.InferenceNet(net.randomise_bn(net.PolicyValueNet(6, 7, 7)))
    n = 1000
    x = _states(n, 6, 7, seed=2).cuda().to(torch.bfloat16)
    want_p, want_v = inf(x)
    g = torch.Generator().manual_seed(1)
    perm = torch.randperm(n, generator=g)
    for k in (0, 1, 2, 371, 1000):
        index = torch.full((n,), -1, dtype=torch.int32)
        index[:k] = perm[:k].to(torch.int32)
        index = index.clamp(min=0).cuda()
        count = torch.tensor([k], dtype=torch.int32, device="cuda")
        p = torch.full((n, 7), -5.0, device="cuda")
        v = torch.full((n,), -5.0, device="cuda")
        inf(x, p, v, index=index, count=count)
        torch.cuda.synchronize()
        listed = torch.zeros(n, dtype=torch.bool)
        listed[perm[:k]] = True
        assert torch.equal(p[listed.cuda()], want_p[listed.cuda()]) and torch.equal(v[listed.cuda()], want_v[listed.cuda()])
        assert (p[~listed.cuda()] == -5.0).all() and (v[~listed.cuda()] == -5.0).all()


def test_whole_net_route_plays_the_same_games_as_the_per_tree_fused_route(layout):
    """SelfPlayRunner routes: az_step_gather + az_net_forward_gathered against az_advance_fused + az_net_tower - same
    weights, same seeds: identical games (the evaluator is a pure function of the position; heads differ by summation
    order only, far below what could flip a visit count here)."""
    from az_b200 import selfplay

    engine, native, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    out = []
    for whole in (True, False):
        torch.manual_seed(0)
        fp32 = net.randomise_bn(net.PolicyValueNet())
        r = selfplay.SelfPlayRunner(rules, n_trees=96, sims_per_move=48, net=fp32, games_target=160, unroll=4, seed=3, whole_net=whole)
        assert r.whole_net == whole
        r.run_until_done(poll_every=64, max_advances=400000)
        fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
        order = np.argsort(fin["game_id"])
        out.append(({k: v[order] for k, v in fin.items()}, r.totals()))
    (a, ta), (b, tb) = out
    assert ta["games"] == tb["games"] == 160
    same = sum(int(np.array_equal(a["action"][g][: a["len"][g]], b["action"][g][: b["len"][g]])) for g in range(160))
    assert same >= 150, same  # a rare near-tie may resolve differently between the two head implementations


@pytest.mark.parametrize("T,G", [(96, 160), (5, 9)])
def test_tree_warps_inside_the_net_kernel_change_no_game(layout, T, G):
    """az_net_forward_trees: trees without a pending leaf go on with evaluator-free simulations inside the net kernel
    (short max_free_sims in az_step, the rest under the net).  The simulations of a tree are the same in the same order,
    only earlier, and the net kernel is batch independent: every game, move for move and visit for visit, and the totals
    must equal the default route's."""
    from az_b200 import selfplay

    engine, native, net = _mods()
    rules = engine.Rules(7, 6, 4, True)
    out = []
    for mf, inside, beside in ((8, 0, 0), (2, 8, 0), (1, 3, 0), (2, 0, 6)):  # beside: az_extra_sims on a side stream
        torch.manual_seed(0)
        fp32 = net.randomise_bn(net.PolicyValueNet())
        r = selfplay.SelfPlayRunner(rules, n_trees=T, sims_per_move=48, net=fp32, games_target=G, unroll=4, seed=3,
                                    max_free_sims=mf, net_tree_sims=inside, extra_sims=beside)
        assert r.whole_net and r.net_tree_sims == inside and r.extra_sims == beside
        r.run_until_done(poll_every=64, max_advances=400000)
        fin = {k: v.cpu().numpy() for k, v in r.finished_device().items()}
        order = np.argsort(fin["game_id"])
        out.append(({k: v[order] for k, v in fin.items()}, r.totals()))
    (a, ta) = out[0]
    assert ta["games"] == G
    for b, tb in out[1:]:
        assert tb["games"] == G and tb["sims"] == ta["sims"] and tb["evals"] == ta["evals"] and tb["moves"] == ta["moves"]
        for k in ("game_id", "len", "result"):
            assert np.array_equal(a[k], b[k]), k
        for g in range(G):
            n = int(a["len"][g])
            assert np.array_equal(a["action"][g][:n], b["action"][g][:n]) and np.array_equal(a["visits"][g][:n], b["visits"][g][:n]), g


def test_tree_warps_need_the_plain_connect4_engine():
    engine, native, net = _mods()
    from az_b200.engine import _ptr, _stream

    inf = net.InferenceNet(net.PolicyValueNet(6, 7, 7))
    eng = engine.TreeEngine(engine.Rules(7, 6, 4, True), n_trees=8, sims_per_move=10, dirichlet_noise=True)
    x = torch.zeros((8, 6, 7, 4), dtype=torch.bfloat16, device="cuda")
    p = torch.zeros((8, 7), dtype=torch.float32, device="cuda")
    v = torch.zeros(8, dtype=torch.float32, device="cuda")
    idx = torch.zeros(8, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(native.NativeError):
        inf(x, p, v, index=idx, count=cnt, trees=(eng._h, 4))
