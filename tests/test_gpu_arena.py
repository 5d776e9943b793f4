"""GPU test of the arena against oracle/arena_ref.py (restatement of evaluation/evaluate.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("W,H,n,gravity", [(7, 6, 4, True), (5, 5, 3, False)])
def test_arena_matches_the_restatement(W, H, n, gravity):
    from az_b200 import arena, engine, net
    from oracle import arena_ref, ref_port

    rules = engine.Rules(W, H, n, gravity)
    rrules = ref_port.Rules(W, H, n, gravity)
    torch.manual_seed(0)
    a = net.randomise_bn(net.PolicyValueNet(H, W, rules.n_actions)).eval()
    torch.manual_seed(1)
    b = net.randomise_bn(net.PolicyValueNet(H, W, rules.n_actions), seed=1).eval()
    with torch.no_grad():  # sharpen the policies so that argmax / sampling is far from ties
        a.policy_fc.weight.mul_(8.0)
        b.policy_fc.weight.mul_(8.0)

    def host(model):
        def f(x):
            with torch.no_grad():
                p, v = model(torch.from_numpy(x))
            return p.numpy(), v.numpy()
        return f

    games = 24
    # deterministic: argmax of the legal probabilities, fp32 on both sides -> identical games
    got = arena.play_arena(a, b, rules, games, deterministic=True, dtype=torch.float32)
    ha, hb = host(a), host(b)
    want = [arena_ref.single_game(rrules, ha, hb, g, True, None) for g in range(games)]
    assert got.tolist() == want
    # games are opened alternately: with deterministic play there are only two distinct games
    assert len(set(got[0::2].tolist())) == 1 and len(set(got[1::2].tolist())) == 1
    # stochastic: same sampling rule; the score is a frequency in [0, 1] or 0.5 when everything is drawn
    score, _ = arena.evaluate_two_models(a, b, rules, games=64, rng=np.random.RandomState(0), dtype=torch.float32)
    assert 0.0 <= score <= 1.0
    same, _ = arena.evaluate_two_models(a, a, rules, games=64, rng=np.random.RandomState(0), dtype=torch.float32)
    assert 0.0 <= same <= 1.0
    assert arena_ref.score([0, 0, 0]) == 0.5 and arena_ref.score([1, -1, 0, 1]) == 2 / 3
