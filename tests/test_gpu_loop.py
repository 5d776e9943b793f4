"""GPU test of the closed loop: self-play -> samples -> training step -> refreshed
inference weights -> next self-play iteration, on one rank."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_selfplay_training_loop_updates_the_weights_used_by_selfplay():
    from az_b200 import engine, net, selfplay, train

    rules = engine.Rules(7, 6, 4, True)
    torch.manual_seed(0)
    fp32 = net.PolicyValueNet()
    runner = selfplay.SelfPlayRunner(rules, n_trees=64, sims_per_move=16, net=fp32, games_target=96, unroll=4, seed=3)
    trainer = train.Trainer(fp32, device="cuda")
    window = train.ReplayWindow(6, 7, 7, capacity=2000)
    train.MIN_TRAINING_SIZE, saved = 300, train.MIN_TRAINING_SIZE
    before = runner.net.flat_weights().clone()
    logs = []
    try:
        hist = train.selfplay_training_loop(runner, trainer, window, iterations=3, train_steps_per_iteration=2,
                                            rng=np.random.RandomState(0), log=lambda *a: logs.append(a))
    finally:
        train.MIN_TRAINING_SIZE = saved
    assert len(logs) == 3 and len(window) > 300
    assert len(hist) >= 2 and all(np.isfinite(h["loss"]) for h in hist) and trainer.steps == len(hist)
    after = runner.net.flat_weights()
    assert not torch.equal(before, after)  # the captured graphs now replay with the trained weights
    # the folded inference weights equal a fresh fold of the trained fp32 module
    fresh = net.InferenceNet(trainer.net, device="cuda").flat_weights()
    assert torch.equal(after, fresh)
    assert runner.totals()["games"] == 96
