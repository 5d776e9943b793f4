"""CPU tests of the training step against oracle/train_ref.py (numpy restatement of
the reference's loss / regulariser / SGD formulas; unpinned: the reference trainer needs TensorFlow)."""
import numpy as np
import torch

from oracle import train_ref


def _batch(n, seed=0):
    g = np.random.RandomState(seed)
    code = g.randint(0, 3, (n, 6, 7))
    x = np.zeros((n, 6, 7, 4), np.float32)
    for c in range(3):
        x[..., c] = code == c
    x[..., 3] = 1
    pi = g.dirichlet(np.ones(7), n)
    z = g.choice([-1.0, 1.0], n)
    return x, pi, z


def test_losses_and_update_match_the_numpy_restatement():
    from az_b200 import net, train

    torch.manual_seed(0)
    m = net.PolicyValueNet()
    tr = train.Trainer(m, device="cpu")
    x, pi, z = _batch(32)
    m.train()
    pl, vl, reg = train.losses(m, torch.from_numpy(x), torch.from_numpy(pi).float(), torch.from_numpy(z).float())
    with torch.no_grad():
        p, v = m(torch.from_numpy(x))  # same batch statistics in train mode
    assert abs(float(pl) - train_ref.policy_loss(pi.astype(np.float32), p.numpy())) < 1e-5
    assert abs(float(vl) - train_ref.value_loss(z.astype(np.float32), v.numpy().ravel())) < 1e-5
    kernels = [w.detach().numpy() for w in train.regularised_parameters(m)]
    assert len(kernels) == 1 + 4 * 3 + 2 + 3  # stem, 4 x (c1, c2, proj), 2 head convs, 3 dense
    assert abs(float(reg) - train_ref.l2_penalty(kernels)) < 1e-6
    # one Keras-SGD step on a single tensor
    w0 = m.value_fc2.weight.detach().clone().numpy()
    loss = pl + vl + reg
    g = torch.autograd.grad(loss, m.value_fc2.weight)[0].numpy()
    out = tr.train_step(x, pi, z)
    w1, _ = train_ref.keras_sgd(w0, g, np.zeros_like(w0), lr=1e-2)
    # BN running stats moved between the two forward passes, the batch statistics did not: same gradient
    np.testing.assert_allclose(m.value_fc2.weight.detach().numpy(), w1, rtol=0, atol=2e-6)
    assert out["lr"] == 1e-2 and tr.steps == 1


def test_learning_rate_schedule_and_step_count():
    from az_b200 import train

    assert train.learning_rate_for(0) == 1e-2 and train.learning_rate_for(149999) == 1e-2
    assert train.learning_rate_for(150000) == 1e-3 and train.learning_rate_for(299999) == 1e-3
    assert train.learning_rate_for(300000) == 1e-4


def test_training_reduces_the_loss_and_window_semantics():
    from az_b200 import net, train

    torch.manual_seed(1)
    tr = train.Trainer(net.PolicyValueNet(), device="cpu")
    win = train.ReplayWindow(6, 7, 7, capacity=300)
    for s in range(5):
        win.append(*_batch(100, seed=s))
    assert len(win) == 300 and win.ready(250) and not win.ready(301)
    last = _batch(100, seed=4)
    np.testing.assert_array_equal(win.states[-100:], last[0])  # newest samples are kept
    rng = np.random.RandomState(0)
    xs, ps, zs = win.sample(64, rng)
    assert xs.shape == (64, 6, 7, 4) and len(np.unique(xs.reshape(64, -1), axis=0)) > 32
    first = tr.train_step(xs, ps, zs)["loss"]
    for _ in range(25):
        out = tr.train_step(xs, ps, zs)
    assert out["loss"] < first and tr.steps == 26
