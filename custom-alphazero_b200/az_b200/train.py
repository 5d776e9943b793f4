"""Training step on the GPU + replay window: closes the loop the weight broadcast and
the sample gather serve.  PyTorch autograd is the library here, as Keras is in the reference.

Reference semantics restated (paths relative to /root/reference/custom_alphazero/):
  train.py:16-84                replay window = last samples_queue_size (10 000) samples, training starts at
                                minimum_training_size (2 500); every iteration samples batch_size (256) indexes
                                WITHOUT replacement and runs model.fit for training_epochs (1) epoch = 1 SGD step
  model/tensorflow/train.py:14-44  after fit: steps += ceil(n / batch_size) * epochs; learning rate looked up in
                                ConfigModel.learning_rates by step, else minimum_learning_rate
  model/tensorflow/base_layers.py:12-17  policy loss = mean_b( sum_a( -pi * log(p + K.epsilon()) ) ), eps = 1e-7;
                                value loss = mean_b( (v - z)^2 ); Keras sums the two
  base_layers.py:36-49, model.py:86-91,129-139  kernel_regularizer l2(1e-4) on every Conv2D / Dense kernel:
                                + 1e-4 * sum(w^2) in the loss (biases and BN parameters are not regularised)
  model.py:163-166              SGD(learning_rate, momentum 0.9), Keras form: v = m*v - lr*g ; w += v
"""
import numpy as np
import torch
import torch.nn as nn

from .net import PolicyValueNet

KERAS_EPSILON = 1e-7
L2 = 1e-4
MOMENTUM = 0.9
LEARNING_RATES = {range(0, 150000): 1e-2, range(150000, 300000): 1e-3}
MIN_LEARNING_RATE = 1e-4
MAX_LEARNING_RATE = 1e-2
BATCH_SIZE = 256
WINDOW = 10000
MIN_TRAINING_SIZE = 2500


def learning_rate_for(steps):
    for rng, lr in LEARNING_RATES.items():
        if steps in rng:
            return lr
    return MIN_LEARNING_RATE


class ReplayWindow:
    """train.py:16-38: append, keep the last `capacity` samples."""

    def __init__(self, height, width, n_actions, capacity=WINDOW, planes=4):
        self.capacity = capacity
        self.states = np.empty((0, height, width, planes), dtype=np.float32)  # 4 planes Connect-N, 118 chess
        self.policies = np.empty((0, n_actions), dtype=np.float64)
        self.values = np.empty(0, dtype=np.float64)

    def __len__(self):
        return len(self.values)

    def append(self, states, policies, values):
        if len(values) == 0:
            return
        self.states = np.concatenate([self.states, states.astype(np.float32)])[-self.capacity:]
        self.policies = np.concatenate([self.policies, policies.astype(np.float64)])[-self.capacity:]
        self.values = np.concatenate([self.values, np.asarray(values, dtype=np.float64)])[-self.capacity:]

    def ready(self, minimum=None):
        return len(self) >= (MIN_TRAINING_SIZE if minimum is None else minimum)

    def sample(self, batch_size=BATCH_SIZE, rng=np.random):
        idx = rng.choice(len(self), batch_size, replace=False)  # train.py:58-62
        return self.states[idx], self.policies[idx], self.values[idx]


def regularised_parameters(net: nn.Module):
    """The kernels Keras puts an l2 penalty on: Conv2D and Dense weights."""
    return [m.weight for m in net.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]


def losses(net: PolicyValueNet, states, policies, values):
    """(policy loss, value loss, l2 penalty) with the reference's definitions."""
    p, v = net(states)
    policy_loss = torch.mean(torch.sum(-policies * torch.log(p + KERAS_EPSILON), dim=-1))
    value_loss = torch.mean((v.reshape(-1) - values) ** 2)
    reg = L2 * sum((w**2).sum() for w in regularised_parameters(net))
    return policy_loss, value_loss, reg


class Trainer:
    def __init__(self, net: PolicyValueNet, device="cuda"):
        self.net = net.to(device)
        self.device = torch.device(device)
        self.steps = 0
        self.lr = MAX_LEARNING_RATE  # model.py:163: SGD starts at maximum_learning_rate
        self.velocity = [torch.zeros_like(p) for p in self.net.parameters()]

    def train_step(self, states, policies, values):
        """One model.fit(batch, epochs=1) of the reference: one SGD step on the batch, then the step count
        and the learning rate advance.  Returns the scalars Keras would log."""
        net = self.net.train()
        x = torch.as_tensor(states, dtype=torch.float32, device=self.device)
        pi = torch.as_tensor(policies, dtype=torch.float32, device=self.device)
        z = torch.as_tensor(values, dtype=torch.float32, device=self.device)
        policy_loss, value_loss, reg = losses(net, x, pi, z)
        loss = policy_loss + value_loss + reg
        grads = torch.autograd.grad(loss, list(net.parameters()))
        with torch.no_grad():
            for p, g, v in zip(net.parameters(), grads, self.velocity):
                v.mul_(MOMENTUM).add_(g, alpha=-self.lr)  # Keras SGD: v = m*v - lr*g ; w += v
                p.add_(v)
        out = {"loss": float(loss.detach()), "policy_loss": float(policy_loss.detach()),
               "value_loss": float(value_loss.detach()), "l2": float(reg.detach()), "lr": self.lr, "steps": self.steps}
        n = len(values)
        self.steps += int(np.ceil(n / BATCH_SIZE))  # model/tensorflow/train.py:32
        self.lr = learning_rate_for(self.steps)
        net.eval()
        return out


def selfplay_training_loop(runner, trainer: Trainer, window: ReplayWindow, iterations, train_steps_per_iteration=1,
                           exclude_null_games=True, rng=np.random, log=None):
    """Self-play -> gather -> train -> broadcast, on one or several ranks (torchrun).  Rank 0 trains; the
    other ranks receive the weights through az_b200.dist.broadcast_weights.  `runner` must have been built
    with a finite games_target per iteration (see custom_alphazero.self_play.play)."""
    from . import dist as azdist
    from .selfplay import decode_samples

    rank, ws = azdist.world()
    history = []
    base0 = sum(int(g.engine.cfg.game_id_base) for g in runner.groups[:1])
    mine = sum(int(g.engine.cfg.games_target) for g in runner.groups)
    stride = mine * ws  # every rank plays `mine` games per iteration: ids never repeat across iterations or ranks
    for it in range(iterations):
        # a new id range every iteration: the ids key the move-sampling counters, so reusing them would replay the
        # same games until the weights change (the reference draws fresh random numbers every iteration)
        runner.reset(game_id_base=base0 + it * stride)
        runner.run_until_done()
        fin = azdist.gather_records({k: v.contiguous() for k, v in runner.finished_device().items()}, dst=0)
        runner.fin_clear()
        if rank == 0:
            s, p, v = decode_samples(runner.rules, fin, exclude_null_games=exclude_null_games)
            window.append(s, p, v)
            if window.ready():
                for _ in range(train_steps_per_iteration):
                    history.append(trainer.train_step(*window.sample(rng=rng)))
            runner.load_weights(trainer.net)
        flat = runner.net.flat_weights()
        azdist.broadcast_weights(flat, src=0)
        if rank != 0:
            runner.load_flat_weights(flat)  # also forgets the evaluations memoised under the old weights
        if log is not None and rank == 0:
            log(it, len(window), history[-1] if history else None)
    return history
