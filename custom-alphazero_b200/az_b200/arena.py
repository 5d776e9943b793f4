"""Arena on the GPU: the reference's evaluate_two_models in policy-only mode
(evaluation/evaluate.py:29-134, the default: ConfigServing.evaluate_with_mcts = False), all games in lock-step.

Per ply and game: the net of the side to move gives probabilities [A]; the legal ones (action-list order) are
normalised (mcts/utils.py:4-16); their argmax (deterministic) or one np.random.choice draw gives a rank j, and the
move is board.moves[j];
Board.play(move, keep_same_player=True); the other net moves next.  Game g is opened by the candidate when g is
even (evaluate.py:39).  Score = candidate wins / decisive games, 0.5 if all games were drawn (evaluate.py:124-129).
Positions live on the device as int8 cells and are advanced by the K2/K3 kernels (az_env_*); torch does the
per-ply bookkeeping (masking, cumulative sums) - this is a once-every-50-training-steps path, not the hot path.
"""
import numpy as np
import torch

from . import env
from .engine import Rules
from .net import InferenceNet, PolicyValueNet

EVALUATION_GAMES = 150  # ConfigServing.evaluation_games_number
REPLACE_MIN_SCORE = 0.55  # ConfigServing.replace_min_score


def _board_order(rules: Rules, device):
    """Permutation: position in board move order -> action index (gravity: identity; free: row-major cells)."""
    if rules.gravity:
        return torch.arange(rules.n_actions, device=device)
    order = sorted(range(rules.n_actions), key=lambda a: (a % rules.height, a // rules.height))
    return torch.tensor(order, device=device)


@torch.no_grad()
def play_arena(net_current, net_previous, rules: Rules, games=EVALUATION_GAMES, deterministic=False, rng=np.random,
               dtype=torch.bfloat16, device="cuda"):
    """Returns the per-game results from the candidate's point of view (+1 win, -1 loss, 0 draw)."""
    device = torch.device(device)
    nets = []
    for n in (net_current, net_previous):
        nets.append(n if isinstance(n, InferenceNet) else InferenceNet(n, dtype=dtype, device=device))
    A = rules.n_actions
    cells = torch.zeros((games, rules.height, rules.width), dtype=torch.int8, device=device)
    mover = (torch.arange(games, device=device) % 2)  # 0 = candidate to move, 1 = previous model to move
    live = torch.ones(games, dtype=torch.bool, device=device)
    results = torch.zeros(games, dtype=torch.int64, device=device)
    order = _board_order(rules, device)
    for _ in range(rules.max_plies):
        if not bool(live.any()):
            break
        idx = torch.nonzero(live).flatten()
        c = cells[idx].contiguous()
        states = env.env_encode(rules, c)  # K3
        probs = torch.empty((idx.numel(), A), dtype=torch.float32, device=device)
        for k, net in enumerate(nets):
            sel = torch.nonzero(mover[idx] == k).flatten()
            if sel.numel():
                x = states[sel].to(net.dtype) if net.dtype != torch.float32 else states[sel]
                p, _ = net(x.contiguous())
                probs[sel] = p.float()
        legal = env.env_legal(rules, c)  # K2: mask in action order
        # probabilities[legal_moves_mask]: the legal probabilities in ACTION-LIST order, normalised in float32
        lp = torch.where(legal, probs, torch.zeros_like(probs))
        s = lp.sum(dim=1, keepdim=True)
        k = legal.sum(dim=1, keepdim=True).float()
        norm = torch.where(s == 0, legal.float() / k, lp / torch.where(s == 0, torch.ones_like(s), s))
        if deterministic:
            score = torch.where(legal, norm, torch.full_like(norm, -1.0))
            slot = torch.argmax(score, dim=1)  # first maximum among the legal entries
        else:
            u = torch.as_tensor(rng.random_sample(idx.numel()), dtype=torch.float64, device=device)
            cdf = torch.cumsum(norm.double(), dim=1)
            cdf = cdf / cdf[:, -1:]
            cdf = torch.where(legal, cdf, torch.full_like(cdf, -1.0))  # an illegal slot can never be the first > u
            slot = (cdf > u[:, None]).float().argmax(dim=1)
        # ... and the chosen RANK j among them indexes board.moves, which is in BOARD order (evaluate.py:45-52):
        # the same pairing by rank as in the search (quirk Q1; identical orders when gravity is on)
        rank = torch.cumsum(legal.long(), dim=1).gather(1, slot[:, None]) - 1
        lm = legal[:, order]
        pick = ((torch.cumsum(lm.long(), dim=1) == rank + 1) & lm).float().argmax(dim=1)
        action = order[pick].to(torch.int32)
        out, status = env.env_play(rules, c, action)  # K2
        assert bool((status >= 0).all())
        cells[idx] = out
        done = status != 0
        won = status == 1
        res = torch.where(won, torch.where(mover[idx] == 0, 1, -1), torch.zeros_like(status, dtype=torch.int64))
        results[idx[done]] = res[done].to(torch.int64)
        live[idx[done]] = False
        mover[idx[~done]] = 1 - mover[idx[~done]]
    return results.cpu().numpy()


def evaluate_two_models(net_current, net_previous, rules=Rules(), games=EVALUATION_GAMES, deterministic=False, rng=np.random,
                        dtype=torch.bfloat16, device="cuda"):
    """(score, None) like the reference (the second element is the solver score, out of scope)."""
    r = play_arena(net_current, net_previous, rules, games, deterministic, rng, dtype, device)
    if np.all(r == 0):
        return 0.5, None
    return float((r == 1).sum() / (r != 0).sum()), None
