"""CPU baseline legs of bench.py (TEST / MEASUREMENT INFRASTRUCTURE - see oracle/__init__.py).

Times oracle/ref_port.py - the object-per-node Python restatement with the reference's own cost
profile - the way the reference runs self-play: os.cpu_count()-1 worker processes, one game each
(self_play.py:98-110), every leaf evaluated by a batch-1 fp32 forward of the policy/value net on one
CPU thread (mcts.py:131-137; ConfigGeneral.self_play_gpu_index = "-1", config.py:11), evaluations
memoised per worker (mcts.py:123-124).  The reference itself is pure Python and is not present on
the GPU box, so this is kind "port".
"""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_state = {}


def _init_worker(rules_tuple, sims, evaluator_kind, seed):
    for p in (ROOT, os.path.join(ROOT, "custom-alphazero_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np

    from oracle import evaluators, ref_port

    rules = ref_port.Rules(*rules_tuple)
    if evaluator_kind == "net":
        import torch

        torch.set_num_threads(1)
        from az_b200.net import PolicyValueNet

        torch.manual_seed(seed)
        net = PolicyValueNet(rules.height, rules.width, rules.n_actions).eval()

        def evaluator(state):
            with torch.no_grad():
                p, v = net(torch.from_numpy(state[None]))
            return p.numpy().ravel().astype(np.float64), float(v.item())
    else:
        evaluator = evaluators.make(evaluator_kind, rules.n_actions)
    _state.update(rules=rules, sims=sims, evaluator=evaluator, cache={}, search=None,
                  rng=np.random.RandomState(seed + os.getpid()), ref_port=ref_port)


def _one_move(_):
    """search(sims) + play for this worker's current game (a new game when the last one ended)."""
    rp = _state["ref_port"]
    s = _state["search"]
    if s is None or s.board.over:
        s = rp.RefSearch(rp.RefBoard(_state["rules"]), _state["evaluator"], _state["cache"])
        _state["search"] = s
    t0 = time.perf_counter()
    e0 = s.evals
    s.search(_state["sims"])
    greedy = s.board.plies >= rp.INDEX_MOVE_GREEDY
    s.play(greedy, deterministic=False, uniform=_state["rng"].random_sample())
    return _state["sims"], s.evals - e0, time.perf_counter() - t0


class CpuSelfPlay:
    """A pool of reference-style workers; step() makes every worker play one move."""

    def __init__(self, rules_tuple=(7, 6, 4, True), sims=800, evaluator="net", workers=None, seed=0):
        self.workers = workers or max(1, (os.cpu_count() or 2) - 1)
        self.sims = sims
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.workers, initializer=_init_worker, initargs=(rules_tuple, sims, evaluator, seed))

    def step(self):
        t0 = time.perf_counter()
        res = self.pool.map(_one_move, range(self.workers), chunksize=1)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), sum(r[1] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def measure(rules_tuple=(7, 6, 4, True), sims=800, evaluator="net", seconds=20.0, workers=None):
    """Bounded sample: every worker plays moves of `sims` simulations until ~`seconds` have passed."""
    sp = CpuSelfPlay(rules_tuple, sims, evaluator, workers)
    try:
        sp.step()  # warm-up move (imports, first torch call)
        total = evals = 0
        wall = 0.0
        moves = 0
        while wall < seconds:
            s, e, w = sp.step()
            total += s
            evals += e
            wall += w
            moves += sp.workers
    finally:
        sp.close()
    return {"sims_per_s": total / wall, "evals_per_s": evals / wall, "workers": sp.workers, "wall_s": wall,
            "moves": moves, "sims": total}
