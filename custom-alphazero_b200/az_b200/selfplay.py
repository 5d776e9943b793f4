"""SelfPlayRunner - the batched self-play loop on one GPU: tree kernels + bf16 net under CUDA graphs.

One lock-step iteration ("advance") of all trees is
    az_step  (consume last evaluation: expand + backup; select next leaf; encode it)
 -> InferenceNet forward on the [T, H, W, 4] bf16 leaf batch   (the only dense contraction)
 -> az_play  (trees whose move budget is spent: record, move, re-root, finish / refill)
`unroll` such iterations are captured into one CUDA graph and replayed; nothing synchronises with the
host inside.  Replaces the per-process loop of the reference's play_game (self_play.py:37-82) and the
joblib fan-out of play (self_play.py:85-119): games are independent, so they become the batch.
"""
import ctypes
import time

import numpy as np
import torch

from . import native
from .engine import Rules, TreeEngine, _ptr, _stream
from .env import _cfg
from .native import check, lib
from .net import InferenceNet, PolicyValueNet, flops_per_eval


def decode_samples(rules: Rules, fin_dev, exclude_null_games=False, with_distance=False):
    """Finished-game records (device tensors, ring layout) -> (states f32 [S,H,W,4], policies f64 [S,A],
    values int64 [S]) on the host, via the az_decode_samples kernel.  Games are emitted in game-id order.
    with_distance: a fourth array, plies from each sample to the end of its game (the exponent of
    ConfigSelfPlay.discounting_factor, self_play.py:77-78 of the reference)."""
    n = fin_dev["len"].numel()
    A = rules.n_actions
    if n == 0:
        out = (np.zeros((0, rules.height, rules.width, 4), np.float32), np.zeros((0, A)), np.zeros(0, np.int64))
        return out + (np.zeros(0, np.int64),) if with_distance else out
    order = torch.argsort(fin_dev["game_id"])
    lens = fin_dev["len"][order].contiguous()
    results = fin_dev["result"][order].contiguous()
    boards = fin_dev["board"][order].contiguous()
    visits = fin_dev["visits"][order].contiguous()
    actions = fin_dev["action"][order].contiguous()
    if exclude_null_games:  # self_play.py:155-162 drops every sample of a drawn game (reward 0)
        lens = torch.where(results == 0, torch.zeros_like(lens), lens)
    offsets = (torch.cumsum(lens, 0) - lens).to(torch.int32)
    S = int(lens.sum())
    dev = lens.device
    states = torch.empty((S, rules.height, rules.width, 4), dtype=torch.float32, device=dev)
    policies = torch.empty((S, A), dtype=torch.float64, device=dev)
    values = torch.empty(S, dtype=torch.int32, device=dev)
    cfg = _cfg(rules)
    check(lib().az_decode_samples(ctypes.byref(cfg), _ptr(boards), _ptr(visits), _ptr(actions), _ptr(lens),
                                  _ptr(results), _ptr(offsets), n, _ptr(states), _ptr(policies), _ptr(values),
                                  _stream()))
    out = (states.cpu().numpy(), policies.cpu().numpy(), values.cpu().numpy().astype(np.int64))
    if with_distance:
        host_lens = lens.cpu().numpy()
        dist_to_end = np.concatenate([np.arange(int(ln))[::-1] for ln in host_lens]) if S else np.zeros(0, np.int64)
        out = out + (dist_to_end.astype(np.int64),)
    return out


class _Group:
    """One slice of the trees with its own engine and evaluator buffers."""

    def __init__(self, engine, rules, dtype, device):
        T, A = engine.n_trees, rules.n_actions
        self.engine = engine
        self.states = torch.zeros((T, rules.height, rules.width, 4), dtype=dtype, device=device)
        self.valid = torch.zeros(T, dtype=torch.int32, device=device)
        self.priors = torch.zeros((T, A), dtype=torch.float32, device=device)
        self.values = torch.zeros(T, dtype=torch.float32, device=device)
        # fused route: the stem's output (tower input) and the tower's output carried between advances
        self.stem_out = None
        self.tower_out = None
        self.tower_carry = None
        # whole-net route: the trees whose leaf awaits evaluation (az_step_gather) and their number
        self.leaf_list = torch.zeros(T, dtype=torch.int32, device=device)
        self.leaf_count = torch.zeros(1, dtype=torch.int32, device=device)


class SelfPlayRunner:
    """`groups` > 1 splits the trees into independent slices whose advances are captured on parallel
    branches of the CUDA graph: while the tensor cores run the net of one slice, the latency-bound tree
    kernels (and the memory-bound stem / heads) of another slice run underneath."""

    def __init__(self, rules=Rules(), n_trees=4096, sims_per_move=800, net=None, *, games_target=None,
                 game_id_base=0, seed=0, move_mode="philox", auto_restart=True, dtype=torch.bfloat16, unroll=8,
                 use_graph=True, max_free_sims=None, node_capacity=None, fin_capacity=None, device=None,
                 index_move_greedy=8, groups=1, fused=True, extra_sims=0, dirichlet_noise=False, dirichlet_alpha=0.03,
                 dirichlet_ratio=0.25, eval_cache_log2=0, whole_net=None, net_tree_sims=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rules = rules
        T, A = int(n_trees), rules.n_actions
        if net is None:
            net = PolicyValueNet(rules.height, rules.width, A)
        self.fp32_net = net
        self.net = net if isinstance(net, InferenceNet) else InferenceNet(net, dtype=dtype, device=self.device)
        groups = max(1, min(int(groups), T))
        if games_target is None:
            games_target = T
        # Routes of one advance.  whole_net (default where the net has it): az_step (the tree step alone) + az_net_forward
        # (stem, tower and heads in ONE tcgen05 kernel: planes in, priors / values out, no activation in HBM).  Else
        # fused: heads + tree step + stem in one per-tree launch (az_advance_fused) around the tower kernel; else the
        # three-kernel route az_step + stem + tower + heads.
        has_net = bool(getattr(self.net, "fused_net", False))
        self.whole_net = has_net if whole_net is None else (bool(whole_net) and has_net)
        # whole-net route on the 6x7 fast path: trees without a pending leaf go on with up to net_tree_sims evaluator-free
        # simulations INSIDE the net kernel (az_net_forward_trees), so az_step only needs a short max_free_sims: the serial
        # tail of the tree step (a few trees running up to 8 terminal-leaf simulations one after the other while 4 000
        # others wait) moves under the net.  Measured (bench.py, profiles/experiments_r2_session2.json): 2 + 16 gives
        # 14.6 M simulations/s against 12.7 M for 8 + 0.
        can_inside = (self.whole_net and not dirichlet_noise
                      and (rules.width, rules.height, rules.n, bool(rules.gravity)) == (7, 6, 4, True))
        self.net_tree_sims = (16 if net_tree_sims is None else int(net_tree_sims)) if can_inside else 0
        if max_free_sims is None:
            max_free_sims = 2 if self.net_tree_sims else 8
        self.max_free_sims = int(max_free_sims)
        self.groups = []
        t0 = g0 = 0
        for i in range(groups):
            ti = T // groups + (1 if i < T % groups else 0)
            gi = games_target // groups + (1 if i < games_target % groups else 0)
            eng = TreeEngine(rules, ti, sims_per_move, eval_mode="external", prior_mode="f32", move_mode=move_mode,
                             games_target=gi, game_id_base=game_id_base + g0, seed=seed, auto_restart=auto_restart,
                             max_free_sims=max_free_sims, node_capacity=node_capacity,
                             fin_capacity=None if fin_capacity is None else max(1, -(-fin_capacity // groups)),
                             device=self.device, index_move_greedy=index_move_greedy, inline_play=True,
                             dirichlet_noise=dirichlet_noise, dirichlet_alpha=dirichlet_alpha, dirichlet_ratio=dirichlet_ratio,
                             eval_cache_log2=eval_cache_log2)
            self.groups.append(_Group(eng, rules, dtype, self.device))
            t0 += ti
            g0 += gi
        self.n_trees = T
        self.fused = bool(fused) and getattr(self.net, "fast", False) and not self.whole_net
        if self.fused:
            for g in self.groups:
                shape = (g.engine.n_trees, rules.height, rules.width, self.net.filters)
                g.stem_out = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
                g.tower_carry = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
        # evaluator-free simulations (terminal leaves, in-line moves) continued on a forked stream while the
        # tower runs: keeps max_free_sims - the tail of the per-tree kernel, which the tower waits for - small
        self.extra_sims = int(extra_sims)
        self._extra_streams = [torch.cuda.Stream(device=self.device, priority=0) for _ in self.groups] if self.extra_sims else []
        self.unroll = int(unroll)
        self.use_graph = use_graph
        self.graph = None
        self.advances = 0
        self.flops_per_eval = flops_per_eval(rules.height, rules.width, A)
        # kernels of libaz_b200 launched per advance and group: az_advance_fused, or az_step + az_net_stem +
        # az_net_heads (+ az_play sweeps); + az_net_tower when the tower is the hand-written kernel
        if self.whole_net:
            self.launches_per_advance = (3 if extra_sims else 2) * groups
        else:
            self.launches_per_advance = ((1 if self.fused else 3) + (1 if extra_sims and self.fused else 0)
                                         + (1 if getattr(self.net, "fused_tower", False) else 0)) * groups
        self._side = [torch.cuda.Stream(device=self.device) for _ in range(groups - 1)]

    # single-group conveniences (tests, compat code)
    @property
    def engine(self):
        return self.groups[0].engine

    @property
    def states(self):
        return self.groups[0].states

    @property
    def valid(self):
        return self.groups[0].valid

    @property
    def priors(self):
        return self.groups[0].priors

    @property
    def values(self):
        return self.groups[0].values

    # one lock-step iteration of one group; everything is enqueued on the current stream
    def _advance(self, g, sweep=True):
        """az_step plays moves itself (inline_play) whenever the re-root fits in place; az_play is only needed
        for the compaction path and for games stalled on a full ring: every advance if the pool can run
        short, else once per captured graph as a safety net (`sweep`)."""
        if self.fused:
            # no tree has a pending leaf right after a reset, so a stale carry buffer is never consumed
            src = g.tower_out if g.tower_out is not None else g.tower_carry
            hw = self.net._heads_arg()
            check(lib().az_advance_fused(g.engine._h, _ptr(src), ctypes.byref(hw), _ptr(self.net.stem_w32),
                                         _ptr(self.net.stem_b32), _ptr(g.stem_out), _ptr(g.valid), _stream()))
            if self.extra_sims:
                cur = torch.cuda.current_stream()
                side = self._extra_streams[self.groups.index(g)]
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    g.engine.extra_sims(self.extra_sims)
                g.tower_out = self.net.tower(g.stem_out)
                cur.wait_stream(side)
            else:
                g.tower_out = self.net.tower(g.stem_out)
        elif self.whole_net:
            # the tree step alone, then the whole net in one kernel on exactly the trees that have a leaf pending
            g.engine.step(g.priors, g.values, g.states, g.valid, g.leaf_list, g.leaf_count)
            if self.net_tree_sims:
                self.net(g.states, g.priors, g.values, index=g.leaf_list, count=g.leaf_count, trees=(g.engine._h, self.net_tree_sims))
            elif self.extra_sims:
                # trees that spent their max_free_sims evaluator-free simulations without meeting a leaf that needs the net
                # have nothing to wait for: they go on simulating on a low-priority side stream, whose blocks the scheduler
                # places on the SMs the net kernel's last, partial round of tiles leaves idle
                cur = torch.cuda.current_stream()
                side = self._extra_streams[self.groups.index(g)]
                side.wait_stream(cur)
                self.net(g.states, g.priors, g.values, index=g.leaf_list, count=g.leaf_count)
                with torch.cuda.stream(side):
                    g.engine.extra_sims(self.extra_sims)
                cur.wait_stream(side)
            else:
                self.net(g.states, g.priors, g.values, index=g.leaf_list, count=g.leaf_count)
        else:
            g.engine.step(g.priors, g.values, g.states, g.valid)
            self.net(g.states, g.priors, g.values)
        if sweep or not g.engine.never_compacts:
            g.engine.play()

    def _advance_all(self, n):
        """n advances of every group: group 0 on the current stream, the others on forked side streams."""
        cur = torch.cuda.current_stream()
        for s in self._side:
            s.wait_stream(cur)
        def chain(g):
            # the first advance of a batch reads the tower output the previous batch left in tower_carry (a fixed
            # address, so a captured graph can be replayed back to back); the last one refreshes it
            g.tower_out = None
            for i in range(n):
                self._advance(g, sweep=i == n - 1)
            if self.fused:
                g.tower_carry.copy_(g.tower_out)
                g.tower_out = None

        for g, s in zip(self.groups[1:], self._side):
            with torch.cuda.stream(s):
                chain(g)
        chain(self.groups[0])
        for s in self._side:
            cur.wait_stream(s)

    def capture(self):
        if self.graph is not None or not self.use_graph:
            return
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm up cuDNN heuristics / workspaces outside capture
            for _ in range(3):
                for g in self.groups:
                    self.net(g.states, g.priors, g.values)
                    if self.fused:
                        self.net.tower(g.stem_out)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        # captured on a high-priority stream: kernel nodes keep the priority of the stream they were captured on, so the
        # net and the tree step win the SMs over the low-priority az_extra_sims branch whenever both have blocks waiting
        with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=self.device, priority=-1)):
            self._advance_all(self.unroll)
        self.graph = graph

    def run(self, advances):
        """Enqueues `advances` lock-step iterations (rounded up to a multiple of `unroll` under graphs)."""
        if self.use_graph:
            self.capture()
            n = (advances + self.unroll - 1) // self.unroll
            for _ in range(n):
                self.graph.replay()
            self.advances += n * self.unroll
            return n * self.unroll
        self._advance_all(advances)
        self.advances += advances
        return advances

    def reset(self, game_id_base=None):
        """Fresh games in every tree.  With game_id_base the groups' id ranges move to [base, base + games_target) and
        the captured graph (which holds the old range in its kernel arguments) is dropped."""
        off = 0
        for g in self.groups:
            g.engine.reset(None if game_id_base is None else int(game_id_base) + off)
            off += int(g.engine.cfg.games_target)
            g.valid.zero_()
            g.tower_out = None
        if game_id_base is not None:
            self.graph = None

    def active_trees(self):
        return sum(int((g.engine.phases() != native.AZ_PHASE_IDLE).sum()) for g in self.groups)

    def run_until_done(self, poll_every=64, max_advances=None):
        """Runs until every tree is idle (all games_target games finished); returns advances executed.  A finished-game
        ring smaller than games_target would stall the trees for ever (they wait for a free slot): that is an error
        here, not a silent spin - size fin_capacity >= games_target or drain with collect() between run() calls."""
        done = 0
        while True:
            done += self.run(poll_every)
            if self.active_trees() == 0:
                break
            if max_advances is not None and done >= max_advances:
                break
            for g in self.groups:
                stalled = int((g.engine.phases() == native.AZ_PHASE_STALLED).sum())
                full = int(g.engine.view("fin_count")[0]) >= int(g.engine.cfg.fin_capacity)
                if stalled and full:
                    raise RuntimeError("finished-game ring full (%d slots) with %d trees waiting for a slot: "
                                       "fin_capacity must cover games_target, or drain it with collect()" %
                                       (int(g.engine.cfg.fin_capacity), stalled))
        self.check_status()
        return done

    def check_status(self):
        for g in self.groups:
            g.engine.check_status()

    def totals(self):
        out = {}
        for g in self.groups:
            for k, v in g.engine.totals().items():
                out[k] = out.get(k, 0) + v
        return out

    def fin_clear(self):
        for g in self.groups:
            g.engine.fin_clear()

    def finished_device(self):
        parts = []
        for g in self.groups:
            e = g.engine
            n = int(e.view("fin_count")[0])
            parts.append({"game_id": e.view("fin_game_id")[:n], "len": e.view("fin_len")[:n],
                          "result": e.view("fin_result")[:n], "visits": e.view("fin_visits")[:n],
                          "action": e.view("fin_action")[:n], "board": e.view("fin_board")[:n]})
        if len(parts) == 1:
            return parts[0]
        return {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0]}

    def collect(self, exclude_null_games=False, with_distance=False):
        """Decodes and drains the finished-game rings: (states, policies, values[, plies to the end]) host arrays."""
        out = decode_samples(self.rules, self.finished_device(), exclude_null_games, with_distance)
        self.fin_clear()
        return out

    def load_flat_weights(self, flat):
        """The inference copy's parameters from one flat float32 vector (what the trainer rank broadcasts,
        InferenceNet.flat_weights), and the memoised evaluations forgotten - they belong to the old weights."""
        off = 0
        with torch.no_grad():
            for q in self.net.parameters():
                q.copy_(flat[off: off + q.numel()].view_as(q))
                off += q.numel()
        for g in self.groups:
            g.engine.cache_clear()

    def load_weights(self, net: PolicyValueNet):
        """New weights: refresh the folded inference copy and forget the memoised evaluations (the reference
        resets plays_inferences when the best-model hash changes, self_play.py:142-150)."""
        self.fp32_net = net
        self.net.load_from(net)
        for g in self.groups:
            g.engine.cache_clear()


def smoke_net_step():
    """Tiny end-to-end: 64 trees x 16 simulations per move through the bf16 net, a few full games."""
    rules = Rules(7, 6, 4, True)
    torch.manual_seed(0)
    r = SelfPlayRunner(rules, n_trees=64, sims_per_move=16, games_target=96, unroll=4, groups=2)
    t0 = time.time()
    r.run_until_done(poll_every=64, max_advances=200000)
    torch.cuda.synchronize()
    tot = r.totals()
    states, policies, values = r.collect()
    assert tot["games"] == 96 and states.shape[0] == policies.shape[0] == values.shape[0] == tot["moves"]
    assert np.isfinite(policies).all() and np.allclose(policies.sum(-1), 1.0)
    assert set(np.unique(values)).issubset({-1, 0, 1})
    print("net smoke ok: %d games, %d samples, %d sims in %.2fs" % (tot["games"], len(values), tot["sims"], time.time() - t0))
