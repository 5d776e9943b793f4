"""ChessSelfPlayRunner - batched chess self-play on one GPU (BASELINE config C5): the chess tree kernels + the bf16
policy/value net under CUDA graphs.

One lock-step advance of all trees is
    az_chess_step  consume the last evaluation (expand + backup), select the next leaf, write its 118 planes
 -> InferenceNet   stem Conv3x3(118 -> 128) + the same 4-block residual tower + heads (policy over the 1 880 actions)
 -> az_chess_move  trees whose move budget is spent: sample-ring entry, move, re-root, game end / next game
`unroll` advances are captured in one CUDA graph.  The host only drains the rings (between graph replays).
Replaces, for chess, the per-process loop of play_game (self_play.py:37-82) and the joblib fan-out (self_play.py:85-119).
"""
import numpy as np
import torch

from . import native
from .chess import N_ACTIONS, PLANES
from .chess_engine import ChessTreeEngine, decode_samples, sample_values
from .net import InferenceNet, PolicyValueNet, flops_per_eval


def chess_net(filters=128, depth=4):
    """The reference's architecture (model/tensorflow/model.py:21-188) at the chess shapes: 8x8x118 in, 1 880 actions."""
    return PolicyValueNet(8, 8, N_ACTIONS, filters=filters, depth=depth, in_planes=PLANES)


class ChessSelfPlayRunner:
    def __init__(self, n_trees=1024, sims_per_move=200, net=None, *, games_target=None, game_id_base=0, seed=0,
                 move_mode="philox", auto_restart=True, unroll=8, use_graph=True, max_free_sims=8, node_capacity=None,
                 max_plies=512, sample_capacity=None, device=None, index_move_greedy=8, stem_from_boards=True, tail_planes=False):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        import os

        route = os.environ.get("AZ_CHESS_STEM")  # experiment knob: boards | tail | planes
        if route:
            stem_from_boards, tail_planes = route == "boards", route == "tail"
        T = int(n_trees)
        if net is None:
            net = chess_net()
        self.fp32_net = net
        self.net = net if isinstance(net, InferenceNet) else InferenceNet(net, dtype=torch.bfloat16, device=self.device)
        if sample_capacity is None:
            sample_capacity = max(256, 8 * T)
        self.engine = ChessTreeEngine(T, sims_per_move, eval_mode="external", prior_mode="f32", move_mode=move_mode,
                                      games_target=games_target, game_id_base=game_id_base, seed=seed,
                                      auto_restart=auto_restart, max_free_sims=max_free_sims, node_capacity=node_capacity,
                                      max_plies=max_plies, sample_capacity=sample_capacity, device=self.device,
                                      index_move_greedy=index_move_greedy)
        self.n_trees = T
        # the step kernel writes the planes with the channel padding the tensor-core stem wants; tail_planes: only planes
        # 84-117 (34 -> 40 channels) - the six older history entries are always empty on the self-play path, so the stem
        # multiplies a third of the planes and the leaf batch is a third of the bytes, with identical results.  Measured:
        # cuDNN's stem is no faster at 40 input channels than at 120 (4.72 vs 4.77 M simulations/s): opt-in
        self.tail_planes = bool(tail_planes) and hasattr(self.net, "stem_w_tail")
        self.plane_first = 84 if self.tail_planes else 0
        width = 40 if self.tail_planes else PLANES + self.net.in_pad
        self.states = torch.zeros((T, 8, 8, width), dtype=torch.bfloat16, device=self.device)
        self.valid = torch.zeros(T, dtype=torch.int32, device=self.device)
        self.priors = torch.zeros((T, N_ACTIONS), dtype=torch.float32, device=self.device)
        self.values = torch.zeros(T, dtype=torch.float32, device=self.device)
        # stem_from_boards (default): az_chess_stem_tc computes the stem from the 64-byte leaf boards on tcgen05 (50 us at
        # 4096 positions against 64.5 us for cuDNN on the 120-plane tensor) and no plane tensor exists at all (63 MB less
        # traffic per advance): 4.89 vs 4.69 M simulations/s on the same box
        self.stem_from_boards = bool(stem_from_boards) and hasattr(self.net, "chess_stem_w")
        self.unroll, self.use_graph, self.graph = int(unroll), use_graph, None
        self.advances = 0
        self.flops_per_eval = flops_per_eval(8, 8, N_ACTIONS, in_planes=PLANES)  # of the reference's net
        # FLOPs the GPU route really executes per evaluation: the structurally zero planes are not multiplied
        self.flops_per_eval_executed = flops_per_eval(8, 8, N_ACTIONS, in_planes=24 if self.stem_from_boards else (34 if self.tail_planes else PLANES))
        self.launches_per_advance = 2  # az_chess_step + az_chess_move (ours); the net's kernels are library calls
        self._games = {}  # game id -> (length, result) of games whose samples may still sit in a later drain

    def _advance(self):
        # the first step after a reset finds no pending leaf, so the stale priors are never consumed
        if self.stem_from_boards:
            # no plane tensor at all: the stem is computed from the 64-byte leaf boards (az_chess_stem)
            self.engine.step(self.priors, self.values, None, self.valid)
            self.net.forward_from_stem(self.net.chess_stem(self.engine.view("leaf_pos")), self.priors, self.values)
        else:
            self.engine.step(self.priors, self.values, self.states, self.valid, self.plane_first)
            self.net(self.states, self.priors, self.values)
        self.engine.move()

    def capture(self):
        if self.graph is not None or not self.use_graph:
            return
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm up cuDNN heuristics / workspaces outside capture
            for _ in range(3):
                self.net(self.states, self.priors, self.values)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(self.unroll):
                self._advance()
        self.graph = graph

    def run(self, advances):
        if self.use_graph:
            self.capture()
            n = (advances + self.unroll - 1) // self.unroll
            for _ in range(n):
                self.graph.replay()
            self.advances += n * self.unroll
            return n * self.unroll
        for _ in range(advances):
            self._advance()
        self.advances += advances
        return advances

    def active_trees(self):
        return int((self.engine.phases() != native.AZ_PHASE_IDLE).sum())

    def ring_fill(self):
        return int(self.engine.view("smp_count")[0]) / self.engine.cfg.sample_capacity

    def collect_all_ranks(self, trainer_rank=0):
        """collect() with the rings of every rank gathered first (torch.distributed).  Only the trainer rank decodes the
        samples; the other ranks take part in the gather and return None."""
        from . import dist as azdist

        d = azdist.all_gather_chess_rings(self.engine.drain(), self.device)
        if azdist.world()[0] != trainer_rank:
            return None
        for g, ln, r in zip(d["fin_game"], d["fin_len"], d["fin_result"]):
            self._games[int(g)] = (int(ln), int(r))
        states, policies = decode_samples(d, self.device)
        values, known = sample_values(d, self._games)
        return states.cpu().numpy(), policies.cpu().numpy(), values, known, d

    def collect(self):
        """Drains the rings: (states f32 [n, 8, 8, 118], policies f64 [n, 1880], values int32 [n], known bool [n]) as
        host arrays.  `known` is False for samples whose game has not finished yet (their value comes with a later
        collect of the same game: keep them and call sample_values again, or use finished games only)."""
        d = self.engine.drain()
        for g, ln, r in zip(d["fin_game"], d["fin_len"], d["fin_result"]):
            self._games[int(g)] = (int(ln), int(r))
        states, policies = decode_samples(d, self.device)
        values, known = sample_values(d, self._games)
        return states.cpu().numpy(), policies.cpu().numpy(), values, known, d

    def run_until_done(self, poll_every=64, max_advances=None):
        """Runs until every tree is idle; returns the samples of all finished games."""
        done, parts = 0, []
        while True:
            done += self.run(poll_every)
            if self.ring_fill() > 0.5 or self.active_trees() == 0:
                parts.append(self.collect())
            if self.active_trees() == 0 or (max_advances is not None and done >= max_advances):
                break
        self.engine.check_status()
        parts.append(self.collect())
        d_all = {k: np.concatenate([p[4][k] for p in parts]) for k in ("game", "ply")}
        states = np.concatenate([p[0] for p in parts])
        policies = np.concatenate([p[1] for p in parts])
        values = np.zeros(len(states), dtype=np.int32)
        known = np.zeros(len(states), dtype=bool)
        for i, (g, ply) in enumerate(zip(d_all["game"], d_all["ply"])):
            if int(g) in self._games:
                ln, r = self._games[int(g)]
                values[i] = r * (1 if (ln - 1 - int(ply)) % 2 == 0 else -1)
                known[i] = True
        return states, policies, values, known

    def totals(self):
        return self.engine.totals()

    def load_weights(self, net: PolicyValueNet):
        self.fp32_net = net
        self.net.load_from(net)


def chess_training_loop(runner: ChessSelfPlayRunner, trainer, window, iterations, train_steps_per_iteration=1,
                        exclude_null_games=True, rng=np.random, log=None, max_advances=None):
    """Self-play -> samples -> train -> new weights for chess on one rank: the loop of train.selfplay_training_loop
    (reference train.py:16-84, self_play.py:122-188) with the chess runner's sample ring in place of the finished-game
    records.  Drawn games contribute no samples when exclude_null_games (self_play.py:155-162)."""
    history = []
    base0, stride = int(runner.engine.cfg.game_id_base), int(runner.engine.cfg.games_target)
    for it in range(iterations):
        # the runner's games_target games, every iteration, under NEW game ids (they key the move-sampling counters:
        # reusing them would replay identical games until the weights change)
        runner.engine.reset(game_id_base=base0 + it * stride)
        runner.graph = None  # the captured launches hold the old id range
        runner.valid.zero_()
        runner._games.clear()
        states, policies, values, known = runner.run_until_done(max_advances=max_advances)
        keep = known & ((values != 0) if exclude_null_games else np.ones(len(values), dtype=bool))
        window.append(states[keep], policies[keep], values[keep])
        if window.ready():
            for _ in range(train_steps_per_iteration):
                from .train import BATCH_SIZE

                history.append(trainer.train_step(*window.sample(min(BATCH_SIZE, len(window)), rng=rng)))
            runner.load_weights(trainer.net)
        if log is not None:
            log(it, len(window), history[-1] if history else None)
    return history
