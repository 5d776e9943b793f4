"""MCTS / UCTNode / UCTEdge - drop-in for the reference's mcts/mcts.py:22-222, as a batch-of-one view
on the GPU tree engine (libaz_b200).

`MCTS(board, all_possible_moves, concurrency, plays_inferences, model=None)` keeps the reference's
constructor, attributes and the evaluator hooks: `model` is called on np.ndarray[1, H, W, 4] and must
return two objects with .numpy() (mcts.py:131-137); with model=None the module-level
`infer_sample(state, concurrency)` is used (mcts.py:138-141) - tests patch it exactly as they would in
the reference.  Select / expand / backup / play run in the CUDA kernels; the host only shuttles the
leaf state out and the evaluator's answer back in (one round trip per simulation: this class is the
correctness/plumbing surface, the throughput path is az_b200.selfplay.SelfPlayRunner).
UCTNode / UCTEdge are read-only snapshots of the device tree for callers such as the visualiser.
"""
from copy import deepcopy
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from az_b200 import native
from az_b200.engine import Rules, TreeEngine
from az_b200.selfplay import decode_samples
from custom_alphazero.config import ConfigGeneral, ConfigMCTS, ConfigSelfPlay
from custom_alphazero.mcts.utils import normalize_probabilities  # noqa: F401  (re-exported like the reference)
from custom_alphazero.serving.factory import infer_sample

if ConfigGeneral.game == "chess":  # mcts.py:12-14 of the reference
    # The module imports like the reference's; MCTS is rebound at the bottom to the chess search object
    # (custom_alphazero/mcts/chess_mcts.py), a batch-of-one view on az_b200.chess_engine.ChessTreeEngine.
    from custom_alphazero.chess.board import Board
    from custom_alphazero.chess.move import Move
elif ConfigGeneral.game == "connect_n":
    from custom_alphazero.connect_n.board import Board
    from custom_alphazero.connect_n.move import Move
else:
    raise NotImplementedError


class UCTEdge:
    def __init__(self, parent: "UCTNode", child: "UCTNode", action: Optional[Move], prior: float,
                 visit_count: int = 0, total_action_value: float = 0.0):
        self.parent, self.child, self.action, self.prior = parent, child, action, prior
        self.visit_count = visit_count
        self.total_action_value = total_action_value
        self.played = False
        self.greedily_played = False

    @property
    def siblings(self):
        return [e for e in self.parent.edges if e is not self]

    def exploitation_term(self) -> float:
        return self.total_action_value / self.visit_count if self.visit_count else 0.0

    def exploration_term(self, override_prior: Optional[float] = None) -> float:
        prior = self.prior if override_prior is None else override_prior
        total = sum(e.visit_count for e in self.parent.edges)
        return ConfigMCTS.exploration_constant * prior * (total**0.5) / (1 + self.visit_count)

    def upper_confidence_bound(self, override_prior: Optional[float] = None) -> float:
        return self.exploitation_term() + self.exploration_term(override_prior)


class _TreeSnapshot:
    """Read access to the device tree of one search object, fetched on demand: the root's child block alone (what
    play_game, the arena and the tests read after a search: k records) costs four small copies; anything deeper pulls
    the live part of the pool once.  Nothing is copied for a tree nobody looks at (the reference's callers usually read
    only `current_root.edges`).  A snapshot that was read is frozen in full before the tree changes; one that was never
    read dies with the tree state it described."""

    def __init__(self, engine):
        self.e = engine
        self.half = int(engine.view("half")[0])
        self.root = int(engine.view("root_node")[0])
        self.full = None       # {"W", "N", "link", "P"} numpy arrays of the used part of the live half
        self.root_block = None
        self.stale = False
        self.touched = False

    def _check(self):
        if self.stale:
            raise RuntimeError("this UCTNode describes a tree state that was searched or played past before anybody read it; "
                               "read mcts.current_root after the search / play you are interested in")

    def freeze(self):
        """Called by the search object right before it changes the tree."""
        if self.touched and self.full is None:
            self._pull()
        elif not self.touched:
            self.stale = True

    def _pull(self):
        self._check()
        e = self.e
        w, n, link = e.node_view()
        used = int(e.view("n_nodes")[0])
        h = self.half
        self.full = {"W": w[0, h, :used].cpu().numpy(), "N": n[0, h, :used].cpu().numpy(),
                     "link": link[0, h, :used].cpu().numpy(), "P": e.view("node_p")[0, h, :used].cpu().numpy()}

    def block(self, index):
        """Child block of node `index`: (k, N[k], W[k], P[k], first child index)."""
        self.touched = True
        if self.full is None and index == self.root:
            if self.root_block is None:
                self._check()
                w, n, link = self.e.node_view()
                lk = int(link[0, self.half, index]) & 0xFFFFFFFF
                base, k = lk & 0xFFFFFF, lk >> 24
                sl = slice(base, base + k)
                self.root_block = (k, n[0, self.half, sl].cpu().numpy(), w[0, self.half, sl].cpu().numpy(),
                                   self.e.view("node_p")[0, self.half, sl].cpu().numpy(), base)
            return self.root_block
        if self.full is None:
            self._pull()
        lk = int(self.full["link"][index]) & 0xFFFFFFFF
        base, k = lk & 0xFFFFFF, lk >> 24
        sl = slice(base, base + k)
        return k, self.full["N"][sl], self.full["W"][sl], self.full["P"][sl], base


class UCTNode:
    """Snapshot of one device node.  `edges` are materialised on first access from the tree export;
    `board` by replaying the path's moves on a copy of the root board (K2 kernel)."""

    def __init__(self, board: Optional[Board], edges: Optional[List[UCTEdge]] = None, _export=None, _index=None,
                 _parent_board=None, _move=None):
        self._board = board
        self._edges = edges
        self._export, self._index = _export, _index
        self._parent_board, self._move = _parent_board, _move
        self.evaluated_value = None  # the engine does not keep per-node evaluations

    @property
    def board(self) -> Board:
        if self._board is None:
            self._board = self._parent_board().play(self._move, on_copy=True, keep_same_player=True)
        return self._board

    @property
    def edges(self) -> List[UCTEdge]:
        if self._edges is None:
            self._edges = []
            ex = self._export
            if ex is not None:
                k, n, w, p, base = ex.block(self._index)
                if k:
                    moves = self.board.moves
                    for j in range(k):
                        child = UCTNode(None, None, ex, base + j, _parent_board=lambda s=self: s.board, _move=moves[j])
                        self._edges.append(UCTEdge(self, child, moves[j], float(p[j]), int(n[j]), float(w[j])))
        return self._edges

    def get_best_edge(self) -> UCTEdge:
        return self.edges[int(np.argmax([e.upper_confidence_bound() for e in self.edges]))]


class MCTS:
    def __init__(self, board: Board, all_possible_moves: List[Move], concurrency: bool, plays_inferences: dict,
                 model=None, use_solver: bool = False) -> None:
        if use_solver:
            raise NotImplementedError("the exact solver back-end is outside the B200 hot path (SURVEY 2 #12)")
        if ConfigGeneral.game != "connect_n":
            raise NotImplementedError("this search object drives the Connect-N engine; chess: custom_alphazero.mcts.chess_mcts")
        # ConfigMCTS.enable_dirichlet_noise (mcts.py:114-115): root noise is drawn on the device (same distribution as
        # np.random.dirichlet, different stream: statistical parity only - SURVEY 8a row a5)
        self.board = deepcopy(board)
        self.all_possible_moves = all_possible_moves
        self.concurrency = concurrency
        self.plays_inferences = plays_inferences if plays_inferences is not None else {}
        self.model = model
        self.use_solver = use_solver
        self.path_cache = []
        self._rules = Rules(board.board_width, board.board_height, board.n, bool(board.gravity))
        self._engine = TreeEngine(self._rules, n_trees=1, sims_per_move=ConfigSelfPlay.mcts_iterations,
                                  eval_mode="external", move_mode="host_uniforms", max_free_sims=64,
                                  index_move_greedy=ConfigMCTS.index_move_greedy,
                                  c_puct=ConfigMCTS.exploration_constant, games_target=1,
                                  dirichlet_noise=bool(ConfigMCTS.enable_dirichlet_noise),
                                  dirichlet_alpha=ConfigMCTS.dirichlet_noise_value,
                                  dirichlet_ratio=ConfigMCTS.dirichlet_noise_ratio,
                                  seed=int(np.random.randint(0, 2**31 - 1)) if ConfigMCTS.enable_dirichlet_noise else 0,
                                  pow_lut_len=self._rules.max_plies * 4096 + 2)
        dev = self._engine.device
        H, W, A = self._rules.height, self._rules.width, self._rules.n_actions
        self._states = torch.zeros((1, H, W, 4), dtype=torch.float32, device=dev)
        self._valid = torch.zeros(1, dtype=torch.int32, device=dev)
        self._staged = None  # (priors tensor [1, A], values tensor [1]) waiting for the next az_step
        self._snapshots = []
        self.root = self.initialize_root()
        self.current_root = self.root

    # ------------------------------------------------------------------ tree views
    def _root_view(self) -> UCTNode:
        snap = _TreeSnapshot(self._engine)
        self._snapshots.append(snap)
        return UCTNode(deepcopy(self.board), None, snap, snap.root)

    def _tree_changes(self):
        """Before the device tree is modified: snapshots somebody has read are completed, the others are dropped."""
        for snap in self._snapshots:
            snap.freeze()
        self._snapshots = []

    def initialize_root(self) -> UCTNode:
        """mcts.py:108-109: an edgeless root at the caller's position (az_set_roots)."""
        self._tree_changes()
        cells = (self.board.array * np.int8(self.board.turn)).astype(np.int8)[None]
        self._engine.set_roots([0], cells, [self.board.fullmove_number])
        self._staged = None
        return self._root_view()

    # ------------------------------------------------------------------ evaluator (mcts.py:122-143)
    def _priors_value_from_state(self, state: np.ndarray) -> Tuple[np.ndarray, float]:
        key = "\n".join("".join("X" if c[1] else ("O" if c[2] else ".") for c in row) for row in state)
        if key in self.plays_inferences:
            return self.plays_inferences[key]
        if self.model is not None:
            probabilities, value = self.model(np.expand_dims(state, axis=0))
            probabilities, value = probabilities.numpy().ravel(), value.numpy().item()
        else:
            probabilities, value = infer_sample(state, concurrency=self.concurrency)
        self.plays_inferences[key] = probabilities, value
        return probabilities, value

    def _stage(self, probabilities, value):
        p = np.asarray(probabilities)
        dt = torch.float32 if p.dtype == np.float32 else torch.float64
        dev = self._engine.device
        self._staged = (torch.as_tensor(p.astype(np.float32 if dt == torch.float32 else np.float64)[None], device=dev).contiguous(),
                        torch.tensor([float(value)], dtype=dt, device=dev))

    # ------------------------------------------------------------------ the reference's step methods
    def select(self) -> Optional[np.ndarray]:
        """Runs az_step: applies the staged evaluation (expand + backup of the previous leaf), selects
        the next leaf and returns its NN input, or None when no evaluation is needed right now."""
        if self._snapshots:
            self._tree_changes()
        pr, va = self._staged if self._staged is not None else (None, None)
        self._engine.step(pr, va, self._states, self._valid)
        self._staged = None
        if int(self._valid[0]):
            return self._states[0].cpu().numpy()
        return None

    def evaluate_and_expand(self, state: np.ndarray) -> float:
        """Evaluates the selected leaf and stages the answer; the expansion itself happens on the
        device at the start of the next az_step (mcts.py:145-161)."""
        probabilities, value = self._priors_value_from_state(state)
        self._stage(probabilities, value)
        return value

    def backup(self, value: float):
        """Backup is fused into az_step (mcts.py:163-168 runs on the device); nothing to do on the host."""
        self.path_cache = []

    def search(self, iterations_number: int):
        self._tree_changes()
        self._engine.begin_search(int(iterations_number))
        while True:
            state = self.select()
            if state is not None:
                self.backup(-self.evaluate_and_expand(state))
                continue
            if int(self._engine.phases()[0]) != native.AZ_PHASE_SEARCH:
                break
        self._engine.check_status()
        self.current_root = self._root_view()

    # ------------------------------------------------------------------ mcts.py:182-222
    def play(self, greedy: bool = False, return_details: bool = False, deterministic: bool = False
             ) -> Union[Tuple[np.ndarray, np.ndarray, np.ndarray, Move], Board]:
        e = self._engine
        self._tree_changes()
        rec = int(e.view("rec_len")[0])
        if not deterministic:
            # np.random.choice(edges, 1, p=pi) consumes exactly one random_sample() of the global stream
            e.view("uniforms")[0, rec] = float(np.random.random_sample())
        e.play(greedy=greedy, move_mode="argmax" if deterministic else "host_uniforms")
        e.check_status()
        assert int(e.view("rec_len")[0]) == rec + 1, "play() needs a searched root"
        one = slice(0, 1)
        fin = {"game_id": e.view("game_id")[one], "len": e.view("rec_len")[one], "result": e.view("result")[one],
               "visits": e.view("rec_visits")[one], "action": e.view("rec_action")[one], "board": e.view("rec_board")[one]}
        states, policies, _ = decode_samples(self._rules, fin)
        action = int(e.view("rec_action")[0, rec]) & 0xFFFF
        move = self.all_possible_moves[action]
        parent_state = states[rec]
        self.board.play(move, keep_same_player=True)
        # the device root must be the position the host board reached (mcts.py:208)
        words = e.view("root_board")[0].cpu().numpy().view(np.uint64)
        assert self._cells_from_bits(words) == self.board.array.tolist()
        if int(e.phases()[0]) == native.AZ_PHASE_SEARCH:
            self.current_root = self._root_view()
        else:  # the game ended: a terminal node has no edges
            self.current_root = UCTNode(deepcopy(self.board), [])
        if return_details:
            return parent_state, self.board.full_state, policies[rec], move
        return self.board

    def _cells_from_bits(self, words):
        H, W = self._rules.height, self._rules.width
        cur = sum(int(w) << (64 * i) for i, w in enumerate(words[0]))
        opp = sum(int(w) << (64 * i) for i, w in enumerate(words[1]))
        return [[1 if (cur >> (y * (W + 1) + x)) & 1 else (-1 if (opp >> (y * (W + 1) + x)) & 1 else 0) for x in range(W)]
                for y in range(H)]


if ConfigGeneral.game == "chess":  # same name, same constructor, same methods: the chess search object
    ConnectNMCTS = MCTS
    from custom_alphazero.mcts.chess_mcts import ChessMCTS as MCTS  # noqa: E402,F811
