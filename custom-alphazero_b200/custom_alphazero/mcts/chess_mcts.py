"""ChessMCTS - the search object of custom_alphazero.mcts.mcts for ConfigGeneral.game == "chess": the reference's MCTS
constructor, attributes and methods (mcts/mcts.py:88-222) as a batch-of-one view on the chess tree engine
(az_chess_step / az_chess_move).

The reference's own MCTS cannot finish a chess simulation (mcts.py:179 passes keep_same_player to a chess
Board.get_result that does not take it), so there is no reference behaviour to reproduce beyond the interface; the
semantics are those of az_b200.chess_engine (include/az_b200.h): children in action order with priors paired by action,
terminal leaves +1 for the mover on checkmate and 0 for every draw.  Evaluator hooks as in the reference: `model` is
called on np.ndarray [1, 8, 8, 118] and returns two objects with .numpy() (mcts.py:131-137); with model=None the
module-level infer_sample(state, concurrency) of custom_alphazero.mcts.mcts is used (mcts.py:138-141).
"""
from copy import deepcopy
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from az_b200 import chess as _chess
from az_b200 import native
from az_b200.chess_engine import ChessTreeEngine, decode_samples
from custom_alphazero.chess.board import Board
from custom_alphazero.chess.move import Move
from custom_alphazero.config import ConfigMCTS, ConfigSelfPlay


class ChessEdge:
    """Read-only snapshot of one root edge (the members visualisers read from UCTEdge, mcts.py:22-55)."""

    def __init__(self, action: Move, prior: float, visit_count: int, total_action_value: float, siblings_visits: int):
        self.action, self.prior, self.visit_count, self.total_action_value = action, prior, visit_count, total_action_value
        self._siblings_visits = siblings_visits
        self.played = self.greedily_played = False

    def exploitation_term(self) -> float:
        return self.total_action_value / self.visit_count if self.visit_count else 0.0

    def exploration_term(self, override_prior: Optional[float] = None) -> float:
        prior = self.prior if override_prior is None else override_prior
        return ConfigMCTS.exploration_constant * prior * self._siblings_visits ** 0.5 / (1 + self.visit_count)

    def upper_confidence_bound(self, override_prior: Optional[float] = None) -> float:
        return self.exploitation_term() + self.exploration_term(override_prior)


class ChessNode:
    def __init__(self, board: Board, edges: List[ChessEdge]):
        self.board, self.edges, self.evaluated_value = board, edges, None

    def get_best_edge(self) -> ChessEdge:
        return self.edges[int(np.argmax([e.upper_confidence_bound() for e in self.edges]))]


class ChessMCTS:
    def __init__(self, board: Board, all_possible_moves: List[Move], concurrency: bool, plays_inferences: dict,
                 model=None, use_solver: bool = False) -> None:
        if use_solver:
            raise NotImplementedError("the exact solver is a Connect-4 program (exact_solvers/c4solver)")
        self.board = deepcopy(board)
        self.all_possible_moves = all_possible_moves
        self.concurrency = concurrency
        self.plays_inferences = plays_inferences if plays_inferences is not None else {}
        self.model = model
        self.use_solver = use_solver
        self.path_cache = []
        self._engine = ChessTreeEngine(n_trees=1, sims_per_move=ConfigSelfPlay.mcts_iterations, eval_mode="external",
                                       move_mode="host_uniforms", max_free_sims=64, max_plies=2048, sample_capacity=4,
                                       fin_capacity=2, index_move_greedy=ConfigMCTS.index_move_greedy,
                                       c_puct=ConfigMCTS.exploration_constant, games_target=1, pow_lut_len=1 << 20)
        dev = self._engine.device
        self._states = torch.zeros((1, 8, 8, _chess.PLANES), dtype=torch.bfloat16, device=dev)
        self._valid = torch.zeros(1, dtype=torch.int32, device=dev)
        self._staged = None
        self.root = self.initialize_root()
        self.current_root = self.root

    # ------------------------------------------------------------------ tree views
    def _root_view(self) -> ChessNode:
        acts, n, w, p = self._engine.root_stats(0)
        total = int(sum(n))
        return ChessNode(deepcopy(self.board), [ChessEdge(self.all_possible_moves[a], pj, nj, wj, total)
                                                for a, nj, wj, pj in zip(acts, n, w, p)])

    def initialize_root(self) -> ChessNode:
        """mcts.py:108-109: an edgeless root at the caller's position (az_chess_set_roots)."""
        self._engine.set_roots([0], self.board._pos[None])
        self._staged = None
        return ChessNode(deepcopy(self.board), [])

    # ------------------------------------------------------------------ evaluator (mcts.py:122-143)
    def _priors_value_from_state(self, state: np.ndarray, key: bytes) -> Tuple[np.ndarray, float]:
        if key in self.plays_inferences:
            return self.plays_inferences[key]
        if self.model is not None:
            probabilities, value = self.model(np.expand_dims(state, axis=0))
            probabilities, value = probabilities.numpy().ravel(), value.numpy().item()
        else:
            from custom_alphazero.mcts import mcts as _m  # the hook tests patch, as in the reference

            probabilities, value = _m.infer_sample(state, concurrency=self.concurrency)
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
        pr, va = self._staged if self._staged is not None else (None, None)
        self._engine.step(pr, va, self._states, self._valid)
        self._staged = None
        if int(self._valid[0]):
            return self._states[0].float().cpu().numpy()  # 0 / 1 planes and small integers: exact in bf16
        return None

    def evaluate_and_expand(self, state: np.ndarray) -> float:
        key = self._engine.view("leaf_pos")[0].cpu().numpy().tobytes()
        probabilities, value = self._priors_value_from_state(state, key)
        self._stage(probabilities, value)
        return value

    def backup(self, value: float):
        self.path_cache = []

    def search(self, iterations_number: int):
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
        assert int(e.phases()[0]) == native.AZ_PHASE_READY, "play() needs a searched root"
        ply = int(e.view("ply")[0])
        if not deterministic:
            e.view("uniforms")[0, ply] = float(np.random.random_sample())  # np.random.choice draws exactly one
        e.move(greedy=greedy, move_mode="argmax" if deterministic else "host_uniforms")
        e.check_status()
        d = e.drain()
        assert len(d["k"]) == 1
        states, policies = decode_samples(d, e.device)
        move = self.all_possible_moves[int(d["choice"][0]) & 0xFFFF]
        self.board.play(move, keep_same_player=True)
        if int(e.phases()[0]) == native.AZ_PHASE_SEARCH:
            # the device root must be the position the host board reached (mcts.py:208)
            dev_root = e.view("root_pos")[0].cpu().numpy().view(np.uint64)
            assert (dev_root[:7] == self.board._pos[:7]).all()
            self.current_root = self._root_view()
        else:  # the game ended: a terminal node has no edges
            self.current_root = ChessNode(deepcopy(self.board), [])
        if return_details:
            return states[0].cpu().numpy(), self.board.full_state, policies[0].cpu().numpy(), move
        return self.board
