"""Board - drop-in for the reference's connect_n/board.py:12-271, backed by the K2/K3 CUDA kernels.

The object keeps what the reference keeps on the host because callers read it (`array` int8 [H, W]
with row 0 on top, `turn`, `fullmove_number`, `game_over`, `is_null`, `played_moves`), but every rule
of the game - where a stone lands, who won, whether the board is full, which moves are legal, the NN
planes - is computed by libaz_b200 (az_env_play / az_env_legal / az_env_encode).  There is no host
implementation of those rules: without the library or without a CUDA device the methods raise.
"""
import hashlib
from copy import deepcopy
from typing import List, Optional

import numpy as np

from az_b200 import env as _env
from az_b200.engine import Rules
from custom_alphazero.config import ConfigConnectN
from custom_alphazero.connect_n.move import Move


def _rules_from_config():
    return Rules(ConfigConnectN.board_width, ConfigConnectN.board_height, ConfigConnectN.n, bool(ConfigConnectN.gravity))


class Board:
    def __init__(self, array: Optional[np.ndarray] = None):
        cfg = ConfigConnectN  # read at construction, so the class attributes can be patched per board
        assert 2 <= cfg.n <= min(cfg.board_width, cfg.board_height)
        for name in ("board_width", "board_height", "n", "gravity", "black", "empty", "white", "pieces"):
            setattr(self, name, getattr(cfg, name))
        self.pieces_to_int = {symbol: value for value, symbol in self.pieces.items()}
        shape = (self.board_height, self.board_width)
        if array is not None:
            assert isinstance(array, np.ndarray) and array.shape == shape
            assert np.unique(array).size <= len(self.pieces)
        self.array = np.zeros(shape, dtype="int8") if array is None else array.astype("int8")
        self.turn, self.fullmove_number = cfg.white, 0
        self.game_over, self.is_null = False, None
        self.played_moves = []

    # ------------------------------------------------------------------ plumbing
    @property
    def _rules(self):
        return Rules(self.board_width, self.board_height, self.n, bool(self.gravity))

    def _relative(self):
        """Cells as the kernels want them: +1 = side to move."""
        return (self.array * np.int8(self.turn)).astype(np.int8)[None]

    def _action_index(self, move: Move) -> int:
        return move.x if self.gravity else move.x * self.board_height + move.y

    def _move_of_action(self, a: int) -> Move:
        return Move(self.gravity, a) if self.gravity else Move(self.gravity, a // self.board_height, a % self.board_height)

    # ------------------------------------------------------------------ identity / text
    def __repr__(self):
        return "\n".join("".join(self.pieces[int(v)] for v in row) for row in self.array)

    def __eq__(self, other):
        return np.array_equal(self.array, other.array)

    def __hash__(self):
        return int(hashlib.md5(repr(self).encode("utf-8")).hexdigest(), 16)

    def repr_graphviz(self) -> str:
        wide = {".": " . "}
        return "\n".join("".join(wide.get(self.pieces[int(v)], self.pieces[int(v)]) for v in row) for row in self.array)

    def repr_list_played_moves(self) -> str:
        if not self.gravity:
            raise NotImplementedError
        return "".join(str(int(str(m)) + 1) for m in self.played_moves)  # 1-indexed columns for the solver

    def display_ascii(self):
        print(repr(self))

    # ------------------------------------------------------------------ derived views
    @property
    def turn_mirror(self) -> int:
        return ConfigConnectN.black if self.turn == ConfigConnectN.white else ConfigConnectN.white

    def mirror(self) -> np.ndarray:
        return (-self.array).astype(self.array.dtype)

    @staticmethod
    def _one_hot(cells) -> np.ndarray:
        # planes in the reference's order: empty (0), white (+1), black (-1 indexes the last row of eye(3))
        return (cells[..., None] == np.asarray([0, 1, -1])).astype(np.float64)

    @property
    def array_one_hot(self) -> np.ndarray:
        return self._one_hot(self.array)

    @property
    def array_one_hot_mirror(self) -> np.ndarray:
        return self._one_hot(self.mirror())

    def _planes(self, cells, turn):
        # K3 kernel: planes (empty, +1 stones, -1 stones, ones); the 4th plane carries the turn sign
        st = _env.env_encode(self._rules, cells[None].astype(np.int8))[0]
        st[:, :, 3] *= np.float32(turn)
        return st

    @property
    def full_state(self) -> np.ndarray:
        return self._planes(self.array, self.turn)

    @property
    def full_state_mirror(self) -> np.ndarray:
        return self._planes(self.mirror(), self.turn_mirror)

    @property
    def odd_moves_number(self) -> bool:
        return bool(self.fullmove_number % 2)

    # ------------------------------------------------------------------ moves
    @property
    def moves(self) -> List[Move]:
        legal = _env.env_legal(self._rules, self._relative())[0]
        return [self._move_of_action(a) for a in _env.board_order_actions(self._rules, legal)]

    def last_move(self) -> Optional[Move]:
        return self.played_moves[-1] if self.played_moves else None

    @staticmethod
    def get_all_possible_moves() -> List[Move]:
        g, w, h = ConfigConnectN.gravity, ConfigConnectN.board_width, ConfigConnectN.board_height
        if g:
            return [Move(g, x) for x in range(w)]
        return [Move(g, x, y) for x in range(w) for y in range(h)]

    @staticmethod
    def from_one_hot(array_oh: np.ndarray) -> np.ndarray:
        plane = np.argmax(array_oh, axis=-1)
        return np.where(plane == 2, -1, plane)  # third plane = black stones

    def legal_moves_mask(self, all_possible_moves: List[Move]) -> np.ndarray:
        legal = _env.env_legal(self._rules, self._relative())[0]
        return np.asarray([bool(legal[self._action_index(m)]) for m in all_possible_moves])

    def update_array(self):
        pass

    def get_random_move(self) -> Optional[Move]:
        try:
            return np.random.choice(self.moves)
        except ValueError:
            return None

    def is_game_over(self) -> bool:
        return self.game_over

    # ------------------------------------------------------------------ playing
    def push(self, move: Move):
        """Places a stone of `self.turn` (reference board.py:210-231) through az_env_play."""
        out, status = _env.env_play(self._rules, self._relative(), np.asarray([self._action_index(move)], dtype=np.int32))
        assert status[0] >= 0, "illegal move"  # the reference asserts too (board.py:221-224, 228)
        # kernel output is mirrored (+1 = new side to move); back to absolute colours
        self.array = (out[0] * np.int8(-self.turn)).astype("int8")
        if not self.game_over and status[0] != 0:
            self.game_over, self.is_null = True, bool(status[0] == 2)
        self.turn = self.turn_mirror

    def update_game_over(self, last_move_x: int, last_move_y: int):
        """Kept for API compatibility: push() already asked the kernel (board.py:178-208)."""

    def play(self, move: Optional[Move], on_copy: bool = False, keep_same_player: bool = False) -> "Board":
        if move is None or self.game_over:  # Q7: unchanged and uncopied
            return self
        board = deepcopy(self) if on_copy else self
        board.push(move)
        board.fullmove_number += 1
        if keep_same_player:
            board.array = board.mirror()
            board.turn = ConfigConnectN.white
        board.played_moves.append(move)
        return board

    def play_random(self, on_copy: bool = False, keep_same_player: bool = False) -> "Board":
        return self.play(self.get_random_move(), on_copy=on_copy, keep_same_player=keep_same_player)

    def get_result(self, keep_same_player: bool = False):
        """None while the game runs, 0 for a draw; otherwise the winner: always +1 under keep_same_player (the
        player who just moved), else by ply parity (white moves on even plies)."""
        if not self.game_over or self.is_null is None:
            return None
        if self.is_null:
            return 0
        white_won = keep_same_player or self.odd_moves_number
        return ConfigConnectN.white if white_won else ConfigConnectN.black
