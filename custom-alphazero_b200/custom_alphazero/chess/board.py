"""Board - drop-in for the reference's chess/board.py:12-196, backed by the az_chess_* CUDA kernels.

The reference subclasses python-chess's Board.  python-chess is not a dependency here: the members the reference
and its callers touch are provided directly - the reference's own (`array`, `array_one_hot`, `moves`, `state`,
`full_state`, `state_history`, `legal_moves_mask`, `play`, `get_result`, `get_random_move`, `play_random`,
`update_array`, `display_ascii`, the FEN / array / one-hot converters) and the python-chess ones it leans on
(`turn`, `fullmove_number`, `halfmove_clock`, `ep_square`, `legal_moves`, `push_uci`, `mirror`, `is_game_over`,
`result`, `board_fen`, `has_kingside_castling_rights`, `has_queenside_castling_rights`, `is_repetition`).
Every rule of chess - legal moves, making a move, the mirror, how a game ends, the 118 planes - is computed by
libaz_b200 on the GPU; there is no host implementation and the methods raise without the library or a device.

Deviations: `is_repetition()` is always False (python-chess needs the move stack, which the keep_same_player
path drops at every mirror(); off that path repetition is not tracked) and fivefold repetition never ends a game.
"""
import copy
from collections import deque
from typing import List, Optional

import numpy as np

from az_b200 import chess as _chess
from custom_alphazero.chess.move import Move
from custom_alphazero.config import ConfigChess


class _UciMove:
    """The little of python-chess's Move the reference touches: `.uci()`."""

    def __init__(self, uci):
        self._uci = uci

    def uci(self):
        return self._uci

    def __repr__(self):
        return "Move.from_uci({!r})".format(self._uci)


class Board:
    def __init__(self, board_fen: Optional[str] = None, array: Optional[np.ndarray] = None, history_size: int = 8):
        self.board_size = ConfigChess.board_size
        self.number_unique_pieces = ConfigChess.number_unique_pieces
        if array is not None:
            assert isinstance(array, np.ndarray)
            assert all(dim == self.board_size for dim in array.shape)
            assert np.unique(array).size <= self.number_unique_pieces + 1
            board_fen = self.array_to_board_fen(array.astype("int8"))
        elif board_fen is None:
            board_fen = ConfigChess.initial_board_fen
        self._pos = _chess.position_from_fen(self.get_fen(board_fen))
        self.array = _chess.unpack_position(self._pos)["array"]
        self._history_size = history_size
        self.state_history = deque(maxlen=history_size)
        self._packed_history = deque(maxlen=history_size)  # the same entries as boards, for az_chess_encode
        for _ in range(history_size):
            self.state_history.append(np.zeros(self.state.shape))
            self._packed_history.append(None)
        self.state_history.append(self.state)
        self._packed_history.append(_chess.pack_position(self.array))

    # ------------------------------------------------------------------ python-chess surface
    def _fields(self):
        return _chess.unpack_position(self._pos)

    @property
    def turn(self) -> bool:
        return bool(self._fields()["turn"])

    @turn.setter
    def turn(self, white: bool):
        meta = int(self._pos[7])
        self._pos[7] = np.uint64(meta & ~_chess.META_TURN if white else meta | _chess.META_TURN)

    @property
    def fullmove_number(self) -> int:
        return self._fields()["fullmove_number"]

    @property
    def halfmove_clock(self) -> int:
        return self._fields()["halfmove_clock"]

    @property
    def ep_square(self) -> Optional[int]:
        return self._fields()["ep_square"]

    def _rights(self, color: bool, king_side: bool) -> bool:
        bit = (1 if king_side else 2) << (0 if color else 2)
        return bool(self._fields()["castling"] & bit)

    def has_kingside_castling_rights(self, color: bool) -> bool:
        return self._rights(bool(color), True)

    def has_queenside_castling_rights(self, color: bool) -> bool:
        return self._rights(bool(color), False)

    def is_repetition(self, count: int = 3) -> bool:
        return False

    def _legal(self):
        mask, count, status = _chess.chess_legal(self._pos[None])
        return np.nonzero(mask[0])[0], int(count[0]), int(status[0])

    @property
    def legal_moves(self):
        return [_UciMove(_chess.action_uci(int(a))) for a in self._legal()[0]]

    def push_uci(self, uci: str):
        a = _chess.uci_action(uci)
        out, status = _chess.chess_play(self._pos[None], [a], keep_same_player=False)
        if a < 0 or int(status[0]) < 0:
            raise ValueError("illegal uci: {!r} in {}".format(uci, self.board_fen()))
        self._pos = out[0].copy()

    def mirror(self) -> "Board":
        """python-chess semantics: a NEW board built through the class constructor (so with a fresh
        state_history holding the initial position), flipped vertically with colours, castling rights, en-passant
        square and turn swapped; the clocks are kept."""
        f = self._fields()
        c = f["castling"]
        ep = f["ep_square"]
        other = type(self)(None)
        other._pos = _chess.pack_position(-f["array"][::-1], turn=not f["turn"], castling=((c & 3) << 2) | (c >> 2),
                                          ep_square=None if ep is None else ep ^ 56,
                                          halfmove_clock=f["halfmove_clock"], fullmove_number=f["fullmove_number"])
        return other

    def is_game_over(self) -> bool:
        return (self._legal()[2] & 3) != 0

    def result(self) -> str:
        status = self._legal()[2] & 3
        if status == 0:
            return "*"
        if status == 2:
            return "1/2-1/2"
        return "0-1" if self.turn else "1-0"  # the side to move is mated

    def board_fen(self) -> str:
        return self.array_to_board_fen(self._fields()["array"])

    # ------------------------------------------------------------------ the reference's own members
    @property
    def array_one_hot(self) -> np.ndarray:
        return np.eye(self.number_unique_pieces + 1)[self.array]

    @property
    def moves(self) -> List[Move]:
        return [Move(uci=move.uci()) for move in self.legal_moves]

    @property
    def state(self) -> np.ndarray:
        return np.dstack([self.array_one_hot, np.full((self.board_size, self.board_size), self.is_repetition())])

    @property
    def full_state(self) -> np.ndarray:
        """118 planes (chess/board.py:58-73), computed by az_chess_encode from the deque's boards."""
        hist = np.zeros((1, 7, 8), dtype=np.uint64)
        older = list(self._packed_history)[:-1][-7:]
        for slot, packed in zip(range(7 - len(older), 7), older):
            if packed is not None:
                hist[0, slot] = packed
        cur = self._pos.copy()
        cur[:7] = self._packed_history[-1][:7]  # the planes show the deque's newest entry; the scalars are live
        return _chess.chess_encode(cur[None], hist)[0].astype(np.float64)

    @staticmethod
    def get_fen(board_fen: str) -> str:
        return " ".join([board_fen, ConfigChess.initial_turn, ConfigChess.initial_castling_rights,
                         ConfigChess.initial_ep_quare, ConfigChess.initial_halfmove_clock,
                         ConfigChess.initial_fullmove_number])

    @staticmethod
    def from_one_hot(array_oh: np.ndarray) -> np.ndarray:
        # kept as the reference computes it (chess/board.py:88-96): indices above 6 map to index - 12 + 1, which is
        # not the inverse of array_one_hot for black pieces (that would be index - 13); nothing on the hot path uses it
        array = np.argmax(array_oh, axis=-1)
        upper = array > ConfigChess.number_unique_pieces / 2
        array[upper] = array[upper] - ConfigChess.number_unique_pieces + 1
        return array

    @staticmethod
    def piece_symbol_to_int(piece_symbol: Optional[str]) -> int:
        if piece_symbol is None:
            return 0
        value = ConfigChess.piece_symbols.index(piece_symbol.lower())
        return value if piece_symbol.isupper() else -value

    @staticmethod
    def int_to_piece_symbol(piece_int: int) -> Optional[str]:
        symbol = ConfigChess.piece_symbols[abs(int(piece_int))]
        if symbol is None:
            return None
        return symbol if piece_int < 0 else symbol.upper()

    def legal_moves_mask(self, all_possible_moves: List[Move]) -> np.ndarray:
        legal = set(self.moves)
        return np.asarray([move in legal for move in all_possible_moves])

    def board_fen_to_array(self, fen: str) -> np.ndarray:
        cells = []
        for ch in fen.replace("/", ""):
            cells.extend([0] * int(ch) if ch.isdigit() else [self.piece_symbol_to_int(ch)])
        return np.asarray(cells).reshape((self.board_size, self.board_size)).astype("int8")

    def array_to_board_fen(self, array: np.ndarray) -> str:
        rows = []
        for row in np.asarray(array):
            text, empty = "", 0
            for value in row:
                symbol = self.int_to_piece_symbol(int(value))
                if symbol is None:
                    empty += 1
                else:
                    text += (str(empty) if empty else "") + symbol
                    empty = 0
            rows.append(text + (str(empty) if empty else ""))
        return "/".join(rows)

    def update_array(self):
        self.array = self.board_fen_to_array(self.board_fen())
        self.state_history.append(self.state)
        self._packed_history.append(_chess.pack_position(self.array))

    def get_random_move(self) -> Optional[Move]:
        try:
            return np.random.choice(self.moves)
        except ValueError:
            return None

    def play(self, move: Move, on_copy: bool = False, keep_same_player: bool = False) -> "Board":
        board = copy.deepcopy(self) if on_copy else self
        board.push_uci(move.uci)
        if keep_same_player:
            board = board.mirror()
            board.turn = True  # virtually, it is always white to play
        board.update_array()
        if not on_copy:
            self.__dict__.update(board.__dict__)
        return board

    def play_random(self) -> "Board":
        return self.play(self.get_random_move())

    def get_result(self, keep_same_player: bool = False):
        """None while the game goes on; 0 for a draw; without keep_same_player +1 / -1 for a white / black win
        (chess/board.py:178-190).  The reference's MCTS passes keep_same_player=True (mcts/mcts.py:179), which its
        chess Board does not accept; here that flag gives the Connect-N meaning (connect_n/board.py:258-268): +1,
        the player who just moved delivered mate."""
        if not self.is_game_over():
            return None
        result = self.result()
        if result == "1/2-1/2":
            return 0
        if keep_same_player:
            return 1
        return 1 if result == "1-0" else -1

    def display_ascii(self):
        for row in self.array:
            print("".join(self.int_to_piece_symbol(v) if v else "." for v in row))
