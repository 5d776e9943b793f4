"""Move - drop-in for the reference's chess/move.py:7-69: a move is ((file, rank), (file, rank, promotion letter)),
ordered, hashed and compared on that pair; `uci` is the usual string."""
from functools import total_ordering
from typing import Optional, Tuple

BOARD_SIZE = 8


@total_ordering
class Move:
    def __init__(self, pos_from: Optional[Tuple[int, int]] = None, pos_to: Optional[Tuple[int, int, str]] = None,
                 uci: Optional[str] = None):
        if uci is not None:
            pos_from, pos_to = self.uci_to_coords(uci)
        assert pos_from is not None and pos_to is not None
        self.pos_from, self.pos_to = tuple(pos_from), tuple(pos_to)

    def _key(self):
        return self.pos_from, self.pos_to

    def __str__(self):
        return "({0}, {1}) -> ({2}, {3}, {4})".format(*self.pos_from, *self.pos_to)

    __repr__ = __str__

    def __eq__(self, other):
        return self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __hash__(self):
        return hash(self._key())

    @property
    def uci(self) -> str:
        (ff, fr), (tf, tr, promo) = self.pos_from, self.pos_to
        return "abcdefgh"[ff] + str(fr + 1) + "abcdefgh"[tf] + str(tr + 1) + promo

    @staticmethod
    def uci_to_coords(uci: str):
        assert 4 <= len(uci) <= 5
        return ("abcdefgh".index(uci[0]), int(uci[1]) - 1), ("abcdefgh".index(uci[2]), int(uci[3]) - 1, uci[4:])

    @staticmethod
    def mirror(move: "Move") -> "Move":
        """Point reflection of both squares, as chess/move.py:57-69 does it."""
        last = BOARD_SIZE - 1
        (ff, fr), (tf, tr, promo) = move.pos_from, move.pos_to
        return Move(pos_from=(last - ff, last - fr), pos_to=(last - tf, last - tr, promo))
