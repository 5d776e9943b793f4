"""Move of the Connect-N environment: drop-in for the reference's connect_n/move.py:6-39
(column x with gravity, cell (x, y) without; ordering, equality and hash on the (x, y) pair;
str() is "x" or "(x, y)")."""


class Move:
    __slots__ = ("gravity", "x", "y")

    def __init__(self, gravity, x, y=None):
        if gravity:
            assert y is None, "a gravity move is a column"
        else:
            assert y is not None, "a free-placement move needs a row"
        self.gravity, self.x, self.y = gravity, x, y

    def _key(self):
        return (self.x, self.y)

    def __eq__(self, other):
        return self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self == other or self < other

    def __gt__(self, other):
        return not self <= other

    def __ge__(self, other):
        return not self < other

    def __hash__(self):
        return hash(self._key())

    def __str__(self):
        return "%d" % self.x if self.gravity else "(%d, %d)" % (self.x, self.y)

    __repr__ = __str__
