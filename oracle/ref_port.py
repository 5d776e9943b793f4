"""Python/numpy restatement of the reference's Connect-N environment, PUCT search and
per-game self-play loop.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
/root/reference/custom_alphazero/).  The data structures mirror the reference on purpose -
an int8 [H, W] array per board, one Python object per node and per edge, child boards
created eagerly by deep copy - so that timing this module on the host CPU is a fair stand-in
for timing the reference itself (bench.py "cpu_baseline", kind "port").  The arithmetic that
decides parity is kept literally the same: Python-float PUCT in the reference's evaluation
order, `** 0.5` (libm pow, NOT sqrt), numpy first-max argmax, numpy sum/divide for the prior
normalisation, numpy cumsum + searchsorted for move sampling.
"""
import copy
from dataclasses import dataclass

import numpy as np

EXPLORATION_CONSTANT = 1.5  # config.py:51
INDEX_MOVE_GREEDY = 8  # config.py:55
DIRECTIONS = ((0, 1), (1, 1), (1, 0), (1, -1))  # config.py:47 (dx, dy)
WHITE, EMPTY, BLACK = 1, 0, -1  # config.py:43-45
SYMBOLS = {BLACK: "O", EMPTY: ".", WHITE: "X"}  # config.py:46


@dataclass(frozen=True)
class Rules:
    """ConfigConnectN (config.py:38-47) as a value instead of a mutable global."""

    width: int = 7
    height: int = 6
    n: int = 4
    gravity: bool = True

    def __post_init__(self):
        assert 2 <= self.n <= min(self.width, self.height)  # connect_n/board.py:14-18

    @property
    def n_actions(self):
        return self.width if self.gravity else self.width * self.height

    def all_actions(self):
        """connect_n/board.py:130-146 - gravity: x ascending; else x-major (x, y) product."""
        if self.gravity:
            return [(x, None) for x in range(self.width)]
        return [(x, y) for x in range(self.width) for y in range(self.height)]

    def action_index(self, move):
        x, y = move
        return x if self.gravity else x * self.height + y


def move_str(move):
    """connect_n/move.py:23-27."""
    x, y = move
    return "{}".format(x) if y is None else "({0}, {1})".format(x, y)


class RefBoard:
    """connect_n/board.py:12-42 state: cells int8 [H, W] (row 0 = top), side to move,
    ply counter, game-over flag, draw flag (None / True / False) and move history."""

    def __init__(self, rules, cells=None):
        self.rules = rules
        if cells is None:
            self.cells = np.zeros((rules.height, rules.width)).astype("int8")
        else:
            cells = np.asarray(cells)
            assert cells.shape == (rules.height, rules.width)
            self.cells = cells.astype("int8")
        self.to_move = WHITE  # board.py:39
        self.plies = 0  # "fullmove_number", counts plies (board.py:40, 243)
        self.over = False
        self.drawn = None  # "is_null"
        self.history = []

    # -- board.py:50-53
    def text(self):
        return "\n".join("".join(SYMBOLS[int(v)] for v in row) for row in self.cells)

    __repr__ = text

    def same_position(self, other):  # board.py:47-48
        return np.array_equal(self.cells, other.cells)

    # -- board.py:113-124: gravity -> columns whose top cell is empty, ascending x;
    #    else empty cells in row-major (y, x) order
    def legal_moves(self):
        r = self.rules
        if r.gravity:
            return [(int(x), None) for x in np.where(self.cells[0, :] == EMPTY)[0]]
        ys, xs = np.where(self.cells == EMPTY)
        return [(int(x), int(y)) for y, x in zip(ys, xs)]

    # -- board.py:154-155 (mask over the action list, built by membership tests)
    def legal_mask(self):
        legal = set(self.legal_moves())
        return np.asarray([a in legal for a in self.rules.all_actions()])

    # -- board.py:169-176
    def swapped_cells(self):
        c = self.cells
        return np.where(c == BLACK, WHITE, np.where(c == WHITE, BLACK, c))

    # -- board.py:83-98: dstack(eye(3)[cells], ones * to_move) float32.
    #    eye(3)[-1] is row 2, so channel 2 is the opponent (-1) plane.
    def full_state(self):
        r = self.rules
        planes = np.eye(3)[self.cells]
        turn = np.ones((r.height, r.width)) * self.to_move
        return np.dstack([planes, turn]).astype("float32")

    # -- board.py:178-208: count from the last stone along the four directions
    def _refresh_over(self, x0, y0):
        if self.over:
            return
        r = self.rules
        colour = self.cells[y0, x0]
        for dx, dy in DIRECTIONS:
            run = 1
            for sx, sy in ((dx, dy), (-dx, -dy)):
                x, y = x0, y0
                while 0 <= x + sx < r.width and 0 <= y + sy < r.height:
                    if self.cells[y + sy, x + sx] != colour:
                        break
                    run += 1
                    x, y = x + sx, y + sy
                    if run >= r.n:
                        self.over, self.drawn = True, False
                        return
        if not self.legal_moves():
            self.over, self.drawn = True, True

    # -- board.py:210-231
    def _place(self, move):
        x, y = move
        if self.rules.gravity:
            # board.py:212-226: the row just above the first non-empty cell from the top
            filled = np.where(self.cells[:, x] != EMPTY)[0]
            y = int(filled.min()) - 1 if len(filled) else self.rules.height - 1
            assert y >= 0, "column is full"
        else:
            assert self.cells[y, x] == EMPTY
        self.cells[y, x] = self.to_move
        self._refresh_over(x, y)
        self.to_move = BLACK if self.to_move == WHITE else WHITE

    # -- board.py:233-250.  Q7: a finished board (or move None) is returned unchanged
    #    and un-copied even when on_copy is requested.
    def play(self, move, on_copy=False, keep_same_player=False):
        if move is None or self.over:
            return self
        b = copy.deepcopy(self) if on_copy else self
        b._place(move)
        b.plies += 1
        if keep_same_player:
            b.cells = b.swapped_cells()
            b.to_move = WHITE
        b.history.append(move)
        return b

    # -- board.py:258-268
    def result(self, keep_same_player=False):
        if self.drawn is None or not self.over:
            return None
        if self.drawn:
            return 0
        if keep_same_player:
            return WHITE
        return WHITE if self.plies % 2 else BLACK


# ----------------------------------------------------------------------------------------
# search
# ----------------------------------------------------------------------------------------
class RefEdge:
    """mcts/mcts.py:22-33."""

    __slots__ = ("parent", "child", "action", "prior", "n", "w")

    def __init__(self, parent, child, action, prior):
        self.parent, self.child, self.action, self.prior = parent, child, action, prior
        self.n = 0
        self.w = 0.0

    def q(self):  # mcts.py:39-43
        try:
            return self.w / self.n
        except ZeroDivisionError:
            return 0.0

    def u(self):  # mcts.py:45-52: ((c * prior) * total ** 0.5) / (1 + n), total over parent.edges
        total = sum(e.n for e in self.parent.edges)
        return EXPLORATION_CONSTANT * self.prior * (total**0.5) / (1 + self.n)

    def puct(self):  # mcts.py:54-55
        return self.q() + self.u()


class RefNode:
    """mcts/mcts.py:58-62."""

    __slots__ = ("board", "edges", "value")

    def __init__(self, board):
        self.board = board
        self.edges = []
        self.value = None

    def best_edge(self):  # mcts.py:64-68: numpy argmax = first maximum
        return self.edges[int(np.argmax([e.puct() for e in self.edges]))]


def normalise(p):
    """mcts/utils.py:4-16."""
    assert len(p) > 0
    s = p.sum()
    if s == 0:
        return np.array([1 / len(p)] * len(p))
    return np.divide(p, s, out=np.zeros_like(p), where=s != 0)


class RefSearch:
    """mcts/mcts.py:88-222 with the evaluator passed in as a plain function
    state[H, W, 4] -> (priors[A], value).  The evaluation cache (mcts.py:123-124, 142) is kept
    because it is part of the reference's cost profile; it never changes results for a
    deterministic evaluator."""

    def __init__(self, board, evaluator, cache=None):
        self.rules = board.rules
        self.board = copy.deepcopy(board)  # mcts.py:98
        self.actions = self.rules.all_actions()
        self.evaluator = evaluator
        self.cache = {} if cache is None else cache
        self.root = RefNode(copy.deepcopy(self.board))  # mcts.py:108-109
        self.current = self.root
        self.path = []
        self.sims = 0
        self.evals = 0

    def select(self):  # mcts.py:111-120 (Dirichlet branch omitted: off by default, config.py:52)
        node = self.current
        while node.edges:
            e = node.best_edge()
            self.path.append(e)
            node = e.child
        return node

    def _evaluate(self, board):  # mcts.py:122-143
        key = board.text()
        hit = self.cache.get(key)
        if hit is None:
            hit = self.evaluator(board.full_state())
            self.cache[key] = hit
            self.evals += 1
        return hit

    def expand(self, node):  # mcts.py:145-161
        priors, value = self._evaluate(node.board)
        node.value = value
        priors = normalise(np.asarray(priors)[node.board.legal_mask()])
        # Q1: priors are in action-list order, moves in board order; zip pairs them by rank
        for prior, move in zip(priors, node.board.legal_moves()):
            child = node.board.play(move, on_copy=True, keep_same_player=True)
            node.edges.append(RefEdge(node, RefNode(child), move, prior))
        return value

    def backup(self, v):  # mcts.py:163-168
        for e in reversed(self.path):
            e.n += 1
            e.w += v
            v = -v
        self.path = []

    def search(self, n):  # mcts.py:170-180
        for _ in range(n):
            leaf = self.select()
            if not leaf.board.over:
                v = -self.expand(leaf)
            else:
                v = leaf.board.result(keep_same_player=True)
            self.backup(v)
            self.sims += 1

    def root_counts(self):
        return [e.n for e in self.current.edges]

    def play(self, greedy=False, deterministic=False, uniform=None):
        """mcts.py:182-222.  `uniform` is the single RandomState.random_sample() draw that
        np.random.choice(edges, 1, p=pi) consumes (mcts.py:201); choice == searchsorted of
        the normalised cumulative sum, side='right'."""
        node = self.current
        counts = [e.n for e in node.edges]
        if greedy:  # mcts.py:189-192
            pi = np.zeros(len(counts)).astype(float)
            pi[np.argmax(counts)] = 1.0
        else:  # mcts.py:193-197
            pi = normalise(np.asarray(counts).astype(float))
        if deterministic:
            k = int(np.argmax(pi))
        else:
            cdf = np.cumsum(pi)
            cdf /= cdf[-1]
            k = int(np.searchsorted(cdf, uniform, side="right"))
        edge = node.edges[k]
        parent_state = self.board.full_state()
        self.board.play(edge.action, keep_same_player=True)
        self.current = edge.child  # mcts.py:207: re-root, subtree kept
        assert self.board.same_position(self.current.board)  # mcts.py:208
        policy = np.zeros(len(self.actions))  # mcts.py:210-214
        policy[[self.rules.action_index(e.action) for e in node.edges]] = pi
        return parent_state, policy, edge.action, counts


def play_game(rules, evaluator, sims, uniforms=None, cache=None, max_plies=None):
    """self_play.py:37-82 for one game.  uniforms=None -> deterministic play (what the
    golden traces use); else one draw per ply.  Returns a dict with the training arrays
    (states f32 [T, H, W, 4], policies f64 [T, A], rewards int [T]) and the per-ply trace."""
    s = RefSearch(RefBoard(rules), evaluator, cache)
    states, policies, trace = [], [], []
    while not s.board.over and (max_plies is None or len(trace) < max_plies):
        s.search(sims)
        greedy = s.board.plies >= INDEX_MOVE_GREEDY  # self_play.py:62
        u = None if uniforms is None else uniforms[len(trace)]
        st, pol, move, counts = s.play(greedy, deterministic=uniforms is None, uniform=u)
        states.append(st)
        policies.append(pol)
        trace.append({"move": rules.action_index(move), "move_str": move_str(move), "N": counts})
    result = s.board.result(keep_same_player=True)
    rewards = None
    if result is not None:  # self_play.py:69-78 (discounting_factor == 1, config.py:20)
        rewards = np.repeat(result, len(states))
        rewards[-2::-2] = -rewards[-2::-2]
    return {
        "states": np.asarray(states), "policies": np.asarray(policies), "rewards": rewards,
        "trace": trace, "result": result, "sims": s.sims, "evals": s.evals, "search": s,
    }
