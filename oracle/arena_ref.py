"""Restatement of the reference's arena (evaluation/evaluate.py:29-134), policy-only mode, on oracle/ref_port
boards.  TEST INFRASTRUCTURE ONLY.  Parity: pinned only through ref_port (whose Board is pinned to the golden
vectors); the arena itself has no golden vectors (its evaluator needs TensorFlow in the reference)."""
import numpy as np

from . import ref_port


def single_game(rules, model_current, model_previous, game_index, deterministic, rng):
    """evaluate.py:29-63 with evaluate_with_mcts=False.  model(state[1,H,W,4]) -> (probabilities[1,A], value)."""
    model = model_current if game_index % 2 == 0 else model_previous
    board = ref_port.RefBoard(rules)
    while not board.over:
        probabilities = np.asarray(model(board.full_state()[None])[0]).ravel()
        legal = ref_port.normalise(probabilities[board.legal_mask()])
        moves = board.legal_moves()
        if deterministic:
            move = moves[int(np.argmax(legal))]
        else:
            cdf = np.cumsum(legal.astype(np.float64))
            cdf /= cdf[-1]
            move = moves[int(np.searchsorted(cdf, rng.random_sample(), side="right"))]
        board.play(move, keep_same_player=True)
        if not board.over:
            model = model_previous if model is model_current else model_current
    result = board.result(keep_same_player=True)
    return (1 if model is model_current else -1) if result else 0


def score(results):
    """evaluate.py:124-129: wins / decisive games; 0.5 when every game was drawn."""
    r = np.asarray(results)
    if np.all(r == 0):
        return 0.5
    return float((r == 1).sum() / (r != 0).sum())
