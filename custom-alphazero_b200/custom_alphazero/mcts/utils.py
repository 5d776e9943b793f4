"""normalize_probabilities: host utility with the reference's exact semantics (mcts/utils.py:4-16).
Inside the search the same normalisation runs in the expand kernel (csrc/az_tree.cuh normalise_sel)."""
import numpy as np


def normalize_probabilities(probabilities: np.ndarray) -> np.ndarray:
    n = len(probabilities)
    assert n > 0
    total = probabilities.sum()
    if total == 0:  # nothing but zeros: uniform
        return np.full(n, 1 / n)
    return np.divide(probabilities, total, out=np.zeros_like(probabilities), where=total != 0)
