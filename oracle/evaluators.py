"""Fixed evaluators used by the parity tests (TEST INFRASTRUCTURE - see oracle/__init__.py).

An evaluator maps the reference's NN input (Board.full_state, [H, W, 4] float32,
/root/reference/custom_alphazero/connect_n/board.py:83-98) to (priors[A] float64, value float),
i.e. what custom_alphazero.serving.factory.infer_sample returns
(/root/reference/custom_alphazero/serving/factory.py:21-55).  The CUDA engine implements the
same two functions in-kernel (custom-alphazero_b200/csrc/az_eval.cuh) so that bit-exact
comparison needs no network.
"""
import numpy as np

MASK64 = (1 << 64) - 1
FNV_OFFSET = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
GOLDEN = 0x9E3779B97F4A7C15
MIX = 0xFF51AFD7ED558CCD


def uniform_evaluator(n_actions):
    """SURVEY 8c fixed evaluator: np.full(A, 1/A) float64, value 0.0."""

    def f(state):
        return np.full(n_actions, 1 / n_actions), 0.0

    return f


def cells_from_state(state):
    """0 empty / 1 side-to-move stone / 2 opponent stone, row-major with row 0 on top."""
    return np.argmax(np.asarray(state)[:, :, :3], axis=-1).ravel()


def hash_of_cells(cells):
    h = FNV_OFFSET
    for c in cells:
        h = ((h ^ (int(c) + 1)) * FNV_PRIME) & MASK64
    return h


def hash_outputs(h, n_actions):
    priors = np.empty(n_actions, dtype=np.float64)
    for a in range(n_actions):
        m = ((h ^ ((a * GOLDEN) & MASK64)) * MIX) & MASK64
        priors[a] = float(((m >> 40) % 1000) + 1)
    value = (float((h >> 20) % 2001) - 1000.0) / 1000.0
    return priors, value


def hash_evaluator(n_actions):
    """State-dependent evaluator with float64-exact outputs (same spec as
    tests/golden/make_golden.py:hash_evaluator, which fed the reference)."""

    def f(state):
        return hash_outputs(hash_of_cells(cells_from_state(state)), n_actions)

    return f


def make(name, n_actions):
    return {"uniform": uniform_evaluator, "hash": hash_evaluator}[name](n_actions)
