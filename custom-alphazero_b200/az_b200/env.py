"""Batched Connect-N environment calls (K2 / K3 kernels through the C ABI).

Inputs may be numpy arrays or torch tensors on any device; they are staged to the GPU, the kernel
runs there, and results come back as the same kind the caller passed (numpy in -> numpy out).
Cell convention is the reference's (connect_n/board.py): int8 [n, H, W], row 0 on top, +1 = side to
move, -1 = opponent, 0 = empty.
"""
import ctypes

import numpy as np
import torch

from . import native
from .engine import Rules, _ptr, _stream
from .native import AzConfig, NativeError, check, lib


def _cfg(rules):
    return AzConfig(abi_version=native.AZ_ABI_VERSION, width=rules.width, height=rules.height, n_connect=rules.n,
                    gravity=int(rules.gravity))


def _dev():
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the Connect-N kernels have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dtype):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=dtype).to(_dev()).contiguous()


def _back(t, like):
    return t if torch.is_tensor(like) and like.is_cuda else t.cpu().numpy()


def env_play(rules: Rules, cells, actions):
    """Board.play(move, keep_same_player=True) for n boards: (cells_out, status); status 0 ongoing,
    1 mover won, 2 draw, -1 illegal action (board returned unchanged)."""
    c = _to_dev(cells, torch.int8)
    a = _to_dev(actions, torch.int32)
    n = a.numel()
    assert c.shape == (n, rules.height, rules.width)
    out = torch.empty_like(c)
    status = torch.empty(n, dtype=torch.int32, device=c.device)
    cfg = _cfg(rules)
    check(lib().az_env_play(ctypes.byref(cfg), _ptr(c), _ptr(a), n, _ptr(out), _ptr(status), _stream()))
    return _back(out, cells), _back(status, cells)


def env_legal(rules: Rules, cells):
    """Board.legal_moves_mask(get_all_possible_moves()) for n boards: bool [n, A]."""
    c = _to_dev(cells, torch.int8)
    n = c.shape[0]
    out = torch.empty((n, rules.n_actions), dtype=torch.uint8, device=c.device)
    cfg = _cfg(rules)
    check(lib().az_env_legal(ctypes.byref(cfg), _ptr(c), n, _ptr(out), _stream()))
    return _back(out.bool(), cells)


def env_encode(rules: Rules, cells):
    """Board.full_state for n boards: float32 [n, H, W, 4]."""
    c = _to_dev(cells, torch.int8)
    n = c.shape[0]
    out = torch.empty((n, rules.height, rules.width, 4), dtype=torch.float32, device=c.device)
    cfg = _cfg(rules)
    check(lib().az_env_encode(ctypes.byref(cfg), _ptr(c), n, _ptr(out), _stream()))
    return _back(out, cells)


def board_order_actions(rules: Rules, legal_row):
    """Legal actions of one board in BOARD move order (connect_n/board.py:113-124): ascending x with
    gravity, row-major (y, x) without - while the action index is x-major (x * H + y)."""
    idx = np.nonzero(np.asarray(legal_row))[0]
    if rules.gravity:
        return [int(a) for a in idx]
    return sorted((int(a) for a in idx), key=lambda a: (a % rules.height, a // rules.height))
