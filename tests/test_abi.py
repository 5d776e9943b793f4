"""CPU checks of the C-ABI library: it loads, exports every symbol include/az_b200.h declares, its
structs match the Python mirror, host-only entry points work and device entry points fail loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from tests.helpers import ROOT


def _header_functions():
    text = open(os.path.join(ROOT, "include", "az_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(az_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from az_b200 import native

    lib = native.lib()
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/az_b200.h but not exported"
    assert sorted(native.SYMBOLS) == names, "az_b200/native.py must bind exactly the header's functions"
    assert lib.az_abi_version() == native.AZ_ABI_VERSION


def test_struct_mirrors_match():
    from az_b200 import native

    cs, ls = ctypes.c_size_t(), ctypes.c_size_t()
    native.lib().az_struct_sizes(ctypes.byref(cs), ctypes.byref(ls))
    assert cs.value == ctypes.sizeof(native.AzConfig) and ls.value == ctypes.sizeof(native.AzLayout)


def _cfg(**kw):
    from az_b200 import native

    base = dict(abi_version=native.AZ_ABI_VERSION, width=7, height=6, n_connect=4, gravity=1, n_trees=4096,
                node_capacity=22457, sims_per_move=800, index_move_greedy=8, eval_mode=0, prior_mode=1, move_mode=2,
                max_free_sims=8, fin_capacity=8192, pow_lut_len=33602, auto_restart=1, inline_play=1, eval_cache_log2=0, dirichlet_noise=0, dirichlet_alpha=0.03, dirichlet_ratio=0.25, c_puct=1.5, seed=0,
                game_id_base=0, games_target=4096)
    base.update(kw)
    return native.AzConfig(**base)


def test_layout_is_host_only_and_consistent():
    from az_b200 import native

    lay = native.AzLayout()
    cfg = _cfg()
    native.check(native.lib().az_query_layout(ctypes.byref(cfg), ctypes.byref(lay)))
    assert (lay.n_actions, lay.max_plies, lay.words) == (7, 42, 1)
    core = [n for n in native.LAYOUT_ARRAYS if not n.startswith("cache_")]
    offs = [getattr(lay, n) for n in core]
    assert len(set(offs)) == len(offs) and all(o % 256 == 0 for o in offs) and lay.total_bytes > max(offs)
    assert lay.cache_meta == lay.cache_key == lay.cache_val == 0  # evaluation memo off: no table in the slab
    cfg_memo = _cfg(eval_cache_log2=20)
    lay2 = native.AzLayout()
    native.check(native.lib().az_query_layout(ctypes.byref(cfg_memo), ctypes.byref(lay2)))
    assert lay2.cache_meta > 0 and lay2.cache_key - lay2.cache_meta >= 4 << 20 and lay2.cache_val - lay2.cache_key >= 16 << 20
    assert lay2.total_bytes - lay.total_bytes >= (4 + 16 + 32) << 20
    # node pools dominate: T * 2 halves * C * 24 B
    assert lay.node_p - lay.node_a >= 4096 * 2 * 22457 * 16
    cfg9 = _cfg(width=9, height=9, n_connect=5, gravity=0)
    native.check(native.lib().az_query_layout(ctypes.byref(cfg9), ctypes.byref(lay)))
    assert (lay.n_actions, lay.max_plies, lay.words) == (81, 81, 2)


@pytest.mark.parametrize("bad", [dict(width=12), dict(height=1), dict(n_connect=8), dict(n_connect=1),
                                 dict(abi_version=99), dict(n_trees=0), dict(node_capacity=1 << 25),
                                 dict(width=11, height=11)])
def test_bad_configuration_is_rejected(bad):
    from az_b200 import native

    lay = native.AzLayout()
    cfg = _cfg(**bad)
    rc = native.lib().az_query_layout(ctypes.byref(cfg), ctypes.byref(lay))
    assert rc == native.AZ_ERR_ARG and native.lib().az_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the engine must refuse to exist (this test is skipped on the GPU box)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from az_b200 import engine, env, native

    with pytest.raises(native.NativeError):
        engine.TreeEngine(engine.Rules(), n_trees=2, sims_per_move=4)
    with pytest.raises(native.NativeError):
        import numpy as np

        env.env_legal(engine.Rules(), np.zeros((1, 6, 7), dtype=np.int8))
    # straight through the ABI: AZ_ERR_NO_DEVICE, not a silent success
    cfg = _cfg(n_trees=2, node_capacity=64, fin_capacity=2, pow_lut_len=8, games_target=2)
    buf = (ctypes.c_char * 256)()
    lut = (ctypes.c_double * 8)()
    h = ctypes.c_void_p()
    rc = native.lib().az_engine_create(ctypes.byref(cfg), buf, 256, lut, None, ctypes.byref(h))
    assert rc == native.AZ_ERR_NO_DEVICE and h.value is None


def _chess_cfg(**kw):
    from az_b200 import native

    base = dict(abi_version=native.AZ_ABI_VERSION, n_trees=1024, node_capacity=60000, sims_per_move=200, index_move_greedy=8,
                eval_mode=0, prior_mode=1, move_mode=2, max_free_sims=8, max_plies=512, sample_capacity=8192,
                fin_capacity=2048, pow_lut_len=80002, auto_restart=1, c_puct=1.5, seed=0, game_id_base=0, games_target=1024)
    base.update(kw)
    return native.AzChessConfig(**base)


def test_chess_layout_and_action_table_are_host_only():
    """The chess engine's layout query and the action list need no device; the struct mirrors match."""
    from az_b200 import chess, native

    cs, ls = ctypes.c_size_t(), ctypes.c_size_t()
    native.lib().az_chess_struct_sizes(ctypes.byref(cs), ctypes.byref(ls))
    assert cs.value == ctypes.sizeof(native.AzChessConfig) and ls.value == ctypes.sizeof(native.AzChessLayout)
    lay = native.AzChessLayout()
    cfg = _chess_cfg()
    native.check(native.lib().az_chess_query_layout(ctypes.byref(cfg), ctypes.byref(lay)))
    offs = [getattr(lay, n) for n in native.CHESS_LAYOUT_ARRAYS]
    assert len(set(offs)) == len(offs) and all(o % 256 == 0 for o in offs) and lay.total_bytes > max(offs)
    assert lay.node_p - lay.node_a >= 1024 * 2 * 60000 * 16 and lay.node_m - lay.node_p >= 1024 * 2 * 60000 * 8
    assert lay.smp_n - lay.smp_act >= 8192 * 224 * 2
    for bad in (dict(node_capacity=100), dict(node_capacity=1 << 24), dict(n_trees=0), dict(max_plies=0),
                dict(eval_mode=3), dict(abi_version=7), dict(sample_capacity=0)):
        rc = native.lib().az_chess_query_layout(ctypes.byref(_chess_cfg(**bad)), ctypes.byref(lay))
        assert rc == native.AZ_ERR_ARG and native.lib().az_last_error()
    table = chess.action_table()
    assert table.shape == (1880,) and chess.action_uci(0) == "a1a2" and chess.uci_action("e7e8q") >= 0
    assert chess.uci_action("e2e4") == int(table.tolist().index(12 | 28 << 6))


def test_chess_has_no_cpu_fallback():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from az_b200 import chess, chess_engine, native

    start = chess.position_from_fen()
    for call in (lambda: chess.chess_legal(start[None]), lambda: chess.chess_play(start[None], [0]),
                 lambda: chess.chess_encode(start[None]), lambda: chess.chess_perft(start[None], 2),
                 lambda: chess_engine.ChessTreeEngine(n_trees=2, sims_per_move=4)):
        with pytest.raises(native.NativeError):
            call()
    # straight through the ABI: AZ_ERR_NO_DEVICE, not a silent success
    buf = (ctypes.c_uint64 * 8)(*[int(x) for x in start])
    out = (ctypes.c_uint64 * 1)()
    assert native.lib().az_chess_perft(buf, 1, 2, out, None) == native.AZ_ERR_NO_DEVICE
    cfg = _chess_cfg(n_trees=2, node_capacity=256, sample_capacity=4, fin_capacity=2, pow_lut_len=8, games_target=2)
    slab = (ctypes.c_char * 256)()
    lut = (ctypes.c_double * 8)()
    h = ctypes.c_void_p()
    rc = native.lib().az_chess_engine_create(ctypes.byref(cfg), slab, 256, lut, None, ctypes.byref(h))
    assert rc == native.AZ_ERR_NO_DEVICE and h.value is None
    # the drop-in classes are GPU-backed too
    from custom_alphazero.chess.board import Board

    with pytest.raises(native.NativeError):
        Board().moves
    assert np.array_equal(Board().array[7], [4, 2, 3, 5, 6, 3, 2, 4])  # the array itself is data-format glue
