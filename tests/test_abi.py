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
