"""ctypes binding of libaz_b200.so (include/az_b200.h).  Mirrors the header one to one."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "libaz_b200.so")

AZ_ABI_VERSION = 1
AZ_MAX_ACTIONS = 128
AZ_MAX_DEPTH = 128

AZ_OK, AZ_ERR_ARG, AZ_ERR_SLAB, AZ_ERR_CUDA, AZ_ERR_NO_DEVICE = range(5)
AZ_EVAL_EXTERNAL, AZ_EVAL_UNIFORM, AZ_EVAL_HASH = range(3)
AZ_PRIOR_F64, AZ_PRIOR_F32 = range(2)
AZ_MOVE_ARGMAX, AZ_MOVE_HOST_UNIFORMS, AZ_MOVE_PHILOX = range(3)
AZ_F32, AZ_F64, AZ_BF16 = range(3)
AZ_PHASE_IDLE, AZ_PHASE_SEARCH, AZ_PHASE_READY, AZ_PHASE_STALLED = range(4)
AZ_PHASE_MASK = 0xFF
AZ_FLAG_POOL_OVERFLOW, AZ_FLAG_LUT_OVERFLOW, AZ_FLAG_ILLEGAL = 1 << 8, 1 << 9, 1 << 10
AZ_CHESS_ACTIONS, AZ_CHESS_MASK_WORDS, AZ_CHESS_PLANES = 1880, 30, 118
AZ_DENSE_HEAD_SPLITS = 4


class NativeError(RuntimeError):
    pass


class AzConfig(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("n_connect", ctypes.c_int32),
        ("gravity", ctypes.c_int32),
        ("n_trees", ctypes.c_int32),
        ("node_capacity", ctypes.c_int32),
        ("sims_per_move", ctypes.c_int32),
        ("index_move_greedy", ctypes.c_int32),
        ("eval_mode", ctypes.c_int32),
        ("prior_mode", ctypes.c_int32),
        ("move_mode", ctypes.c_int32),
        ("max_free_sims", ctypes.c_int32),
        ("fin_capacity", ctypes.c_int32),
        ("pow_lut_len", ctypes.c_int32),
        ("auto_restart", ctypes.c_int32),
        ("inline_play", ctypes.c_int32),
        ("eval_cache_log2", ctypes.c_int32),
        ("dirichlet_noise", ctypes.c_int32),
        ("dirichlet_alpha", ctypes.c_double),
        ("dirichlet_ratio", ctypes.c_double),
        ("c_puct", ctypes.c_double),
        ("seed", ctypes.c_uint64),
        ("game_id_base", ctypes.c_int64),
        ("games_target", ctypes.c_int64),
    ]


LAYOUT_ARRAYS = [
    "status", "ply", "game_id", "root_board", "half", "root_node", "n_nodes", "sims_done", "pending", "path_len", "path",
    "leaf_board", "counters", "uniforms", "node_a", "node_p", "rec_visits", "rec_action", "rec_board", "rec_len",
    "result", "fin_count", "fin_game_id", "fin_len", "fin_result", "fin_visits", "fin_action", "fin_board", "pow_lut",
    "cache_meta", "cache_key", "cache_val",
]


class AzLayout(ctypes.Structure):
    _fields_ = (
        [("total_bytes", ctypes.c_size_t)]
        + [(n, ctypes.c_int32) for n in ("n_actions", "max_plies", "words", "max_depth")]
        + [(n, ctypes.c_size_t) for n in LAYOUT_ARRAYS]
    )


class AzChessConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "abi_version", "n_trees", "node_capacity", "sims_per_move", "index_move_greedy", "eval_mode", "prior_mode",
        "move_mode", "max_free_sims", "max_plies", "sample_capacity", "fin_capacity", "pow_lut_len", "auto_restart")] + [
        ("c_puct", ctypes.c_double), ("seed", ctypes.c_uint64), ("game_id_base", ctypes.c_int64),
        ("games_target", ctypes.c_int64)]


CHESS_LAYOUT_ARRAYS = [
    "status", "ply", "game_id", "root_pos", "half", "root_node", "n_nodes", "sims_done", "pending", "path_len", "path",
    "leaf_pos", "leaf_mask", "counters", "uniforms", "node_a", "node_p", "node_m", "smp_count", "smp_game", "smp_ply",
    "smp_pos", "smp_k", "smp_act", "smp_n", "smp_choice", "fin_count", "fin_game", "fin_len", "fin_result", "pow_lut",
]
AZ_CHESS_MAX_CHILDREN = 224


class AzChessLayout(ctypes.Structure):
    _fields_ = [("total_bytes", ctypes.c_size_t)] + [(n, ctypes.c_size_t) for n in CHESS_LAYOUT_ARRAYS]


class AzNetHeadParams(ctypes.Structure):
    """az_net_head_params (include/az_b200.h): plain row-major float32 head weights of az_net_forward."""

    _fields_ = [(n, ctypes.c_void_p) for n in ("conv_w", "conv_b", "policy_w", "policy_b", "value1_w", "value1_b",
                                               "value2_w", "value2_b")]


class AzHeadWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("conv_w", "conv_b", "policy_w", "policy_b", "value1_w", "value1_b",
                                                "value2_w", "value2_b")]


# every symbol include/az_b200.h declares: (restype, argtypes)
_P, _I, _S = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
SYMBOLS = {
    "az_last_error": (ctypes.c_char_p, []),
    "az_abi_version": (ctypes.c_int, []),
    "az_struct_sizes": (None, [ctypes.POINTER(_S), ctypes.POINTER(_S)]),
    "az_query_layout": (ctypes.c_int, [ctypes.POINTER(AzConfig), ctypes.POINTER(AzLayout)]),
    "az_engine_create": (ctypes.c_int, [ctypes.POINTER(AzConfig), _P, _S, _P, _P, ctypes.POINTER(_P)]),
    "az_engine_destroy": (None, [_P]),
    "az_reset_games": (ctypes.c_int, [_P, _P]),
    "az_set_game_id_base": (ctypes.c_int, [_P, ctypes.c_int64]),
    "az_set_roots": (ctypes.c_int, [_P, _P, _P, _P, _I, _P]),
    "az_begin_search": (ctypes.c_int, [_P, _I, _P]),
    "az_step": (ctypes.c_int, [_P, _P, _P, _I, _P, _I, _P, _P]),
    "az_step_gather": (ctypes.c_int, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "az_extra_sims": (ctypes.c_int, [_P, _I, _P]),
    "az_search": (ctypes.c_int, [_P, _P]),
    "az_play": (ctypes.c_int, [_P, _I, _I, _P]),
    "az_fin_clear": (ctypes.c_int, [_P, _P]),
    "az_cache_clear": (ctypes.c_int, [_P, _P]),
    "az_env_play": (ctypes.c_int, [ctypes.POINTER(AzConfig), _P, _P, _I, _P, _P, _P]),
    "az_env_legal": (ctypes.c_int, [ctypes.POINTER(AzConfig), _P, _I, _P, _P]),
    "az_env_encode": (ctypes.c_int, [ctypes.POINTER(AzConfig), _P, _I, _P, _P]),
    "az_net_stem": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "az_net_stem_tc": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "az_net_heads": (ctypes.c_int, [_P, ctypes.POINTER(AzHeadWeights), _I, _I, _I, _I, _P, _P, _P]),
    "az_net_heads_dense": (ctypes.c_int, [_P, ctypes.POINTER(AzHeadWeights), _I, _I, _I, _P, _P, _P]),
    "az_net_conv1x1": (ctypes.c_int, [_P, _P, ctypes.c_int64, _I, _P, _P]),
    "az_net_tower_timing": (ctypes.c_int, [_P]),
    "az_net_tower": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "az_net_forward": (ctypes.c_int, [_P, _P, _P, _P, ctypes.POINTER(AzNetHeadParams), _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "az_net_forward_gathered": (ctypes.c_int, [_P, _P, _P, _P, ctypes.POINTER(AzNetHeadParams), _P, _P, _I, _I, _I, _I, _I,
                                               _I, _I, _P, _P, _P]),
    "az_net_forward_trees": (ctypes.c_int, [_P, _P, _P, _P, ctypes.POINTER(AzNetHeadParams), _P, _P, _I, _I, _I, _I, _I,
                                            _I, _I, _P, _P, _P, _I, _P]),
    "az_net_dense_heads": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "az_net_head_convs": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "az_advance_fused": (ctypes.c_int, [_P, _P, ctypes.POINTER(AzHeadWeights), _P, _P, _P, _P, _P]),
    "az_debug_dirichlet": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_double, _I, _I, _P, _P]),
    "az_debug_timeline": (ctypes.c_int, [_P, _P, _I]),
    "az_net_debug_timeline": (ctypes.c_int, [_P, _I]),
    "az_decode_samples": (ctypes.c_int, [ctypes.POINTER(AzConfig), _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P]),
    # chess (az_chess_pos = 8 x uint64, passed as plain device pointers)
    "az_chess_action_table": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint16)]),
    "az_chess_legal": (ctypes.c_int, [_P, _I, _P, _P, _P, _P]),
    "az_chess_play": (ctypes.c_int, [_P, _P, _I, _I, _P, _P, _P]),
    "az_chess_encode": (ctypes.c_int, [_P, _P, _I, _I, _P, _P]),
    "az_chess_perft": (ctypes.c_int, [_P, _I, _I, _P, _P]),
    "az_chess_struct_sizes": (None, [ctypes.POINTER(_S), ctypes.POINTER(_S)]),
    "az_chess_query_layout": (ctypes.c_int, [ctypes.POINTER(AzChessConfig), ctypes.POINTER(AzChessLayout)]),
    "az_chess_engine_create": (ctypes.c_int, [ctypes.POINTER(AzChessConfig), _P, _S, _P, _P, ctypes.POINTER(_P)]),
    "az_chess_engine_destroy": (None, [_P]),
    "az_chess_reset_games": (ctypes.c_int, [_P, _P]),
    "az_chess_set_game_id_base": (ctypes.c_int, [_P, ctypes.c_int64]),
    "az_chess_set_roots": (ctypes.c_int, [_P, _P, _P, _I, _P]),
    "az_chess_begin_search": (ctypes.c_int, [_P, _I, _P]),
    "az_chess_search": (ctypes.c_int, [_P, _P]),
    "az_chess_step": (ctypes.c_int, [_P, _P, _P, _I, _P, _I, _I, _P, _P]),
    "az_chess_stem": (ctypes.c_int, [_P, _I, _P, _P, _P, _P]),
    "az_chess_stem_tc": (ctypes.c_int, [_P, _I, _P, _P, _P, _P]),
    "az_chess_move": (ctypes.c_int, [_P, _I, _I, _P]),
    "az_chess_rings_clear": (ctypes.c_int, [_P, _P]),
    "az_chess_decode_samples": (ctypes.c_int, [_P, _P, _P, _P, _P, _I, _P, _P, _P]),
}

_lib = None


def lib():
    """Loads libaz_b200.so.  No fallback: a missing library is an error, loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with custom-alphazero_b200/build.sh "
                "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.az_abi_version() != AZ_ABI_VERSION:
            raise NativeError("libaz_b200.so ABI version mismatch")
        cs, ls = _S(), _S()
        handle.az_struct_sizes(ctypes.byref(cs), ctypes.byref(ls))
        if cs.value != ctypes.sizeof(AzConfig) or ls.value != ctypes.sizeof(AzLayout):
            raise NativeError("az_config / az_layout mirror out of sync with include/az_b200.h")
        handle.az_chess_struct_sizes(ctypes.byref(cs), ctypes.byref(ls))
        if cs.value != ctypes.sizeof(AzChessConfig) or ls.value != ctypes.sizeof(AzChessLayout):
            raise NativeError("az_chess_config / az_chess_layout mirror out of sync with include/az_b200.h")
        _lib = handle
    return _lib


def check(rc):
    if rc != AZ_OK:
        msg = lib().az_last_error().decode()
        raise NativeError(f"libaz_b200 error {rc}: {msg}")
