"""ctypes loader for the C restatement (oracle/c/az_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libaz_oracle.so")

EVAL_KINDS = {"uniform": 0, "hash": 1, "callback": 2}
PRIOR_F64, PRIOR_F32 = 0, 1


class Rules(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("n", ctypes.c_int), ("gravity", ctypes.c_int)]

    @property
    def n_actions(self):
        return self.width if self.gravity else self.width * self.height


EVAL_CB = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double),
                           ctypes.POINTER(ctypes.c_double), ctypes.c_void_p)


def build(force=False):
    src = os.path.join(HERE, "c", "az_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "_build/libaz_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.azo_pow_half.restype = ctypes.c_double
        _lib.azo_pow_half.argtypes = [ctypes.c_longlong]
    return _lib


def make_rules(width=7, height=6, n=4, gravity=True):
    return Rules(width, height, n, int(gravity))


def _wrap_callback(rules, fn):
    """fn(state[H, W, 4] float32) -> (priors[A], value)"""
    if fn is None:
        return ctypes.cast(None, EVAL_CB)
    H, W, A = rules.height, rules.width, rules.n_actions

    def cb(state_p, priors_p, value_p, _user):
        state = np.ctypeslib.as_array(state_p, shape=(H, W, 4)).copy()
        p, v = fn(state)
        for a in range(A):
            priors_p[a] = float(p[a])
        value_p[0] = float(v)

    return EVAL_CB(cb)


def normalise(p, prior_mode=PRIOR_F64):
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(p)
    lib().azo_normalise(p.ctypes.data_as(ctypes.c_void_p), len(p), prior_mode, out.ctypes.data_as(ctypes.c_void_p))
    return out


def env_playout(rules, lcg_state):
    n = rules.width * rules.height
    picked = (ctypes.c_int * n)()
    n_picked, result = ctypes.c_int(), ctypes.c_int()
    cells = np.zeros(n, dtype=np.int8)
    rc = lib().azo_env_playout(ctypes.byref(rules), ctypes.c_uint64(lcg_state), picked, ctypes.byref(n_picked),
                               ctypes.byref(result), cells.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return list(picked[: n_picked.value]), result.value, cells.reshape(rules.height, rules.width)


def search_once(rules, prefix, sims, evaluator="uniform", prior_mode=PRIOR_F64, callback=None):
    A = rules.n_actions
    act, n = (ctypes.c_int * A)(), (ctypes.c_int * A)()
    w, p = (ctypes.c_double * A)(), (ctypes.c_double * A)()
    evals = ctypes.c_longlong()
    pre = (ctypes.c_int * max(1, len(prefix)))(*prefix)
    cb = _wrap_callback(rules, callback)
    k = lib().azo_search_once(ctypes.byref(rules), pre, len(prefix), sims, EVAL_KINDS[evaluator], prior_mode, cb, None,
                              act, n, w, p, ctypes.byref(evals))
    assert k >= 0
    return {"actions": list(act[:k]), "N": list(n[:k]), "W": list(w[:k]), "P": list(p[:k]), "evals": evals.value}


def play_game(rules, sims, evaluator="uniform", uniforms=None, prior_mode=PRIOR_F64, callback=None, max_plies=None):
    """Returns dict(moves, visits[T, A] (-1 illegal), policies[T, A] f64, result, sims, evals)."""
    A = rules.n_actions
    cap = rules.width * rules.height if max_plies is None else max_plies
    moves = np.zeros(cap, dtype=np.int32)
    visits = np.zeros((cap, A), dtype=np.int32)
    policy = np.zeros((cap, A), dtype=np.float64)
    result = ctypes.c_int()
    sims_done, evals = ctypes.c_longlong(), ctypes.c_longlong()
    if uniforms is not None:
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert len(u) >= cap
        u_p = u.ctypes.data_as(ctypes.c_void_p)
    else:
        u_p = None
    cb = _wrap_callback(rules, callback)
    T = lib().azo_play_game(ctypes.byref(rules), sims, EVAL_KINDS[evaluator], prior_mode, cb, None, u_p, cap,
                            moves.ctypes.data_as(ctypes.c_void_p), visits.ctypes.data_as(ctypes.c_void_p),
                            policy.ctypes.data_as(ctypes.c_void_p), ctypes.byref(result), ctypes.byref(sims_done),
                            ctypes.byref(evals))
    assert T >= 0, "board mismatch after re-root"
    return {"moves": moves[:T].copy(), "visits": visits[:T].copy(), "policies": policy[:T].copy(),
            "result": result.value, "sims": sims_done.value, "evals": evals.value}
