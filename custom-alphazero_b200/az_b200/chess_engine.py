"""ChessTreeEngine - Python handle on one GPU's batch of chess game trees (az_chess_* engine entry points).

Same division of labour as engine.TreeEngine: torch provides the device slab, typed views and streams; every
simulation, move, re-root and game end is a kernel behind the C ABI (include/az_b200.h, "chess search engine").
"""
import ctypes

import numpy as np
import torch

from . import native
from .chess import N_ACTIONS, PLANES
from .engine import _ptr, _stream, pow_half_table
from .native import AzChessConfig, AzChessLayout, NativeError, check, lib

K = native.AZ_CHESS_MAX_CHILDREN

_DTYPES = {
    "status": torch.int32, "ply": torch.int32, "game_id": torch.int64, "root_pos": torch.int64, "half": torch.int32,
    "root_node": torch.int32, "n_nodes": torch.int32, "sims_done": torch.int32, "pending": torch.int32,
    "path_len": torch.int32, "path": torch.int32, "leaf_pos": torch.int64, "leaf_mask": torch.int64,
    "counters": torch.int64, "uniforms": torch.float64, "node_p": torch.float64, "node_m": torch.int16,
    "smp_count": torch.int32, "smp_game": torch.int64, "smp_ply": torch.int32, "smp_pos": torch.int64,
    "smp_k": torch.int32, "smp_act": torch.int16, "smp_n": torch.int32, "smp_choice": torch.int32,
    "fin_count": torch.int32, "fin_game": torch.int64, "fin_len": torch.int32, "fin_result": torch.int32,
    "pow_lut": torch.float64,
}


class ChessTreeEngine:
    def __init__(self, n_trees=1, sims_per_move=250, *, eval_mode="external", prior_mode="f32", move_mode="argmax",
                 node_capacity=None, games_target=None, game_id_base=0, seed=0, auto_restart=False, max_plies=512,
                 sample_capacity=None, fin_capacity=None, max_free_sims=8, index_move_greedy=8, c_puct=1.5,
                 pow_lut_len=None, device=None):
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the chess search engine has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_trees = T = int(n_trees)
        if node_capacity is None:
            # a search adds about 32 nodes per simulation (35 legal moves on average, 8-aligned blocks); room for the
            # kept subtree plus a few searches, bounded by a quarter of the free HBM.  Exhaustion is flagged, never silent.
            want = 6 * (sims_per_move + 1) * 48 + 1024
            free_bytes, _ = torch.cuda.mem_get_info(self.device)
            node_capacity = max(4096, min(0xFFFFFF, want, (free_bytes // 4) // (T * 2 * 26)))
        if games_target is None:
            games_target = T
        if sample_capacity is None:
            sample_capacity = max(64, 4 * T)
        if fin_capacity is None:
            fin_capacity = max(int(games_target), 1) if not auto_restart else max(2 * T, 1)
        if pow_lut_len is None:
            pow_lut_len = min(max_plies, 400) * max(sims_per_move, 1) + 2
        cfg = AzChessConfig(
            abi_version=native.AZ_ABI_VERSION, n_trees=T, node_capacity=int(node_capacity), sims_per_move=int(sims_per_move),
            index_move_greedy=int(index_move_greedy), eval_mode={"external": 0, "uniform": 1, "hash": 2}[eval_mode],
            prior_mode={"f64": 0, "f32": 1}[prior_mode], move_mode={"argmax": 0, "host_uniforms": 1, "philox": 2}[move_mode],
            max_free_sims=int(max_free_sims), max_plies=int(max_plies), sample_capacity=int(sample_capacity),
            fin_capacity=int(fin_capacity), pow_lut_len=int(pow_lut_len), auto_restart=int(auto_restart),
            c_puct=float(c_puct), seed=int(seed), game_id_base=int(game_id_base), games_target=int(games_target))
        self.cfg = cfg
        self.layout = AzChessLayout()
        check(lib().az_chess_query_layout(ctypes.byref(cfg), ctypes.byref(self.layout)))
        with torch.cuda.device(self.device):
            self.slab = torch.zeros(self.layout.total_bytes, dtype=torch.uint8, device=self.device)
            lut = pow_half_table(cfg.pow_lut_len)
            handle = ctypes.c_void_p()
            check(lib().az_chess_engine_create(ctypes.byref(cfg), _ptr(self.slab), self.layout.total_bytes,
                                               lut.ctypes.data_as(ctypes.c_void_p), _stream(), ctypes.byref(handle)))
        self._h = handle
        self.sims_per_move = int(sims_per_move)
        S, F, C, P = cfg.sample_capacity, cfg.fin_capacity, cfg.node_capacity, cfg.max_plies
        self._shapes = {
            "root_pos": (T, 8), "path": (T, native.AZ_MAX_DEPTH), "leaf_pos": (T, 8), "leaf_mask": (T, 32),
            "counters": (T, 8), "uniforms": (T, P), "node_p": (T, 2, C), "node_m": (T, 2, C), "smp_count": (4,),
            "smp_game": (S,), "smp_ply": (S,), "smp_pos": (S, 8), "smp_k": (S,), "smp_act": (S, K), "smp_n": (S, K),
            "smp_choice": (S,), "fin_count": (4,), "fin_game": (F,), "fin_len": (F,), "fin_result": (F,),
            "pow_lut": (cfg.pow_lut_len,),
        }

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:
            try:
                lib().az_chess_engine_destroy(h)
            except Exception:
                pass
            self._h = None

    def view(self, name):
        dt = _DTYPES[name]
        shape = self._shapes.get(name, (self.n_trees,))
        n = int(np.prod(shape))
        off = getattr(self.layout, name)
        return self.slab[off: off + n * dt.itemsize].view(dt).view(*shape)

    def node_view(self):
        T, C = self.n_trees, self.cfg.node_capacity
        off = self.layout.node_a
        raw = self.slab[off: off + 16 * T * 2 * C]
        w = raw.view(torch.float64).view(T, 2, C, 2)[..., 0]
        ints = raw.view(torch.int32).view(T, 2, C, 4)
        return w, ints[..., 2], ints[..., 3]

    # ------------------------------------------------------------------ C ABI calls
    def reset(self, game_id_base=None):
        if game_id_base is not None:  # see TreeEngine.reset
            check(lib().az_chess_set_game_id_base(self._h, int(game_id_base)))
            self.cfg.game_id_base = int(game_id_base)
        check(lib().az_chess_reset_games(self._h, _stream()))

    def set_roots(self, tree_ids, positions):
        ids = torch.as_tensor(tree_ids, dtype=torch.int32, device=self.device).contiguous()
        pos = np.ascontiguousarray(np.asarray(positions, dtype=np.uint64)).view(np.int64)
        pos = torch.from_numpy(pos).to(self.device).reshape(-1, 8)
        assert pos.shape[0] == ids.numel()
        check(lib().az_chess_set_roots(self._h, _ptr(ids), _ptr(pos), ids.numel(), _stream()))

    def begin_search(self, sims=0):
        if sims:
            self.sims_per_move = int(sims)
        check(lib().az_chess_begin_search(self._h, int(sims), _stream()))

    def search(self):
        check(lib().az_chess_search(self._h, _stream()))

    def step(self, priors, values, states_out, leaf_valid_out, plane_first=0):
        """One lock-step advance.  priors [T, 1880] / values [T] float32 or float64 (or None on the first call);
        states_out bf16 [T, 8, 8, 118 or more]; leaf_valid_out int32 [T]."""
        eval_dtype = native.AZ_F32
        if priors is not None:
            assert priors.is_contiguous() and values.is_contiguous() and priors.dtype == values.dtype
            assert priors.shape == (self.n_trees, N_ACTIONS)
            eval_dtype = {torch.float32: native.AZ_F32, torch.float64: native.AZ_F64}[priors.dtype]
        assert leaf_valid_out.dtype == torch.int32
        stride = PLANES
        if states_out is not None:  # None: no planes, the caller runs az_chess_stem on view("leaf_pos")
            assert states_out.dtype == torch.bfloat16 and states_out.is_contiguous()
            stride = states_out.shape[-1]  # extra channels are written as zeros (channel padding for the stem)
            assert states_out.shape == (self.n_trees, 8, 8, stride) and stride >= PLANES - plane_first
        check(lib().az_chess_step(self._h, _ptr(priors), _ptr(values), eval_dtype, _ptr(states_out), stride, int(plane_first),
                                  _ptr(leaf_valid_out), _stream()))

    def move(self, greedy=None, move_mode=None):
        g = -1 if greedy is None else int(bool(greedy))
        m = -1 if move_mode is None else {"argmax": 0, "host_uniforms": 1, "philox": 2}[move_mode]
        check(lib().az_chess_move(self._h, g, m, _stream()))

    def rings_clear(self):
        check(lib().az_chess_rings_clear(self._h, _stream()))

    # ------------------------------------------------------------------ host-side conveniences
    def set_uniforms(self, uniforms):
        u = torch.as_tensor(np.asarray(uniforms, dtype=np.float64), device=self.device)
        self.view("uniforms")[:, : u.shape[1]].copy_(u)

    def phases(self):
        return self.view("status") & native.AZ_PHASE_MASK

    def check_status(self):
        st = self.view("status")
        bad = (st & 0xFF00).ne(0)
        if bool(bad.any()):
            t = int(torch.nonzero(bad)[0])
            flags = int(st[t]) & 0xFF00
            names = [n for n, b in (("node pool exhausted", native.AZ_FLAG_POOL_OVERFLOW),
                                    ("pow-half table exceeded", native.AZ_FLAG_LUT_OVERFLOW),
                                    ("tree deeper than AZ_MAX_DEPTH / edgeless root", native.AZ_FLAG_ILLEGAL)) if flags & b]
            raise NativeError(f"chess tree {t}: " + ", ".join(names))

    def totals(self):
        c = self.view("counters")
        s = c.sum(dim=0).tolist()
        return {"sims": s[0], "evals": s[1], "moves": s[2], "games": s[3], "depth_sum": s[4], "children": s[5],
                "reroot_nodes": s[6], "pool_high_water": int(c[:, 7].max())}

    def drain(self):
        """Copies the sample ring and the finished-game ring to the host and empties both."""
        n = int(self.view("smp_count")[0])
        f = int(self.view("fin_count")[0])
        out = {
            "game": self.view("smp_game")[:n].cpu().numpy(), "ply": self.view("smp_ply")[:n].cpu().numpy(),
            "pos": self.view("smp_pos")[:n].cpu().numpy().view(np.uint64), "k": self.view("smp_k")[:n].cpu().numpy(),
            "act": self.view("smp_act")[:n].cpu().numpy().view(np.uint16), "n": self.view("smp_n")[:n].cpu().numpy(),
            "choice": self.view("smp_choice")[:n].cpu().numpy(),
            "fin_game": self.view("fin_game")[:f].cpu().numpy(), "fin_len": self.view("fin_len")[:f].cpu().numpy(),
            "fin_result": self.view("fin_result")[:f].cpu().numpy(),
        }
        self.rings_clear()
        return out

    def root_stats(self, tree=0):
        """Root edges of one tree: (actions, N, W, P) in ascending action order."""
        w, n, link = self.node_view()
        half = int(self.view("half")[tree])
        root = int(self.view("root_node")[tree])
        lk = int(link[tree, half, root]) & 0xFFFFFFFF
        base, k = lk & 0xFFFFFF, lk >> 24
        sl = slice(base, base + k)
        acts = (self.view("node_m")[tree, half, sl].cpu().numpy().view(np.uint16)).tolist()
        return acts, n[tree, half, sl].tolist(), w[tree, half, sl].tolist(), self.view("node_p")[tree, half, sl].tolist()


def decode_samples(drained, device=None):
    """Sample-ring entries -> (states float32 [n, 8, 8, 118], policies float64 [n, 1880]) on the device
    (az_chess_decode_samples)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = len(drained["k"])
    states = torch.empty((n, 8, 8, PLANES), dtype=torch.float32, device=dev)
    policies = torch.empty((n, N_ACTIONS), dtype=torch.float64, device=dev)
    if n == 0:
        return states, policies
    pos = torch.from_numpy(np.ascontiguousarray(drained["pos"]).view(np.int64)).to(dev)
    k = torch.from_numpy(np.ascontiguousarray(drained["k"])).to(dev)
    act = torch.from_numpy(np.ascontiguousarray(drained["act"]).view(np.int16)).to(dev)
    nv = torch.from_numpy(np.ascontiguousarray(drained["n"])).to(dev)
    ch = torch.from_numpy(np.ascontiguousarray(drained["choice"])).to(dev)
    check(lib().az_chess_decode_samples(_ptr(pos), _ptr(k), _ptr(act), _ptr(nv), _ptr(ch), n, _ptr(states), _ptr(policies),
                                        _stream()))
    return states, policies


def sample_values(drained, games=None):
    """self_play.py:66-78: the value of the sample at ply i of a finished game of L plies with result r is
    r * (+1 for the last ply, alternating backwards).  Returns (values int32 [n], known bool [n]); samples of games
    that are not in the finished ring yet are marked unknown."""
    fin = {int(g): (int(ln), int(r)) for g, ln, r in zip(drained["fin_game"], drained["fin_len"], drained["fin_result"])}
    if games:
        fin.update(games)
    vals = np.zeros(len(drained["k"]), dtype=np.int32)
    known = np.zeros(len(drained["k"]), dtype=bool)
    for i, (g, p) in enumerate(zip(drained["game"], drained["ply"])):
        if int(g) in fin:
            ln, r = fin[int(g)]
            vals[i] = r * (1 if (ln - 1 - int(p)) % 2 == 0 else -1)
            known[i] = True
    return vals, known
