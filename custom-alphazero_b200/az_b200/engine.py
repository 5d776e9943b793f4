"""TreeEngine - Python handle on one GPU's batch of game trees (libaz_b200.so).

torch is used for what the header calls plumbing: the device slab, typed views into it, streams.
Everything that computes is a kernel behind the C ABI (include/az_b200.h).
"""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import native
from .native import AzConfig, AzLayout, NativeError, check, lib


@dataclass(frozen=True)
class Rules:
    """ConfigConnectN of the reference (config.py:38-47)."""

    width: int = 7
    height: int = 6
    n: int = 4
    gravity: bool = True

    @property
    def n_actions(self):
        return self.width if self.gravity else self.width * self.height

    @property
    def max_plies(self):
        return self.width * self.height


_VIEW_DTYPES = {
    "status": torch.int32, "ply": torch.int32, "game_id": torch.int64, "root_board": torch.int64,
    "half": torch.int32, "root_node": torch.int32, "n_nodes": torch.int32, "sims_done": torch.int32, "pending": torch.int32,
    "path_len": torch.int32, "path": torch.int32, "leaf_board": torch.int64, "counters": torch.int64,
    "uniforms": torch.float64, "node_p": torch.float64, "rec_visits": torch.int32, "rec_action": torch.int32,
    "rec_board": torch.int64, "rec_len": torch.int32, "result": torch.int32, "fin_count": torch.int32,
    "fin_game_id": torch.int64, "fin_len": torch.int32, "fin_result": torch.int32, "fin_visits": torch.int32,
    "fin_action": torch.int32, "fin_board": torch.int64, "pow_lut": torch.float64,
}


def pow_half_table(n):
    """n -> n ** 0.5 exactly as CPython computes it for the reference (mcts/mcts.py:50); libm pow, not sqrt."""
    return np.array([i**0.5 for i in range(n)], dtype=np.float64)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(None)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class TreeEngine:
    def __init__(self, rules=Rules(), n_trees=1, sims_per_move=250, *, eval_mode="external", prior_mode="f32",
                 move_mode="argmax", node_capacity=None, games_target=None, game_id_base=0, seed=0,
                 auto_restart=False, fin_capacity=None, max_free_sims=8, index_move_greedy=8, c_puct=1.5,
                 pow_lut_len=None, device=None, inline_play=False, dirichlet_noise=False, dirichlet_alpha=0.03,
                 dirichlet_ratio=0.25, eval_cache_log2=0):
        if not torch.cuda.is_available():
            raise NativeError("no CUDA device: the self-play engine has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rules = rules
        self.n_trees = int(n_trees)
        A, P = rules.n_actions, rules.max_plies
        if node_capacity is None:
            # Enough for every expansion of a whole game (re-root then never has to compact), unless that
            # does not fit in half of the free HBM: then at least 4 searches' worth and the kept subtree is
            # compacted into the other pool half whenever the live half runs short.
            A8 = (A + 7) & ~7  # child blocks are 8-node aligned
            worst = (P * sims_per_move + 2) * A8 + 16
            floor = (4 * sims_per_move + 8) * A8 + 16
            free_bytes, _ = torch.cuda.mem_get_info(self.device)
            budget = (free_bytes // 2) // (self.n_trees * 2 * 24)
            node_capacity = min(0xFFFFFF, worst, max(floor, budget))
        if games_target is None:
            games_target = n_trees
        if fin_capacity is None:
            fin_capacity = max(int(games_target), 1) if not auto_restart else max(2 * n_trees, 1)
        if pow_lut_len is None:
            pow_lut_len = P * max(sims_per_move, 1) + 2
        cfg = AzConfig(
            abi_version=native.AZ_ABI_VERSION, width=rules.width, height=rules.height, n_connect=rules.n,
            gravity=int(rules.gravity), n_trees=self.n_trees, node_capacity=int(node_capacity),
            sims_per_move=int(sims_per_move), index_move_greedy=int(index_move_greedy),
            eval_mode={"external": 0, "uniform": 1, "hash": 2}[eval_mode],
            prior_mode={"f64": 0, "f32": 1}[prior_mode],
            move_mode={"argmax": 0, "host_uniforms": 1, "philox": 2}[move_mode],
            max_free_sims=int(max_free_sims), fin_capacity=int(fin_capacity), pow_lut_len=int(pow_lut_len),
            auto_restart=int(auto_restart), inline_play=int(inline_play), eval_cache_log2=int(eval_cache_log2),
            dirichlet_noise=int(dirichlet_noise),
            dirichlet_alpha=float(dirichlet_alpha), dirichlet_ratio=float(dirichlet_ratio), c_puct=float(c_puct), seed=int(seed), game_id_base=int(game_id_base),
            games_target=int(games_target),
        )
        self.cfg = cfg
        self.layout = AzLayout()
        check(lib().az_query_layout(ctypes.byref(cfg), ctypes.byref(self.layout)))
        with torch.cuda.device(self.device):
            self.slab = torch.zeros(self.layout.total_bytes, dtype=torch.uint8, device=self.device)
            lut = pow_half_table(cfg.pow_lut_len)
            handle = ctypes.c_void_p()
            check(lib().az_engine_create(ctypes.byref(cfg), _ptr(self.slab), self.layout.total_bytes,
                                         lut.ctypes.data_as(ctypes.c_void_p), _stream(), ctypes.byref(handle)))
        self._h = handle
        self.sims_per_move = int(sims_per_move)
        A8 = (A + 7) & ~7
        # with room for every expansion of a whole game the re-root never needs the compaction path
        self.never_compacts = cfg.node_capacity >= (P * sims_per_move + 2) * A8 + 16
        T, F, WD, C = self.n_trees, cfg.fin_capacity, self.layout.words, cfg.node_capacity
        self._shapes = {
            "root_board": (T, 2, WD), "path": (T, native.AZ_MAX_DEPTH), "leaf_board": (T, 2, WD), "counters": (T, 8),
            "uniforms": (T, P), "node_p": (T, 2, C), "rec_visits": (T, P, A), "rec_action": (T, P),
            "rec_board": (T, P, 2, WD), "fin_visits": (F, P, A), "fin_action": (F, P), "fin_board": (F, P, 2, WD),
            "fin_game_id": (F,), "fin_len": (F,), "fin_result": (F,), "fin_count": (4,), "pow_lut": (cfg.pow_lut_len,),
        }

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:  # module globals may already be gone at interpreter shutdown
            try:
                lib().az_engine_destroy(h)
            except Exception:
                pass
            self._h = None

    # ------------------------------------------------------------------ typed views into the slab
    def view(self, name):
        dt = _VIEW_DTYPES[name]
        shape = self._shapes.get(name, (self.n_trees,))
        n = int(np.prod(shape))
        off = getattr(self.layout, name)
        return self.slab[off: off + n * dt.itemsize].view(dt).view(*shape)

    def node_view(self):
        """(W float64, N int32, link uint32-as-int32) views [T, 2, C] over the 16-byte node records."""
        T, C = self.n_trees, self.cfg.node_capacity
        off = self.layout.node_a
        raw = self.slab[off: off + 16 * T * 2 * C]
        w = raw.view(torch.float64).view(T, 2, C, 2)[..., 0]
        ints = raw.view(torch.int32).view(T, 2, C, 4)
        return w, ints[..., 2], ints[..., 3]

    # ------------------------------------------------------------------ C ABI calls
    def reset(self, game_id_base=None):
        """Fresh games in every tree.  game_id_base moves the id range first (a new iteration of a self-play loop must
        not reuse ids: they key the sampling counters); graphs captured before are stale then."""
        if game_id_base is not None:
            check(lib().az_set_game_id_base(self._h, int(game_id_base)))
            self.cfg.game_id_base = int(game_id_base)
        check(lib().az_reset_games(self._h, _stream()))

    def set_roots(self, tree_ids, cells, plies):
        ids = torch.as_tensor(tree_ids, dtype=torch.int32, device=self.device).contiguous()
        cells = torch.as_tensor(cells, dtype=torch.int8, device=self.device).contiguous()
        plies = torch.as_tensor(plies, dtype=torch.int32, device=self.device).contiguous()
        assert cells.shape == (ids.numel(), self.rules.height, self.rules.width)
        check(lib().az_set_roots(self._h, _ptr(ids), _ptr(cells), _ptr(plies), ids.numel(), _stream()))

    def begin_search(self, sims):
        self.sims_per_move = int(sims)
        check(lib().az_begin_search(self._h, int(sims), _stream()))

    def step(self, priors, values, states_out, leaf_valid_out, leaf_list_out=None, leaf_count_out=None):
        """One lock-step advance (az_step).  priors [T, A] / values [T] float32 or float64 (or None).  With leaf_list_out
        (int32 [T]) and leaf_count_out (int32 [1]) the trees whose leaf awaits evaluation are also listed
        (az_step_gather), for InferenceNet.forward(..., index=, count=)."""
        eval_dtype = native.AZ_F32
        if priors is not None:
            assert priors.is_contiguous() and values.is_contiguous() and priors.dtype == values.dtype
            eval_dtype = {torch.float32: native.AZ_F32, torch.float64: native.AZ_F64}[priors.dtype]
        state_dtype = {torch.bfloat16: native.AZ_BF16, torch.float32: native.AZ_F32}[states_out.dtype]
        assert states_out.is_contiguous() and leaf_valid_out.dtype == torch.int32
        if leaf_list_out is not None:
            assert leaf_list_out.dtype == torch.int32 and leaf_count_out.dtype == torch.int32
            check(lib().az_step_gather(self._h, _ptr(priors), _ptr(values), eval_dtype, _ptr(states_out), state_dtype,
                                       _ptr(leaf_valid_out), _ptr(leaf_list_out), _ptr(leaf_count_out), _stream()))
            return
        check(lib().az_step(self._h, _ptr(priors), _ptr(values), eval_dtype, _ptr(states_out), state_dtype,
                            _ptr(leaf_valid_out), _stream()))

    def search(self):
        check(lib().az_search(self._h, _stream()))

    def extra_sims(self, max_sims):
        check(lib().az_extra_sims(self._h, int(max_sims), _stream()))

    def play(self, greedy=None, move_mode=None):
        g = -1 if greedy is None else int(bool(greedy))
        m = -1 if move_mode is None else {"argmax": 0, "host_uniforms": 1, "philox": 2}[move_mode]
        check(lib().az_play(self._h, g, m, _stream()))

    def fin_clear(self):
        check(lib().az_fin_clear(self._h, _stream()))

    def cache_clear(self):
        check(lib().az_cache_clear(self._h, _stream()))

    # ------------------------------------------------------------------ host-side conveniences
    def set_uniforms(self, uniforms):
        """uniforms [T, <=P] float64: the np.random draws for AZ_MOVE_HOST_UNIFORMS."""
        u = torch.as_tensor(np.asarray(uniforms, dtype=np.float64), device=self.device)
        self.view("uniforms")[:, : u.shape[1]].copy_(u)

    def check_status(self):
        st = self.view("status")
        bad = (st & ~native.AZ_PHASE_MASK).ne(0)
        if bool(bad.any()):
            t = int(torch.nonzero(bad)[0])
            flags = int(st[t]) & ~native.AZ_PHASE_MASK
            names = [n for n, b in (("node pool exhausted", native.AZ_FLAG_POOL_OVERFLOW),
                                    ("pow-half table exceeded", native.AZ_FLAG_LUT_OVERFLOW),
                                    ("illegal action / edgeless root", native.AZ_FLAG_ILLEGAL)) if flags & b]
            raise NativeError(f"tree {t}: " + ", ".join(names))

    def phases(self):
        return self.view("status") & native.AZ_PHASE_MASK

    def totals(self):
        c = self.view("counters").sum(dim=0).tolist()
        return {"sims": c[0], "evals": c[1], "moves": c[2], "games": c[3], "depth_sum": c[4], "children": c[5],
                "reroot_nodes": c[6], "memo_hits": c[7]}

    def drain_finished(self):
        """Copies the finished-game ring to the host and empties it."""
        n = int(self.view("fin_count")[0])
        out = {
            "game_id": self.view("fin_game_id")[:n].cpu().numpy(),
            "len": self.view("fin_len")[:n].cpu().numpy(),
            "result": self.view("fin_result")[:n].cpu().numpy(),
            "visits": self.view("fin_visits")[:n].cpu().numpy(),
            "action": self.view("fin_action")[:n].cpu().numpy(),
            "board": self.view("fin_board")[:n].cpu().numpy().view(np.uint64),
        }
        self.fin_clear()
        return out

    def records(self):
        """Per-tree record of the game in progress (host copies)."""
        return {
            "len": self.view("rec_len").cpu().numpy(),
            "visits": self.view("rec_visits").cpu().numpy(),
            "action": self.view("rec_action").cpu().numpy(),
            "board": self.view("rec_board").cpu().numpy().view(np.uint64),
            "ply": self.view("ply").cpu().numpy(),
            "result": self.view("result").cpu().numpy(),
        }

    def root_stats(self, tree=0):
        """Root edges of one tree: (N, W, P) lists in board move order."""
        w, n, link = self.node_view()
        half = int(self.view("half")[tree])
        root = int(self.view("root_node")[tree])
        lk = int(link[tree, half, root]) & 0xFFFFFFFF
        base, k = lk & 0xFFFFFF, lk >> 24
        p = self.view("node_p")
        sl = slice(base, base + k)
        return (n[tree, half, sl].tolist(), w[tree, half, sl].tolist(), p[tree, half, sl].tolist())
