// az_kernels.cu - kernels and C ABI of libaz_b200.so (see include/az_b200.h).
//
// Kernel inventory (one warp per tree unless noted):
//   k_step    lock-step advance with an external evaluator: consume evaluation (K4 expand + K5 backup),
//             K1 select, K3 encode the leaf for the policy/value net
//   k_search  all remaining simulations of a move with an in-kernel evaluator (uniform / hash)
//   k_play    K6: root policy, edge choice, game record, move on the live board, re-root with
//             breadth-first compaction into the other pool half, game finish + refill
//   k_env_*   K2/K3 standalone, one thread per board (compat Board, unit tests)
// Built for sm_100a only.  There is no CPU implementation behind this ABI.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/az_b200.h"
#include "az_mma.cuh"
#include "az_tree.cuh"

namespace az {

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

int fail_net(int code, const char* msg) { return fail(code, "%s", msg); }

#define AZ_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t err__ = (call);                                                      \
        if (err__ != cudaSuccess) return fail(AZ_ERR_CUDA, #call ": %s", cudaGetErrorString(err__)); \
    } while (0)

// ------------------------------------------------------------------------------------------ header kernels
__global__ void k_reset(Eng e) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        *e.fin_count = 0;
        long long started = e.T < e.games_target ? e.T : e.games_target;
        *e.games_started = (unsigned long long)started;
    }
    if (t >= e.T) return;
    const int nw2 = (e.r.bits > 64 ? 2 : 1) * 2;
    e.status[t] = t < e.games_target ? AZ_PHASE_SEARCH : AZ_PHASE_IDLE;
    e.ply[t] = 0;
    e.game_id[t] = e.game_base + t;
    for (int i = 0; i < nw2; ++i) e.root_board[(size_t)t * nw2 + i] = 0;
    e.half[t] = 0;
    e.root_node[t] = 0;
    e.n_nodes[t] = 1;
    e.sims_done[t] = 0;
    e.pending[t] = 0;
    e.path_len[t] = 0;
    for (int i = 0; i < 8; ++i) e.counters[(size_t)t * 8 + i] = 0;
    NodeA z;
    z.w = 0.0;
    z.n = 0;
    z.link = 0;
    store_node(e.node_a + (size_t)t * 2 * e.C, z);
    e.node_p[(size_t)t * 2 * e.C] = 0.0;
}

// header arrays only k_play / k_set_roots touch
struct Aux {
    int32_t* rec_len;
    int32_t* result;
};

__global__ void k_reset_aux(Aux a, int T) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    a.rec_len[t] = 0;
    a.result[t] = 0;
}

__global__ void k_set_roots(Eng e, Aux aux, const int32_t* ids, const int8_t* cells, const int32_t* plies, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = ids[i];
    if (t < 0 || t >= e.T) return;
    const int nw = e.r.bits > 64 ? 2 : 1;
    uint64_t cur[2] = {0, 0}, opp[2] = {0, 0};
    const int8_t* c = cells + (size_t)i * e.r.cells;
    for (int y = 0; y < e.r.H; ++y)
        for (int x = 0; x < e.r.W; ++x) {
            int v = c[y * e.r.W + x], b = y * e.r.stride + x;
            if (v > 0) cur[b >> 6] |= 1ull << (b & 63);
            if (v < 0) opp[b >> 6] |= 1ull << (b & 63);
        }
    uint64_t* rb = e.root_board + (size_t)t * 2 * nw;
    for (int w = 0; w < nw; ++w) {
        rb[w] = cur[w];
        rb[nw + w] = opp[w];
    }
    e.status[t] = AZ_PHASE_SEARCH;
    e.ply[t] = plies[i];
    e.half[t] = 0;
    e.root_node[t] = 0;
    e.n_nodes[t] = 1;
    e.sims_done[t] = 0;
    e.pending[t] = 0;
    e.path_len[t] = 0;
    aux.rec_len[t] = 0;
    aux.result[t] = 0;
    NodeA z;
    z.w = 0.0;
    z.n = 0;
    z.link = 0;
    store_node(e.node_a + (size_t)t * 2 * e.C, z);
    e.node_p[(size_t)t * 2 * e.C] = 0.0;
}

__global__ void k_begin_search(Eng e) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= e.T) return;
    int st = e.status[t], ph = st & AZ_PHASE_MASK;
    if (ph == AZ_PHASE_SEARCH || ph == AZ_PHASE_READY) {
        e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
        if (!e.pending[t]) e.sims_done[t] = 0;
    }
}

__global__ void k_fin_clear(Eng e) { *e.fin_count = 0; }

// per-tree statistics: a reduction the warp never waits for (no read-modify-write round trip)
__device__ __forceinline__ void bump(long long* p, long long v) {
    if (v) atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

// (RulesView / C4Rules: az_tree.cuh)

// ------------------------------------------------------------------------------------------ k_play
// rank (edge index) of a legal action among the moves in board order
template <int NW, class R>
__device__ __forceinline__ int edge_of_action(const R& r, const BB<NW>& legal, int a) {
    if (r.gravity) return __popcll(legal.w[0] & ((1ull << a) - 1ull));
    int x = a / r.H, y = a - x * r.H, bit = y * r.stride + x;
    if (NW == 2 && bit >= 64) return popc64(legal.w[0]) + __popcll(legal.w[NW - 1] & ((1ull << (bit - 64)) - 1ull));
    return __popcll(legal.w[0] & ((1ull << bit) - 1ull));
}

template <int NW>
__device__ __forceinline__ void finish_game(const Eng& e, const Aux& aux, int t, int st, int lane) {
    // finished-game ring slot (self_play.py:66-78 hands the game's arrays to the caller)
    int slot = 0;
    if (lane == 0) {
        slot = atomicAdd(e.fin_count, 1);
        if (slot >= e.F) {
            atomicSub(e.fin_count, 1);
            slot = -1;
        }
    }
    slot = __shfl_sync(kFull, slot, 0);
    if (slot < 0) {
        if (lane == 0) e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_STALLED;
        return;
    }
    const int len = aux.rec_len[t], A = e.r.A;
    for (int i = lane; i < len * A; i += 32)
        e.fin_visits[(size_t)slot * e.P * A + i] = e.rec_visits[(size_t)t * e.P * A + i];
    for (int i = lane; i < len; i += 32) e.fin_action[(size_t)slot * e.P + i] = e.rec_action[(size_t)t * e.P + i];
    for (int i = lane; i < len * 2 * NW; i += 32)
        e.fin_board[(size_t)slot * e.P * 2 * NW + i] = e.rec_board[(size_t)t * e.P * 2 * NW + i];
    unsigned long long g = 0;
    if (lane == 0) {
        e.fin_game_id[slot] = e.game_id[t];
        e.fin_len[slot] = len;
        e.fin_result[slot] = aux.result[t];
        e.counters[(size_t)t * 8 + 3] += 1;
        if (e.auto_restart) g = atomicAdd(e.games_started, 1ull);
    }
    g = __shfl_sync(kFull, g, 0);
    if (e.auto_restart && (long long)g < e.games_target) {  // next game in this slot: Board() + fresh MCTS
        if (lane < 2 * NW) e.root_board[(size_t)t * 2 * NW + lane] = 0;
        if (lane == 0) {
            e.game_id[t] = e.game_base + (long long)g;
            e.ply[t] = 0;
            aux.rec_len[t] = 0;
            e.n_nodes[t] = 1;
            e.root_node[t] = 0;
            e.sims_done[t] = 0;
            e.pending[t] = 0;
            NodeA z;
            z.w = 0.0;
            z.n = 0;
            z.link = 0;
            store_node(e.node_a + ((size_t)t * 2 + e.half[t]) * e.C, z);
            e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
        }
    } else if (lane == 0) {
        e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_IDLE;
    }
}

// K6 for one tree in AZ_PHASE_READY (MCTS.play, mcts.py:182-222): root policy from the visit counts, edge
// choice, game record, move on the live board, re-root.  Called by az_play for every READY tree and, with
// inline_play, by az_step itself the moment a tree's budget is spent (then only when the re-root fits in place).
template <int NW, int KC, class R>
__device__ __forceinline__ bool play_tree(const Eng& e, const Aux& aux, const R& r, int t, int st, WarpScratch& ws,
                                          int lane, int greedy_override, int move_mode, bool allow_compact) {
    const int h = e.half[t];
    const size_t pool = ((size_t)t * 2 + h) * e.C, pool2 = ((size_t)t * 2 + (h ^ 1)) * e.C;
    NodeA* As = e.node_a + pool;
    double* Ps = e.node_p + pool;
    NodeA* Ad = e.node_a + pool2;
    double* Pd = e.node_p + pool2;
    const int root = e.root_node[t];
    const uint32_t rlink = load_node(As + root).link;
    const int base = (int)(rlink & 0xffffffu), k = (int)(rlink >> 24);
    const int ply = e.ply[t], rec = aux.rec_len[t];
    const int used = e.n_nodes[t];
    const bool fits = (long long)used + (long long)(e.sims_target + 1) * ((r.A + 7) & ~7) + 8 <= (long long)e.C;
    if (!fits && !allow_compact) return false;  // leave the tree READY for az_play
    if (k == 0 || rec >= e.P) {  // the reference would raise on an edgeless root (np.argmax of [])
        if (lane == 0) e.status[t] = st | AZ_FLAG_ILLEGAL;
        return true;
    }
    // root visit counts (mcts.py:189-197)
    for (int j = lane; j < k; j += 32) ws.sel[j] = (double)load_node(As + base + j).n;
    __syncwarp();
    int am = 0;
    {
        double bv = ws.sel[0];
        for (int j = 1; j < k; ++j)
            if (ws.sel[j] > bv) {
                bv = ws.sel[j];
                am = j;
            }
    }
    const bool greedy = greedy_override >= 0 ? greedy_override != 0 : ply >= e.greedy_idx;  // self_play.py:62
    int pick = am;  // deterministic=True, and every greedy draw (one-hot policy)
    if (move_mode != AZ_MOVE_ARGMAX && !greedy) {
        // np.random.choice(edges, 1, p=pi): cdf = cumsum(pi); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
        const double u = move_mode == AZ_MOVE_HOST_UNIFORMS ? e.uniforms[(size_t)t * e.P + rec]
                                                            : philox_uniform(e.seed, e.game_id[t], ply);
        double total = 0.0;
        for (int j = 0; j < k; ++j) total = __dadd_rn(total, ws.sel[j]);  // integers: exact in any order
        double last = 0.0;
        for (int j = 0; j < k; ++j) {
            double pj = total == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn(ws.sel[j], total);
            last = j == 0 ? pj : __dadd_rn(last, pj);
        }
        double acc = 0.0;
        pick = k - 1;
        for (int j = 0; j < k; ++j) {
            double pj = total == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn(ws.sel[j], total);
            acc = j == 0 ? pj : __dadd_rn(acc, pj);
            if (__ddiv_rn(acc, last) > u) {
                pick = j;
                break;
            }
        }
    }
    // record: parent position, visit counts by action, chosen action (self_play.py:63-66)
    Pos<NW> pos = load_pos<NW>(e.root_board + (size_t)t * 2 * NW);
    const BB<NW> legal = legal_set(r, pos);
    for (int a = lane; a < r.A; a += 32) {
        int v = -1;
        if (action_legal(r, pos, legal, a)) v = (int)ws.sel[edge_of_action<NW>(r, legal, a)];
        e.rec_visits[((size_t)t * e.P + rec) * r.A + a] = v;
    }
    store_pos<NW>(e.rec_board + ((size_t)t * e.P + rec) * 2 * NW, pos, lane);
    int bit, action;
    edge_move(r, pos, legal, pick, bit, action);
    const int term = place(r, pos, bit);  // mcts.py:205
    store_pos<NW>(e.root_board + (size_t)t * 2 * NW, pos, lane);
    if (lane == 0) {
        e.rec_action[(size_t)t * e.P + rec] = action | (greedy ? 1 << 16 : 0);
        aux.rec_len[t] = rec + 1;
        e.ply[t] = ply + 1;
        e.counters[(size_t)t * 8 + 2] += 1;
        e.sims_done[t] = 0;
    }
    __syncwarp();
    if (term) {
        if (lane == 0) aux.result[t] = term == 1 ? 1 : 0;  // board.py:258-268 with keep_same_player
        __syncwarp();
        finish_game<NW>(e, aux, t, st, lane);
        return true;
    }
    // re-root to the chosen child, keeping its subtree (mcts.py:207).  In place while this half still
    // has room for a whole search (sims_target expansions of at most A children each) ...
    if (fits) {
        if (lane == 0) {
            e.root_node[t] = base + pick;
            e.pending[t] = 0;
            e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
        }
        return true;
    }
    // ... otherwise the kept subtree is copied breadth-first into the other half: 32 queue nodes per
    // wave, children re-based with a warp prefix sum, dead siblings left behind.
    {
        NodeA z;
        z.w = 0.0;
        z.n = 0;
        z.link = 0;
        if (lane == 0) {
            store_node(Ad, load_node(As + base + pick));
            Pd[0] = Ps[base + pick];
        } else if (lane < 8) {  // the root's block is padded to 8 slots so that every child block stays aligned
            store_node(Ad + lane, z);
        }
    }
    __syncwarp();
    int n_dst = 8, head = 0;
    while (head < n_dst) {
        const int cnt = min(32, n_dst - head);
        uint32_t lk = 0;
        if (lane < cnt) lk = load_node(Ad + head + lane).link;
        const int kk = (int)(lk >> 24), ob = (int)(lk & 0xffffffu);
        const int kp = (kk + 7) & ~7;  // child blocks are 8-node aligned in the new half too
        int incl = kp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - kp, total = __shfl_sync(kFull, incl, 31);
        if (n_dst + total > e.C) {  // cannot happen when both halves have the same capacity; never write past it
            if (lane == 0) e.status[t] = st | AZ_FLAG_POOL_OVERFLOW;
            return true;
        }
        if (kk) {
            NodeA r2 = load_node(Ad + head + lane);
            r2.link = (uint32_t)(n_dst + excl) | ((uint32_t)kk << 24);
            store_node(Ad + head + lane, r2);
        }
        ws.off[lane] = excl;
        ws.ob[lane] = ob;
        ws.path[lane] = kk;
        __syncwarp();
        for (int idx = lane; idx < total; idx += 32) {
            int lo = 0, hi = 32;  // last lane whose exclusive offset is <= idx
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (ws.off[mid] <= idx) lo = mid; else hi = mid;
            }
            const int j = idx - ws.off[lo];
            if (j < ws.path[lo]) {
                const int src = ws.ob[lo] + j;
                store_node(Ad + n_dst + idx, load_node(As + src));
                Pd[n_dst + idx] = Ps[src];
            } else {  // alignment padding: an edgeless dummy the queue skips over
                NodeA z;
                z.w = 0.0;
                z.n = 0;
                z.link = 0;
                store_node(Ad + n_dst + idx, z);
            }
        }
        n_dst += total;
        head += cnt;
        __syncwarp();
    }
    if (lane == 0) {
        e.counters[(size_t)t * 8 + 6] += n_dst;
        e.half[t] = h ^ 1;
        e.root_node[t] = 0;
        e.n_nodes[t] = n_dst;
        e.pending[t] = 0;
        e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
    }
    return true;
}

template <int NW, int KC, class R>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_play(Eng e, Aux aux, int greedy_override, int move_mode) {
    __shared__ WarpScratch s_ws[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= e.T) return;
    const auto r = RulesView<R>::get(e);
    const int st = e.status[t];
    const int phase = st & AZ_PHASE_MASK;
    if (phase == AZ_PHASE_STALLED) {
        finish_game<NW>(e, aux, t, st, lane);
        return;
    }
    if (phase != AZ_PHASE_READY) return;
    play_tree<NW, KC>(e, aux, r, t, st, s_ws[warp], lane, greedy_override, move_mode, true);
}

// ------------------------------------------------------------------------------------------ k_step
// One advance of one tree (warp-uniform): consume the evaluation of the pending leaf (K4 + K5), then
// simulate until a leaf needs the evaluator: terminal leaves are finished on the spot, a spent budget is
// turned into a move right here when inline_play allows (K6).  Returns 1 with the leaf position in
// `leaf` when an evaluation is wanted, else 0.
template <int NW, int KC, bool NOISE = false, class R, class PriorFn>
__device__ __forceinline__ int step_tree(const Eng& e, const Aux& aux, const R& r, int t, WarpScratch& ws, int lane,
                                         bool have_eval, int prior_mode, double value, PriorFn prior_of, Pos<NW>& leaf,
                                         int submit = 1) {
    int st = e.status[t];
    if ((st & AZ_PHASE_MASK) != AZ_PHASE_SEARCH) return 0;
    uint32_t flags = 0;
    const size_t pool = ((size_t)t * 2 + e.half[t]) * e.C;
    NodeA* A = e.node_a + pool;
    double* Pr = e.node_p + pool;
    int sims = e.sims_done[t];
    int root = e.root_node[t];
    long long nsim = 0, neval = 0, ndepth = 0, nchild = 0, nhit = 0;
    // pending: 0 = no leaf in flight, 1 = leaf handed to the evaluator, 2 = leaf selected by az_extra_sims but
    // not handed out yet (it is submitted by the next az_step / az_advance_fused, which finds no answer for it)
    const int was_pending = e.pending[t];
    if (was_pending == 1 && have_eval) {
        const int depth = e.path_len[t];
        for (int i = lane; i < depth; i += 32) ws.path[i] = e.path[(size_t)t * kMaxDepth + i];
        const Pos<NW> at = load_pos<NW>(e.leaf_board + (size_t)t * 2 * NW);
        __syncwarp();
        const uint32_t link = expand_leaf<NW>(e, r, A, Pr, at, t, ws, lane, flags, prior_mode, prior_of);
        backup_path(A, root, ws, depth, -value, link, lane);  // mcts.py:175: value seen by the player who moved in
        if (e.cache_meta && prior_mode == AZ_PRIOR_F32)  // memoise float32 evaluations (exact round trip)
            cache_insert<NW>(e, r, at, lane, [&](int a) { return (float)prior_of(a); }, (float)value);
        ++sims;
        ++nsim;
        ++neval;
        nchild += link >> 24;
    }
    int pend = (was_pending == 2 || (was_pending == 1 && !have_eval)) ? 1 : 0;
    if (pend) leaf = load_pos<NW>(e.leaf_board + (size_t)t * 2 * NW);  // not answered yet: hand the same leaf out (again)
    int freed = 0;
    int staged_root = -1;  // node whose child block select_leaf has staged in ws.root_* during this launch (none yet)
    for (;;) {
        if (!pend && sims >= e.sims_target && e.inline_play) {
            // budget spent: play the move right here (K6) and carry on with the first simulation of the
            // next search, unless the re-root needs the compaction path (left to az_play)
            if (lane == 0) {
                e.sims_done[t] = sims;
                e.pending[t] = 0;
            }
            __syncwarp();
            if (!play_tree<NW, KC>(e, aux, r, t, (st & ~AZ_PHASE_MASK) | AZ_PHASE_READY | (int)flags, ws, lane, -1,
                                   e.move_mode, false))
                break;
            __syncwarp();
            st = e.status[t];
            if ((st & AZ_PHASE_MASK) != AZ_PHASE_SEARCH) {  // game over and no further game for this tree
                sims = -1;
                break;
            }
            root = e.root_node[t];
            sims = 0;
            staged_root = -1;  // new root (and possibly the other pool half): the staged block is stale
            if (++freed >= e.max_free) break;
        }
        if (pend || sims >= e.sims_target) break;
        Pos<NW> pos = load_pos<NW>(e.root_board + (size_t)t * 2 * NW);
        int depth, term;
        select_leaf<NW, KC, NOISE>(e, r, A, Pr, root, pos, ws, lane, depth, term, flags, e.game_id[t], e.ply[t], sims,
                                   &staged_root);
        ndepth += depth;
        if (term) {  // mcts.py:179: terminal leaf, result 1 (win of the mover) or 0 (draw)
            backup_path(A, root, ws, depth, term == 1 ? 1.0 : 0.0, 0u, lane, staged_root);
            ++sims;
            ++nsim;
            if (++freed >= e.max_free) break;
            continue;
        }
        if (e.cache_meta) {  // plays_inferences (mcts.py:123-124): this position was evaluated before
            float cv;
            if (cache_lookup<NW>(e, r, pos, lane, ws.cpri, cv)) {
                const float* cp = ws.cpri;
                const uint32_t link = expand_leaf<NW>(e, r, A, Pr, pos, t, ws, lane, flags, AZ_PRIOR_F32,
                                                      [cp](int a) { return (double)cp[a]; });
                backup_path(A, root, ws, depth, -(double)cv, link, lane, staged_root);
                if (depth == 0) staged_root = -1;  // the root itself was expanded: its link changed
                ++sims;
                ++nsim;
                ++nhit;
                nchild += link >> 24;
                if (++freed >= e.max_free) break;
                continue;
            }
        }
        for (int i = lane; i < depth; i += 32) e.path[(size_t)t * kMaxDepth + i] = ws.path[i];
        store_pos<NW>(e.leaf_board + (size_t)t * 2 * NW, pos, lane);
        if (lane == 0) e.path_len[t] = depth;
        leaf = pos;
        pend = 1;
    }
    if (lane == 0) {
        bump(e.counters + (size_t)t * 8 + 0, nsim);
        bump(e.counters + (size_t)t * 8 + 1, neval);
        bump(e.counters + (size_t)t * 8 + 4, ndepth);
        bump(e.counters + (size_t)t * 8 + 5, nchild);
        bump(e.counters + (size_t)t * 8 + 7, nhit);
        if (sims >= 0) {
            e.sims_done[t] = sims;
            e.pending[t] = pend ? submit : 0;
            int ph = (sims >= e.sims_target && !pend) ? AZ_PHASE_READY : AZ_PHASE_SEARCH;
            e.status[t] = (st & ~AZ_PHASE_MASK) | ph | (int)flags;
        } else if (flags) {
            e.status[t] = st | (int)flags;
        }
    }
    return pend;
}

// 7 blocks of 4 warps per SM (<= 72 registers) for the one-word boards: 148 x 28 = 4 144 warp slots hold all 4 096 trees of the
// headline configuration in ONE wave.  At 96 registers (5 blocks, 2 960 slots) the kernel ran 1.38 waves and the SMs were
// busy 45-55 % of its duration (ncu, profiles/ncu_kstep_r2.csv): the second, partial wave doubled the tail.
template <int NW, int KC, class R, bool NOISE = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, (NW == 1 && KC == 1 && !NOISE) ? 7 : 1)
    k_step(Eng e, Aux aux, const void* __restrict__ priors, const void* __restrict__ values, int eval_dtype, void* states,
           int state_dtype, int32_t* leaf_valid, int32_t* leaf_list = nullptr, int32_t* leaf_count = nullptr,
           unsigned long long* timeline = nullptr) {
    __shared__ WarpScratch s_ws[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= e.T) return;
    // optional timeline slot {first start, last end} in globaltimer ns (az_debug_timeline): measurement aid
    auto stamp = [] {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        return t_;
    };
    if (timeline && lane == 0) atomicMin(timeline, stamp());
    WarpScratch& ws = s_ws[warp];
    const auto r = RulesView<R>::get(e);
    Pos<NW> leaf;
    int pend;
    // the dtype of the evaluator output decides the arithmetic of normalize_probabilities, as in the
    // reference: float64 from infer_sample (factory.py:55), float32 from the model (mcts.py:131-137)
    if (priors == nullptr) {
        pend = step_tree<NW, KC, NOISE>(e, aux, r, t, ws, lane, false, AZ_PRIOR_F64, 0.0, [](int) { return 0.0; }, leaf);
    } else if (eval_dtype == AZ_F64) {
        const double* p = static_cast<const double*>(priors) + (size_t)t * r.A;
        pend = step_tree<NW, KC, NOISE>(e, aux, r, t, ws, lane, true, AZ_PRIOR_F64, static_cast<const double*>(values)[t],
                                 [p](int a) { return p[a]; }, leaf);
    } else {
        const float* p = static_cast<const float*>(priors) + (size_t)t * r.A;
        // value.numpy().item() (mcts.py:136): the float32 value widened
        pend = step_tree<NW, KC, NOISE>(e, aux, r, t, ws, lane, true, AZ_PRIOR_F32, (double)static_cast<const float*>(values)[t],
                                 [p](int a) { return (double)p[a]; }, leaf);
    }
    if (pend) {
        if (state_dtype == AZ_BF16)
            encode_state_bf16<NW>(r, leaf, static_cast<__nv_bfloat16*>(states) + (size_t)t * r.cells * 4, lane);
        else
            encode_state_f32<NW>(r, leaf, static_cast<float*>(states) + (size_t)t * r.cells * 4, lane);
    }
    if (lane == 0) {
        leaf_valid[t] = pend;
        // the leaf batch as a dense list of tree indices (az_step_gather): the net then runs on the trees that really
        // have a leaf pending - 3 900 of 4 096 in steady state, which is 9 instead of 10 rounds of the net kernel's tiles
        if (pend && leaf_list) leaf_list[atomicAdd(leaf_count, 1)] = t;
        if (timeline) atomicMax(timeline + 1, stamp());
    }
}

// ------------------------------------------------------------------------------------------ k_advance
// The three per-tree stages that sit between two tower passes, fused into one launch, one warp per tree:
//   A  heads   tower output of the tree's pending leaf -> 1x1 convs (mma.sync) -> Dense/softmax priors, Dense/tanh value
//   B  tree    step_tree: expand + backup with those priors, select the next leaf (moves played in line)
//   C  stem    the new leaf's planes -> Conv3x3(4 -> 128)+BN+ReLU (mma.sync) -> stem_out[t], the tower's input
// Priors, value and the leaf's planes never touch global memory, two kernel boundaries disappear, and warps in
// the latency-bound stage B share their SM with warps issuing the math of A and C.  Same arithmetic as
// az_net_heads / az_step / az_net_stem (tests compare the two routes bit for bit).
constexpr int kAdvWarps = 16;

struct StemParams {
    const float* w;  // [128][4][3][3]
    const float* b;  // [128]
};

template <int NW, int KC, class R, bool NOISE = false>
__global__ void __launch_bounds__(kAdvWarps * 32, 2)
    k_advance(Eng e, Aux aux, const __nv_bfloat16* __restrict__ x, HeadParams hp, StemParams sp,
              __nv_bfloat16* __restrict__ stem_out, int32_t* leaf_valid, unsigned long long* timeline) {
    constexpr int C = 128;
    // optional timeline slot {first start, last end} in globaltimer ns (az_debug_timeline): measurement aid
    auto now = []() {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        return t_;
    };
    if (timeline && threadIdx.x == 0) atomicMin(timeline, now());
    extern __shared__ float s_f[];
    const auto r = RulesView<R>::get(e);
    const int cells = r.cells, A = r.A, PW = r.W + 2;
    const int ps = 2 * cells + 1;
    float* s_pw = s_f;                                           // policy dense [A][ps]
    float* s_vw = s_pw + ((A * ps + 3) & ~3);                    // value dense 1, transposed [cells][256]
    uint32_t* s_sw = reinterpret_cast<uint32_t*>(s_vw + kHidden * cells);  // stem B fragments [4][3][4][2][32]
    float* s_sb = reinterpret_cast<float*>(s_sw + 4 * 3 * 4 * 2 * 32);               // stem bias [128]
    int* s_base = reinterpret_cast<int*>(s_sb + C);                                   // pixel -> padded cell [cells + 16]
    WarpScratch* s_ws = reinterpret_cast<WarpScratch*>(s_base + ((cells + 16 + 3) & ~3));
    {
        const int np4 = (A * ps) >> 2, nv4 = (kHidden * cells) >> 2;
        const float4* gp = reinterpret_cast<const float4*>(hp.policy_w);
        const float4* gv = reinterpret_cast<const float4*>(hp.value1_w);
        for (int i = threadIdx.x; i < np4; i += blockDim.x) reinterpret_cast<float4*>(s_pw)[i] = gp[i];
        for (int i = (np4 << 2) + threadIdx.x; i < A * ps; i += blockDim.x) s_pw[i] = hp.policy_w[i];
        for (int i = threadIdx.x; i < nv4; i += blockDim.x) reinterpret_cast<float4*>(s_vw)[i] = gv[i];
        for (int i = (nv4 << 2) + threadIdx.x; i < kHidden * cells; i += blockDim.x) s_vw[i] = hp.value1_w[i];
        for (int i = threadIdx.x; i < 4 * 3 * 4 * 2 * 32; i += blockDim.x) {  // pre-packed stem B fragments
            const int ln = i & 31, h = (i >> 5) & 1, nt = (i >> 6) & 3, ks = (i >> 8) % 3, q = i / 768;
            const int co = q * 32 + nt * 8 + (ln >> 2), kk = ks * 16 + (ln & 3) * 2 + h * 8, tap = kk >> 2, ci = kk & 3;
            s_sw[i] = tap < 9 ? pack_bf16(sp.w[co * 36 + ci * 9 + tap], sp.w[co * 36 + (ci + 1) * 9 + tap]) : 0u;
        }
        for (int i = threadIdx.x; i < C; i += blockDim.x) s_sb[i] = sp.b[i];
        for (int p = threadIdx.x; p < cells + 16; p += blockDim.x) {
            const int q = p < cells ? p : cells - 1;
            s_base[p] = (q / r.W) * PW + q % r.W;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int t = blockIdx.x * kAdvWarps + warp;
    if (t >= e.T) return;
    WarpScratch& ws = s_ws[warp];
    const int mtiles = (cells + 15) >> 4;

    // ---- A: heads for the pending leaf of this tree
    const bool have_eval = x != nullptr && e.pending[t] == 1 && (e.status[t] & AZ_PHASE_MASK) == AZ_PHASE_SEARCH;
    float pr[4] = {0.f, 0.f, 0.f, 0.f};
    float value = 0.f;
    if (have_eval) {
        float* h = reinterpret_cast<float*>(ws.sel);  // [cells][2] policy planes, then [cells] value plane
        uint32_t bf[C / 16][2];
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k = ks * 16 + t4 * 2 + hh * 8;
                bf[ks][hh] = g < 3 ? pack_bf16(hp.conv_w[g * C + k], hp.conv_w[g * C + k + 1]) : 0u;
            }
        const float cb0 = hp.conv_b[0], cb1 = hp.conv_b[1], cb2 = hp.conv_b[2];
        const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + (size_t)t * cells * C);
        for (int mt = 0; mt < mtiles; ++mt) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const int q0 = r0 < cells ? r0 : cells - 1, q1 = r1 < cells ? r1 : cells - 1;
            uint32_t a[C / 16][4];
#pragma unroll
            for (int ks = 0; ks < C / 16; ++ks) {
                a[ks][0] = xb[q0 * (C / 2) + ks * 8 + t4];
                a[ks][1] = xb[q1 * (C / 2) + ks * 8 + t4];
                a[ks][2] = xb[q0 * (C / 2) + ks * 8 + t4 + 4];
                a[ks][3] = xb[q1 * (C / 2) + ks * 8 + t4 + 4];
            }
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < C / 16; ++ks) mma_16816(acc, a[ks], bf[ks]);
            if (t4 == 0) {
                if (r0 < cells) { h[r0 * 2] = fmaxf(acc[0] + cb0, 0.f); h[r0 * 2 + 1] = fmaxf(acc[1] + cb1, 0.f); }
                if (r1 < cells) { h[r1 * 2] = fmaxf(acc[2] + cb0, 0.f); h[r1 * 2 + 1] = fmaxf(acc[3] + cb1, 0.f); }
            } else if (t4 == 1) {
                if (r0 < cells) h[2 * cells + r0] = fmaxf(acc[0] + cb2, 0.f);
                if (r1 < cells) h[2 * cells + r1] = fmaxf(acc[2] + cb2, 0.f);
            }
        }
        __syncwarp();
        float logit[4], mx = -INFINITY;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int a = lane + 32 * m;
            logit[m] = -INFINITY;
            if (a < A) {
                float acc = hp.policy_b[a];
                const float* wrow = s_pw + a * ps;
                for (int i = 0; i < 2 * cells; ++i) acc = fmaf(h[i], wrow[i], acc);
                logit[m] = acc;
                mx = fmaxf(mx, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            pr[m] = lane + 32 * m < A ? expf(logit[m] - mx) : 0.f;
            sum += pr[m];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
#pragma unroll
        for (int m = 0; m < 4; ++m) pr[m] = pr[m] / sum;
        // Dense(256): the weights sit transposed [cell][256] in shared memory; lane l owns hidden units
        // 4l..4l+3 and 128+4l..128+4l+3, so every cell costs two conflict-free 128-bit loads and 8 FMAs
        float hacc[8];
        {
            const float4 b0 = reinterpret_cast<const float4*>(hp.value1_b)[lane];
            const float4 b1 = reinterpret_cast<const float4*>(hp.value1_b)[32 + lane];
            hacc[0] = b0.x; hacc[1] = b0.y; hacc[2] = b0.z; hacc[3] = b0.w;
            hacc[4] = b1.x; hacc[5] = b1.y; hacc[6] = b1.z; hacc[7] = b1.w;
        }
        const float* hv = h + 2 * cells;
        for (int p = 0; p < cells; ++p) {
            const float hvp = hv[p];
            const float4 w0 = reinterpret_cast<const float4*>(s_vw + p * kHidden)[lane];
            const float4 w1 = reinterpret_cast<const float4*>(s_vw + p * kHidden)[32 + lane];
            hacc[0] = fmaf(hvp, w0.x, hacc[0]); hacc[1] = fmaf(hvp, w0.y, hacc[1]);
            hacc[2] = fmaf(hvp, w0.z, hacc[2]); hacc[3] = fmaf(hvp, w0.w, hacc[3]);
            hacc[4] = fmaf(hvp, w1.x, hacc[4]); hacc[5] = fmaf(hvp, w1.y, hacc[5]);
            hacc[6] = fmaf(hvp, w1.z, hacc[6]); hacc[7] = fmaf(hvp, w1.w, hacc[7]);
        }
        float part = 0.f;
        {
            const float4 v0 = reinterpret_cast<const float4*>(hp.value2_w)[lane];
            const float4 v1 = reinterpret_cast<const float4*>(hp.value2_w)[32 + lane];
            part = fmaf(fmaxf(hacc[0], 0.f), v0.x, part); part = fmaf(fmaxf(hacc[1], 0.f), v0.y, part);
            part = fmaf(fmaxf(hacc[2], 0.f), v0.z, part); part = fmaf(fmaxf(hacc[3], 0.f), v0.w, part);
            part = fmaf(fmaxf(hacc[4], 0.f), v1.x, part); part = fmaf(fmaxf(hacc[5], 0.f), v1.y, part);
            part = fmaf(fmaxf(hacc[6], 0.f), v1.z, part); part = fmaf(fmaxf(hacc[7], 0.f), v1.w, part);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
        value = tanhf(part + hp.value2_b[0]);
        __syncwarp();
    }

    // ---- B: the tree step with the priors in registers (lane l holds actions l, l+32, ...)
    Pos<NW> leaf;
    const float p0 = pr[0], p1 = pr[1], p2 = pr[2], p3 = pr[3];
    const int pend = step_tree<NW, KC, NOISE>(e, aux, r, t, ws, lane, have_eval, AZ_PRIOR_F32, (double)value,
                                       [p0, p1, p2, p3](int a) {
                                           const int m = a >> 5;
                                           return (double)(m == 0 ? p0 : (m == 1 ? p1 : (m == 2 ? p2 : p3)));
                                       },
                                       leaf);
    if (lane == 0) leaf_valid[t] = pend;
    if (!pend) {
        if (timeline && lane == 0) atomicMax(timeline + 1, now());
        return;
    }

    // ---- C: stem of the new leaf
    __syncwarp();
    uint2* board = reinterpret_cast<uint2*>(ws.sel);  // zero-bordered planes [(H+2)][(W+2)] x 4 bf16
    for (int i = lane; i < (r.H + 2) * PW; i += 32) {
        const int py = i / PW, px = i - py * PW;
        uint2 v = make_uint2(0u, 0u);
        if (py >= 1 && py <= r.H && px >= 1 && px <= r.W) {
            const int code = cell_code(r, leaf, (py - 1) * r.W + (px - 1));
            v.x = (code == 0 ? 0x3F80u : 0u) | (code == 1 ? 0x3F800000u : 0u);  // planes: empty, side to move,
            v.y = (code == 2 ? 0x3F80u : 0u) | 0x3F800000u;                      //         opponent, turn (+1)
        }
        board[i] = v;
    }
    __syncwarp();
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(board);
    int toff[3][2], wsel[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int kk = ks * 16 + t4 * 2 + hh * 8;
            int tap = kk >> 2;
            if (tap > 8) tap = 0;
            toff[ks][hh] = (tap / 3) * PW + tap % 3;
            wsel[ks][hh] = (kk & 3) >> 1;
        }
    for (int q = 0; q < 4; ++q) {
        uint32_t bq[3][4][2];
        float bv[4][2];
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) bq[ks][nt][hh] = s_sw[((((q * 3 + ks) * 4 + nt) * 2 + hh) << 5) + lane];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            bv[nt][0] = s_sb[q * 32 + nt * 8 + t4 * 2];
            bv[nt][1] = s_sb[q * 32 + nt * 8 + t4 * 2 + 1];
        }
        __nv_bfloat16* o = stem_out + (size_t)t * cells * C + q * 32 + t4 * 2;
        for (int mt = 0; mt < mtiles; ++mt) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const int c0 = s_base[r0], c1 = s_base[r1];
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                uint32_t a[4];
                a[0] = sw[(c0 + toff[ks][0]) * 2 + wsel[ks][0]];
                a[1] = sw[(c1 + toff[ks][0]) * 2 + wsel[ks][0]];
                a[2] = sw[(c0 + toff[ks][1]) * 2 + wsel[ks][1]];
                a[3] = sw[(c1 + toff[ks][1]) * 2 + wsel[ks][1]];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_16816(acc[nt], a, bq[ks][nt]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                if (r0 < cells)
                    *reinterpret_cast<uint32_t*>(o + (size_t)r0 * C + nt * 8) =
                        pack_bf16(fmaxf(acc[nt][0] + bv[nt][0], 0.f), fmaxf(acc[nt][1] + bv[nt][1], 0.f));
                if (r1 < cells)
                    *reinterpret_cast<uint32_t*>(o + (size_t)r1 * C + nt * 8) =
                        pack_bf16(fmaxf(acc[nt][2] + bv[nt][0], 0.f), fmaxf(acc[nt][3] + bv[nt][1], 0.f));
            }
        }
    }
    if (timeline && lane == 0) atomicMax(timeline + 1, now());
}

// ------------------------------------------------------------------------------------------ k_extra
// Simulations that need no evaluator, for trees without a leaf in flight (their last simulation hit a terminal
// leaf or spent the move budget).  Runs beside the tower on a forked stream: such trees keep finishing
// terminal-leaf simulations and playing moves (up to e.max_free of them) until a leaf does need the net; that
// leaf is parked (pending = 2) and submitted by the next az_advance_fused / az_step.
template <int NW, int KC, class R, bool NOISE = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_extra(Eng e, Aux aux) {
    __shared__ WarpScratch s_ws[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= e.T) return;
    if (e.pending[t] != 0 || (e.status[t] & AZ_PHASE_MASK) != AZ_PHASE_SEARCH) return;
    const auto r = RulesView<R>::get(e);
    Pos<NW> leaf;
    step_tree<NW, KC, NOISE>(e, aux, r, t, s_ws[warp], lane, false, AZ_PRIOR_F64, 0.0, [](int) { return 0.0; }, leaf, 2);
}

// ------------------------------------------------------------------------------------------ k_search
template <int NW, int KC, class R, bool NOISE = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_search(Eng e) {
    __shared__ WarpScratch s_ws[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= e.T) return;
    WarpScratch& ws = s_ws[warp];
    const auto r = RulesView<R>::get(e);
    const int st = e.status[t];
    if ((st & AZ_PHASE_MASK) != AZ_PHASE_SEARCH) return;
    uint32_t flags = 0;
    const size_t pool = ((size_t)t * 2 + e.half[t]) * e.C;
    NodeA* A = e.node_a + pool;
    double* Pr = e.node_p + pool;
    const Pos<NW> root_pos = load_pos<NW>(e.root_board + (size_t)t * 2 * NW);
    const int root = e.root_node[t];
    const double uniform_prior = __ddiv_rn(1.0, (double)r.A);  // np.full(A, 1 / A)
    int sims = e.sims_done[t];
    long long nsim = 0, neval = 0, ndepth = 0, nchild = 0;
    while (sims < e.sims_target && !(flags & AZ_FLAG_POOL_OVERFLOW)) {
        Pos<NW> pos = root_pos;
        int depth, term;
        select_leaf<NW, KC, NOISE>(e, r, A, Pr, root, pos, ws, lane, depth, term, flags, e.game_id[t], e.ply[t], sims);
        ndepth += depth;
        if (term) {
            backup_path(A, root, ws, depth, term == 1 ? 1.0 : 0.0, 0u, lane);
        } else {
            double v = 0.0;
            uint32_t link;
            if (e.eval_mode == AZ_EVAL_HASH) {
                const uint64_t h = hash_position<NW>(r, pos);
                v = hash_value(h);
                link = expand_leaf<NW>(e, r, A, Pr, pos, t, ws, lane, flags, e.prior_mode, [h](int a) { return hash_prior(h, a); });
            } else {
                link = expand_leaf<NW>(e, r, A, Pr, pos, t, ws, lane, flags, e.prior_mode, [uniform_prior](int) { return uniform_prior; });
            }
            backup_path(A, root, ws, depth, -v, link, lane);
            ++neval;
            nchild += link >> 24;
        }
        ++sims;
        ++nsim;
    }
    if (lane == 0) {
        e.sims_done[t] = sims;
        e.counters[(size_t)t * 8 + 0] += nsim;
        e.counters[(size_t)t * 8 + 1] += neval;
        e.counters[(size_t)t * 8 + 4] += ndepth;
        e.counters[(size_t)t * 8 + 5] += nchild;
        int ph = sims >= e.sims_target ? AZ_PHASE_READY : AZ_PHASE_SEARCH;
        e.status[t] = (st & ~AZ_PHASE_MASK) | ph | (int)flags;
    }
}

// ------------------------------------------------------------------------------------------ env kernels
__device__ __forceinline__ Pos<2> pos_from_cells(const Rules& r, const int8_t* c) {
    Pos<2> p;
    p.cur.w[0] = p.cur.w[1] = p.opp.w[0] = p.opp.w[1] = 0;
    for (int y = 0; y < r.H; ++y)
        for (int x = 0; x < r.W; ++x) {
            int v = c[y * r.W + x];
            if (v > 0) bb_set(p.cur, y * r.stride + x);
            if (v < 0) bb_set(p.opp, y * r.stride + x);
        }
    return p;
}

__global__ void k_env_play(Rules r, const int8_t* cells_in, const int32_t* actions, int n, int8_t* cells_out,
                           int32_t* status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int8_t* c = cells_in + (size_t)i * r.cells;
    int8_t* o = cells_out + (size_t)i * r.cells;
    Pos<2> p = pos_from_cells(r, c);
    const BB<2> legal = legal_set(r, p);
    const int a = actions[i];
    if (a < 0 || !action_legal(r, p, legal, a)) {  // board.py:221-224,228: AssertionError in the reference
        for (int j = 0; j < r.cells; ++j) o[j] = c[j];
        status[i] = -1;
        return;
    }
    int bit, action;
    edge_move<false>(r, p, legal, edge_of_action<2>(r, legal, a), bit, action);
    status[i] = place(r, p, bit);
    for (int j = 0; j < r.cells; ++j) {
        int code = cell_code(r, p, j);
        o[j] = code == 1 ? 1 : (code == 2 ? -1 : 0);
    }
}

__global__ void k_env_legal(Rules r, const int8_t* cells, int n, uint8_t* legal_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos<2> p = pos_from_cells(r, cells + (size_t)i * r.cells);
    const BB<2> legal = legal_set(r, p);
    for (int a = 0; a < r.A; ++a) legal_out[(size_t)i * r.A + a] = action_legal(r, p, legal, a) ? 1 : 0;
}

__global__ void k_env_encode(Rules r, const int8_t* cells, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos<2> p = pos_from_cells(r, cells + (size_t)i * r.cells);
    for (int j = 0; j < r.cells; ++j) {
        int code = cell_code(r, p, j);
        reinterpret_cast<float4*>(out + (size_t)i * r.cells * 4)[j] = make_float4(code == 0, code == 1, code == 2, 1.0f);
    }
}

// one block per finished game, one warp per ply (round-robin): states, policy targets, rewards
template <int NW>
__global__ void __launch_bounds__(128) k_decode(Rules r, int P, const uint64_t* boards, const int32_t* visits,
                                                 const int32_t* actions, const int32_t* lens, const int32_t* results,
                                                 const int32_t* offsets, float* states, double* policies,
                                                 int32_t* values) {
    const int g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int len = lens[g], res = results[g], off = offsets[g];
    for (int i = warp; i < len; i += 4) {
        const size_t s = (size_t)off + i;
        const Pos<NW> pos = load_pos<NW>(boards + ((size_t)g * P + i) * 2 * NW);
        encode_state_f32<NW>(r, pos, states + s * r.cells * 4, lane);
        const int32_t* v = visits + ((size_t)g * P + i) * r.A;
        const bool greedy = (actions[(size_t)g * P + i] >> 16) & 1;
        // sum, number of legal actions and first maximum in EDGE (board move) order
        int sum = 0, k = 0, bestn = -1, bestkey = 0x7fffffff;
        for (int a = lane; a < r.A; a += 32) {
            int n = v[a];
            if (n < 0) continue;
            sum += n;
            ++k;
            int key = r.gravity ? a : (a % r.H) * r.W + a / r.H;  // row-major rank of the cell
            if (n > bestn || (n == bestn && key < bestkey)) {
                bestn = n;
                bestkey = key;
            }
        }
        sum = __reduce_add_sync(kFull, sum);
        k = __reduce_add_sync(kFull, k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            int on = __shfl_xor_sync(kFull, bestn, o), ok = __shfl_xor_sync(kFull, bestkey, o);
            if (on > bestn || (on == bestn && ok < bestkey)) {
                bestn = on;
                bestkey = ok;
            }
        }
        for (int a = lane; a < r.A; a += 32) {
            int n = v[a];
            double p = 0.0;
            if (n >= 0) {
                int key = r.gravity ? a : (a % r.H) * r.W + a / r.H;
                if (greedy)
                    p = key == bestkey ? 1.0 : 0.0;
                else
                    p = sum == 0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn((double)n, (double)sum);
            }
            policies[s * r.A + a] = p;
        }
        if (lane == 0) values[s] = ((len - 1 - i) & 1) ? -res : res;
    }
}

__global__ void k_debug_dirichlet(uint64_t seed, double alpha, int k, int n, double* out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        if (j < k) out[(size_t)i * k + j] = gamma_variate(seed ^ 0xD1B54A32D192ED03ull, (uint32_t)i, 0u, 0u, (uint32_t)j, alpha);
    }
    __syncwarp();
    double s = 0.0;
    for (int j = lane; j < k; j += 32) s += out[(size_t)i * k + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    for (int j = lane; j < k; j += 32) out[(size_t)i * k + j] /= s;
}

}  // namespace az

// ========================================================================================== C ABI
using namespace az;

struct az_engine {
    az_config cfg;
    az_layout lay;
    Eng eng;
    Aux aux;
    int nw, kc;
    bool c4;  // headline configuration: 6x7, connect 4, gravity -> compile-time rules
    unsigned long long* timeline = nullptr;  // az_debug_timeline
    int timeline_slots = 0, timeline_next = 0;
};

// the engine's device view for az_net_forward_trees (az_tower.cu); 1 = plain 6x7 connect-4 fast path (compile-time rules,
// one child chunk, no root noise), 0 = anything else
int az::engine_view(const az_engine* e, az::Eng* out) {
    *out = e->eng;
    return (e->c4 && e->eng.dirichlet == 0) ? 1 : 0;
}

static int check_cfg(const az_config* c) {
    if (!c) return fail(AZ_ERR_ARG, "null config%s");
    if (c->abi_version != AZ_ABI_VERSION) return fail(AZ_ERR_ARG, "abi_version mismatch%s");
    if (c->width < 2 || c->height < 2 || c->width > kMaxDim || c->height > kMaxDim)
        return fail(AZ_ERR_ARG, "board dimensions out of range%s");
    if (c->height * (c->width + 1) > 128) return fail(AZ_ERR_ARG, "board needs more than 128 bits%s");
    int m = c->width < c->height ? c->width : c->height;
    if (c->n_connect < 2 || c->n_connect > m) return fail(AZ_ERR_ARG, "n_connect must be in [2, min(W, H)]%s");
    if (c->eval_cache_log2 < 0 || c->eval_cache_log2 > 30) return fail(AZ_ERR_ARG, "eval_cache_log2 must be in [0, 30]%s");
    if (c->dirichlet_noise && (!(c->dirichlet_alpha > 0.0) || c->dirichlet_ratio < 0.0 || c->dirichlet_ratio > 1.0))
        return fail(AZ_ERR_ARG, "dirichlet_alpha must be > 0 and dirichlet_ratio in [0, 1]%s");
    return AZ_OK;
}

static size_t take(size_t& off, size_t bytes) {
    size_t at = (off + 255) & ~(size_t)255;
    off = at + bytes;
    return at;
}

#define AZ_API extern "C" __attribute__((visibility("default")))
AZ_API const char* az_last_error(void) { return g_err; }
AZ_API int az_abi_version(void) { return AZ_ABI_VERSION; }
AZ_API void az_struct_sizes(size_t* c, size_t* l) {
    if (c) *c = sizeof(az_config);
    if (l) *l = sizeof(az_layout);
}

AZ_API int az_query_layout(const az_config* c, az_layout* L) {
    if (int rc = check_cfg(c)) return rc;
    if (!L) return fail(AZ_ERR_ARG, "null layout%s");
    if (c->n_trees < 1 || c->node_capacity < 2 || c->node_capacity > 0xffffff || c->fin_capacity < 1 ||
        c->pow_lut_len < 2 || c->sims_per_move < 1 || c->max_free_sims < 1)
        return fail(AZ_ERR_ARG, "n_trees / node_capacity / fin_capacity / pow_lut_len / sims_per_move out of range%s");
    memset(L, 0, sizeof(*L));
    const size_t T = c->n_trees, C = c->node_capacity, F = c->fin_capacity;
    const size_t A = c->gravity ? c->width : c->width * c->height, P = (size_t)c->width * c->height;
    const size_t WD = c->height * (c->width + 1) > 64 ? 2 : 1;
    L->n_actions = (int)A;
    L->max_plies = (int)P;
    L->words = (int)WD;
    L->max_depth = kMaxDepth;
    size_t off = 0;
    L->status = take(off, 4 * T);
    L->ply = take(off, 4 * T);
    L->game_id = take(off, 8 * T);
    L->root_board = take(off, 8 * T * 2 * WD);
    L->half = take(off, 4 * T);
    L->root_node = take(off, 4 * T);
    L->n_nodes = take(off, 4 * T);
    L->sims_done = take(off, 4 * T);
    L->pending = take(off, 4 * T);
    L->path_len = take(off, 4 * T);
    L->path = take(off, 4 * T * kMaxDepth);
    L->leaf_board = take(off, 8 * T * 2 * WD);
    L->counters = take(off, 8 * T * 8);
    L->uniforms = take(off, 8 * T * P);
    L->node_a = take(off, 16 * T * 2 * C);
    L->node_p = take(off, 8 * T * 2 * C);
    L->rec_visits = take(off, 4 * T * P * A);
    L->rec_action = take(off, 4 * T * P);
    L->rec_board = take(off, 8 * T * P * 2 * WD);
    L->fin_count = take(off, 16);
    L->fin_game_id = take(off, 8 * F);
    L->fin_len = take(off, 4 * F);
    L->fin_result = take(off, 4 * F);
    L->fin_visits = take(off, 4 * F * P * A);
    L->fin_action = take(off, 4 * F * P);
    L->fin_board = take(off, 8 * F * P * 2 * WD);
    L->pow_lut = take(off, 8 * (size_t)c->pow_lut_len);
    L->rec_len = take(off, 4 * T);
    L->result = take(off, 4 * T);
    if (c->eval_cache_log2 > 0) {
        const size_t S = (size_t)1 << c->eval_cache_log2;
        L->cache_meta = take(off, 4 * S);
        L->cache_key = take(off, 8 * S * 2 * WD);
        L->cache_val = take(off, 4 * S * (A + 1));
    }
    L->total_bytes = (off + 255) & ~(size_t)255;
    return AZ_OK;
}

AZ_API int az_engine_create(const az_config* c, void* slab, size_t bytes, const double* host_lut, void* stream,
                                az_engine** out) {
    if (!out) return fail(AZ_ERR_ARG, "null out%s");
    *out = nullptr;
    az_layout L;
    if (int rc = az_query_layout(c, &L)) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback%s");
    }
    if (!slab || !host_lut) return fail(AZ_ERR_ARG, "null slab / pow table%s");
    if (bytes < L.total_bytes || (reinterpret_cast<uintptr_t>(slab) & 255)) return fail(AZ_ERR_SLAB, "slab too small or not 256-byte aligned%s");
    az_engine* e = new (std::nothrow) az_engine();
    if (!e) return fail(AZ_ERR_ARG, "out of host memory%s");
    e->cfg = *c;
    e->lay = L;
    char* b = static_cast<char*>(slab);
    Eng& g = e->eng;
    g.r = make_rules(c->width, c->height, c->n_connect, c->gravity ? 1 : 0);
    g.T = c->n_trees;
    g.C = c->node_capacity;
    g.P = L.max_plies;
    g.F = c->fin_capacity;
    g.sims_target = c->sims_per_move;
    g.greedy_idx = c->index_move_greedy;
    g.eval_mode = c->eval_mode;
    g.prior_mode = c->prior_mode;
    g.move_mode = c->move_mode;
    g.max_free = c->max_free_sims;
    g.lut_len = c->pow_lut_len;
    g.auto_restart = c->auto_restart;
    g.inline_play = c->inline_play;
    g.dirichlet = c->dirichlet_noise;
    g.dir_alpha = c->dirichlet_alpha;
    g.dir_ratio = c->dirichlet_ratio;
    g.c_puct = c->c_puct;
    g.seed = c->seed;
    g.game_base = c->game_id_base;
    {   // NS1 (north_star): root child block staged in shared memory across the simulations of a launch; AZ_ROOT_SMEM=0 off
        const char* rs = getenv("AZ_ROOT_SMEM");
        g.root_smem = rs ? (atoi(rs) != 0) : 1;
    }
    g.games_target = c->games_target;
    g.status = reinterpret_cast<int32_t*>(b + L.status);
    g.ply = reinterpret_cast<int32_t*>(b + L.ply);
    g.game_id = reinterpret_cast<long long*>(b + L.game_id);
    g.root_board = reinterpret_cast<uint64_t*>(b + L.root_board);
    g.half = reinterpret_cast<int32_t*>(b + L.half);
    g.root_node = reinterpret_cast<int32_t*>(b + L.root_node);
    g.n_nodes = reinterpret_cast<int32_t*>(b + L.n_nodes);
    g.sims_done = reinterpret_cast<int32_t*>(b + L.sims_done);
    g.pending = reinterpret_cast<int32_t*>(b + L.pending);
    g.path_len = reinterpret_cast<int32_t*>(b + L.path_len);
    g.path = reinterpret_cast<int32_t*>(b + L.path);
    g.leaf_board = reinterpret_cast<uint64_t*>(b + L.leaf_board);
    g.counters = reinterpret_cast<long long*>(b + L.counters);
    g.uniforms = reinterpret_cast<double*>(b + L.uniforms);
    g.node_a = reinterpret_cast<NodeA*>(b + L.node_a);
    g.node_p = reinterpret_cast<double*>(b + L.node_p);
    g.rec_visits = reinterpret_cast<int32_t*>(b + L.rec_visits);
    g.rec_action = reinterpret_cast<int32_t*>(b + L.rec_action);
    g.rec_board = reinterpret_cast<uint64_t*>(b + L.rec_board);
    g.fin_count = reinterpret_cast<int32_t*>(b + L.fin_count);
    g.games_started = reinterpret_cast<unsigned long long*>(b + L.fin_count + 8);
    g.fin_game_id = reinterpret_cast<long long*>(b + L.fin_game_id);
    g.fin_len = reinterpret_cast<int32_t*>(b + L.fin_len);
    g.fin_result = reinterpret_cast<int32_t*>(b + L.fin_result);
    g.fin_visits = reinterpret_cast<int32_t*>(b + L.fin_visits);
    g.fin_action = reinterpret_cast<int32_t*>(b + L.fin_action);
    g.fin_board = reinterpret_cast<uint64_t*>(b + L.fin_board);
    g.pow_lut = reinterpret_cast<const double*>(b + L.pow_lut);
    g.cache_meta = nullptr;
    g.cache_key = nullptr;
    g.cache_val = nullptr;
    g.cache_mask = 0;
    if (c->eval_cache_log2 > 0) {
        g.cache_meta = reinterpret_cast<unsigned int*>(b + L.cache_meta);
        g.cache_key = reinterpret_cast<uint64_t*>(b + L.cache_key);
        g.cache_val = reinterpret_cast<float*>(b + L.cache_val);
        g.cache_mask = (1u << c->eval_cache_log2) - 1u;
    }
    e->aux.rec_len = reinterpret_cast<int32_t*>(b + L.rec_len);
    e->aux.result = reinterpret_cast<int32_t*>(b + L.result);
    e->nw = L.words;
    e->kc = L.n_actions <= 32 ? 1 : 4;
    e->c4 = c->width == 7 && c->height == 6 && c->n_connect == 4 && c->gravity != 0;
    cudaError_t err = cudaMemcpyAsync(b + L.pow_lut, host_lut, 8 * (size_t)c->pow_lut_len, cudaMemcpyHostToDevice,
                                      static_cast<cudaStream_t>(stream));
    if (err == cudaSuccess) err = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));  // host_lut may be freed
    if (err != cudaSuccess) {
        delete e;
        return fail(AZ_ERR_CUDA, "pow table upload: %s", cudaGetErrorString(err));
    }
    *out = e;
    return az_reset_games(e, stream);
}

AZ_API void az_engine_destroy(az_engine* e) { delete e; }

static inline dim3 tree_grid(const az_engine* e) { return dim3((e->eng.T + kWarpsPerBlock - 1) / kWarpsPerBlock); }
static inline dim3 flat_grid(int n) { return dim3((n + 127) / 128); }

AZ_API int az_reset_games(az_engine* e, void* stream) {
    if (!e) return fail(AZ_ERR_ARG, "null engine%s");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    e->eng.sims_target = e->cfg.sims_per_move;
    k_reset<<<flat_grid(e->eng.T), 128, 0, s>>>(e->eng);
    k_reset_aux<<<flat_grid(e->eng.T), 128, 0, s>>>(e->aux, e->eng.T);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_set_game_id_base(az_engine* e, int64_t game_id_base) {
    if (!e || game_id_base < 0) return fail(AZ_ERR_ARG, "az_set_game_id_base: bad argument%s");
    e->cfg.game_id_base = game_id_base;
    e->eng.game_base = game_id_base;
    return AZ_OK;
}

AZ_API int az_set_roots(az_engine* e, const int32_t* ids, const int8_t* cells, const int32_t* plies, int32_t n,
                            void* stream) {
    if (!e || n < 0) return fail(AZ_ERR_ARG, "az_set_roots: bad argument%s");
    if (n == 0) return AZ_OK;
    if (!ids || !cells || !plies) return fail(AZ_ERR_ARG, "az_set_roots: null pointer%s");
    k_set_roots<<<flat_grid(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(e->eng, e->aux, ids, cells, plies, n);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_begin_search(az_engine* e, int32_t sims, void* stream) {
    if (!e || sims < 1) return fail(AZ_ERR_ARG, "az_begin_search: bad argument%s");
    e->eng.sims_target = sims;
    k_begin_search<<<flat_grid(e->eng.T), 128, 0, static_cast<cudaStream_t>(stream)>>>(e->eng);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

// Kernel variants: compile-time rules for the headline board; generic rules by bitboard words (NW) and children
// chunks (KC); root-noise variants (generic rules only: the sampler would cost the fast path its registers).
#define AZ_DISPATCH(kernel, ...)                                                                          \
    do {                                                                                                  \
        cudaStream_t s__ = static_cast<cudaStream_t>(stream);                                             \
        dim3 g__ = tree_grid(e), b__(kWarpsPerBlock * 32);                                                \
        const bool nz__ = e->eng.dirichlet != 0;                                                          \
        if (e->c4 && !nz__) kernel<1, 1, C4Rules><<<g__, b__, 0, s__>>>(__VA_ARGS__);                     \
        else if (e->nw == 1 && e->kc == 1 && !nz__) kernel<1, 1, Rules><<<g__, b__, 0, s__>>>(__VA_ARGS__); \
        else if (e->nw == 1 && e->kc == 1) kernel<1, 1, Rules, true><<<g__, b__, 0, s__>>>(__VA_ARGS__);  \
        else if (e->nw == 1 && !nz__) kernel<1, 4, Rules><<<g__, b__, 0, s__>>>(__VA_ARGS__);             \
        else if (e->nw == 1) kernel<1, 4, Rules, true><<<g__, b__, 0, s__>>>(__VA_ARGS__);                \
        else if (e->kc == 1 && !nz__) kernel<2, 1, Rules><<<g__, b__, 0, s__>>>(__VA_ARGS__);             \
        else if (e->kc == 1) kernel<2, 1, Rules, true><<<g__, b__, 0, s__>>>(__VA_ARGS__);                \
        else if (!nz__) kernel<2, 4, Rules><<<g__, b__, 0, s__>>>(__VA_ARGS__);                           \
        else kernel<2, 4, Rules, true><<<g__, b__, 0, s__>>>(__VA_ARGS__);                                \
        AZ_CUDA(cudaGetLastError());                                                                      \
    } while (0)

static unsigned long long* next_timeline_slot(az_engine* e) {
    if (!e->timeline || e->timeline_slots <= 0) return nullptr;
    unsigned long long* tl = e->timeline + 2 * (e->timeline_next % e->timeline_slots);
    e->timeline_next += 1;
    return tl;
}

AZ_API int az_step(az_engine* e, const void* priors, const void* values, int32_t eval_dtype, void* states,
                       int32_t state_dtype, int32_t* leaf_valid, void* stream) {
    if (!e || !states || !leaf_valid) return fail(AZ_ERR_ARG, "az_step: null pointer%s");
    if ((priors == nullptr) != (values == nullptr)) return fail(AZ_ERR_ARG, "az_step: priors and values go together%s");
    if ((eval_dtype != AZ_F32 && eval_dtype != AZ_F64) || (state_dtype != AZ_BF16 && state_dtype != AZ_F32))
        return fail(AZ_ERR_ARG, "az_step: unsupported dtype%s");
    AZ_DISPATCH(k_step, e->eng, e->aux, priors, values, eval_dtype, states, state_dtype, leaf_valid, nullptr, nullptr,
                next_timeline_slot(e));
    return AZ_OK;
}

AZ_API int az_step_gather(az_engine* e, const void* priors, const void* values, int32_t eval_dtype, void* states,
                          int32_t state_dtype, int32_t* leaf_valid, int32_t* leaf_list, int32_t* leaf_count, void* stream) {
    if (!e || !states || !leaf_valid || !leaf_list || !leaf_count) return fail(AZ_ERR_ARG, "az_step_gather: null pointer%s");
    if ((priors == nullptr) != (values == nullptr)) return fail(AZ_ERR_ARG, "az_step_gather: priors and values go together%s");
    if ((eval_dtype != AZ_F32 && eval_dtype != AZ_F64) || (state_dtype != AZ_BF16 && state_dtype != AZ_F32))
        return fail(AZ_ERR_ARG, "az_step_gather: unsupported dtype%s");
    AZ_CUDA(cudaMemsetAsync(leaf_count, 0, sizeof(int32_t), static_cast<cudaStream_t>(stream)));
    AZ_DISPATCH(k_step, e->eng, e->aux, priors, values, eval_dtype, states, state_dtype, leaf_valid, leaf_list, leaf_count,
                next_timeline_slot(e));
    return AZ_OK;
}

AZ_API int az_extra_sims(az_engine* e, int32_t max_sims, void* stream) {
    if (!e || max_sims < 1) return fail(AZ_ERR_ARG, "az_extra_sims: bad argument%s");
    Eng g = e->eng;
    g.max_free = max_sims;
    AZ_DISPATCH(k_extra, g, e->aux);
    return AZ_OK;
}

AZ_API int az_search(az_engine* e, void* stream) {
    if (!e) return fail(AZ_ERR_ARG, "null engine%s");
    if (e->eng.eval_mode != AZ_EVAL_UNIFORM && e->eng.eval_mode != AZ_EVAL_HASH)
        return fail(AZ_ERR_ARG, "az_search needs an in-kernel evaluator (eval_mode UNIFORM or HASH)%s");
    AZ_DISPATCH(k_search, e->eng);
    return AZ_OK;
}

AZ_API int az_play(az_engine* e, int32_t greedy_override, int32_t move_mode_override, void* stream) {
    if (!e) return fail(AZ_ERR_ARG, "null engine%s");
    int mode = move_mode_override >= 0 ? move_mode_override : e->eng.move_mode;
    if (mode < AZ_MOVE_ARGMAX || mode > AZ_MOVE_PHILOX) return fail(AZ_ERR_ARG, "az_play: bad move mode%s");
    {
        cudaStream_t s__ = static_cast<cudaStream_t>(stream);
        dim3 g__ = tree_grid(e), b__(kWarpsPerBlock * 32);
        if (e->c4) k_play<1, 1, C4Rules><<<g__, b__, 0, s__>>>(e->eng, e->aux, greedy_override, mode);
        else if (e->nw == 1 && e->kc == 1) k_play<1, 1, Rules><<<g__, b__, 0, s__>>>(e->eng, e->aux, greedy_override, mode);
        else if (e->nw == 1) k_play<1, 4, Rules><<<g__, b__, 0, s__>>>(e->eng, e->aux, greedy_override, mode);
        else if (e->kc == 1) k_play<2, 1, Rules><<<g__, b__, 0, s__>>>(e->eng, e->aux, greedy_override, mode);
        else k_play<2, 4, Rules><<<g__, b__, 0, s__>>>(e->eng, e->aux, greedy_override, mode);
        AZ_CUDA(cudaGetLastError());
    }
    return AZ_OK;
}

AZ_API int az_fin_clear(az_engine* e, void* stream) {
    if (!e) return fail(AZ_ERR_ARG, "null engine%s");
    k_fin_clear<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(e->eng);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

static int env_rules(const az_config* c, Rules* r) {
    if (int rc = check_cfg(c)) return rc;
    *r = make_rules(c->width, c->height, c->n_connect, c->gravity ? 1 : 0);
    return AZ_OK;
}

AZ_API int az_env_play(const az_config* c, const int8_t* cells_in, const int32_t* actions, int32_t n,
                           int8_t* cells_out, int32_t* status, void* stream) {
    Rules r;
    if (int rc = env_rules(c, &r)) return rc;
    if (n == 0) return AZ_OK;  // an empty batch has no buffers to point at
    if (!cells_in || !actions || !cells_out || !status || n < 0) return fail(AZ_ERR_ARG, "az_env_play: bad argument%s");
    k_env_play<<<flat_grid(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(r, cells_in, actions, n, cells_out, status);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_env_legal(const az_config* c, const int8_t* cells, int32_t n, uint8_t* legal, void* stream) {
    Rules r;
    if (int rc = env_rules(c, &r)) return rc;
    if (n == 0) return AZ_OK;
    if (!cells || !legal || n < 0) return fail(AZ_ERR_ARG, "az_env_legal: bad argument%s");
    k_env_legal<<<flat_grid(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(r, cells, n, legal);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_env_encode(const az_config* c, const int8_t* cells, int32_t n, float* states, void* stream) {
    Rules r;
    if (int rc = env_rules(c, &r)) return rc;
    if (n == 0) return AZ_OK;
    if (!cells || !states || n < 0) return fail(AZ_ERR_ARG, "az_env_encode: bad argument%s");
    k_env_encode<<<flat_grid(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(r, cells, n, states);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_decode_samples(const az_config* c, const uint64_t* boards, const int32_t* visits, const int32_t* actions,
                             const int32_t* lens, const int32_t* results, const int32_t* offsets, int32_t n_games,
                             float* states, double* policies, int32_t* values, void* stream) {
    Rules r;
    if (int rc = env_rules(c, &r)) return rc;
    if (n_games == 0) return AZ_OK;
    if (!boards || !visits || !actions || !lens || !results || !offsets || n_games < 0)
        return fail(AZ_ERR_ARG, "az_decode_samples: bad argument%s");
    // states / policies / values may be null only if every game is empty; the kernel then writes nothing
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (r.bits > 64)
        k_decode<2><<<n_games, 128, 0, s>>>(r, r.cells, boards, visits, actions, lens, results, offsets, states, policies, values);
    else
        k_decode<1><<<n_games, 128, 0, s>>>(r, r.cells, boards, visits, actions, lens, results, offsets, states, policies, values);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_advance_fused(az_engine* e, const void* tower_out, const az_head_weights* hw, const float* stem_w,
                            const float* stem_b, void* stem_out, int32_t* leaf_valid, void* stream) {
    if (!e || !hw || !stem_w || !stem_b || !stem_out || !leaf_valid) return fail(AZ_ERR_ARG, "az_advance_fused: null pointer%s");
    const Rules& r = e->eng.r;
    const size_t A = r.A, cells = r.cells;
    const size_t smem = sizeof(float) * (((A * (2 * cells + 1) + 3) & ~(size_t)3) + (size_t)kHidden * cells) +
                        sizeof(uint32_t) * 4 * 3 * 4 * 2 * 32 + sizeof(float) * 128 + sizeof(int) * ((cells + 16 + 3) & ~(size_t)3) +
                        sizeof(WarpScratch) * kAdvWarps;
    if (smem > 227 * 1024) return fail(AZ_ERR_ARG, "az_advance_fused: head weights do not fit in shared memory for this board%s");
    HeadParams hp{hw->conv_w, hw->conv_b, hw->policy_w, hw->policy_b, hw->value1_w, hw->value1_b, hw->value2_w, hw->value2_b,
                  e->eng.T, (int)cells, (int)A};
    StemParams sp{stem_w, stem_b};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid((e->eng.T + kAdvWarps - 1) / kAdvWarps), block(kAdvWarps * 32);
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(tower_out);
    __nv_bfloat16* so = static_cast<__nv_bfloat16*>(stem_out);
    unsigned long long* tl = nullptr;
    if (e->timeline && e->timeline_slots > 0) {
        tl = e->timeline + 2 * (e->timeline_next % e->timeline_slots);
        e->timeline_next += 1;
    }
#define AZ_ADV(NWv, KCv, Rv, NZv)                                                                                         \
    do {                                                                                                             \
        auto kfn = k_advance<NWv, KCv, Rv, NZv>;                                                                          \
        AZ_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        kfn<<<grid, block, smem, s>>>(e->eng, e->aux, x, hp, sp, so, leaf_valid, tl);                                    \
    } while (0)
    const bool nz = e->eng.dirichlet != 0;
    if (e->c4 && !nz) AZ_ADV(1, 1, C4Rules, false);
    else if (e->nw == 1 && e->kc == 1 && !nz) AZ_ADV(1, 1, Rules, false);
    else if (e->nw == 1 && e->kc == 1) AZ_ADV(1, 1, Rules, true);
    else if (e->nw == 1 && !nz) AZ_ADV(1, 4, Rules, false);
    else if (e->nw == 1) AZ_ADV(1, 4, Rules, true);
    else if (e->kc == 1 && !nz) AZ_ADV(2, 1, Rules, false);
    else if (e->kc == 1) AZ_ADV(2, 1, Rules, true);
    else if (!nz) AZ_ADV(2, 4, Rules, false);
    else AZ_ADV(2, 4, Rules, true);
#undef AZ_ADV
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_debug_timeline(az_engine* e, void* dev_slots, int32_t n_slots) {
    if (!e || n_slots < 0) return fail(AZ_ERR_ARG, "az_debug_timeline: bad argument%s");
    e->timeline = static_cast<unsigned long long*>(dev_slots);
    e->timeline_slots = dev_slots ? n_slots : 0;
    e->timeline_next = 0;
    return AZ_OK;
}

AZ_API int az_debug_dirichlet(uint64_t seed, double alpha, int32_t k, int32_t n, double* out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!out || k < 1 || k > AZ_MAX_ACTIONS || n < 0 || !(alpha > 0.0)) return fail(AZ_ERR_ARG, "az_debug_dirichlet: bad argument%s");
    k_debug_dirichlet<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(seed, alpha, k, n, out);
    AZ_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_cache_clear(az_engine* e, void* stream) {
    if (!e) return fail(AZ_ERR_ARG, "null engine%s");
    if (!e->eng.cache_meta) return AZ_OK;
    AZ_CUDA(cudaMemsetAsync(e->eng.cache_meta, 0, sizeof(unsigned int) * ((size_t)e->eng.cache_mask + 1),
                            static_cast<cudaStream_t>(stream)));
    return AZ_OK;
}
