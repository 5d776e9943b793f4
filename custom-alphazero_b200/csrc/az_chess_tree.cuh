// az_chess_tree.cuh - warp-per-tree PUCT search over chess positions (included by az_chess.cu).
//
// Same search as az_tree.cuh (reference mcts/mcts.py:39-222: Q = W/N, U = ((c * prior) * (sum N) ** 0.5) / (1 + N),
// first-maximum argmax, eager expansion of all children, alternating-sign backup, play + re-root keeping the
// subtree), with what chess changes:
//   * up to 218 children per node: lanes score them in chunks of 32 (kKC = 7 chunks), a node's move is stored
//     (node_m, uint16 action index) because the j-th legal move of a chess position is not a cheap bit trick;
//   * the position is eight 64-bit words replayed in registers (move + mirror per level, every lane redundantly -
//     one instruction stream per warp, so this costs what one thread would);
//   * a leaf is classified by the legal-move generator (az_chess.cuh): no moves / insufficient material / 75-move
//     rule end the game; the legal mask IS the child list (ascending action order);
//   * np.sum over more than 128 priors takes numpy's pairwise split once (n <= 224).
// Bit-exactness: IEEE double with explicit round-to-nearest intrinsics, (sum N) ** 0.5 from the host-built table.
#pragma once

namespace azc {

using az::kFull;
using az::load_node;
using az::NodeA;
using az::store_node;

constexpr int kMaxKids = AZ_CHESS_MAX_CHILDREN;
constexpr int kKC = kMaxKids / 32;
constexpr int kCDepth = AZ_MAX_DEPTH;
constexpr int kCWarps = 4;

struct CEng {
    int T, C, S, F, P;
    int sims_target, greedy_idx, eval_mode, prior_mode, move_mode, max_free, lut_len, auto_restart;
    double c_puct;
    uint64_t seed;
    long long game_base, games_target;
    int32_t* status;
    int32_t* ply;
    long long* game_id;
    Pos* root_pos;
    int32_t* half;
    int32_t* root_node;
    int32_t* n_nodes;
    int32_t* sims_done;
    int32_t* pending;
    int32_t* path_len;
    int32_t* path;
    Pos* leaf_pos;
    u64* leaf_mask;
    long long* counters;
    double* uniforms;
    NodeA* node_a;
    double* node_p;
    uint16_t* node_m;
    int32_t* smp_count;
    long long* smp_game;
    int32_t* smp_ply;
    Pos* smp_pos;
    int32_t* smp_k;
    uint16_t* smp_act;
    int32_t* smp_n;
    int32_t* smp_choice;
    int32_t* fin_count;
    unsigned long long* games_started;
    long long* fin_game;
    int32_t* fin_len;
    int32_t* fin_result;
    const double* pow_lut;
};

struct CScratch {
    double sel[kMaxKids];   // legal priors in action order / root visit counts
    uint16_t act[kMaxKids]; // their actions
    int32_t path[kCDepth];
    int32_t off[32], ob[32], kk[32];
    u64 mask[32];
    Pos e8[8];  // history entries for the encoder
};

__device__ __forceinline__ void cbump(long long* p, long long v) { *p += v; }

// ------------------------------------------------------------------------------------------ K1 select
// pos: root position in, leaf position out.  Returns the leaf node; ws.path[0..depth) holds the path.
__device__ __forceinline__ int c_select(const CEng& e, const NodeA* A, const double* Pr, const uint16_t* Mv, int root,
                                        Pos& pos, CScratch& ws, int lane, int& depth, uint32_t& flags) {
    int node = root;
    uint32_t link = load_node(A + root).link;
    depth = 0;
    while (link) {
        const int base = (int)(link & 0xffffffu), k = (int)(link >> 24);
        NodeA rec[kKC];
        double pr[kKC];
        uint16_t mvv[kKC];  // the children's actions ride along with their records: no dependent load after the argmax
        int ln = 0;
#pragma unroll
        for (int c = 0; c < kKC; ++c) {
            const int j = lane + 32 * c;
            mvv[c] = 0;
            if (j < k) {
                rec[c] = load_node(A + base + j);
                pr[c] = Pr[base + j];
                mvv[c] = Mv[base + j];
                ln += rec[c].n;
            } else {
                rec[c].n = 0;
                rec[c].w = 0.0;
                rec[c].link = 0;
                pr[c] = 0.0;
            }
        }
        const int total = __reduce_add_sync(kFull, ln);  // mcts.py:50: sum over the node's edges
        double s;
        if (total < e.lut_len) {
            s = __ldg(e.pow_lut + total);
        } else {
            s = sqrt((double)total);
            flags |= AZ_FLAG_LUT_OVERFLOW;
        }
        double best = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < kKC; ++c) {
            const int j = lane + 32 * c;
            if (j < k) {
                double q = rec[c].n ? __ddiv_rn(rec[c].w, (double)rec[c].n) : 0.0;  // mcts.py:39-43
                double u = __dmul_rn(e.c_puct, pr[c]);                               // mcts.py:47-48
                u = __dmul_rn(u, s);
                u = __ddiv_rn(u, (double)(1 + rec[c].n));
                double v = __dadd_rn(q, u);
                if (v > best || bi == 0x7fffffff) {
                    best = v;
                    bi = j;
                }
            }
        }
        double mx = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(kFull, mx, o);
            mx = ov > mx ? ov : mx;
        }
        bi = __reduce_min_sync(kFull, (bi != 0x7fffffff && best == mx) ? bi : 0x7fffffff);  // first maximum
        uint32_t clink = 0, cact = 0;
#pragma unroll
        for (int c = 0; c < kKC; ++c)
            if ((bi >> 5) == c) {
                clink = rec[c].link;
                cact = mvv[c];
            }
        clink = __shfl_sync(kFull, clink, bi & 31);
        cact = __shfl_sync(kFull, cact, bi & 31);
        node = base + bi;
        if (depth >= kCDepth) {  // the stored path is full: the tree is deeper than AZ_MAX_DEPTH
            flags |= AZ_FLAG_ILLEGAL;
            break;
        }
        if (lane == 0) ws.path[depth] = node;
        ++depth;
        const int mv = act_move((int)cact);
        pos = play(pos, mv & 63, (mv >> 6) & 63, mv >> 12, true);  // chess/board.py:162-173
        link = clink;
    }
    __syncwarp();
    return node;
}

// numpy add.reduce (pairwise) for n <= 256: one split above 128 elements
__device__ __forceinline__ double c_sum_f64(const double* a, int n) {
    if (n <= 128) return az::np_sum_f64(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(az::np_sum_f64(a, n2), az::np_sum_f64(a + n2, n - n2));
}
__device__ __forceinline__ float c_sum_f32(const double* a, int n) {
    if (n <= 128) return az::np_sum_f32(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(az::np_sum_f32(a, n2), az::np_sum_f32(a + n2, n - n2));
}

// mcts/utils.py:4-16 on ws.sel[0..k)
__device__ __forceinline__ void c_normalise(CScratch& ws, int k, int prior_mode, int lane) {
    __syncwarp();
    if (prior_mode == AZ_PRIOR_F32) {
        const float s = c_sum_f32(ws.sel, k);
        __syncwarp();
        for (int j = lane; j < k; j += 32)
            ws.sel[j] = s == 0.0f ? __ddiv_rn(1.0, (double)k) : (double)__fdiv_rn((float)ws.sel[j], s);
    } else {
        const double s = c_sum_f64(ws.sel, k);
        __syncwarp();
        for (int j = lane; j < k; j += 32) ws.sel[j] = s == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn(ws.sel[j], s);
    }
    __syncwarp();
}

// legal mask in ws.mask[0..30) -> ws.act[0..k) ascending; returns k.  Lane w owns word w.
__device__ __forceinline__ int c_list_actions(CScratch& ws, int lane) {
    const u64 word = lane < kMaskWords ? ws.mask[lane] : 0ull;
    const int cnt = popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    int j = incl - cnt;
    for (u64 b = word; b; b &= b - 1) ws.act[j++] = (uint16_t)(lane * 64 + lsb(b));
    const int k = __shfl_sync(kFull, incl, 31);
    __syncwarp();
    return k;
}

// ------------------------------------------------------------------------------------------ K4 expand
// ws.mask holds the leaf's legal mask.  prior_of(a) = evaluator's prior of action a (double).  Returns the new link
// (0 when the pool is exhausted).
template <typename PriorFn>
__device__ __forceinline__ uint32_t c_expand(const CEng& e, NodeA* A, double* Pr, uint16_t* Mv, int t, CScratch& ws, int lane,
                                             uint32_t& flags, int prior_mode, PriorFn prior_of) {
    const int k = c_list_actions(ws, lane);
    for (int j = lane; j < k; j += 32) ws.sel[j] = prior_of((int)ws.act[j]);
    c_normalise(ws, k, prior_mode, lane);
    const int base = (e.n_nodes[t] + 7) & ~7;
    if (k > 255 || base + k > e.C || base + k > 0xffffff) {
        flags |= AZ_FLAG_POOL_OVERFLOW;
        return 0;
    }
    NodeA fresh;
    fresh.w = 0.0;
    fresh.n = 0;
    fresh.link = 0;
    for (int j = lane; j < k; j += 32) {
        store_node(A + base + j, fresh);
        Pr[base + j] = ws.sel[j];
        Mv[base + j] = ws.act[j];
    }
    __syncwarp();
    if (lane == 0) {
        e.n_nodes[t] = base + k;
        cbump(e.counters + (size_t)t * 8 + 5, k);
        long long* hw = e.counters + (size_t)t * 8 + 7;
        if (base + k > *hw) *hw = base + k;
    }
    return (uint32_t)base | ((uint32_t)k << 24);
}

// ------------------------------------------------------------------------------------------ K5 backup
__device__ __forceinline__ void c_backup(NodeA* A, int root, const CScratch& ws, int depth, double v0, uint32_t new_link,
                                         int lane) {
    for (int i = lane; i < depth; i += 32) {
        NodeA* p = A + ws.path[depth - 1 - i];
        NodeA rec = load_node(p);
        rec.n += 1;
        rec.w = __dadd_rn(rec.w, (i & 1) ? -v0 : v0);
        if (i == 0 && new_link) rec.link = new_link;
        store_node(p, rec);
    }
    if (depth == 0 && new_link && lane == 0) {  // first simulation on an edgeless root: nothing to back up
        NodeA rec = load_node(A + root);
        rec.link = new_link;
        store_node(A + root, rec);
    }
    __syncwarp();
}

// in-kernel hash evaluator (oracle/chess_ref.py: hash_evaluator): FNV-1a over the 64 squares' piece codes + 7, then
// castling rights and en-passant square; priors / value as oracle/evaluators.py derives them from the hash
__device__ __forceinline__ uint64_t c_hash_position(const Pos& p) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int sq = 0; sq < 64; ++sq) h = (h ^ (uint64_t)(piece_at(p, sq) + 7)) * 0x100000001B3ull;
    h = (h ^ (p.meta & 0x7ffull)) * 0x100000001B3ull;
    return h;
}

// Warp-collective move sink: every lane runs the generator on the same position (one instruction stream), and when a
// piece's target set is known the lanes share out its moves - lane j takes the j-th target, looks up the action index
// and sets its bit in the shared-memory mask - instead of each lane walking all targets.
struct WarpSink {
    uint32_t* m32;  // ws.mask viewed as 64 x 32-bit words
    int lane;
    __device__ __forceinline__ void clear() {
        m32[lane] = 0u;
        m32[lane + 32] = 0u;
        __syncwarp();
    }
    __device__ __forceinline__ void set(int a) { atomicOr(m32 + (a >> 5), 1u << (a & 31)); }
    __device__ __forceinline__ void one(int from, int to, int promo) {
        if (lane == 0) set(act_index(from, to) + promo);
    }
    __device__ __forceinline__ void targets(int from, u64 t) {
        if (lane < popc(t)) set(act_index(from, az::nth_set64(t, lane)));
    }
    __device__ __forceinline__ void pawn_targets(int from, u64 t) {  // at most 3 targets, promotions x 4
        if (lane < 4 * popc(t)) {
            const int to = az::nth_set64(t, lane >> 2), pr = lane & 3;
            if (to >= 56) set(act_index(from, to) + pr + 1);
            else if (pr == 0) set(act_index(from, to));
        }
    }
    __device__ __forceinline__ int count() {
        __syncwarp();
        return __reduce_add_sync(kFull, __popc(m32[lane]) + __popc(m32[lane + 32]));
    }
};

// leaf classification: fills ws.mask (legal moves as a mask over the action list), returns game_status
// (0 ongoing, 1 checkmate, 2 draw)
__device__ __forceinline__ int c_classify(const Pos& pos, CScratch& ws, int lane) {
    WarpSink sink{reinterpret_cast<uint32_t*>(ws.mask), lane};
    const GenInfo gi = gen_white_to(pos, sink);
    __syncwarp();
    return game_status(pos, gi);
}

// ------------------------------------------------------------------------------------------ one advance of one tree
// Consume the evaluation of the pending leaf (K4 + K5), then simulate until a leaf needs the evaluator; terminal
// leaves are finished on the spot (at most max_free per call).  Returns 1 with the leaf in e.leaf_pos / e.leaf_mask
// when an evaluation is wanted.  EXTERNAL = false runs the in-kernel evaluator instead and never returns 1 before
// the budget is spent.
template <bool EXTERNAL, typename PriorFn>
__device__ __forceinline__ int c_step_tree(const CEng& e, int t, CScratch& ws, int lane, bool have_eval, double value_in,
                                           int eval_prior_mode, PriorFn ext_prior, Pos& leaf_out) {
    int st = e.status[t];
    if ((st & AZ_PHASE_MASK) != AZ_PHASE_SEARCH) return 0;
    uint32_t flags = 0;
    const int h = e.half[t];
    const size_t pool = ((size_t)t * 2 + h) * e.C;
    NodeA* A = e.node_a + pool;
    double* Pr = e.node_p + pool;
    uint16_t* Mv = e.node_m + pool;
    const int root = e.root_node[t];
    int sims = e.sims_done[t];
    long long n_sims = 0, n_evals = 0, sum_depth = 0;
    int want = 0;
    {
        // the root's child block (records, priors, actions) is the first thing the next selection reads: ask L2 for it
        // now, so that the HBM round trip runs underneath the expansion and backup of the pending leaf
        const uint32_t rl = load_node(A + root).link;
        const int rb = (int)(rl & 0xffffffu), rk = (int)(rl >> 24);
        if (rl) {
            if (lane * 8 < rk) az::prefetch_l2(A + rb + lane * 8);          // 8 records per 128-byte line
            if (lane * 16 < rk) az::prefetch_l2(Pr + rb + lane * 16);       // 16 priors per line
            if (lane * 64 < rk) az::prefetch_l2(Mv + rb + lane * 64);       // 64 actions per line
        }
    }
    if (EXTERNAL && e.pending[t] == 1) {
        if (!have_eval) return 0;  // first call after a reset without priors: nothing to consume yet
        const int depth = e.path_len[t];
        for (int i = lane; i < depth; i += 32) ws.path[i] = e.path[(size_t)t * kCDepth + i];
        if (lane < 32) ws.mask[lane] = e.leaf_mask[(size_t)t * 32 + lane];
        __syncwarp();
        const uint32_t link = c_expand(e, A, Pr, Mv, t, ws, lane, flags, eval_prior_mode, ext_prior);
        c_backup(A, root, ws, depth, -value_in, link, lane);  // mcts.py:175
        ++sims;
        ++n_sims;
        ++n_evals;
        sum_depth += depth;
        if (lane == 0) e.pending[t] = 0;
    }
    int free_left = e.max_free;
    while (!(flags & (AZ_FLAG_POOL_OVERFLOW | AZ_FLAG_ILLEGAL))) {
        if (sims >= e.sims_target) {
            st = (st & ~AZ_PHASE_MASK) | AZ_PHASE_READY;
            break;
        }
        Pos pos = load_cpos(e.root_pos + t);
        int depth;
        c_select(e, A, Pr, Mv, root, pos, ws, lane, depth, flags);
        if (flags & AZ_FLAG_ILLEGAL) break;
        const int term = c_classify(pos, ws, lane);
        if (term) {
            c_backup(A, root, ws, depth, term == 1 ? 1.0 : 0.0, 0, lane);  // mcts.py:179
            ++sims;
            ++n_sims;
            sum_depth += depth;
            if (EXTERNAL && --free_left <= 0) break;
            continue;
        }
        if constexpr (!EXTERNAL) {
            uint32_t link;
            double v;
            if (e.eval_mode == AZ_EVAL_HASH) {
                const uint64_t hh = c_hash_position(pos);
                link = c_expand(e, A, Pr, Mv, t, ws, lane, flags, e.prior_mode, [&](int a) { return az::hash_prior(hh, a); });
                v = az::hash_value(hh);
            } else {
                link = c_expand(e, A, Pr, Mv, t, ws, lane, flags, e.prior_mode,
                                [&](int) { return 1.0 / (double)kActions; });  // np.full(A, 1 / A)
                v = 0.0;
            }
            c_backup(A, root, ws, depth, -v, link, lane);
            ++sims;
            ++n_sims;
            ++n_evals;
            sum_depth += depth;
            continue;
        }
        if constexpr (EXTERNAL) {  // hand the leaf to the evaluator
            for (int i = lane; i < depth; i += 32) e.path[(size_t)t * kCDepth + i] = ws.path[i];
            e.leaf_mask[(size_t)t * 32 + lane] = ws.mask[lane];
            if (lane == 0) {
                e.path_len[t] = depth;
                e.pending[t] = 1;
                store_cpos(e.leaf_pos + t, pos);
            }
            leaf_out = pos;
            want = 1;
            break;
        }
    }
    if (lane == 0) {
        e.sims_done[t] = sims;
        e.status[t] = st | (int)flags;
        cbump(e.counters + (size_t)t * 8 + 0, n_sims);
        cbump(e.counters + (size_t)t * 8 + 1, n_evals);
        cbump(e.counters + (size_t)t * 8 + 4, sum_depth);
    }
    __syncwarp();
    return want;
}

__global__ void __launch_bounds__(kCWarps * 32) k_chess_search(CEng e) {
    __shared__ CScratch s_ws[kCWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kCWarps + warp;
    if (t >= e.T) return;
    Pos dummy;
    c_step_tree<false>(e, t, s_ws[warp], lane, false, 0.0, e.prior_mode, [](int) { return 0.0; }, dummy);
}

// MINB = resident blocks per SM the register allocation aims for: 7 x 4 warps = 28 warps hold 4096 trees on 148 SMs in
// one wave (72 registers, some spills); 4 keeps everything in registers (112) but needs two waves.
template <typename PT, int MINB>
__global__ void __launch_bounds__(kCWarps * 32, MINB) k_chess_step(CEng e, const PT* priors, const PT* values, int have_eval,
                                                             __nv_bfloat16* states_out, int plane_stride, int plane_first, int32_t* leaf_valid) {
    __shared__ CScratch s_ws[kCWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kCWarps + warp;
    if (t >= e.T) return;
    CScratch& ws = s_ws[warp];
    const PT* pt = priors + (size_t)t * kActions;
    const double v = have_eval ? (double)values[t] : 0.0;
    Pos leaf;
    const int mode = sizeof(PT) == 4 ? AZ_PRIOR_F32 : AZ_PRIOR_F64;
    const int want = c_step_tree<true>(e, t, ws, lane, have_eval != 0, v, mode, [&](int a) { return (double)pt[a]; }, leaf);
    if (lane == 0) leaf_valid[t] = want;
    if (want && states_out) {  // states_out null: the caller computes the stem from leaf_pos (az_chess_stem)
        stage_history(leaf, nullptr, ws.e8, lane);
        encode_planes_strided<__nv_bfloat16>(ws.e8, states_out + (size_t)t * 64 * plane_stride, lane, plane_stride, plane_first);
    }
}

// ------------------------------------------------------------------------------------------ game bookkeeping
__device__ __forceinline__ void c_fresh_tree(const CEng& e, int t, const Pos& root, long long game, int lane) {
    NodeA z;
    z.w = 0.0;
    z.n = 0;
    z.link = 0;
    if (lane < 8) store_node(e.node_a + ((size_t)t * 2) * e.C + lane, z);
    if (lane == 0) {
        store_cpos(e.root_pos + t, root);
        e.half[t] = 0;
        e.root_node[t] = 0;
        e.n_nodes[t] = 8;
        e.sims_done[t] = 0;
        e.pending[t] = 0;
        e.path_len[t] = 0;
        e.ply[t] = 0;
        e.game_id[t] = game;
        e.status[t] = AZ_PHASE_SEARCH;
    }
    __syncwarp();
}

__global__ void k_chess_reset(CEng e) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        *e.smp_count = 0;
        *e.fin_count = 0;
        const long long started = (long long)e.T < e.games_target ? (long long)e.T : e.games_target;
        *e.games_started = (unsigned long long)started;
    }
    const int t = warp;
    if (t >= e.T) return;
    if (lane < 8) e.counters[(size_t)t * 8 + lane] = 0;
    if ((long long)t < e.games_target) {
        c_fresh_tree(e, t, start_position(), e.game_base + t, lane);
    } else if (lane == 0) {
        e.status[t] = AZ_PHASE_IDLE;
        e.game_id[t] = -1;
        e.pending[t] = 0;
    }
}

__global__ void k_chess_set_roots(CEng e, const int32_t* ids, const Pos* positions, int n) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const int t = ids[i];
    if (t < 0 || t >= e.T) return;
    Pos p = load_cpos(positions + i);
    if (black_to_move(p)) {  // the engine's trees always have white to move: mirror, like Board.play would have
        p = mirror(p);
    }
    const long long game = e.game_id[t] >= 0 ? e.game_id[t] : e.game_base + t;
    c_fresh_tree(e, t, p, game, lane);
}

__global__ void k_chess_begin(CEng e) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= e.T) return;
    const int st = e.status[t], ph = st & AZ_PHASE_MASK;
    if (ph == AZ_PHASE_SEARCH || ph == AZ_PHASE_READY) {
        e.sims_done[t] = 0;
        e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
    }
}

__global__ void k_chess_rings_clear(CEng e) {
    *e.smp_count = 0;
    *e.fin_count = 0;
}

// the game of tree t is over (result: 1 = last mover won, 0 = draw): finished ring, then the next game or idle.
// sticky = the tree's AZ_FLAG_* bits.  Returns false (tree left STALLED, result parked in bit 16) when the ring is full.
__device__ __forceinline__ bool c_finish_game(const CEng& e, int t, int sticky, int result, int lane) {
    int slot = 0;
    if (lane == 0) {
        slot = atomicAdd(e.fin_count, 1);
        if (slot >= e.F) atomicSub(e.fin_count, 1);
    }
    slot = __shfl_sync(kFull, slot, 0);
    if (slot >= e.F) {
        if (lane == 0) e.status[t] = sticky | AZ_PHASE_STALLED | (result << 16);
        return false;
    }
    long long next = -1;
    if (lane == 0) {
        e.fin_game[slot] = e.game_id[t];
        e.fin_len[slot] = e.ply[t];
        e.fin_result[slot] = result;
        cbump(e.counters + (size_t)t * 8 + 3, 1);
        if (e.auto_restart) {
            const unsigned long long g = atomicAdd(e.games_started, 1ull);
            if ((long long)g < e.games_target) next = e.game_base + (long long)g;
            else atomicAdd(e.games_started, ~0ull);  // -1
        }
    }
    next = __shfl_sync(kFull, next, 0);
    if (next >= 0) {
        c_fresh_tree(e, t, start_position(), next, lane);
        if (lane == 0) e.status[t] = AZ_PHASE_SEARCH | sticky;
    } else if (lane == 0) {
        e.status[t] = AZ_PHASE_IDLE | sticky;
    }
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------------------------------ K6 play / re-root
__global__ void __launch_bounds__(kCWarps * 32) k_chess_move(CEng e, int greedy_override, int move_mode) {
    __shared__ CScratch s_ws[kCWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kCWarps + warp;
    if (t >= e.T) return;
    CScratch& ws = s_ws[warp];
    const int st = e.status[t], phase = st & AZ_PHASE_MASK;
    if (phase == AZ_PHASE_STALLED) {
        c_finish_game(e, t, st & 0xff00, (st >> 16) & 1, lane);
        return;
    }
    if (phase != AZ_PHASE_READY) return;
    const int h = e.half[t];
    const size_t pool = ((size_t)t * 2 + h) * e.C, pool2 = ((size_t)t * 2 + (h ^ 1)) * e.C;
    NodeA* As = e.node_a + pool;
    double* Ps = e.node_p + pool;
    uint16_t* Ms = e.node_m + pool;
    NodeA* Ad = e.node_a + pool2;
    double* Pd = e.node_p + pool2;
    uint16_t* Md = e.node_m + pool2;
    const int root = e.root_node[t];
    const uint32_t rlink = load_node(As + root).link;
    const int base = (int)(rlink & 0xffffffu), k = (int)(rlink >> 24);
    const int ply = e.ply[t];
    if (k == 0) {  // the reference would raise on an edgeless root (np.argmax of [])
        if (lane == 0) e.status[t] = st | AZ_FLAG_ILLEGAL;
        return;
    }
    // sample-ring slot first: without one the tree stays READY and the host has to drain the ring
    int slot = 0;
    if (lane == 0) {
        slot = atomicAdd(e.smp_count, 1);
        if (slot >= e.S) atomicSub(e.smp_count, 1);
    }
    slot = __shfl_sync(kFull, slot, 0);
    if (slot >= e.S) return;
    // root visit counts (mcts.py:189-197)
    for (int j = lane; j < k; j += 32) {
        ws.sel[j] = (double)load_node(As + base + j).n;
        ws.act[j] = Ms[base + j];
    }
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
    int pick = am;
    if (move_mode != AZ_MOVE_ARGMAX && !greedy) {
        // np.random.choice(edges, 1, p=pi): cdf = cumsum(pi); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
        const double u = move_mode == AZ_MOVE_HOST_UNIFORMS ? e.uniforms[(size_t)t * e.P + (ply < e.P ? ply : e.P - 1)]
                                                            : az::philox_uniform(e.seed, e.game_id[t], ply);
        double total = 0.0;
        for (int j = 0; j < k; ++j) total = __dadd_rn(total, ws.sel[j]);
        double last = 0.0;
        for (int j = 0; j < k; ++j) {
            const double pj = total == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn(ws.sel[j], total);
            last = j == 0 ? pj : __dadd_rn(last, pj);
        }
        double acc = 0.0;
        pick = k - 1;
        for (int j = 0; j < k; ++j) {
            const double pj = total == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn(ws.sel[j], total);
            acc = j == 0 ? pj : __dadd_rn(acc, pj);
            if (__ddiv_rn(acc, last) > u) {
                pick = j;
                break;
            }
        }
    }
    // sample: parent position, legal actions, their visit counts, the chosen action (self_play.py:63-66)
    Pos pos = load_cpos(e.root_pos + t);
    for (int j = lane; j < kMaxKids; j += 32) {
        e.smp_act[(size_t)slot * kMaxKids + j] = j < k ? ws.act[j] : (uint16_t)0xffff;
        e.smp_n[(size_t)slot * kMaxKids + j] = j < k ? (int)ws.sel[j] : 0;
    }
    const int action = ws.act[pick];
    if (lane == 0) {
        store_cpos(e.smp_pos + slot, pos);
        e.smp_game[slot] = e.game_id[t];
        e.smp_ply[slot] = ply;
        e.smp_k[slot] = k;
        e.smp_choice[slot] = action | (greedy ? 1 << 16 : 0);
    }
    const int mv = act_move(action);
    pos = play(pos, mv & 63, (mv >> 6) & 63, mv >> 12, true);  // mcts.py:205
    if (lane == 0) {
        store_cpos(e.root_pos + t, pos);
        e.ply[t] = ply + 1;
        cbump(e.counters + (size_t)t * 8 + 2, 1);
        e.sims_done[t] = 0;
        e.pending[t] = 0;
    }
    __syncwarp();
    int term = c_classify(pos, ws, lane);
    if (!term && e.P > 0 && ply + 1 >= e.P) term = 2;  // cut-off: recorded as a draw
    if (term) {
        c_finish_game(e, t, st & 0xff00, term == 1 ? 1 : 0, lane);
        return;
    }
    // re-root to the chosen child, keeping its subtree (mcts.py:207): in place while this half has room for a whole
    // search at the worst-case fan-out ...
    const int used = e.n_nodes[t];
    const bool fits = (long long)used + (long long)(e.sims_target + 1) * kMaxKids + 8 <= (long long)e.C;
    if (fits) {
        if (lane == 0) {
            e.root_node[t] = base + pick;
            e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
        }
        return;
    }
    // ... otherwise the kept subtree is copied breadth-first into the other half
    {
        NodeA z;
        z.w = 0.0;
        z.n = 0;
        z.link = 0;
        if (lane == 0) {
            store_node(Ad, load_node(As + base + pick));
            Pd[0] = Ps[base + pick];
            Md[0] = Ms[base + pick];
        } else if (lane < 8) {
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
        const int kp = (kk + 7) & ~7;
        int incl = kp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - kp, total = __shfl_sync(kFull, incl, 31);
        if (n_dst + total > e.C) {
            if (lane == 0) e.status[t] = st | AZ_FLAG_POOL_OVERFLOW;
            return;
        }
        if (kk) {
            NodeA r2 = load_node(Ad + head + lane);
            r2.link = (uint32_t)(n_dst + excl) | ((uint32_t)kk << 24);
            store_node(Ad + head + lane, r2);
        }
        ws.off[lane] = excl;
        ws.ob[lane] = ob;
        ws.kk[lane] = kk;
        __syncwarp();
        for (int idx = lane; idx < total; idx += 32) {
            int lo = 0, hi = 32;  // last lane whose exclusive offset is <= idx
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (ws.off[mid] <= idx) lo = mid; else hi = mid;
            }
            const int j = idx - ws.off[lo];
            if (j < ws.kk[lo]) {
                const int src = ws.ob[lo] + j;
                store_node(Ad + n_dst + idx, load_node(As + src));
                Pd[n_dst + idx] = Ps[src];
                Md[n_dst + idx] = Ms[src];
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
        cbump(e.counters + (size_t)t * 8 + 6, n_dst);
        e.half[t] = h ^ 1;
        e.root_node[t] = 0;
        e.n_nodes[t] = n_dst;
        e.status[t] = (st & ~AZ_PHASE_MASK) | AZ_PHASE_SEARCH;
    }
}

// sample-ring entries -> (states f32 [n][8][8][118], policies f64 [n][1880]); one warp per sample
__global__ void __launch_bounds__(128) k_chess_decode(const Pos* pos, const int32_t* ks, const uint16_t* act, const int32_t* ns,
                                                      const int32_t* choice, int n, float* states, double* policies) {
    __shared__ Pos s_e[4][8];
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (i >= n) return;
    const Pos cur = load_cpos(pos + i);
    stage_history(cur, nullptr, s_e[warp], lane);
    encode_planes<float>(s_e[warp], states + (size_t)i * 64 * kPlanes, lane);
    double* pol = policies + (size_t)i * kActions;
    for (int a = lane; a < kActions; a += 32) pol[a] = 0.0;
    __syncwarp();
    const int k = ks[i];
    const uint16_t* ac = act + (size_t)i * kMaxKids;
    const int32_t* nv = ns + (size_t)i * kMaxKids;
    const int ch = choice[i];
    if (ch >> 16) {  // greedy ply: one-hot at the first maximum (mcts.py:189-193)
        if (lane == 0) {
            int am = 0;
            for (int j = 1; j < k; ++j)
                if (nv[j] > nv[am]) am = j;
            pol[ac[am]] = 1.0;
        }
    } else {
        double total = 0.0;
        for (int j = 0; j < k; ++j) total = __dadd_rn(total, (double)nv[j]);
        for (int j = lane; j < k; j += 32)
            pol[ac[j]] = total == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn((double)nv[j], total);
    }
}

}  // namespace azc
