// az_tree.cuh - warp-per-tree PUCT search over HBM node pools (K1 select, K4 expand, K5 backup,
// K6 play / re-root).  One warp owns one game tree; all control flow is warp-uniform, lanes
// parallelise over the children of a node (select, expand), the nodes of a path (backup) and the
// nodes of a BFS wave (re-root compaction).
//
// Reference behaviour restated (paths relative to /root/reference/custom_alphazero/):
//   mcts/mcts.py:39-55    Q = W/N (0 when N == 0), U = ((c * prior) * (sum N) ** 0.5) / (1 + N)
//   mcts/mcts.py:64-68    argmax of Q + U, first maximum wins
//   mcts/mcts.py:111-120  select
//   mcts/mcts.py:145-161  evaluate_and_expand (all children created at once)
//   mcts/mcts.py:163-168  backup with alternating sign
//   mcts/mcts.py:182-222  play: policy from root visits, edge choice, re-root keeping the subtree
//   mcts/utils.py:4-16    normalize_probabilities (numpy pairwise sum + divide)
//
// Bit-exactness: everything that feeds a comparison is IEEE double with explicit round-to-nearest
// intrinsics (no FMA contraction), in the reference's evaluation order; (sum N) ** 0.5 is read
// from a host-built table of CPython's `n ** 0.5` because libm pow(n, .5) != sqrt(n) for some n.
//
// Node pool (per tree, two halves of C nodes): a node is at once a position and the edge that
// leads to it.  Children of an expanded node are contiguous, in board move order, so edge j of
// node v is node first_child(v) + j and its move is the j-th legal move of v's position: neither
// moves nor boards are stored per node - the position is replayed in registers while descending.
//   node_a[v] = {double W; int32 N; uint32 link}   16 B, one 128-bit load per lane
//   node_p[v] = double prior                        8 B
//   link = first_child | (k << 24); 0 = no edges (unexpanded or terminal)
#pragma once
#include <cuda_bf16.h>

#include "az_bitboard.cuh"

namespace az {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxActions = 128;
constexpr int kMaxDepth = 128;
constexpr int kWarpsPerBlock = 4;

struct __align__(16) NodeA {
    double w;
    int32_t n;
    uint32_t link;
};

__device__ __forceinline__ NodeA load_node(const NodeA* p) {
    int4 v = *reinterpret_cast<const int4*>(p);
    NodeA r;
    r.w = __hiloint2double(v.y, v.x);
    r.n = v.z;
    r.link = (uint32_t)v.w;
    return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void store_node(NodeA* p, const NodeA& r) {
    int4 v;
    v.x = __double2loint(r.w);
    v.y = __double2hiint(r.w);
    v.z = r.n;
    v.w = (int)r.link;
    *reinterpret_cast<int4*>(p) = v;
}

// Philox4x32-10 block: 128 random bits for (key, counter)
__device__ __forceinline__ uint4 philox4x32(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {  // [0, 1) with 53 bits, numpy's random_sample recipe
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

// One Gamma(alpha, 1) variate, any alpha > 0 (Marsaglia-Tsang squeeze; alpha < 1 through G(alpha + 1) * U^(1/alpha)).
// Independent stream per (c0, c1, c2, lane-specific c3 base): used for the Dirichlet root noise (mcts.py:70-85).
__device__ __forceinline__ double gamma_variate(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t stream,
                                                double alpha) {
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double g = d;
    for (uint32_t it = 0; it < 64; ++it) {
        const uint4 rb = philox4x32(seed, c0, c1, c2, (stream << 8) | (2 * it + 1));
        const uint4 rc = philox4x32(seed, c0, c1, c2, (stream << 8) | (2 * it + 2));
        const double u1 = u53(rb.x, rb.y), u2 = u53(rb.z, rb.w), u3 = u53(rc.x, rc.y);
        const double x = sqrt(-2.0 * log(1.0 - u1)) * cospi(2.0 * u2);  // standard normal (Box-Muller)
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(1.0 - u3) < 0.5 * x * x + d - d * v + d * log(v)) {
            g = d * v;
            break;
        }
    }
    if (alpha < 1.0) {
        const uint4 rd = philox4x32(seed, c0, c1, c2, (stream << 8) | 255u);
        g *= pow(1.0 - u53(rd.x, rd.y), 1.0 / alpha);  // U in (0, 1]
    }
    return g;
}

// Device view of the engine: rules, knobs and typed pointers into the slab (az_layout).
struct Eng {
    Rules r;
    int T, C, P, F;  // trees, node capacity per half, max plies, finished-ring entries
    int sims_target, greedy_idx, eval_mode, prior_mode, move_mode, max_free, lut_len, auto_restart, inline_play;
    int root_smem;  // 1: select_leaf reads the root's child block from WarpScratch after the first simulation of a launch
    int dirichlet;
    double dir_alpha, dir_ratio;
    double c_puct;
    uint64_t seed;
    long long game_base, games_target;
    int32_t* status;
    int32_t* ply;
    long long* game_id;
    uint64_t* root_board;
    int32_t* half;
    int32_t* root_node;
    int32_t* n_nodes;
    int32_t* sims_done;
    int32_t* pending;
    int32_t* path_len;
    int32_t* path;
    uint64_t* leaf_board;
    long long* counters;
    double* uniforms;
    NodeA* node_a;
    double* node_p;
    int32_t* rec_visits;
    int32_t* rec_action;
    uint64_t* rec_board;
    int32_t* fin_count;
    unsigned long long* games_started;
    long long* fin_game_id;
    int32_t* fin_len;
    int32_t* fin_result;
    int32_t* fin_visits;
    int32_t* fin_action;
    uint64_t* fin_board;
    const double* pow_lut;
    // evaluation memo (null when off)
    unsigned int* cache_meta;
    uint64_t* cache_key;
    float* cache_val;
    unsigned int cache_mask;
};

// per-warp scratch in shared memory
struct WarpScratch {
    double sel[kMaxActions];  // legal priors in action-list order / visit counts
    int32_t path[kMaxDepth];
    int32_t off[32];
    int32_t ob[32];
    float cpri[kMaxActions];  // priors of an evaluation-memo hit
    // north_star "hot root nodes staged in shared memory": the root's child block (records + priors, <= 32 children) kept
    // here across the simulations one launch runs on a tree, written through by the backup (Eng::root_smem)
    NodeA root_rec[32];
    double root_pr[32];
    uint32_t root_link;
};

template <int NW>
__device__ __forceinline__ Pos<NW> load_pos(const uint64_t* p) {
    Pos<NW> r;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        r.cur.w[i] = p[i];
        r.opp.w[i] = p[NW + i];
    }
    return r;
}

template <int NW>
__device__ __forceinline__ void store_pos(uint64_t* p, const Pos<NW>& v, int lane) {
    if (lane < 2 * NW) p[lane] = lane < NW ? v.cur.w[lane % NW] : v.opp.w[lane % NW];
}

// ------------------------------------------------------------------------------------------
// K1 select (mcts.py:111-120).  pos: root position in, leaf position out.
// Returns the leaf node; depth/ws.path receive the path; term = 0 none, 1 mover won, 2 draw.
// ------------------------------------------------------------------------------------------
template <int NW, int KC, bool NOISE = false, class R, class WS = WarpScratch>
__device__ __forceinline__ int select_leaf(const Eng& e, const R& r, const NodeA* A, const double* Pr, int root, Pos<NW>& pos,
                                           WS& ws, int lane, int& depth, int& term, uint32_t& flags,
                                           long long noise_game = 0, int noise_ply = 0, int noise_sim = 0,
                                           int* staged_root = nullptr) {
    int node = root;
    // staged_root (warp-uniform, owned by the caller's simulation loop): the node whose child block sits in ws.root_*
    const bool use_stage = KC == 1 && staged_root != nullptr && e.root_smem;
    const bool hit = use_stage && *staged_root == root;
    uint32_t link = hit ? ws.root_link : load_node(A + root).link;
    depth = 0;
    term = 0;
    while (link) {
        const int base = (int)(link & 0xffffffu), k = (int)(link >> 24);
        NodeA rec[KC];
        double pr[KC];
        int ln = 0;
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            int j = lane + 32 * c;
            if (j < k) {
                if (KC == 1 && hit && depth == 0) {  // the root block: shared memory instead of an L2 / HBM round trip
                    rec[c] = ws.root_rec[j];
                    pr[c] = ws.root_pr[j];
                } else {
                    rec[c] = load_node(A + base + j);
                    pr[c] = Pr[base + j];
                }
                ln += rec[c].n;
            } else {
                rec[c].n = 0;
                rec[c].w = 0.0;
                rec[c].link = 0;
                pr[c] = 0.0;
            }
        }
        if (KC == 1 && use_stage && !hit && depth == 0) {  // first descent of this launch from this root: stage its block
            if (lane < k) {
                ws.root_rec[lane] = rec[0];
                ws.root_pr[lane] = pr[0];
            }
            if (lane == 0) ws.root_link = link;
            *staged_root = root;
        }
        // one level ahead: ask L2 for the child blocks of every child while this level is being scored,
        // so the next level's dependent loads find their lines on chip instead of paying an HBM round trip
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const uint32_t cl = rec[c].link;
            if (cl) {  // child blocks start on 8-node boundaries: 128 B of records, 64 B of priors per 8 children
                const int cb = (int)(cl & 0xffffffu), ck = (int)(cl >> 24);
                for (int o = 0; o < ck; o += 8) {
                    prefetch_l2(A + cb + o);
                    prefetch_l2(Pr + cb + o);
                }
            }
        }
        if (NOISE && depth == 0) {  // compiled only into the noise instantiations: the sampler is register-hungry
            // get_best_edge_with_noise (mcts.py:70-85): at the root the prior is replaced, for this simulation only,
            // by (1 - ratio) * prior + ratio * Dirichlet(alpha * 1_k); a Dirichlet draw is k Gamma(alpha) variates
            // normalised by their sum
            double gsum = 0.0, gam[KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                const int j = lane + 32 * c;
                gam[c] = j < k ? gamma_variate(e.seed ^ 0xD1B54A32D192ED03ull, (uint32_t)noise_game,
                                               (uint32_t)((unsigned long long)noise_game >> 32),
                                               ((uint32_t)noise_ply << 20) | (uint32_t)noise_sim, (uint32_t)j, e.dir_alpha)
                               : 0.0;
                gsum += gam[c];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(kFull, gsum, o);
#pragma unroll
            for (int c = 0; c < KC; ++c)
                if (lane + 32 * c < k)
                    pr[c] = __dadd_rn(__dmul_rn(1.0 - e.dir_ratio, pr[c]), __dmul_rn(e.dir_ratio, gam[c] / gsum));
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
        for (int c = 0; c < KC; ++c) {
            int j = lane + 32 * c;
            if (j < k) {
                double q = rec[c].n ? __ddiv_rn(rec[c].w, (double)rec[c].n) : 0.0;  // mcts.py:39-43
                double u = __dmul_rn(e.c_puct, pr[c]);                               // mcts.py:47-48
                u = __dmul_rn(u, s);                                                 // :49-50
                u = __ddiv_rn(u, (double)(1 + rec[c].n));                            // :51
                double v = __dadd_rn(q, u);                                          // :55
                if (v > best || bi == 0x7fffffff) {
                    best = v;
                    bi = j;
                }
            }
        }
        // first maximum wins (np.argmax, mcts.py:65-67): butterfly max of the score, then the lowest
        // edge index among the lanes that hold it
        double mx = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            if (KC == 1 && r.A <= 8 && o >= 8) continue;  // at most 8 candidates sit in lanes 0..7
            const double ov = __shfl_xor_sync(kFull, mx, o);
            mx = ov > mx ? ov : mx;
        }
        if (KC == 1 && r.A <= 8) mx = __shfl_sync(kFull, mx, 0);
        if (KC == 1) {
            bi = __ffs((int)__ballot_sync(kFull, bi != 0x7fffffff && best == mx)) - 1;
        } else {
            bi = __reduce_min_sync(kFull, (bi != 0x7fffffff && best == mx) ? bi : 0x7fffffff);
        }
        uint32_t clink = 0;
#pragma unroll
        for (int c = 0; c < KC; ++c)
            if ((bi >> 5) == c) clink = rec[c].link;
        clink = __shfl_sync(kFull, clink, bi & 31);
        node = base + bi;
        if (lane == 0) ws.path[depth] = node;
        ++depth;
        // replay the move of edge bi on the register position (board.py:233-250)
        BB<NW> legal = legal_set(r, pos);
        int bit, action;
        edge_move(r, pos, legal, bi, bit, action);
        term = place(r, pos, bit);
        if (term) break;
        link = clink;
    }
    __syncwarp();
    return node;
}

// numpy add.reduce over a contiguous 1-D array (pairwise sum): left fold for n < 8, otherwise
// eight strided accumulators combined as a balanced tree, then the tail (n <= 128 here).
__device__ __forceinline__ double np_sum_f64(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__device__ __forceinline__ float np_sum_f32(const double* a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, (float)a[i]);
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (float)a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], (float)a[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, (float)a[i]);
    return res;
}

// mcts/utils.py:4-16 on ws.sel[0..k): result written back into ws.sel.
__device__ __forceinline__ void normalise_sel(WarpScratch& ws, int k, int prior_mode, int lane) {
    __syncwarp();
    if (prior_mode == AZ_PRIOR_F32) {
        float s = np_sum_f32(ws.sel, k);
        __syncwarp();
        for (int j = lane; j < k; j += 32)
            ws.sel[j] = s == 0.0f ? __ddiv_rn(1.0, (double)k) : (double)__fdiv_rn((float)ws.sel[j], s);
    } else {
        double s = np_sum_f64(ws.sel, k);
        __syncwarp();
        for (int j = lane; j < k; j += 32) ws.sel[j] = s == 0.0 ? __ddiv_rn(1.0, (double)k) : __ddiv_rn(ws.sel[j], s);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K4 expand (mcts.py:145-161).  prior_of(a) yields the evaluator's prior for action a (double).
// Gathers the legal priors in action-list order, normalises them, appends k children in board
// order (Q1: the j-th normalised prior goes to the j-th move in board order).  Returns the new
// link (0 when the pool is exhausted).
// ------------------------------------------------------------------------------------------
template <int NW, class R, typename PriorFn>
__device__ __forceinline__ uint32_t expand_leaf(const Eng& e, const R& r, NodeA* A, double* Pr, const Pos<NW>& pos, int t,
                                                WarpScratch& ws, int lane, uint32_t& flags, int prior_mode,
                                                PriorFn prior_of) {
    BB<NW> legal = legal_set(r, pos);
    int k = 0;
    for (int a0 = 0; a0 < r.A; a0 += 32) {
        int a = a0 + lane;
        bool ok = action_legal(r, pos, legal, a);
        unsigned m = __ballot_sync(kFull, ok);
        if (ok) ws.sel[k + __popc(m & ((1u << lane) - 1u))] = prior_of(a);
        k += __popc(m);
    }
    normalise_sel(ws, k, prior_mode, lane);
    const int base = (e.n_nodes[t] + 7) & ~7;  // 8-node alignment: a block of <= 8 children is one 128 B line
    if (base + k > e.C || base + k > 0xffffff) {
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
    }
    if (lane == 0) e.n_nodes[t] = base + k;
    return (uint32_t)base | ((uint32_t)k << 24);
}

// ------------------------------------------------------------------------------------------
// Evaluation memo (plays_inferences of the reference, mcts.py:122-143): direct-mapped, shared by all trees,
// lossy (a colliding insert overwrites).  Entries are guarded by a seqlock word so that a reader never combines
// the key of one position with the numbers of another: writers take the slot with a CAS (or skip the insert),
// readers re-check the version after reading and treat any change as a miss.  All table accesses bypass L1
// (ld.cg / st.cg): other SMs write these lines during the same kernel.
template <int NW>
__device__ __forceinline__ unsigned int cache_slot(const Eng& e, const Pos<NW>& p) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        h = (h ^ p.cur.w[i]) * 0xFF51AFD7ED558CCDull;
        h = (h ^ (h >> 32) ^ p.opp.w[i]) * 0xC4CEB9FE1A85EC53ull;
    }
    h ^= h >> 29;
    return (unsigned int)h & e.cache_mask;
}

// Insert (whole warp): prior_of(a) for a < A and the value.
template <int NW, class R, typename PriorFn>
__device__ __forceinline__ void cache_insert(const Eng& e, const R& r, const Pos<NW>& p, int lane, PriorFn prior_of, float value) {
    const unsigned int s = cache_slot<NW>(e, p);
    unsigned int m = 0;
    int got = 0;
    if (lane == 0) {
        m = __ldcg(e.cache_meta + s);
        got = !(m & 1u) && atomicCAS(e.cache_meta + s, m, m | 1u) == m;
    }
    got = __shfl_sync(kFull, got, 0);
    if (!got) return;
    uint64_t* key = e.cache_key + (size_t)s * 2 * NW;
    float* val = e.cache_val + (size_t)s * (r.A + 1);
    if (lane < 2 * NW) __stcg(key + lane, lane < NW ? p.cur.w[lane % NW] : p.opp.w[lane % NW]);
    for (int a = lane; a < r.A; a += 32) __stcg(val + a, prior_of(a));
    if (lane == 0) __stcg(val + r.A, value);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(e.cache_meta + s, (m | 1u) + 1u);  // even again, new version
}

// Lookup (whole warp): on a hit the priors land in out_priors[0..A) and value is set.
template <int NW, class R>
__device__ __forceinline__ bool cache_lookup(const Eng& e, const R& r, const Pos<NW>& p, int lane, float* out_priors,
                                             float& value) {
    const unsigned int s = cache_slot<NW>(e, p);
    const unsigned int m1 = __ldcg(e.cache_meta + s);
    if (m1 & 1u) return false;
    const uint64_t* key = e.cache_key + (size_t)s * 2 * NW;
    bool same = true;
#pragma unroll
    for (int i = 0; i < NW; ++i) same = same && __ldcg(key + i) == p.cur.w[i] && __ldcg(key + NW + i) == p.opp.w[i];
    if (!same || m1 == 0u) return false;  // version 0 = never written
    const float* val = e.cache_val + (size_t)s * (r.A + 1);
    for (int a = lane; a < r.A; a += 32) out_priors[a] = __ldcg(val + a);
    const float v = __ldcg(val + r.A);
    __threadfence();
    const unsigned int m2 = __ldcg(e.cache_meta + s);
    const bool ok = __all_sync(kFull, m2 == m1);
    if (!ok) return false;
    value = v;
    __syncwarp();
    return true;
}

// ------------------------------------------------------------------------------------------
// K5 backup (mcts.py:163-168): path nodes are distinct, so each lane owns one read-modify-write;
// no atomics.  v0 is the value for the player who moved into the leaf; sign alternates upward.
// new_link != 0 also publishes the leaf's fresh children in the same 16-byte store.
// ------------------------------------------------------------------------------------------
template <class WS>
__device__ __forceinline__ void backup_path(NodeA* A, int root, WS& ws, int depth, double v0,
                                            uint32_t new_link, int lane, int staged_root = -1) {
    for (int i = lane; i < depth; i += 32) {
        NodeA* p = A + ws.path[depth - 1 - i];
        NodeA rec = load_node(p);
        rec.n += 1;
        rec.w = __dadd_rn(rec.w, (i & 1) ? -v0 : v0);
        if (i == 0 && new_link) rec.link = new_link;
        store_node(p, rec);
        // write through: path[0] is a child of the root, i.e. an entry of the staged block
        if (i == depth - 1 && staged_root == root) ws.root_rec[ws.path[0] - (int)(ws.root_link & 0xffffffu)] = rec;
    }
    if (depth == 0 && new_link && lane == 0) {  // first simulation on an edgeless root: nothing to back up
        NodeA rec = load_node(A + root);
        rec.link = new_link;
        store_node(A + root, rec);
    }
    __syncwarp();
}

// Scratch of a warp that runs evaluator-free simulations only (select + backup): the stored path.  (The staging
// members exist for select_leaf / backup_path to compile; such a warp never stages a root block.)
struct PathScratch {
    int32_t path[64];
    NodeA root_rec[1];
    double root_pr[1];
    uint32_t root_link;
};

// Evaluator-free simulations of ONE tree without a leaf in flight, for warps that have nothing else to do while the net
// runs (az_net_forward_trees): up to `cap` simulations that end in a terminal leaf (mcts.py:179) are finished on the
// spot; the first leaf that needs the evaluator is parked exactly as az_extra_sims parks it (pending = 2: the next
// az_step hands it out).  Moves are never played here: a tree whose budget is spent stays in SEARCH with pending = 0 and
// the next az_step plays its move in line.  Same simulations in the same order as step_tree would run them - only
// earlier - so per-tree results cannot change.  Boards of up to 63 cells (PathScratch::path).  `stop`: a shared-memory
// word the caller raises when no further simulation should be started.
template <int NW, int KC, class R>
__device__ __forceinline__ void free_sims_tree(const Eng& e, const R& r, int t, PathScratch& ws, int lane, int cap,
                                               const volatile int* stop = nullptr) {
    const int st = e.status[t];
    if ((st & AZ_PHASE_MASK) != AZ_PHASE_SEARCH || e.pending[t] != 0) return;
    int sims = e.sims_done[t];
    if (sims >= e.sims_target) return;
    uint32_t flags = 0;
    const size_t pool = ((size_t)t * 2 + e.half[t]) * e.C;
    NodeA* A = e.node_a + pool;
    const double* Pr = e.node_p + pool;
    const int root = e.root_node[t];
    long long nsim = 0, ndepth = 0;
    int pend = 0;
    for (int done = 0; done < cap && sims < e.sims_target; ++done) {
        if (stop && *stop) break;  // the host kernel is about to finish: never keep it waiting (warp-uniform: one shared word)
        Pos<NW> pos = load_pos<NW>(e.root_board + (size_t)t * 2 * NW);
        int depth, term;
        select_leaf<NW, KC, false>(e, r, A, Pr, root, pos, ws, lane, depth, term, flags);
        ndepth += depth;
        if (!term) {  // needs the evaluator: park the leaf
            for (int i = lane; i < depth; i += 32) e.path[(size_t)t * kMaxDepth + i] = ws.path[i];
            store_pos<NW>(e.leaf_board + (size_t)t * 2 * NW, pos, lane);
            if (lane == 0) e.path_len[t] = depth;
            pend = 2;
            break;
        }
        backup_path(A, root, ws, depth, term == 1 ? 1.0 : 0.0, 0u, lane);
        ++sims;
        ++nsim;
    }
    if (lane == 0) {
        if (nsim) atomicAdd(reinterpret_cast<unsigned long long*>(e.counters + (size_t)t * 8 + 0), (unsigned long long)nsim);
        if (ndepth) atomicAdd(reinterpret_cast<unsigned long long*>(e.counters + (size_t)t * 8 + 4), (unsigned long long)ndepth);
        e.sims_done[t] = sims;
        e.pending[t] = pend;
        if (flags) e.status[t] = st | (int)flags;
    }
    __syncwarp();
}

int engine_view(const ::az_engine* e, Eng* out);  // az_kernels.cu (host): device view of an engine, 1 = plain 6x7 fast path

// Rules as seen by a kernel instance: the runtime struct, or the compile-time headline configuration.
using C4Rules = FixedRules<7, 6, 4, 1>;
template <class R>
struct RulesView {
    __device__ static __forceinline__ const Rules& get(const Eng& e) { return e.r; }
};
template <int W_, int H_, int N_, int G_>
struct RulesView<FixedRules<W_, H_, N_, G_>> {
    __device__ static __forceinline__ FixedRules<W_, H_, N_, G_> get(const Eng&) { return {}; }
};

// ------------------------------------------------------------------------------------------
// in-kernel evaluators (oracle/evaluators.py)
// ------------------------------------------------------------------------------------------
template <int NW, class R>
__device__ __forceinline__ uint64_t hash_position(const R& r, const Pos<NW>& p) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int c = 0; c < r.cells; ++c) h = (h ^ (uint64_t)(cell_code(r, p, c) + 1)) * 0x100000001B3ull;
    return h;
}

__device__ __forceinline__ double hash_prior(uint64_t h, int a) {
    uint64_t m = (h ^ ((uint64_t)a * 0x9E3779B97F4A7C15ull)) * 0xFF51AFD7ED558CCDull;
    return (double)(((m >> 40) % 1000ull) + 1ull);
}

__device__ __forceinline__ double hash_value(uint64_t h) {
    return __ddiv_rn((double)((h >> 20) % 2001ull) - 1000.0, 1000.0);
}

// Philox4x32-10, counter (game id lo, game id hi, ply, 0), key = seed; 53-bit uniform in [0, 1)
// built like numpy's random_sample: (a >> 5) * 2^26 + (b >> 6), / 2^53.
__device__ __forceinline__ double philox_uniform(uint64_t seed, long long game, int ply) {
    const uint4 rv = philox4x32(seed, (uint32_t)game, (uint32_t)((uint64_t)game >> 32), (uint32_t)ply, 0u);
    return u53(rv.x, rv.y);
}

// K3: Board.full_state (board.py:83-98) of `pos` into out[H][W][4]; lanes stride over cells.
template <int NW, class R>
__device__ __forceinline__ void encode_state_bf16(const R& r, const Pos<NW>& pos, __nv_bfloat16* out, int lane) {
    for (int c = lane; c < r.cells; c += 32) {
        int code = cell_code(r, pos, c);
        // bf16 1.0 = 0x3F80; planes: empty, side to move, opponent, turn (+1 under keep_same_player)
        uint2 v;
        v.x = (code == 0 ? 0x3F80u : 0u) | (code == 1 ? 0x3F800000u : 0u);
        v.y = (code == 2 ? 0x3F80u : 0u) | 0x3F800000u;
        reinterpret_cast<uint2*>(out)[c] = v;
    }
}

template <int NW, class R>
__device__ __forceinline__ void encode_state_f32(const R& r, const Pos<NW>& pos, float* out, int lane) {
    for (int c = lane; c < r.cells; c += 32) {
        int code = cell_code(r, pos, c);
        reinterpret_cast<float4*>(out)[c] = make_float4(code == 0, code == 1, code == 2, 1.0f);
    }
}

}  // namespace az
