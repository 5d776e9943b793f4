// az_chess.cu - chess kernels and their C ABI (include/az_b200.h, section "chess").
//
//   k_chess_legal   Board.moves / legal_moves_mask / is_game_over (chess/board.py:46-48, :116-117, :178-190), thread per board
//   k_chess_play    Board.play (chess/board.py:162-173): push + mirror, thread per board
//   k_chess_encode  Board.full_state (chess/board.py:58-73): 118 planes, warp per board, coalesced writes
//   k_chess_perft   move-path enumeration on the self-play path (move, mirror, move ...): the known-answer test of
//                   the generator and its throughput measurement, thread per root position
//   k_chess_search / k_chess_step / k_chess_move   warp-per-tree PUCT search over chess positions (further down)
// Rules: az_chess.cuh.  Built for sm_100a only; there is no CPU implementation behind this ABI.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/az_b200.h"
#include "az_chess.cuh"
#include "az_mma.cuh"
#include "az_tree.cuh"

namespace az {
int fail_net(int code, const char* msg);
}

namespace azc {

static_assert(sizeof(Pos) == sizeof(az_chess_pos), "az_chess_pos mirrors azc::Pos");
static_assert(kActions == AZ_CHESS_ACTIONS && kMaskWords == AZ_CHESS_MASK_WORDS && kPlanes == AZ_CHESS_PLANES, "header constants");

#define AZC_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t err__ = (call);                                           \
        if (err__ != cudaSuccess) {                                           \
            char buf__[256];                                                  \
            snprintf(buf__, sizeof(buf__), #call ": %s", cudaGetErrorString(err__)); \
            return az::fail_net(AZ_ERR_CUDA, buf__);                          \
        }                                                                     \
    } while (0)

__device__ __forceinline__ Pos load_cpos(const Pos* p) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
    ulonglong2 a = q[0], b = q[1], c = q[2], d = q[3];
    Pos r;
    r.pawns = a.x; r.knights = a.y; r.bishops = b.x; r.rooks = b.y;
    r.queens = c.x; r.kings = c.y; r.white = d.x; r.meta = d.y;
    return r;
}
__device__ __forceinline__ void store_cpos(Pos* p, const Pos& r) {
    ulonglong2* q = reinterpret_cast<ulonglong2*>(p);
    q[0] = make_ulonglong2(r.pawns, r.knights);
    q[1] = make_ulonglong2(r.bishops, r.rooks);
    q[2] = make_ulonglong2(r.queens, r.kings);
    q[3] = make_ulonglong2(r.white, r.meta);
}

__global__ void __launch_bounds__(128) k_chess_legal(const Pos* pos, int n, u64* mask_out, int32_t* count_out,
                                                      int32_t* status_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos p = load_cpos(pos + i);
    MoveMask mm;
    int unlisted = 0;
    GenInfo gi = legal_moves(p, mm, &unlisted);
    if (mask_out)
        for (int w = 0; w < kMaskWords; ++w) mask_out[(size_t)i * kMaskWords + w] = mm.w[w];
    if (count_out) count_out[i] = gi.n_moves;
    if (status_out) status_out[i] = game_status(p, gi) | (gi.in_check ? 4 : 0) | (unlisted << 8);
}

__global__ void __launch_bounds__(128) k_chess_play(const Pos* pos_in, const int32_t* actions, int n, int keep,
                                                     Pos* pos_out, int32_t* status_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos p = load_cpos(pos_in + i);
    const int a = actions[i];
    MoveMask mm;
    legal_moves(p, mm, nullptr);
    if (a < 0 || a >= kActions || !((mm.w[a >> 6] >> (a & 63)) & 1ull)) {
        store_cpos(pos_out + i, p);  // python-chess raises on an illegal uci: the board stays as it was
        status_out[i] = -1;
        return;
    }
    const int mv = act_move(a);
    Pos q = play(p, mv & 63, (mv >> 6) & 63, mv >> 12, keep != 0);
    store_cpos(pos_out + i, q);
    GenInfo gi = legal_moves(q, mm, nullptr);
    status_out[i] = game_status(q, gi);
}

// The 8 deque entries of Board.full_state for `cur` (oldest first; entry 7 is the current state) into e8.
// hist: 7 older entries or null for the self-play path's deque (six empty entries, then the state of the initial
// position - see oracle/chess_ref.py on python-chess's mirror(); seven empty entries for the un-mirrored ply-0 root).
__device__ __forceinline__ void stage_history(const Pos& cur, const Pos* hist, Pos* e8, int lane) {
    if (lane < 8) {
        Pos e;
        if (lane == 7) {
            e = cur;
            e.meta = (e.meta & ~META_REP) | META_VALID;  // is_repetition() needs the move stack: see az_chess.cuh
        } else if (hist) {
            e = load_cpos(hist + lane);
        } else {
            e = start_position();
            e.meta = (lane == 6 && !fresh_root(cur)) ? (e.meta | META_VALID) : 0;  // Board() itself: 7 empty entries
        }
        e8[lane] = e;
    }
    __syncwarp();
}

template <typename T>
__device__ __forceinline__ T plane_cast(float v);
template <>
__device__ __forceinline__ float plane_cast<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 plane_cast<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// Board.full_state (chess/board.py:58-73) of the deque e8 (shared memory) into out[8][8][stride], stride >= 118 and
// 16-byte rows (stride * sizeof(T) % 16 == 0) or stride == 118.  Almost every value is zero (per cell at most one piece
// plane and one repetition plane per deque entry plus the six scalar planes), so the warp first clears the block with
// coalesced 128-bit stores and then each lane scatters the few non-zero values of its two cells.
template <typename T>
__device__ __forceinline__ void encode_planes_strided(const Pos* e8, T* out, int lane, int stride, int first = 0) {
    // planes [first, 118) land at out[cell][plane - first]; `first` is a multiple of 14 (whole history entries are dropped:
    // on the self-play path entries 0-5 are always empty, so first = 84 loses nothing)
    const int total = 64 * stride;  // elements; the block starts 16-byte aligned (n * 64 * stride * sizeof(T))
    constexpr int per16 = 16 / (int)sizeof(T);
    if ((total % per16) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        uint4* o4 = reinterpret_cast<uint4*>(out);
        for (int i = lane; i < total / per16; i += 32) o4[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (int i = lane; i < total; i += 32) out[i] = plane_cast<T>(0.0f);
    }
    __syncwarp();
    const Pos& cur = e8[7];
    const bool black = black_to_move(cur);
    const int own_k = black ? 4 : 1, own_q = black ? 8 : 2, opp_k = black ? 1 : 4, opp_q = black ? 2 : 8;
    const T one = plane_cast<T>(1.0f);
    const int h0 = first / 14;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cell = lane + 32 * half;
        const int sq = ((7 - (cell >> 3)) << 3) | (cell & 7);  // array row 0 is rank 8 (chess/board.py:119-131)
        T* row = out + (size_t)cell * stride - first;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            if (h < h0) continue;
            const Pos& e = e8[h];
            if (!(e.meta & META_VALID)) continue;
            row[h * 14 + piece_plane(piece_at(e, sq))] = one;  // plane 0 = empty square
            if (e.meta & META_REP) row[h * 14 + 13] = one;
        }
        if (cur.meta & own_q) row[112] = one;
        if (cur.meta & own_k) row[113] = one;
        if (cur.meta & opp_q) row[114] = one;
        if (cur.meta & opp_k) row[115] = one;
        row[116] = plane_cast<T>((float)fullmove(cur));
        row[117] = plane_cast<T>((float)halfmove(cur));
    }
}

template <typename T>
__device__ __forceinline__ void encode_planes(const Pos* e8, T* out, int lane) {
    encode_planes_strided<T>(e8, out, lane, kPlanes);
}

template <typename T>
__global__ void __launch_bounds__(128) k_chess_encode(const Pos* pos, const Pos* hist, int n, T* out) {
    __shared__ Pos s_e[4][8];
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (i >= n) return;
    const Pos cur = load_cpos(pos + i);
    stage_history(cur, hist ? hist + (size_t)i * 7 : nullptr, s_e[warp], lane);
    encode_planes<T>(s_e[warp], out + (size_t)i * 64 * kPlanes, lane);
}

// ------------------------------------------------------------------------------------------ stem from bitboards
// Stem of the net (Conv3x3(118 -> 128) + BN + ReLU, model/tensorflow/model.py:36-46) for leaves of the self-play
// path, computed straight from the 64-byte positions - the 15 kB plane tensor is never written or read.  On that path
// Board.full_state is six all-zero history entries, the state of the initial position and the current state (see
// oracle/chess_ref.py), so of the 118 input planes only the current entry's 14 and the 6 scalar planes vary: the
// initial position's contribution and the bias are a per-cell constant map (host-computed from the weights), and the
// convolution that remains has 20 input planes (24 with padding): K = 9 taps x 24 instead of 9 x 118.
// Implicit GEMM per position on mma.sync (m16n8k16 for planes 0-15 of a tap, m16n8k8 for planes 16-23), A fragments
// read from a zero-bordered 10 x 10 copy of the position in shared memory (12 words per cell: conflict free), the
// warp's weights (32 output channels) in 108 registers; block = 4 warps = the 128 output channels of one position.
__device__ __forceinline__ void mma_1688(float (&d)[4], const uint32_t (&a)[2], uint32_t b) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(b));
}

constexpr int kStemPlanes = 24;                 // 14 piece / repetition planes, 6 scalar planes, 4 zero planes
constexpr int kStemWords = kStemPlanes / 2;     // 32-bit words per cell
constexpr int kStemPad = 100;                   // 10 x 10 zero-bordered board

__global__ void __launch_bounds__(128) k_chess_stem(const Pos* __restrict__ pos, int n, const float* __restrict__ w,
                                                    const float* __restrict__ cmap, __nv_bfloat16* __restrict__ out) {
    __shared__ uint32_t s_in[4][kStemPad * kStemWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    uint32_t* sw = s_in[warp];
    for (int i = lane; i < kStemPad * kStemWords; i += 32) sw[i] = 0u;
    // B fragments of this warp's 32 output channels: w [128][24 planes][9 taps] float32 (BN folded, reduced plane set)
    uint32_t b16[9][4][2], b8[9][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float* wc = w + (size_t)(warp * 32 + nt * 8 + g) * kStemPlanes * 9;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            b16[tap][nt][0] = az::pack_bf16(wc[(2 * t4) * 9 + tap], wc[(2 * t4 + 1) * 9 + tap]);
            b16[tap][nt][1] = az::pack_bf16(wc[(2 * t4 + 8) * 9 + tap], wc[(2 * t4 + 9) * 9 + tap]);
            b8[tap][nt] = az::pack_bf16(wc[(2 * t4 + 16) * 9 + tap], wc[(2 * t4 + 17) * 9 + tap]);
        }
    }
    __syncwarp();
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        const Pos p = load_cpos(pos + t);
        __syncwarp();
        // the position's 20 planes into the padded board: lane = cell (two per lane)
        const bool black = black_to_move(p);
        const int own_k = black ? 4 : 1, own_q = black ? 8 : 2, opp_k = black ? 1 : 4, opp_q = black ? 2 : 8;
        const uint32_t one = 0x3f80u;  // bf16 1.0
        const uint32_t s01 = ((p.meta & own_q) ? one : 0u) | ((p.meta & own_k) ? one << 16 : 0u);          // planes 14, 15
        const uint32_t s23 = ((p.meta & opp_q) ? one : 0u) | ((p.meta & opp_k) ? one << 16 : 0u);          // planes 16, 17
        const uint32_t s45 = az::pack_bf16((float)fullmove(p), (float)halfmove(p));                           // planes 18, 19
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int cell = lane + 32 * half;
            const int sq = ((7 - (cell >> 3)) << 3) | (cell & 7);
            uint32_t* c = sw + (((cell >> 3) + 1) * 10 + (cell & 7) + 1) * kStemWords;
            const int pl = piece_plane(piece_at(p, sq));  // 0 = empty ... 12; plane 13 (repetition) stays 0
#pragma unroll
            for (int wd = 0; wd < 7; ++wd) c[wd] = (pl >> 1) == wd ? (one << (16 * (pl & 1))) : 0u;
            c[7] = s01;
            c[8] = s23;
            c[9] = s45;
        }
        __syncwarp();
        __nv_bfloat16* o = out + (size_t)t * 64 * 128 + warp * 32 + t4 * 2;
        const float* cm = cmap + (fresh_root(p) ? 64 * 128 : 0) + warp * 32 + t4 * 2;  // map 1: no initial-position entry
#pragma unroll 1
        for (int mt = 0; mt < 4; ++mt) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const uint32_t* a0p = sw + ((r0 >> 3) * 10 + (r0 & 7)) * kStemWords;
            const uint32_t* a1p = sw + ((r1 >> 3) * 10 + (r1 & 7)) * kStemWords;
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int toff = ((tap / 3) * 10 + tap % 3) * kStemWords;
                uint32_t a[4], a8[2];
                a[0] = a0p[toff + t4];
                a[1] = a1p[toff + t4];
                a[2] = a0p[toff + t4 + 4];
                a[3] = a1p[toff + t4 + 4];
                a8[0] = a0p[toff + t4 + 8];
                a8[1] = a1p[toff + t4 + 8];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    az::mma_16816(acc[nt], a, b16[tap][nt]);
                    mma_1688(acc[nt], a8, b8[tap][nt]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const float2 c0 = *reinterpret_cast<const float2*>(cm + (size_t)r0 * 128 + nt * 8);
                const float2 c1 = *reinterpret_cast<const float2*>(cm + (size_t)r1 * 128 + nt * 8);
                *reinterpret_cast<uint32_t*>(o + (size_t)r0 * 128 + nt * 8) =
                    az::pack_bf16(fmaxf(acc[nt][0] + c0.x, 0.f), fmaxf(acc[nt][1] + c0.y, 0.f));
                *reinterpret_cast<uint32_t*>(o + (size_t)r1 * 128 + nt * 8) =
                    az::pack_bf16(fmaxf(acc[nt][2] + c1.x, 0.f), fmaxf(acc[nt][3] + c1.y, 0.f));
            }
        }
    }
}

constexpr int kPerftMaxDepth = 8;

__global__ void __launch_bounds__(64) k_chess_perft(const Pos* pos, int n, int depth, unsigned long long* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos st[kPerftMaxDepth];
    MoveMask mm[kPerftMaxDepth];
    int wi[kPerftMaxDepth];
    u64 cur[kPerftMaxDepth];
    Pos p = load_cpos(pos + i);
    if (black_to_move(p)) p = mirror(p);
    unsigned long long total = 0;
    if (depth <= 0) {
        out[i] = 1;
        return;
    }
    st[0] = p;
    GenInfo g0 = gen_white(st[0], mm[0]);
    if (depth == 1) {
        out[i] = (unsigned long long)g0.n_moves;
        return;
    }
    int d = 0;
    wi[0] = 0;
    cur[0] = mm[0].w[0];
    while (d >= 0) {
        while (cur[d] == 0 && wi[d] + 1 < kMaskWords) cur[d] = mm[d].w[++wi[d]];
        if (cur[d] == 0) {
            --d;
            continue;
        }
        const int a = wi[d] * 64 + lsb(cur[d]);
        cur[d] &= cur[d] - 1;
        const int mv = act_move(a);
        Pos q = play(st[d], mv & 63, (mv >> 6) & 63, mv >> 12, true);
        if (d + 2 == depth) {  // children of q are leaves: count them without visiting
            MoveMask tmp;
            total += (unsigned long long)gen_white(q, tmp).n_moves;
        } else {
            ++d;
            st[d] = q;
            gen_white(q, mm[d]);
            wi[d] = 0;
            cur[d] = mm[d].w[0];
        }
    }
    out[i] = total;
}

}  // namespace azc

#include "az_chess_tree.cuh"

namespace azc {

static inline dim3 flat_grid(int n, int block) { return dim3((unsigned)((n + block - 1) / block)); }
static int have_device() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    return AZ_OK;
}

}  // namespace azc

using namespace azc;
#define AZ_API extern "C" __attribute__((visibility("default")))

AZ_API int az_chess_action_table(uint16_t* host_moves_out) {
    if (!host_moves_out) return az::fail_net(AZ_ERR_ARG, "az_chess_action_table: null output");
    for (int a = 0; a < kActions; ++a) host_moves_out[a] = (uint16_t)host_tables::ACT_MOVE[a];
    return AZ_OK;
}

AZ_API int az_chess_legal(const az_chess_pos* pos, int32_t n, uint64_t* mask_out, int32_t* count_out, int32_t* status_out,
                          void* stream) {
    if (n == 0) return AZ_OK;
    if (!pos || n < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_legal: bad argument");
    if (int rc = have_device()) return rc;
    k_chess_legal<<<flat_grid(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const Pos*>(pos), n, reinterpret_cast<u64*>(mask_out), count_out, status_out);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_play(const az_chess_pos* pos_in, const int32_t* actions, int32_t n, int32_t keep_same_player,
                         az_chess_pos* pos_out, int32_t* status_out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!pos_in || !actions || !pos_out || !status_out || n < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_play: bad argument");
    if (int rc = have_device()) return rc;
    k_chess_play<<<flat_grid(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const Pos*>(pos_in), actions, n, keep_same_player, reinterpret_cast<Pos*>(pos_out), status_out);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_encode(const az_chess_pos* pos, const az_chess_pos* history, int32_t n, int32_t dtype, void* states_out,
                           void* stream) {
    if (n == 0) return AZ_OK;
    if (!pos || !states_out || n < 0 || (dtype != AZ_F32 && dtype != AZ_BF16))
        return az::fail_net(AZ_ERR_ARG, "az_chess_encode: bad argument (dtype must be AZ_F32 or AZ_BF16)");
    if (int rc = have_device()) return rc;
    const dim3 grid = flat_grid(n * 32, 128);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == AZ_F32)
        k_chess_encode<float><<<grid, 128, 0, s>>>(reinterpret_cast<const Pos*>(pos), reinterpret_cast<const Pos*>(history), n,
                                                   static_cast<float*>(states_out));
    else
        k_chess_encode<__nv_bfloat16><<<grid, 128, 0, s>>>(reinterpret_cast<const Pos*>(pos),
                                                           reinterpret_cast<const Pos*>(history), n,
                                                           static_cast<__nv_bfloat16*>(states_out));
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_perft(const az_chess_pos* pos, int32_t n, int32_t depth, uint64_t* nodes_out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!pos || !nodes_out || n < 0 || depth < 0 || depth > kPerftMaxDepth)
        return az::fail_net(AZ_ERR_ARG, "az_chess_perft: bad argument (depth 0..8)");
    if (int rc = have_device()) return rc;
    k_chess_perft<<<flat_grid(n, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const Pos*>(pos), n, depth, reinterpret_cast<unsigned long long*>(nodes_out));
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

// ------------------------------------------------------------------------------------------ chess search engine
struct az_chess_engine {
    az_chess_config cfg;
    az_chess_layout lay;
    CEng eng;
};

static size_t ctake(size_t& off, size_t bytes) {
    size_t at = (off + 255) & ~(size_t)255;
    off = at + bytes;
    return at;
}

AZ_API void az_chess_struct_sizes(size_t* c, size_t* l) {
    if (c) *c = sizeof(az_chess_config);
    if (l) *l = sizeof(az_chess_layout);
}

AZ_API int az_chess_query_layout(const az_chess_config* c, az_chess_layout* L) {
    if (!c || !L) return az::fail_net(AZ_ERR_ARG, "az_chess_query_layout: null argument");
    if (c->abi_version != AZ_ABI_VERSION) return az::fail_net(AZ_ERR_ARG, "abi_version mismatch");
    if (c->n_trees < 1 || c->node_capacity < 256 || c->node_capacity > 0xffffff || c->sims_per_move < 1 ||
        c->max_free_sims < 1 || c->max_plies < 1 || c->sample_capacity < 1 || c->fin_capacity < 1 || c->pow_lut_len < 2)
        return az::fail_net(AZ_ERR_ARG,
                            "az_chess_config: n_trees / node_capacity (256 .. 2^24-1) / sims_per_move / max_free_sims / "
                            "max_plies / sample_capacity / fin_capacity / pow_lut_len out of range");
    if (c->eval_mode < AZ_EVAL_EXTERNAL || c->eval_mode > AZ_EVAL_HASH || c->move_mode < AZ_MOVE_ARGMAX ||
        c->move_mode > AZ_MOVE_PHILOX || c->prior_mode < AZ_PRIOR_F64 || c->prior_mode > AZ_PRIOR_F32)
        return az::fail_net(AZ_ERR_ARG, "az_chess_config: eval_mode / move_mode / prior_mode out of range");
    memset(L, 0, sizeof(*L));
    const size_t T = c->n_trees, C = c->node_capacity, S = c->sample_capacity, F = c->fin_capacity, P = c->max_plies;
    const size_t K = kMaxKids, D = kCDepth, PB = sizeof(Pos);
    size_t off = 0;
    L->status = ctake(off, 4 * T);
    L->ply = ctake(off, 4 * T);
    L->game_id = ctake(off, 8 * T);
    L->root_pos = ctake(off, PB * T);
    L->half = ctake(off, 4 * T);
    L->root_node = ctake(off, 4 * T);
    L->n_nodes = ctake(off, 4 * T);
    L->sims_done = ctake(off, 4 * T);
    L->pending = ctake(off, 4 * T);
    L->path_len = ctake(off, 4 * T);
    L->path = ctake(off, 4 * T * D);
    L->leaf_pos = ctake(off, PB * T);
    L->leaf_mask = ctake(off, 8 * T * 32);
    L->counters = ctake(off, 8 * T * 8);
    L->uniforms = ctake(off, 8 * T * P);
    L->node_a = ctake(off, 16 * T * 2 * C);
    L->node_p = ctake(off, 8 * T * 2 * C);
    L->node_m = ctake(off, 2 * T * 2 * C);
    L->smp_count = ctake(off, 16);
    L->smp_game = ctake(off, 8 * S);
    L->smp_ply = ctake(off, 4 * S);
    L->smp_pos = ctake(off, PB * S);
    L->smp_k = ctake(off, 4 * S);
    L->smp_act = ctake(off, 2 * S * K);
    L->smp_n = ctake(off, 4 * S * K);
    L->smp_choice = ctake(off, 4 * S);
    L->fin_count = ctake(off, 16);
    L->fin_game = ctake(off, 8 * F);
    L->fin_len = ctake(off, 4 * F);
    L->fin_result = ctake(off, 4 * F);
    L->pow_lut = ctake(off, 8 * (size_t)c->pow_lut_len);
    L->total_bytes = (off + 255) & ~(size_t)255;
    return AZ_OK;
}

AZ_API int az_chess_reset_games(az_chess_engine* e, void* stream) {
    if (!e) return az::fail_net(AZ_ERR_ARG, "null engine");
    k_chess_reset<<<flat_grid(e->eng.T * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(e->eng);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_set_game_id_base(az_chess_engine* e, int64_t game_id_base) {
    if (!e || game_id_base < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_set_game_id_base: bad argument");
    e->eng.game_base = game_id_base;
    return AZ_OK;
}

AZ_API int az_chess_engine_create(const az_chess_config* c, void* slab, size_t bytes, const double* host_lut, void* stream,
                                  az_chess_engine** out) {
    if (!out) return az::fail_net(AZ_ERR_ARG, "null out");
    *out = nullptr;
    az_chess_layout L;
    if (int rc = az_chess_query_layout(c, &L)) return rc;
    if (int rc = have_device()) return rc;
    if (!slab || !host_lut) return az::fail_net(AZ_ERR_ARG, "null slab / pow table");
    if (bytes < L.total_bytes || (reinterpret_cast<uintptr_t>(slab) & 255))
        return az::fail_net(AZ_ERR_SLAB, "slab too small or not 256-byte aligned");
    az_chess_engine* e = new (std::nothrow) az_chess_engine();
    if (!e) return az::fail_net(AZ_ERR_ARG, "out of host memory");
    e->cfg = *c;
    e->lay = L;
    char* b = static_cast<char*>(slab);
    CEng& g = e->eng;
    g.T = c->n_trees;
    g.C = c->node_capacity;
    g.S = c->sample_capacity;
    g.F = c->fin_capacity;
    g.P = c->max_plies;
    g.sims_target = c->sims_per_move;
    g.greedy_idx = c->index_move_greedy;
    g.eval_mode = c->eval_mode;
    g.prior_mode = c->prior_mode;
    g.move_mode = c->move_mode;
    g.max_free = c->max_free_sims;
    g.lut_len = c->pow_lut_len;
    g.auto_restart = c->auto_restart;
    g.c_puct = c->c_puct;
    g.seed = c->seed;
    g.game_base = c->game_id_base;
    g.games_target = c->games_target;
#define AZC_PTR(field, type) g.field = reinterpret_cast<type*>(b + L.field)
    AZC_PTR(status, int32_t);
    AZC_PTR(ply, int32_t);
    AZC_PTR(game_id, long long);
    AZC_PTR(root_pos, Pos);
    AZC_PTR(half, int32_t);
    AZC_PTR(root_node, int32_t);
    AZC_PTR(n_nodes, int32_t);
    AZC_PTR(sims_done, int32_t);
    AZC_PTR(pending, int32_t);
    AZC_PTR(path_len, int32_t);
    AZC_PTR(path, int32_t);
    AZC_PTR(leaf_pos, Pos);
    AZC_PTR(leaf_mask, u64);
    AZC_PTR(counters, long long);
    AZC_PTR(uniforms, double);
    AZC_PTR(node_a, NodeA);
    AZC_PTR(node_p, double);
    AZC_PTR(node_m, uint16_t);
    AZC_PTR(smp_count, int32_t);
    AZC_PTR(smp_game, long long);
    AZC_PTR(smp_ply, int32_t);
    AZC_PTR(smp_pos, Pos);
    AZC_PTR(smp_k, int32_t);
    AZC_PTR(smp_act, uint16_t);
    AZC_PTR(smp_n, int32_t);
    AZC_PTR(smp_choice, int32_t);
    AZC_PTR(fin_count, int32_t);
    AZC_PTR(fin_game, long long);
    AZC_PTR(fin_len, int32_t);
    AZC_PTR(fin_result, int32_t);
#undef AZC_PTR
    g.games_started = reinterpret_cast<unsigned long long*>(b + L.fin_count + 8);
    g.pow_lut = reinterpret_cast<const double*>(b + L.pow_lut);
    cudaError_t err = cudaMemcpyAsync(b + L.pow_lut, host_lut, 8 * (size_t)c->pow_lut_len, cudaMemcpyHostToDevice,
                                      static_cast<cudaStream_t>(stream));
    if (err == cudaSuccess) err = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));  // host_lut may be freed
    if (err != cudaSuccess) {
        delete e;
        return az::fail_net(AZ_ERR_CUDA, cudaGetErrorString(err));
    }
    *out = e;
    return az_chess_reset_games(e, stream);
}

AZ_API void az_chess_engine_destroy(az_chess_engine* e) { delete e; }

static inline dim3 ctree_grid(const az_chess_engine* e) { return dim3((unsigned)((e->eng.T + kCWarps - 1) / kCWarps)); }

AZ_API int az_chess_set_roots(az_chess_engine* e, const int32_t* ids, const az_chess_pos* positions, int32_t n, void* stream) {
    if (!e) return az::fail_net(AZ_ERR_ARG, "null engine");
    if (n == 0) return AZ_OK;
    if (!ids || !positions || n < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_set_roots: bad argument");
    k_chess_set_roots<<<flat_grid(n * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        e->eng, ids, reinterpret_cast<const Pos*>(positions), n);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_begin_search(az_chess_engine* e, int32_t sims, void* stream) {
    if (!e || sims < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_begin_search: bad argument");
    if (sims > 0) e->eng.sims_target = sims;
    k_chess_begin<<<flat_grid(e->eng.T, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(e->eng);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_search(az_chess_engine* e, void* stream) {
    if (!e) return az::fail_net(AZ_ERR_ARG, "null engine");
    if (e->eng.eval_mode == AZ_EVAL_EXTERNAL)
        return az::fail_net(AZ_ERR_ARG, "az_chess_search needs an in-kernel evaluator (eval_mode uniform / hash); use az_chess_step");
    k_chess_search<<<ctree_grid(e), kCWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(e->eng);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_step(az_chess_engine* e, const void* priors, const void* values, int32_t eval_dtype, void* states_out,
                         int32_t plane_stride, int32_t plane_first, int32_t* leaf_valid_out, void* stream) {
    if (!e || !leaf_valid_out) return az::fail_net(AZ_ERR_ARG, "az_chess_step: null argument");
    if (states_out && (plane_first < 0 || plane_first > 98 || plane_first % 14 || plane_stride < kPlanes - plane_first ||
                       plane_stride > 256))
        return az::fail_net(AZ_ERR_ARG, "az_chess_step: plane_first must be a multiple of 14 <= 98, plane_stride in [118 - plane_first, 256]");
    if ((priors == nullptr) != (values == nullptr)) return az::fail_net(AZ_ERR_ARG, "az_chess_step: priors and values go together");
    if (eval_dtype != AZ_F32 && eval_dtype != AZ_F64) return az::fail_net(AZ_ERR_ARG, "az_chess_step: eval_dtype must be AZ_F32 or AZ_F64");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int have = priors != nullptr;
    __nv_bfloat16* so = static_cast<__nv_bfloat16*>(states_out);
    static const int minb = [] {
        // 4 (default) = 112 registers, no spills, two waves of warps at 4096 trees; 7 = 72 registers with spills, one
        // wave.  Measured on B200 at 4096 trees x 800 simulations: 4.45 vs 4.43 M simulations/s (profiles/README.md).
        const char* v = getenv("AZ_CHESS_STEP_BLOCKS");
        return v ? atoi(v) : 4;
    }();
    const dim3 grid = ctree_grid(e);
#define AZC_STEP(PT, MINB)                                                                                              \
    k_chess_step<PT, MINB><<<grid, kCWarps * 32, 0, s>>>(e->eng, static_cast<const PT*>(priors), static_cast<const PT*>(values), \
                                                         have, so, plane_stride, plane_first, leaf_valid_out)
    if (eval_dtype == AZ_F32) {
        if (minb <= 4) AZC_STEP(float, 4); else AZC_STEP(float, 7);
    } else {
        if (minb <= 4) AZC_STEP(double, 4); else AZC_STEP(double, 7);
    }
#undef AZC_STEP
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_move(az_chess_engine* e, int32_t greedy_override, int32_t move_mode_override, void* stream) {
    if (!e) return az::fail_net(AZ_ERR_ARG, "null engine");
    const int mode = move_mode_override >= 0 ? move_mode_override : e->eng.move_mode;
    if (mode > AZ_MOVE_PHILOX) return az::fail_net(AZ_ERR_ARG, "az_chess_move: bad move mode");
    k_chess_move<<<ctree_grid(e), kCWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(e->eng, greedy_override, mode);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_rings_clear(az_chess_engine* e, void* stream) {
    if (!e) return az::fail_net(AZ_ERR_ARG, "null engine");
    k_chess_rings_clear<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(e->eng);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_decode_samples(const az_chess_pos* pos, const int32_t* k, const uint16_t* act, const int32_t* nv,
                                   const int32_t* choice, int32_t n, float* states_out, double* policies_out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!pos || !k || !act || !nv || !choice || !states_out || !policies_out || n < 0)
        return az::fail_net(AZ_ERR_ARG, "az_chess_decode_samples: bad argument");
    if (int rc = have_device()) return rc;
    k_chess_decode<<<flat_grid(n * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const Pos*>(pos), k, act, nv, choice, n, states_out, policies_out);
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}

AZ_API int az_chess_stem(const az_chess_pos* pos, int32_t n, const float* w_reduced, const float* cell_map, void* out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!pos || !w_reduced || !cell_map || !out || n < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_stem: bad argument");
    if (int rc = have_device()) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n < sms * 3 ? n : sms * 3;
    k_chess_stem<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const Pos*>(pos), n, w_reduced, cell_map,
                                                                      static_cast<__nv_bfloat16*>(out));
    AZC_CUDA(cudaGetLastError());
    return AZ_OK;
}
