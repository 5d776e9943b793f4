// az_gemm.cu - the tower's 1x1 projection shortcut (base_layers.py:105-113 of the reference) as a hand-written
// tcgen05 GEMM for sm_100a:   Y[rows][128] = X[rows][128] . W[128][128]^T   (bf16 in, float32 accumulate, bf16 out)
//
// The shortcut is bandwidth bound (33 kFLOP against 512 B of traffic per cell).  Measured on B200 at 172 032 cells with
// the input still in L2 (tools/explore_conv1x1.py): this kernel 14.2 us = 6.2 TB/s of algorithmic traffic (0.96 of the
// measured HBM peak), cuDNN's 1x1 convolution 12.9 us; inside the tower cuDNN's is 2 us faster per call, so the tower
// uses this kernel only on request (AZ_TC_SHORTCUT=1).  Design:
//   * persistent CTAs (2 per SM), the 32 KB weight matrix staged once per CTA in the canonical K-major SWIZZLE_128B
//     layout, activation tiles of 128 cells streamed with cp.async into a double buffer in the same layout
//     (16-byte chunk c of row r lands at chunk c ^ (r & 7) of its 128-byte row; 8-row groups are 1024 B apart);
//   * one elected thread issues 8 tcgen05.mma (M = 128, N = 128, K = 16 each) per tile into a 128-column TMEM
//     accumulator and commits to an mbarrier; the next tile's cp.async is already in flight underneath;
//   * epilogue: the four warps read their 32 TMEM lanes with tcgen05.ld (32x32b.x32), pack to bf16, stage the tile in
//     the shared-memory buffer the MMA has just finished with (XOR-swizzled, conflict free) and write it out with
//     fully coalesced 128-bit stores.
// No TMA descriptors are needed (rows are 256 B and contiguous), so nothing here depends on the driver API.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/az_b200.h"

namespace az {
int fail_net(int code, const char* msg);

namespace gemm {

constexpr int kC = 128;              // channels in and out (config.py:71)
constexpr int kTileM = 128;          // cells per tile = UMMA M
constexpr int kKBlockBytes = 16384;  // one 64-channel K block of a 128-row tile: 128 rows x 128 B
constexpr int kTileBytes = 2 * kKBlockBytes;
constexpr int kSmemBytes = 3 * kTileBytes + 1024;  // W + two activation buffers + alignment slack
constexpr uint32_t kTmemCols = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits 0-13,
// leading byte offset (unused for swizzled K-major) bits 16-29, stride byte offset = 1024 B between 8-row groups in
// bits 32-45, descriptor version 1 in bits 46-47, layout type 2 (SWIZZLE_128B) in bits 61-63.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A = B = BF16 (bits 7, 10), both K-major,
// N >> 3 in bits 17-22, M >> 4 in bits 24-28
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kC >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();  // a lost arrival must fail loudly, never hang the GPU
    }
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// position of 16-byte chunk c (0..15) of row r (0..127) inside a tile in the UMMA layout
__device__ __forceinline__ uint32_t umma_chunk_offset(int r, int c) {
    return (uint32_t)((c >> 3) * kKBlockBytes + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4));
}
// ... and inside the row-major staging tile of the epilogue (256 B rows, chunks XOR-ed with the row: conflict free)
__device__ __forceinline__ uint32_t stage_chunk_offset(int r, int c) { return (uint32_t)(r * 256 + ((c ^ (r & 15)) << 4)); }

__device__ __forceinline__ void load_tile_async(uint32_t sdst, const __nv_bfloat16* x, long long row0, long long rows, int tid) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = tid + 128 * i, r = idx >> 4, c = idx & 15;
        const long long row = row0 + r;
        const uint32_t dst = sdst + umma_chunk_offset(r, c);
        const __nv_bfloat16* src = x + (row < rows ? row : rows - 1) * kC + c * 8;
        const int bytes = row < rows ? 16 : 0;  // rows past the end are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

__global__ void __launch_bounds__(128, 2) k_conv1x1(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                    __nv_bfloat16* __restrict__ y, long long rows, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-byte aligned tiles
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW = base, sA0 = base + kTileBytes;
    const uint32_t bar = smem_u32(&s_bar);

    // one-time: weights -> UMMA layout (row = output channel, K = input channel), TMEM, barrier
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = tid + 128 * i, r = idx >> 4, c = idx & 15;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(w + r * kC) + c);
        *reinterpret_cast<uint4*>(gen + umma_chunk_offset(r, c)) = v;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    int tile = blockIdx.x, buf = 0;
    if (tile < n_tiles) load_tile_async(sA0, x, (long long)tile * kTileM, rows, tid);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t phase = 0;

    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const uint32_t sA = sA0 + buf * kTileBytes;
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        const int next = tile + gridDim.x;
        if (next < n_tiles) load_tile_async(sA0 + (buf ^ 1) * kTileBytes, x, (long long)next * kTileM, rows, tid);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int k = 0; k < kC / 16; ++k) {  // UMMA K = 16 bf16 = 32 B: step inside the 128-byte swizzle row
                const uint32_t koff = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                mma_bf16(tmem, umma_desc(sA + koff), umma_desc(sW + koff), k > 0);
            }
            // arrives on the barrier when the MMAs above have finished (implies tcgen05.fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // epilogue: thread = row (TMEM lane) warp * 32 + lane; 4 x 32 columns
        uint8_t* stage = gen + (sA - base);
        const int r = warp * 32 + lane;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 8 columns = one 16-byte chunk of bf16
                uint4 o;
                o.x = pack_bf16(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
                o.y = pack_bf16(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
                o.z = pack_bf16(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
                o.w = pack_bf16(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
                *reinterpret_cast<uint4*>(stage + stage_chunk_offset(r, cb * 4 + q)) = o;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();  // staging complete; TMEM reads done before the next tile's MMA overwrites the accumulator
        const long long row0 = (long long)tile * kTileM;
#pragma unroll
        for (int i = 0; i < 16; ++i) {  // a warp writes two whole rows (512 contiguous bytes) per instruction
            const int idx = tid + 128 * i, rr = idx >> 4, c = idx & 15;
            if (row0 + rr < rows)
                *(reinterpret_cast<uint4*>(y + (row0 + rr) * kC) + c) =
                    *reinterpret_cast<const uint4*>(stage + stage_chunk_offset(rr, c));  // default policy: the next convolution reads y from L2
        }
        // the staging buffer is refilled by cp.async two iterations from now, after the __syncthreads at the loop top
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTmemCols) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Dense head layers for wide action spaces (chess: Dense(1880) + softmax on the 128 policy features, Dense(256) + ReLU
// + Dense(1) + tanh on the 64 value features; model/tensorflow/model.py:86-103, 129-149) as ONE tcgen05 kernel.
// A tile is 128 positions = the 128 TMEM lanes, so a thread owns a whole output row: bias, ReLU, the value layer's
// second dot product and the softmax statistics (running maximum and sum over the 15 chunks of 128 logits) are
// per-thread scalars - no shuffles, no shared-memory reductions.  The policy GEMM is simply issued twice (it is 0.5
// GFLOP for 4096 positions): pass 1 collects max / sum, pass 2 writes exp(l - max) / sum, staged through shared
// memory so that the 30 MB of priors leave in coalesced 128-bit stores.  Replaces eleven library kernels (bf16 cast,
// cuBLAS GEMMs, bias add, float conversion, softmax, value MLP): 116 us -> see profiles/README.md.
constexpr int kHA_P = 0;                 // policy features  [128 rows][128 K] bf16, two K blocks
constexpr int kHA_V = 32768;             // value features   [128 rows][ 64 K] bf16
constexpr int kHW_V = 49152;             // value fc1        [256 rows][ 64 K] bf16
constexpr int kHW_0 = 81920;             // policy weight chunk, double buffered [128 rows][128 K] bf16
constexpr int kHStage = 147456;          // [128 rows][128] float32 staging of one output chunk
constexpr int kHeadsSmem = kHStage + 65536 + 1024;
constexpr uint32_t kHeadsTmemCols = 512;  // columns 0-127 policy chunk, 256-511 value hidden layer

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_i(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void commit_to(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

struct DenseHeadParams {
    const float* hd;           // [n][64][3] ReLU'd head-convolution outputs (az_net_head_convs)
    const __nv_bfloat16* wp;   // [n_chunks * 128][128] policy weights, rows beyond n_actions are zero
    const float* bp;           // [n_actions]
    const __nv_bfloat16* w1;   // [256][64] value fc1
    const float* b1;           // [256]
    const float* w2;           // [256] value fc2
    const float* b2;           // [1]
    float* priors;             // [n][n_actions]
    float* values;             // [n]
    float2* stats;             // [n][n_splits] partial softmax statistics (running maximum, sum of exp) per column split
    int n, n_actions, n_chunks, chunks_per_split, n_splits;
};

// grid = (row tiles of 128 positions, column splits).  PASS 0 computes this split's logits and leaves their (max, sum
// exp) in P.stats (split 0 also does the value head); PASS 1 combines the splits' statistics, recomputes the logits and
// writes the normalised priors.  Thread-per-row alone would be 4096 threads for 7.7 M logits: the column splits are
// what fills the 148 SMs.
template <int PASS>
__global__ void __launch_bounds__(128, 1) k_dense_heads(DenseHeadParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[8 * 128];  // chunks_per_split <= 8
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar = smem_u32(&s_bar);
    const long long row0 = (long long)blockIdx.x * 128;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)),
                     "n"(kHeadsTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // weights in flight: value fc1 (256 rows x 128 B) and the first policy chunk
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int idx = tid + 128 * i, r = idx >> 3, c = idx & 7;
        const uint32_t dst = base + kHW_V + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(P.w1 + r * 64 + c * 8) : "memory");
    }
    const int c_begin = blockIdx.y * P.chunks_per_split;
    const int c_end = min(P.n_chunks, c_begin + P.chunks_per_split);
    const bool do_value = PASS == 0 && blockIdx.y == 0;
    load_tile_async(base + kHW_0, P.wp, (long long)c_begin * 128, (long long)P.n_chunks * 128, tid);
    // this split's slice of the policy bias -> shared memory (read 128 times per chunk by every thread)
    for (int i = tid; i < (c_end - c_begin) * 128; i += 128) {
        const int col = c_begin * 128 + i;
        s_bias[i] = col < P.n_actions ? __ldg(P.bp + col) : 0.0f;
    }
    // features: hd [row][cell][plane] float32 -> bf16 operands (policy K index = cell * 2 + plane, the NHWC flatten
    // order of the Dense layer; value K index = cell).  Thread = row: twelve floats (four cells) make exactly one 16-byte
    // policy chunk and half a value chunk, so every shared-memory store is a 128-bit store at a compile-time position of
    // the row (the first version scattered 2-byte stores with idx / 192, rem / 3 arithmetic per element: half the kernel).
    {
        const long long grow = row0 + tid;
        const bool live = grow < P.n;
        const float4* src = reinterpret_cast<const float4*>(P.hd + (live ? grow : 0) * 192);
        uint8_t* prow = gen + kHA_P + (tid >> 3) * 1024 + (tid & 7) * 128;
        uint8_t* vrow = gen + kHA_V + (tid >> 3) * 1024 + (tid & 7) * 128;
        const int r7 = tid & 7;
#pragma unroll
        for (int g0 = 0; g0 < 16; g0 += 4) {
            float4 f[12];
#pragma unroll
            for (int u = 0; u < 12; ++u) f[u] = live ? __ldg(src + 3 * g0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                const int g = g0 + gg;  // cells 4g .. 4g + 3
                const float4 a = f[3 * gg], b = f[3 * gg + 1], c = f[3 * gg + 2];
                // a = (c0p0, c0p1, c0v, c1p0)  b = (c1p1, c1v, c2p0, c2p1)  c = (c2v, c3p0, c3p1, c3v)
                const uint4 pc = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.w, b.x), pack_bf16(b.z, b.w), pack_bf16(c.y, c.z));
                *reinterpret_cast<uint4*>(prow + (g >> 3) * kKBlockBytes + (((g & 7) ^ r7) << 4)) = pc;
                if (do_value) {
                    const uint2 vh = make_uint2(pack_bf16(a.z, b.y), pack_bf16(c.x, c.w));
                    *reinterpret_cast<uint2*>(vrow + ((((g >> 1) & 7) ^ r7) << 4) + (g & 1) * 8) = vh;
                }
            }
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t phase = 0;
    const int r = tid;
    const long long row = row0 + r;

    // ---- value head: hidden = relu(features . W1^T + b1) in TMEM columns 256-511, value = tanh(hidden . w2 + b2)
    if (do_value) {
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_bf16_i(tmem + 256, umma_desc(base + kHA_V + k * 32), umma_desc(base + kHW_V + k * 32), idesc_bf16(128, 256), k > 0);
        commit_to(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    {
        float acc = 0.0f;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + 256 + cb * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = cb * 32 + j;
                acc = fmaf(fmaxf(__uint_as_float(v[j]) + __ldg(P.b1 + col), 0.0f), __ldg(P.w2 + col), acc);
            }
        }
        if (row < P.n) P.values[row] = tanhf(acc + __ldg(P.b2));
    }
    }

    // ---- policy head: this split's chunks of 128 logits
    float m = -INFINITY, ssum = 0.0f, inv = 0.0f;
    if (PASS == 1 && row < P.n) {  // combine the splits' statistics of this row
        for (int sp = 0; sp < P.n_splits; ++sp) m = fmaxf(m, P.stats[row * P.n_splits + sp].x);
        for (int sp = 0; sp < P.n_splits; ++sp) {
            const float2 st = P.stats[row * P.n_splits + sp];
            if (st.x > -INFINITY) ssum += st.y * __expf(st.x - m);
        }
        inv = 1.0f / ssum;
        ssum = 0.0f;
    }
    int counter = 0;
    const int total = c_end - c_begin;
    constexpr int pass = PASS;
    {
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c, ++counter) {
            const uint32_t sW = base + kHW_0 + (uint32_t)(counter & 1) * kTileBytes;
            asm volatile("cp.async.wait_all;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();  // this chunk's weights landed; everybody is done with the previous accumulator and staging
            if (counter + 1 < total)
                load_tile_async(base + kHW_0 + (uint32_t)((counter + 1) & 1) * kTileBytes, P.wp, (long long)(c + 1) * 128,
                                (long long)P.n_chunks * 128, tid);
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t koff = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                    mma_bf16_i(tmem, umma_desc(base + kHA_P + koff), umma_desc(sW + koff), idesc_bf16(128, 128), k > 0);
                }
                commit_to(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int col0 = c * 128;
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_base + cb * 32, v);
                float l[32];
                float bmax = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = col0 + cb * 32 + j;
                    l[j] = col < P.n_actions ? __uint_as_float(v[j]) + s_bias[col - c_begin * 128] : -INFINITY;
                    bmax = fmaxf(bmax, l[j]);
                }
                if (pass == 0) {
                    if (bmax > -INFINITY) {
                        const float mn = fmaxf(m, bmax);
                        float part = 0.0f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) part += __expf(l[j] - mn);
                        ssum = ssum * __expf(m - mn) + part;
                        m = mn;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 o;
                        o.x = __expf(l[4 * q + 0] - m) * inv;
                        o.y = __expf(l[4 * q + 1] - m) * inv;
                        o.z = __expf(l[4 * q + 2] - m) * inv;
                        o.w = __expf(l[4 * q + 3] - m) * inv;
                        const int ch = cb * 8 + q;
                        *reinterpret_cast<float4*>(gen + kHStage + r * 512 + ((ch ^ (r & 31)) << 4)) = o;
                    }
                }
            }
            if (pass == 1) {
                __syncthreads();
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {  // one warp instruction = one whole row of the chunk (512 contiguous bytes)
                    const int idx = tid + 128 * i, rr = idx >> 5, ch = idx & 31;
                    const int col = col0 + ch * 4;
                    if (row0 + rr < P.n && col < P.n_actions)
                        *reinterpret_cast<float4*>(P.priors + (row0 + rr) * P.n_actions + col) =
                            *reinterpret_cast<const float4*>(gen + kHStage + rr * 512 + ((ch ^ (rr & 31)) << 4));
                }
            }
        }
    }
    if (PASS == 0 && row < P.n) P.stats[row * P.n_splits + blockIdx.y] = make_float2(m, ssum);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kHeadsTmemCols) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Chess stem from the 64-byte boards on tcgen05 (the product-path version of az_chess.cu's k_chess_stem).  On the
// self-play path only 20 of the 118 input planes vary (current entry + scalars; az_chess.cu explains why), so the stem is
// out[cell][ch] = relu(cmap[cell][ch] + sum_{tap, plane < 24} x[cell + tap][plane] * w[ch][tap * 24 + plane]):
// an implicit GEMM with K = 9 x 24 = 216 (padded to 256).  A tile is two positions = 128 cells = the 128 TMEM lanes.
//   * the reduced weights [128][256] bf16 stay resident in shared memory (K-major SWIZZLE_128B, 64 KB);
//   * per tile the CTA first writes each cell's 24 planes as three 16-byte chunks (from the bitboards: one-hot piece
//     plane, castling flags, clocks), then builds the im2col tile: row = cell, chunks 3 * tap .. 3 * tap + 2 = the
//     neighbour's three chunks (zeros off the board) - 128-bit shared-memory copies into the swizzled layout;
//   * one thread issues 16 tcgen05.mma (128 x 128 x 16) into TMEM; the A tile is double buffered so the next tile is
//     built while the tensor core runs; epilogue as in k_conv1x1 (+ the per-cell constant, ReLU), 8 warps.
constexpr int kStemK = 256;
constexpr int kStemATile = 4 * kKBlockBytes;        // [128 rows][256 K] bf16 = 64 KB
constexpr int kStemSmem = 3 * kStemATile + 2 * 64 * 48 + 1024;  // weights + two A tiles + the two positions' cell chunks

struct StemTcParams {
    const uint64_t* pos;       // [n][8] az_chess_pos
    const __nv_bfloat16* w;    // [128][256] reduced stem weights, K = tap * 24 + plane, zero beyond 216
    const float* cmap;         // [2][64][128]: bias + initial-position contribution per cell; [1] = bias only (fresh ply-0 root)
    __nv_bfloat16* out;        // [n][64][128]
    int n;
};

__device__ __forceinline__ int stem_piece_plane(const uint64_t* p, int sq) {
    const uint64_t s = 1ull << sq;
    int v = 0;
    if (p[0] & s) v = 1;
    else if (p[1] & s) v = 2;
    else if (p[2] & s) v = 3;
    else if (p[3] & s) v = 4;
    else if (p[4] & s) v = 5;
    else if (p[5] & s) v = 6;
    if (v && !(p[6] & s)) v = 13 - v;  // black pieces: np.eye(13)[negative] wraps (chess/board.py:50-56)
    return v;
}

template <int H>
__device__ __forceinline__ void im2col_half(uint8_t* A, const uint4* cellchunks, int r) {
    const int which = r >> 6, cell = r & 63, cy = cell >> 3, cx = cell & 7, r7 = r & 7;
    uint8_t* row = A + (r >> 3) * 1024 + r7 * 128;
    const uint4* cc = cellchunks + which * 64 * 3;
#pragma unroll
    for (int g = 0; g < (H ? 13 : 14); ++g) {
        const int gc = H * 14 + g, tap = gc / 3, j = gc - tap * 3, dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int y = cy + dy, x = cx + dx;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)y < 8u && (unsigned)x < 8u) v = cc[(y * 8 + x) * 3 + j];
        *reinterpret_cast<uint4*>(row + (gc >> 3) * kKBlockBytes + (((gc & 7) ^ r7) << 4)) = v;
    }
}

__global__ void __launch_bounds__(256, 1) k_chess_stem_tc(StemTcParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW = base, sA0 = base + kStemATile;
    uint4* cellchunks = reinterpret_cast<uint4*>(gen + 3 * kStemATile);  // [2 positions][64 cells][3 chunks]
    __shared__ int s_fresh[2][2];  // [A buffer][position of the tile]: un-mirrored ply-0 root (seven empty history entries)
    const uint32_t bar = smem_u32(&s_bar);
    const int n_tiles = (P.n + 1) / 2;

    // weights: [128 rows][256 K] -> four K blocks of the UMMA layout; both A tiles start as zeros (their K padding stays)
    for (int idx = tid; idx < 128 * 32; idx += 256) {  // cp.async: in flight underneath the first tile's build
        const int r = idx >> 5, c = idx & 31;          // 32 chunks of 8 K per row
        const uint32_t dst = sW + (uint32_t)((c >> 3) * kKBlockBytes + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(P.w + r * kStemK + c * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int idx = tid; idx < 2 * kStemATile / 16; idx += 256)
        reinterpret_cast<uint4*>(gen + kStemATile)[idx] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t phase = 0;

    // the boards of a tile are fetched one tile ahead (registers), so their HBM latency is off the per-tile critical path
    uint64_t q[8];
    bool q_valid = false;
    auto fetch = [&](int tile) {
        q_valid = false;
        if (tid < 128 && tile < n_tiles) {
            const long long pidx = 2LL * tile + (tid >> 6);
            if (pidx < P.n) {
                q_valid = true;
#pragma unroll
                for (int i = 0; i < 8; ++i) q[i] = __ldg(P.pos + pidx * 8 + i);
            }
        }
    };
    // builds the im2col tile of the fetched boards in A buffer `buf` (all 256 threads)
    auto build = [&](int buf) {
        // (1) the 24 planes of every cell of the two positions as three 16-byte chunks
        if (tid < 128) {
            const int which = tid >> 6, cell = tid & 63;
            uint32_t w[12] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (cell == 0)  // the start position (white set = ranks 1-2, ...) with a zero halfmove clock, white to move
                s_fresh[buf][which] = q_valid && q[0] == 0x00ff00000000ff00ull && q[1] == 0x4200000000000042ull &&
                                      q[2] == 0x2400000000000024ull && q[3] == 0x8100000000000081ull && q[4] == 0x0800000000000008ull &&
                                      q[5] == 0x1000000000000010ull && q[6] == 0xffffull && ((q[7] >> 16) & 0xffff) == 0 &&
                                      !((q[7] >> 11) & 1);
            if (q_valid) {
                const int sq = ((7 - (cell >> 3)) << 3) | (cell & 7);  // array row 0 is rank 8
                const int pl = stem_piece_plane(q, sq);
                const uint64_t meta = q[7];
                const bool black = (meta >> 11) & 1;
                const int own_k = black ? 4 : 1, own_q = black ? 8 : 2, opp_k = black ? 1 : 4, opp_q = black ? 2 : 8;
                const uint32_t one = 0x3f80u;
#pragma unroll
                for (int wd = 0; wd < 7; ++wd) w[wd] = (pl >> 1) == wd ? (one << (16 * (pl & 1))) : 0u;
                w[7] = ((meta & own_q) ? one : 0u) | ((meta & own_k) ? one << 16 : 0u);
                w[8] = ((meta & opp_q) ? one : 0u) | ((meta & opp_k) ? one << 16 : 0u);
                w[9] = pack_bf16((float)((meta >> 32) & 0xffff), (float)((meta >> 16) & 0xffff));  // fullmove, halfmove
            }
            uint4* cc = cellchunks + (which * 64 + cell) * 3;
            cc[0] = make_uint4(w[0], w[1], w[2], w[3]);
            cc[1] = make_uint4(w[4], w[5], w[6], w[7]);
            cc[2] = make_uint4(w[8], w[9], w[10], w[11]);
        }
        __syncthreads();
        // (2) row r = (position, cell), chunk 3 * tap + j <- chunk j of the neighbour cell (zeros off the board).  Warps 0-3
        // copy chunks 0-13 of rows 0-127, warps 4-7 chunks 14-26: every index below is a compile-time constant after
        // unrolling (the first version spent two thirds of the kernel's instructions on idx / 27, gc / 3, tap / 3 ...)
        uint8_t* A = gen + kStemATile + buf * kStemATile;
        if (tid < 128) im2col_half<0>(A, cellchunks, tid);
        else im2col_half<1>(A, cellchunks, tid - 128);
    };

    int tile = blockIdx.x, buf = 0;
    fetch(tile);
    if (tile < n_tiles) build(0);
    fetch(tile + gridDim.x);
    asm volatile("cp.async.wait_all;\n" ::: "memory");  // the weights
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();  // A[buf] complete; the previous epilogue is done with TMEM and its staging buffer
        const uint32_t sA = sA0 + buf * kStemATile;
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int k = 0; k < kStemK / 16; ++k) {
                const uint32_t koff = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                mma_bf16_i(tmem, umma_desc(sA + koff), umma_desc(sW + koff), idesc_bf16(128, 128), k > 0);
            }
            commit_to(bar);
        }
        const int next = tile + gridDim.x;
        if (next < n_tiles) build(buf ^ 1);  // while the tensor core works on this tile (contains a __syncthreads)
        fetch(next + gridDim.x);
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // epilogue: warp w reads TMEM lanes 32 * (w & 3), columns 64 * (w >> 2) .. + 63 and parks the raw float32
        // accumulators in this tile's A buffer (the MMA has finished reading it; 128 rows x 512 B, chunks XOR-ed with the
        // row: conflict free).  The per-cell constant is added in the coalesced pass below - read per row it would cost
        // 32 cache lines per load instruction.
        uint8_t* stage = gen + kStemATile + buf * kStemATile;
        const int r = (warp & 3) * 32 + lane, ch0 = (warp >> 2) * 64;
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(ch0 + cb * 32), v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int ch4 = (ch0 >> 2) + cb * 8 + q;  // 16-byte chunk of four float32
                *reinterpret_cast<uint4*>(stage + r * 512 + ((ch4 ^ (r & 31)) << 4)) =
                    make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        const long long row0 = (long long)tile * 128, rows = (long long)P.n * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // 16 lanes = one output row of 256 B: coalesced stores, coalesced constant loads
            const int idx = tid + 256 * i, rr = idx >> 4, c = idx & 15;
            const float4 a0 = *reinterpret_cast<const float4*>(stage + rr * 512 + (((2 * c) ^ (rr & 31)) << 4));
            const float4 a1 = *reinterpret_cast<const float4*>(stage + rr * 512 + (((2 * c + 1) ^ (rr & 31)) << 4));
            const float4* cm = reinterpret_cast<const float4*>(P.cmap + (s_fresh[buf][rr >> 6] ? 64 * 128 : 0) + (rr & 63) * 128 + c * 8);
            const float4 c0 = __ldg(cm), c1 = __ldg(cm + 1);
            uint4 o;
            o.x = pack_bf16(fmaxf(a0.x + c0.x, 0.f), fmaxf(a0.y + c0.y, 0.f));
            o.y = pack_bf16(fmaxf(a0.z + c0.z, 0.f), fmaxf(a0.w + c0.w, 0.f));
            o.z = pack_bf16(fmaxf(a1.x + c1.x, 0.f), fmaxf(a1.y + c1.y, 0.f));
            o.w = pack_bf16(fmaxf(a1.z + c1.z, 0.f), fmaxf(a1.w + c1.w, 0.f));
            if (row0 + rr < rows) *(reinterpret_cast<uint4*>(P.out + (row0 + rr) * 128) + c) = o;
        }
        __syncthreads();
        // the float32 staging covered the whole A buffer: its K padding (chunks 27-31 of every row) must be zero again
        // before the buffer is an operand the next time; the 27 real chunks per row are rewritten by the next build
        if (tid < 128) {
            uint8_t* prow = stage + 3 * kKBlockBytes + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
            for (int gc = 27; gc < 32; ++gc) *reinterpret_cast<uint4*>(prow + (((gc & 7) ^ (tid & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTmemCols) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Connect-N stem on tcgen05: Conv3x3(4 -> 128) + BN + ReLU (model/tensorflow/model.py:36-46) on the [n][H][W][4] bf16 leaf
// batch az_step writes - the tensor-core replacement of az_net_stem's mma.sync kernel.  K = 9 taps x 4 planes = 36
// (one 64-wide K block); a 128-row tile holds floor(128 / cells) whole positions (3 at 6x7); a cell's four planes are one
// 8-byte word, so an im2col chunk (two taps) is two neighbour words (zeros off the board).
constexpr int kC4StemSmem = 16384 + 2 * 16384 + 32768 + 2 * 1024 + 1024;

struct C4StemParams {
    const uint2* in;           // [n][cells] four bf16 planes per cell
    const __nv_bfloat16* w;    // [128][64], K index = tap * 4 + plane, zero beyond 36
    const float* bias;         // [128]
    __nv_bfloat16* out;        // [n][cells][128]
    int n, H, W, cells, ppt;   // ppt = positions per tile
};

__global__ void __launch_bounds__(256, 2) k_stem_tc(C4StemParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW = base, sA0 = base + 16384;
    uint8_t* stage = gen + 3 * 16384;
    uint2* cellbuf = reinterpret_cast<uint2*>(gen + 3 * 16384 + 32768);  // [2 buffers][128 cells]
    const uint32_t bar = smem_u32(&s_bar);
    const int rows_used = P.ppt * P.cells;
    const int n_tiles = (P.n + P.ppt - 1) / P.ppt;

    for (int idx = tid; idx < 128 * 8; idx += 256) {  // weights: one K block
        const int r = idx >> 3, c = idx & 7;
        const uint32_t dst = sW + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(P.w + r * 64 + c * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (tid < 128) s_bias[tid] = P.bias[tid];
    for (int idx = tid; idx < 2 * 16384 / 16; idx += 256) reinterpret_cast<uint4*>(gen + 16384)[idx] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t phase = 0;

    // this thread's im2col row is the same in every tile: row r = tid & 127, chunks 0-2 (taps 0-5) for warps 0-3,
    // chunks 3-4 (taps 6-8) for warps 4-7
    const int r = tid & 127, half = tid >> 7, r7 = r & 7;
    const int lp = r / P.cells, cell = r - lp * P.cells, cy = cell / P.W, cx = cell - cy * P.W;
    const bool row_live = r < rows_used;
    int nb[6];  // neighbour indices into cellbuf for this thread's taps (-1 = off the board / dead row)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int tap = half * 6 + i, y = cy + tap / 3 - 1, x = cx + tap % 3 - 1;
        nb[i] = (row_live && tap < 9 && (unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W) ? lp * P.cells + y * P.W + x : -1;
    }
    uint2 pre = make_uint2(0u, 0u);  // this thread's cell of the next tile, fetched one tile ahead
    auto fetch = [&](int tile) {
        pre = make_uint2(0u, 0u);
        if (tid < rows_used && tile < n_tiles) {
            const long long g = (long long)tile * rows_used + tid;
            if (g < (long long)P.n * P.cells) pre = __ldg(P.in + g);
        }
    };
    auto build = [&](int buf) {
        uint2* cb = cellbuf + buf * 128;
        if (tid < 128) cb[tid] = pre;
        __syncthreads();
        uint8_t* row = gen + 16384 + buf * 16384 + (r >> 3) * 1024 + r7 * 128;
        const uint2 z = make_uint2(0u, 0u);
        if (half == 0) {
#pragma unroll
            for (int gc = 0; gc < 3; ++gc) {
                const uint2 a = nb[2 * gc] >= 0 ? cb[nb[2 * gc]] : z, b = nb[2 * gc + 1] >= 0 ? cb[nb[2 * gc + 1]] : z;
                *reinterpret_cast<uint4*>(row + ((gc ^ r7) << 4)) = make_uint4(a.x, a.y, b.x, b.y);
            }
        } else {
#pragma unroll
            for (int gc = 3; gc < 5; ++gc) {
                const uint2 a = nb[2 * (gc - 3)] >= 0 ? cb[nb[2 * (gc - 3)]] : z, b = nb[2 * (gc - 3) + 1] >= 0 ? cb[nb[2 * (gc - 3) + 1]] : z;
                *reinterpret_cast<uint4*>(row + ((gc ^ r7) << 4)) = make_uint4(a.x, a.y, b.x, b.y);
            }
        }
    };

    int tile = blockIdx.x, buf = 0;
    fetch(tile);
    if (tile < n_tiles) build(0);
    fetch(tile + gridDim.x);
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        const uint32_t sA = sA0 + buf * 16384;
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_i(tmem, umma_desc(sA + k * 32), umma_desc(sW + k * 32), idesc_bf16(128, 128), k > 0);
            commit_to(bar);
        }
        const int next = tile + gridDim.x;
        if (next < n_tiles) build(buf ^ 1);
        fetch(next + gridDim.x);
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int er = (warp & 3) * 32 + lane, ch0 = (warp >> 2) * 64;
#pragma unroll
        for (int cb2 = 0; cb2 < 2; ++cb2) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(ch0 + cb2 * 32), v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* bb = s_bias + ch0 + cb2 * 32 + 8 * q;
                uint4 o;
                o.x = pack_bf16(fmaxf(__uint_as_float(v[8 * q + 0]) + bb[0], 0.f), fmaxf(__uint_as_float(v[8 * q + 1]) + bb[1], 0.f));
                o.y = pack_bf16(fmaxf(__uint_as_float(v[8 * q + 2]) + bb[2], 0.f), fmaxf(__uint_as_float(v[8 * q + 3]) + bb[3], 0.f));
                o.z = pack_bf16(fmaxf(__uint_as_float(v[8 * q + 4]) + bb[4], 0.f), fmaxf(__uint_as_float(v[8 * q + 5]) + bb[5], 0.f));
                o.w = pack_bf16(fmaxf(__uint_as_float(v[8 * q + 6]) + bb[6], 0.f), fmaxf(__uint_as_float(v[8 * q + 7]) + bb[7], 0.f));
                *reinterpret_cast<uint4*>(stage + stage_chunk_offset(er, (ch0 >> 3) + cb2 * 4 + q)) = o;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        const long long row0 = (long long)tile * rows_used, rows = (long long)P.n * P.cells;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + 256 * i, rr = idx >> 4, c = idx & 15;
            if (rr < rows_used && row0 + rr < rows)
                *(reinterpret_cast<uint4*>(P.out + (row0 + rr) * 128) + c) = *reinterpret_cast<const uint4*>(stage + stage_chunk_offset(rr, c));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

}  // namespace gemm
}  // namespace az

extern "C" __attribute__((visibility("default"))) int az_net_conv1x1(const void* x, const void* w, int64_t rows, int32_t channels,
                                                                      void* y, void* stream) {
    using namespace az::gemm;
    if (rows == 0) return AZ_OK;
    if (!x || !w || !y || rows < 0) return az::fail_net(AZ_ERR_ARG, "az_net_conv1x1: bad argument");
    if (channels != kC) return az::fail_net(AZ_ERR_ARG, "az_net_conv1x1: built for 128 filters (config.py:71)");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15)
        return az::fail_net(AZ_ERR_ARG, "az_net_conv1x1: pointers must be 16-byte aligned");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_conv1x1, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
            return az::fail_net(AZ_ERR_CUDA, "az_net_conv1x1: shared memory request refused");
        configured = true;
    }
    const long long n_tiles = (rows + kTileM - 1) / kTileM;
    long long grid = n_tiles < 2LL * sms ? n_tiles : 2LL * sms;
    k_conv1x1<<<(unsigned)grid, 128, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w), static_cast<__nv_bfloat16*>(y), rows,
        (int)n_tiles);
    if (cudaGetLastError() != cudaSuccess) return az::fail_net(AZ_ERR_CUDA, "az_net_conv1x1: launch failed");
    return AZ_OK;
}

extern "C" __attribute__((visibility("default"))) int az_net_dense_heads(const float* hd, const void* policy_w, const float* policy_b,
                                                                          const void* value1_w, const float* value1_b,
                                                                          const float* value2_w, const float* value2_b, int32_t n,
                                                                          int32_t cells, int32_t n_actions, float* priors,
                                                                          float* values, float* stats_scratch, void* stream) {
    using namespace az::gemm;
    if (n == 0) return AZ_OK;
    if (!hd || !policy_w || !policy_b || !value1_w || !value1_b || !value2_w || !value2_b || !priors || !values ||
        !stats_scratch || n < 0)
        return az::fail_net(AZ_ERR_ARG, "az_net_dense_heads: bad argument");
    if (cells != 64 || n_actions < 1 || n_actions > 4096 || (n_actions & 3))
        return az::fail_net(AZ_ERR_ARG, "az_net_dense_heads: built for 64 cells (8x8 boards) and a multiple of 4 actions <= 4096");
    if ((reinterpret_cast<uintptr_t>(policy_w) | reinterpret_cast<uintptr_t>(value1_w) | reinterpret_cast<uintptr_t>(priors)) & 15)
        return az::fail_net(AZ_ERR_ARG, "az_net_dense_heads: weights and priors must be 16-byte aligned");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_dense_heads<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadsSmem) != cudaSuccess ||
            cudaFuncSetAttribute(k_dense_heads<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadsSmem) != cudaSuccess)
            return az::fail_net(AZ_ERR_CUDA, "az_net_dense_heads: shared memory request refused");
        configured = true;
    }
    if (reinterpret_cast<uintptr_t>(stats_scratch) & 7) return az::fail_net(AZ_ERR_ARG, "az_net_dense_heads: scratch must be 8-byte aligned");
    const int n_chunks = (n_actions + 127) / 128, n_splits = AZ_DENSE_HEAD_SPLITS;
    const int cps = (n_chunks + n_splits - 1) / n_splits;  // <= 8 since n_actions <= 4096
    DenseHeadParams P{hd, static_cast<const __nv_bfloat16*>(policy_w), policy_b, static_cast<const __nv_bfloat16*>(value1_w),
                      value1_b, value2_w, value2_b, priors, values, reinterpret_cast<float2*>(stats_scratch), n, n_actions,
                      n_chunks, cps, n_splits};
    const dim3 grid((unsigned)((n + 127) / 128), (unsigned)n_splits);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    k_dense_heads<0><<<grid, 128, kHeadsSmem, s>>>(P);
    k_dense_heads<1><<<grid, 128, kHeadsSmem, s>>>(P);
    if (cudaGetLastError() != cudaSuccess) return az::fail_net(AZ_ERR_CUDA, "az_net_dense_heads: launch failed");
    return AZ_OK;
}

extern "C" __attribute__((visibility("default"))) int az_chess_stem_tc(const void* pos, int32_t n, const void* w_reduced_bf16,
                                                                        const float* cell_map, void* out, void* stream) {
    using namespace az::gemm;
    if (n == 0) return AZ_OK;
    if (!pos || !w_reduced_bf16 || !cell_map || !out || n < 0) return az::fail_net(AZ_ERR_ARG, "az_chess_stem_tc: bad argument");
    if ((reinterpret_cast<uintptr_t>(w_reduced_bf16) | reinterpret_cast<uintptr_t>(cell_map) | reinterpret_cast<uintptr_t>(out)) & 15)
        return az::fail_net(AZ_ERR_ARG, "az_chess_stem_tc: pointers must be 16-byte aligned");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_chess_stem_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem) != cudaSuccess)
            return az::fail_net(AZ_ERR_CUDA, "az_chess_stem_tc: shared memory request refused");
        configured = true;
    }
    const int n_tiles = (n + 1) / 2;
    StemTcParams P{static_cast<const uint64_t*>(pos), static_cast<const __nv_bfloat16*>(w_reduced_bf16), cell_map,
                   static_cast<__nv_bfloat16*>(out), n};
    k_chess_stem_tc<<<n_tiles < sms ? n_tiles : sms, 256, kStemSmem, static_cast<cudaStream_t>(stream)>>>(P);
    if (cudaGetLastError() != cudaSuccess) return az::fail_net(AZ_ERR_CUDA, "az_chess_stem_tc: launch failed");
    return AZ_OK;
}

extern "C" __attribute__((visibility("default"))) int az_net_stem_tc(const void* states, const void* w_bf16, const float* bias, int32_t n,
                                                                      int32_t H, int32_t W, int32_t channels, void* out, void* stream) {
    using namespace az::gemm;
    if (n == 0) return AZ_OK;
    if (!states || !w_bf16 || !bias || !out || n < 0 || H < 1 || W < 1) return az::fail_net(AZ_ERR_ARG, "az_net_stem_tc: bad argument");
    if (channels != kC) return az::fail_net(AZ_ERR_ARG, "az_net_stem_tc: built for 128 filters (config.py:71)");
    if (H * W > 128) return az::fail_net(AZ_ERR_ARG, "az_net_stem_tc: a position must fit one 128-row tile (H * W <= 128)");
    if ((reinterpret_cast<uintptr_t>(states) & 7) || ((reinterpret_cast<uintptr_t>(w_bf16) | reinterpret_cast<uintptr_t>(out)) & 15))
        return az::fail_net(AZ_ERR_ARG, "az_net_stem_tc: misaligned pointer");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_stem_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kC4StemSmem) != cudaSuccess)
            return az::fail_net(AZ_ERR_CUDA, "az_net_stem_tc: shared memory request refused");
        configured = true;
    }
    const int cells = H * W, ppt = 128 / cells, n_tiles = (n + ppt - 1) / ppt;
    C4StemParams P{static_cast<const uint2*>(states), static_cast<const __nv_bfloat16*>(w_bf16), bias,
                   static_cast<__nv_bfloat16*>(out), n, H, W, cells, ppt};
    k_stem_tc<<<n_tiles < 2 * sms ? n_tiles : 2 * sms, 256, kC4StemSmem, static_cast<cudaStream_t>(stream)>>>(P);
    if (cudaGetLastError() != cudaSuccess) return az::fail_net(AZ_ERR_CUDA, "az_net_stem_tc: launch failed");
    return AZ_OK;
}
