// mma_rate.cu - how fast does the tensor pipe of one SM (or SM pair) run a stream of 128 x N x 16 bf16 tcgen05.mma
// instructions whose operands already sit in shared memory, as a function of the operand layout?  No loads, no epilogue:
// one elected thread issues `count` MMAs (4 per "stage", stages rotate over a ring like the weight ring of az_tower.cu,
// A steps over the chunk columns like a filter tap's window), commits, waits.  Prints cycles per MMA (clock64 inside
// the kernel, max over CTAs) for each variant.  Timing experiment only: the operands are whatever is in shared memory.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

struct Variant {
    const char* name;
    int pair;         // cta_group::2
    int n;            // UMMA N
    uint32_t a_lbo;   // bytes between the two 8-channel chunk columns of a K = 16 step (no-swizzle) ...
    uint32_t a_kstep; // ... and between K = 16 steps (bytes)
    uint32_t a_half;  // ... and between the two 64-channel halves
    uint32_t a_shift; // start of the A window in bytes (row shift of a filter tap)
    uint32_t b_lbo, b_kstep, b_stage;  // B: the same, and bytes between ring slots
    uint32_t a_hi, b_hi;               // high descriptor words (SBO, version, layout type)
    int two_acc;      // alternate between two accumulators per stage
};

template <int PAIR>
__global__ void __launch_bounds__(128, 1) k_rate(Variant v, int count, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_done;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    // something non-trivial in the operands (denormal-free bf16 patterns)
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ ((uint32_t)(i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    long long t = 0;
    if (warp == 1 && rank == 0) {
        if (threadIdx.x == 32) {
            const uint32_t base = smem_u32(smem);
            const uint32_t a_base = base + v.a_shift, b_base = base + 163840u;  // B ring in the upper 64 KB
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(v.n >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
            uint32_t a_lo[8], b_lo[8], dd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // two stages of four K steps: channel half 0 / 1, ring slots rotate over four
                const uint32_t a = a_base + (uint32_t)(j >> 2) * v.a_half + (uint32_t)(j & 3) * v.a_kstep;
                a_lo[j] = ((a & 0x3ffff) >> 4) | ((v.a_lbo >> 4) << 16);
                dd[j] = tmem + ((v.two_acc && (j >> 2)) ? 256u : 0u);
            }
            const long long t0 = clock64();
            for (int it = 0; it < count / 8; ++it) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t b = b_base + (uint32_t)((2 * it + (j >> 2)) & 3) * v.b_stage + (uint32_t)(j & 3) * v.b_kstep;
                    b_lo[j] = ((b & 0x3ffff) >> 4) | ((v.b_lbo >> 4) << 16);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t acc = it > 0 ? 1u : 0u;
                    if (PAIR)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\t"
                            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(dd[j]),
                            "r"(a_lo[j]), "r"(b_lo[j]), "r"(idesc), "r"(acc), "r"(v.a_hi), "r"(v.b_hi)
                            : "memory");
                    else
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %6};\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(dd[j]),
                            "r"(a_lo[j]), "r"(b_lo[j]), "r"(idesc), "r"(acc), "r"(v.a_hi), "r"(v.b_hi)
                            : "memory");
                }
            }
            if (PAIR)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(&s_done)), "h"((uint16_t)3) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_done)) : "memory");
            mbar_wait(smem_u32(&s_done), 0);
            t = clock64() - t0;
            cycles[blockIdx.x] = t;
        }
    } else if (PAIR && rank != 0 && threadIdx.x == 32) {
        mbar_wait(smem_u32(&s_done), 0);  // keep the peer (and its shared memory / TMEM) alive until the MMAs have finished
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync();
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
    }
}


// ---- latencies of the hops a weight-ring handshake is made of (cycles, one CTA pair, nothing else running)
template <bool TEST>
__device__ __forceinline__ void wait_poll(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        if (TEST)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void arrive_local(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar), "r"(cta) : "memory");
}
// mode 0: one MMA -> commit -> wait (commit latency); 1: ping-pong between two warps of a CTA; 2: between the two CTAs of a
// pair (remote arrive); 3: 8 KB cp.async.bulk from global (L2) -> wait; 4: one pair MMA -> multicast commit -> wait
template <bool TEST>
__global__ void __launch_bounds__(128, 1) k_hops(int mode, int rounds, const uint8_t* src, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_b[2];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&s_b[0]), 1);
        mbar_init(smem_u32(&s_b[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem, b0 = smem_u32(&s_b[0]), b1 = smem_u32(&s_b[1]);
    const uint32_t kNone = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = ((smem_u32(smem) & 0x3ffff) >> 4) | ((2048u >> 4) << 16), b_lo = (((smem_u32(smem) + 8192) & 0x3ffff) >> 4) | ((2048u >> 4) << 16);
    if ((mode == 0 || mode == 4) && threadIdx.x == 32) {
        if (rank == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((mode == 4 ? 256 : 128) >> 4) << 24);
            const long long t0 = clock64();
            for (int i = 0; i < rounds; ++i) {
                if (mode == 4) {
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(0u), "r"(kNone) : "memory");
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(b0), "h"((uint16_t)3) : "memory");
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(0u), "r"(kNone) : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(b0) : "memory");
                }
                wait_poll<TEST>(b0, i & 1);
            }
            out[0] = clock64() - t0;
        } else if (mode == 4) {
            for (int i = 0; i < rounds; ++i) wait_poll<TEST>(b0, i & 1);
        }
    } else if ((mode == 5 || mode == 6) && rank == 0 && warp == 1 && (mode == 6 || threadIdx.x == 32)) {
        // polls of a phase that has already completed (parity 1 of a fresh barrier): one lane, or the whole warp
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < rounds; ++i) {
            uint32_t done;
            if (TEST)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(b0 + (acc & 8u)), "r"(1u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(b0 + (acc & 8u)), "r"(1u) : "memory");
            acc += done << 3;  // the next poll's address depends on this answer
        }
        if (threadIdx.x == 32) out[0] = clock64() - t0 + (acc == 12345u);
    } else if (mode == 1 && rank == 0 && (threadIdx.x == 0 || threadIdx.x == 32)) {
        const long long t0 = clock64();
        for (int i = 0; i < rounds; ++i) {
            if (warp == 0) { arrive_local(b0); wait_poll<TEST>(b1, i & 1); }
            else { wait_poll<TEST>(b0, i & 1); arrive_local(b1); }
        }
        if (warp == 0) out[0] = clock64() - t0;
    } else if (mode == 2 && threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < rounds; ++i) {
            if (rank == 0) { arrive_remote(b0, 1u); wait_poll<TEST>(b1, i & 1); }
            else { wait_poll<TEST>(b0, i & 1); arrive_remote(b1, 0u); }
        }
        if (rank == 0) out[0] = clock64() - t0;
    } else if (mode == 3 && rank == 0 && threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < rounds; ++i) {
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(b0), "r"(8192u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem) + 16384u),
                         "l"(src + (size_t)(i & 63) * 8192), "r"(8192u), "r"(b0) : "memory");
            wait_poll<TEST>(b0, i & 1);
        }
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

template <bool TEST>
static void run_hops(const uint8_t* d_src, long long* d_cyc) {
    const char* names[7] = {"one MMA (64 cycles) -> commit -> wait          ", "arrive -> wait, two warps of one CTA (per hop) ",
                            "remote arrive -> wait, two CTAs of a pair (hop)", "8 KB cp.async.bulk from L2 -> wait             ",
                            "one pair MMA -> multicast commit -> wait       ",
                            "poll of a completed phase, one lane, dependent ", "poll of a completed phase, whole warp, dependent"};
    cudaFuncSetAttribute(k_hops<TEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode = 0; mode < 7; ++mode) {
        const int rounds = 2000;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 64 * 1024;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        long long best = 1ll << 60;
        for (int rep = 0; rep < 3; ++rep) {
            cudaError_t err = cudaLaunchKernelEx(&cfg, k_hops<TEST>, mode, rounds, d_src, d_cyc);
            if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
                printf("hops mode %d FAILED: %s\n", mode, cudaGetErrorString(cudaGetLastError()));
                return;
            }
            long long c;
            cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
            best = c < best ? c : best;
        }
        const double per = (double)best / rounds / ((mode == 1 || mode == 2) ? 2.0 : 1.0);
        printf("  %s %s  %7.1f cycles\n", TEST ? "test_wait spin" : "try_wait      ", names[mode], per);
    }
}

int main() {
    const uint32_t kNone = (128u >> 4) | (1u << 14);                  // SBO 128 B (rows linear at 16 B), version 1, no swizzle
    const uint32_t kSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t kLboA = 622 * 16;                                  // az_tower.cu: chunk columns of 622 rows
    std::vector<Variant> vs = {
        // name, pair, N, a_lbo, a_kstep, a_half, a_shift, b_lbo, b_kstep, b_stage, a_hi, b_hi, two_acc
        {"g1 N128 none(tower layout)           ", 0, 128, kLboA, 2 * kLboA, 8 * kLboA, 352, 2048, 4096, 16384, kNone, kNone, 0},
        {"g1 N128 none, window shifted 21 rows ", 0, 128, kLboA, 2 * kLboA, 8 * kLboA, 352 + 336, 2048, 4096, 16384, kNone, kNone, 0},
        {"g1 N128 none, A compact (LBO 2048)   ", 0, 128, 2048, 4096, 16384, 0, 2048, 4096, 16384, kNone, kNone, 0},
        {"g1 N128 A sw128, B none              ", 0, 128, 16, 32, 16384, 0, 2048, 4096, 16384, kSw128, kNone, 0},
        {"g1 N128 A sw128, B sw128             ", 0, 128, 16, 32, 16384, 0, 16, 32, 16384, kSw128, kSw128, 0},
        {"g1 N128 A sw128 shifted 3 rows, B sw ", 0, 128, 16, 32, 16384, 384, 16, 32, 16384, kSw128, kSw128, 0},
        {"g1 N256 A sw128, B sw128             ", 0, 256, 16, 32, 16384, 0, 16, 32, 8192, kSw128, kSw128, 0},
        {"g1 N128 none, two accumulators       ", 0, 128, kLboA, 2 * kLboA, 8 * kLboA, 352, 2048, 4096, 16384, kNone, kNone, 1},
        {"g2 N128 none(tower layout)           ", 1, 128, kLboA, 2 * kLboA, 8 * kLboA, 352, 1024, 2048, 8192, kNone, kNone, 0},
        {"g2 N128 none, window shifted 21 rows ", 1, 128, kLboA, 2 * kLboA, 8 * kLboA, 352 + 336, 1024, 2048, 8192, kNone, kNone, 0},
        {"g2 N128 none, A compact              ", 1, 128, 2048, 4096, 16384, 0, 1024, 2048, 8192, kNone, kNone, 0},
        {"g2 N128 A sw128, B none              ", 1, 128, 16, 32, 16384, 0, 1024, 2048, 8192, kSw128, kNone, 0},
        {"g2 N128 A sw128, B sw128             ", 1, 128, 16, 32, 16384, 0, 16, 32, 8192, kSw128, kSw128, 0},
        {"g2 N256 none(tower layout A)         ", 1, 256, kLboA, 2 * kLboA, 8 * kLboA, 352, 2048, 4096, 16384, kNone, kNone, 0},
        {"g2 N256 A sw128, B sw128             ", 1, 256, 16, 32, 16384, 0, 16, 32, 16384, kSw128, kSw128, 0},
        {"g2 N128 none, two accumulators       ", 1, 128, kLboA, 2 * kLboA, 8 * kLboA, 352, 1024, 2048, 8192, kNone, kNone, 1},
        {"g2 N64  none(tower layout)           ", 1, 64, kLboA, 2 * kLboA, 8 * kLboA, 352, 512, 1024, 4096, kNone, kNone, 0},
        {"g1 N64  none(tower layout)           ", 0, 64, kLboA, 2 * kLboA, 8 * kLboA, 352, 1024, 2048, 8192, kNone, kNone, 0},
    };
    const int count = 4000, smem_bytes = 225 * 1024;
    cudaFuncSetAttribute(k_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    long long* d_cyc;
    cudaMalloc(&d_cyc, 148 * sizeof(long long));
    for (int grid : {148}) {
        printf("grid %d CTAs, %d MMAs per issuing thread\n", grid, count);
        for (const Variant& v : vs) {
            float best_ms = 1e9f;
            long long worst = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaMemset(d_cyc, 0, 148 * sizeof(long long));
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(128);
                cfg.dynamicSmemBytes = smem_bytes;
                cudaLaunchAttribute attr;
                attr.id = cudaLaunchAttributeClusterDimension;
                attr.val.clusterDim.x = v.pair ? 2 : 1;
                attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
                cfg.attrs = &attr;
                cfg.numAttrs = 1;
                cudaEventRecord(e0);
                cudaError_t err = v.pair ? cudaLaunchKernelEx(&cfg, k_rate<1>, v, count, d_cyc) : cudaLaunchKernelEx(&cfg, k_rate<0>, v, count, d_cyc);
                cudaEventRecord(e1);
                if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
                    printf("%s FAILED: %s\n", v.name, cudaGetErrorString(cudaGetLastError()));
                    return 1;
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best_ms = ms < best_ms ? ms : best_ms;
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), d_cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                worst = 0;
                for (long long c : h) worst = c > worst ? c : worst;
            }
            printf("  %s  %7.1f cycles per MMA   (ideal %3d; kernel %.3f ms)\n", v.name, (double)worst / count, v.n / 2, best_ms);
        }
    }
    {   // sustained: the pair variant in the tower's layout launched back to back for about a second, per-CTA rates of the last launch
        const Variant& v = vs[8];
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        for (int phase = 0; phase < 2; ++phase) {
            const int reps = phase == 0 ? 1 : 6000;
            for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, k_rate<1>, v, count, d_cyc);
            cudaDeviceSynchronize();
            std::vector<long long> h(148);
            cudaMemcpy(h.data(), d_cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            printf("%s: per-pair cycles per MMA after %d back-to-back launches:", v.name, reps);
            for (int i = 0; i < 148; i += 2) printf(" %.1f", (double)h[i] / count);
            printf("\n");
        }
    }
    uint8_t* d_src;
    cudaMalloc(&d_src, 64 * 8192);
    cudaMemset(d_src, 0, 64 * 8192);
    printf("handshake hops\n");
    run_hops<false>(d_src, d_cyc);
    run_hops<true>(d_src, d_cyc);
    return 0;
}
