// az_tower.cu - the whole residual tower of the policy/value net (base_layers.py:85-125 of the reference, four times:
// model/tensorflow/model.py:48-66) as ONE persistent tcgen05 kernel for sm_100a.
//
//   per block:   h = relu(conv3x3(x, w1) + b1)        y = relu(conv3x3(h, w2) + conv1x1(x, wp) + b2p)
//
// cuDNN runs this as 12 kernels that stream 817 MB through DRAM per 4096-position forward (profiles/
// advance_traffic_r1.json).  Here a CTA takes a tile of whole positions (3 boards of 6x7 = 126 of the 128 rows of a UMMA
// tile, 2 boards of 8x8 = 128) through all blocks: activations never leave shared memory / TMEM, the only DRAM traffic
// is the tile in (bf16 [cells][128]) and out, and the 2.4 MB of weights stream from L2 through a bulk-copy ring.
//
// Implicit GEMM without an im2col buffer.  Activations live in shared memory in the canonical K-major *no-swizzle*
// UMMA layout with the rows linear: 16-byte chunk kc (8 channels) of row r at  kc * LBO + r * 16.  With the rows of a
// tile ordered (y, position, x) a filter tap (dy, dx) of a 3x3 convolution is the SAME buffer read through a descriptor
// whose start address is moved by  (dy * positions * W + dx) * 16  bytes:
//   * rows above / below the board fall into zero padding rows before / after the tile (22 each);
//   * a tap that leaves the board sideways lands on the neighbouring cell in row order, which is always an x = W-1
//     cell (dx = -1) or an x = 0 cell (dx = +1): the dx = -1 taps read a copy of the activations whose x = W-1 rows
//     are zero, the dx = +1 taps a copy whose x = 0 rows are zero.  The epilogue that produces a layer's output writes
//     the three copies (plain, left-masked, right-masked), so no tap ever needs its own im2col tile.
// Four activation buffers (x, x-left, x-right, h) of 128 + 22 rows share one row space: 159 KB.
//
// Roles (320 threads, 1 CTA per SM, persistent over tiles):
//   warp 0   producer: the whole warp runs the loop, one elected lane streams the weight stages (16 KB = one tap x 64 input
//            channels x 128 output channels, pre-packed by the host in the UMMA layout, in the order they are consumed: per
//            convolution the input-channel half is the outer loop, the tap the inner one) from global memory with
//            cp.async.bulk into a 4-deep ring (pairs: 8 x 8 KB), full / empty mbarriers.  It must hand over a stage every
//            256 cycles: everything in its loop stays on the uniform datapath (see the comment there);
//   warp 1   MMA issuer: ONE elected lane runs the role and issues 4 tcgen05.mma (128 x 128 x 16) per stage from one block of
//            PTX that also polls the next stage's barrier (issue_stage); conv1 -> TMEM columns 0-127, the 1x1 shortcut
//            (needs only x) and then conv2 -> columns 128-255, so the shortcut runs under epilogue 1;
//   warps 2-9  epilogue: tcgen05.ld (32 lanes x 32 columns), + bias, ReLU + bf16 pair in one cvt.rn.relu.bf16x2, up to three
//            16-byte stores per chunk (the masked copies never write their zero rows, which stay zero from the start;
//            conflict free: a warp writes 512 contiguous bytes), fence.proxy.async, arrive on the "activations ready"
//            barrier.  The same warps load a tile from global memory at its start and store it after the last block.
//   warps 10-11 (az_net_forward_trees only): evaluator-free simulations of the engine's trees that have no leaf in flight
//            (free_sims_tree, az_tree.cuh) - the kernel owns its SM but issues in ~15 % of its cycles.
// SASS: UTCHMMA (tcgen05.mma), UBLKCP (cp.async.bulk), LDTM (tcgen05.ld).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/az_b200.h"
#include "az_tree.cuh"

namespace az {
int fail_net(int code, const char* msg);

namespace tower {

constexpr int kC = 128;                           // filters (config.py:71)
constexpr int kPad = 22;                          // zero rows either side of a tile: >= positions * W + 1
constexpr int kTileRows = 128;                    // UMMA M
constexpr int kBufRows = kTileRows + kPad;        // a buffer = its tile rows + the pad that follows them
constexpr int kNBuf = 4;                          // x, x-left-masked, x-right-masked, h
constexpr int kRows = kPad + kNBuf * kBufRows;    // 622 rows of 16 bytes per chunk column
constexpr int kLboA = kRows * 16;                 // byte distance between adjacent 8-channel chunk columns
constexpr int kActBytes = 16 * kLboA;             // 159 232
constexpr int kStageBytes = 16384;                // [8 chunks][128 output channels][8 input channels] bf16
constexpr int kLboB = 2048;                       // 128 rows x 16 bytes
constexpr int kStages = 4;                        // 64 KB of weight ring: 4 x 16 KB, or 8 x 8 KB per CTA of a pair.  (Tried 3 / 6
                                                  // slots so that a block of the per-tree kernel of ANOTHER tree group fits beside
                                                  // this CTA: 3 % slower alone, and --groups 2..4 gained nothing - one 4-warp block
                                                  // per SM cannot carry a group's trees within a net launch.  DESIGN.md section 8.)
constexpr int kMaxDepth = 4;                      // config.py:63
constexpr int kBiasBytes = ((1 + 2 * kMaxDepth) * kC + 3 * kC) * 4;  // stem + tower biases, the 1x1 head convolutions
constexpr int kSmemBytes = kActBytes + kStages * kStageBytes + kBiasBytes;
constexpr int kStagesPerBlock = 38;               // conv1: 9 taps x 2, shortcut: 2, conv2: 9 taps x 2
constexpr int kThreads = 320;
constexpr int kTreeWarps = 2;                     // az_net_forward_trees: warps 10-13 run evaluator-free simulations beside the net
constexpr uint32_t kTmemCols = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start >> 4 in
// bits 0-13, leading byte offset (between the two 8-element core matrices of one K = 16 step) bits 16-29, stride byte
// offset (between 8-row groups; 128 = rows linear at 16 bytes) bits 32-45, version 1 bits 46-47, layout type 0.
// The issuing thread keeps only the low word per operand (address + LBO) and steps it by constants: ncu showed the
// first version issue-bound on 64-bit descriptor arithmetic (93 cycles of instructions per 64-cycle MMA).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr & 0x3ffff) >> 4) | ((lbo_bytes >> 4) << 16);
}
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1
// D = F32 (bit 4), A = B = BF16 (bits 7, 10), both K-major, N >> 3 in bits 17-22, M >> 4 in bits 24-28
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(kIdesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
// The same MMA across a CTA pair (cta_group::2, M = 256): each CTA contributes its own 128 rows of A and 64 of the 128 rows
// of B from the same shared-memory offsets, and receives its 128 rows of D in its own TMEM.  Per SM that is 4 + 2 KB of
// operand reads and 2 KB of weight refill per MMA instead of 4 + 4 + 4: ncu showed the single-CTA kernel bound by
// shared-memory bandwidth at 88 cycles per 64-cycle MMA.
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(kIdesc2), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair when the MMAs issued so far have finished
__device__ __forceinline__ void commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at offset `bar` in the shared memory of CTA `cta` of the cluster.  Default semantics, as in
// CUTLASS's ClusterBarrier::arrive(cta_id): the .release.cluster / .acquire.cluster forms compile to MEMBAR.ALL.GPU and
// CCTL.IVALL around every stage (the first pair kernel ran 2.4x slower than the single-CTA one because of them); the data
// these barriers guard lives in shared memory, which no cache shadows.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ uint32_t opaque(uint32_t v) {  // a value the compiler must keep in a register, not rematerialise
    asm volatile("mov.u32 %0, %0;\n" : "+r"(v));
    return v;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(is_leader));
    return is_leader != 0;
}
__device__ __forceinline__ void commit_to(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes)
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
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
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
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
}
// Pins 32 registers behind the preceding volatile asm (tcgen05.wait::ld): no use of them may be scheduled above it.
__device__ __forceinline__ void pin32(uint32_t (&v)[32]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                      "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
    asm volatile("" : "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                      "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}
// relu(hi) : relu(lo) as one bf16x2 word, round to nearest even (one instruction instead of two FMNMX + F2FP)
__device__ __forceinline__ uint32_t pack_relu_bf16(float lo, float hi) {
    uint32_t v;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(v) : "f"(hi), "f"(lo));
    return v;
}

// One weight stage of the tower as ONE block of PTX: the poll of the NEXT stage's "full" barrier is issued first, the four
// K = 16 MMAs and the commit that frees this stage's slot follow, and only then is the poll's answer read - so the ~70 cycles
// a poll takes (tools/mma_rate.cu) overlap the issue of the MMAs instead of preceding them.  Before this the issuing warp
// spent 222 of its 392 cycles per stage in two serial polls (counters of az_net_tower_timing: never once was a stage
// missing).  `leader` = 1 in the one elected lane that runs the MMA role (the predicate is kept so that the block can also be
// executed by a whole warp with one issuing lane).  Returns the poll's answer.
template <bool PAIR>
__device__ __forceinline__ uint32_t issue_stage(uint32_t d_tmem, uint32_t a_lo, uint32_t a_step, uint32_t b_lo, uint32_t b_step,
                                                uint32_t first_accumulate, uint32_t empty_bar, uint32_t next_full_bar,
                                                uint32_t next_parity, uint32_t leader) {
    uint32_t ready;
    if (PAIR)
        asm volatile(
            "{\n\t.reg .pred pn, pl, pa, pt;\n\t.reg .b64 da, db;\n\t.reg .b32 ra, rb;\n\t"
            "setp.ne.b32 pl, %10, 0;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 pn, [%8], %9;\n\t"
            "mov.b64 da, {%2, %12};\n\tmov.b64 db, {%4, %12};\n\t"
            "@pl tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %11, pa;\n\t"
            "add.u32 ra, %2, %3;\n\tadd.u32 rb, %4, %5;\n\tmov.b64 da, {ra, %12};\n\tmov.b64 db, {rb, %12};\n\t"
            "@pl tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %11, pt;\n\t"
            "add.u32 ra, ra, %3;\n\tadd.u32 rb, rb, %5;\n\tmov.b64 da, {ra, %12};\n\tmov.b64 db, {rb, %12};\n\t"
            "@pl tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %11, pt;\n\t"
            "add.u32 ra, ra, %3;\n\tadd.u32 rb, rb, %5;\n\tmov.b64 da, {ra, %12};\n\tmov.b64 db, {rb, %12};\n\t"
            "@pl tcgen05.mma.cta_group::2.kind::f16 [%1], da, db, %11, pt;\n\t"
            "@pl tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%7], %13;\n\t"
            "selp.u32 %0, 1, 0, pn;\n\t}\n"
            : "=r"(ready)
            : "r"(d_tmem), "r"(a_lo), "r"(a_step), "r"(b_lo), "r"(b_step), "r"(first_accumulate), "r"(empty_bar), "r"(next_full_bar),
              "r"(next_parity), "r"(leader), "r"(kIdesc2), "r"(kDescHi), "h"((uint16_t)3)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred pn, pl, pa, pt;\n\t.reg .b64 da, db;\n\t.reg .b32 ra, rb;\n\t"
            "setp.ne.b32 pl, %10, 0;\n\tsetp.ne.b32 pa, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 pn, [%8], %9;\n\t"
            "mov.b64 da, {%2, %12};\n\tmov.b64 db, {%4, %12};\n\t"
            "@pl tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %11, pa;\n\t"
            "add.u32 ra, %2, %3;\n\tadd.u32 rb, %4, %5;\n\tmov.b64 da, {ra, %12};\n\tmov.b64 db, {rb, %12};\n\t"
            "@pl tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %11, pt;\n\t"
            "add.u32 ra, ra, %3;\n\tadd.u32 rb, rb, %5;\n\tmov.b64 da, {ra, %12};\n\tmov.b64 db, {rb, %12};\n\t"
            "@pl tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %11, pt;\n\t"
            "add.u32 ra, ra, %3;\n\tadd.u32 rb, rb, %5;\n\tmov.b64 da, {ra, %12};\n\tmov.b64 db, {rb, %12};\n\t"
            "@pl tcgen05.mma.cta_group::1.kind::f16 [%1], da, db, %11, pt;\n\t"
            "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
            "selp.u32 %0, 1, 0, pn;\n\t}\n"
            : "=r"(ready)
            : "r"(d_tmem), "r"(a_lo), "r"(a_step), "r"(b_lo), "r"(b_step), "r"(first_accumulate), "r"(empty_bar), "r"(next_full_bar),
              "r"(next_parity), "r"(leader), "r"(kIdesc), "r"(kDescHi)
            : "memory");
    return ready;
}

struct HeadParams {
    const float* conv_w;    // [3][128] 1x1 head convolutions (rows 0-1 policy, row 2 value), BN folded
    const float* conv_b;    // [3]
    const float* policy_w;  // [A][2 * cells] Dense(A) on the NHWC-flattened policy planes
    const float* policy_b;  // [A]
    const float* value1_w;  // [256][cells] Dense(256)
    const float* value1_b;  // [256]
    const float* value2_w;  // [256] Dense(1)
    const float* value2_b;  // [1]
    float* priors;          // [n][A] softmax
    float* values;          // [n] tanh
    int A;
};

struct TowerParams {
    const __nv_bfloat16* x;    // tower mode: [n][cells][128] stem output; net mode: [n][cells][4] leaf planes (az_step)
    const uint8_t* w_img;      // stages of 16 KB in consumption order (az_b200/net.py: pack_tower_weights / pack_stem_weights):
                               // net mode 3 stem stages first, then [depth][38]
    const float* bias;         // [depth][2][128]: conv1 bias, conv2 bias + shortcut bias
    const float* stem_bias;    // net mode: [128]
    __nv_bfloat16* y;          // tower mode: [n][cells][128]
    HeadParams heads;          // net mode
    const int32_t* index;      // net mode, optional: position i of the batch is tree index[i] (az_step_gather) ...
    const int32_t* count;      // ... and the batch holds *count positions (read on the device)
    int n, W, cells, ppt, depth, n_tiles;
    long long* timing;         // AZ_TOWER_DEBUG bit 3: per CTA {cycles total, MMA warp waiting for activations, for weights, epilogue warp 2
                               // waiting for the accumulator, its body} (az_net_tower_timing)
    Eng eng;                   // TREES: the engine whose trees without a pending leaf go on simulating beside the net ...
    int tree_cap;              // ... up to this many evaluator-free simulations per tree and launch
    unsigned long long* timeline;  // az_net_debug_timeline: {first CTA start, last CTA end} of this launch, globaltimer ns
    int debug;                 // timing experiments only (AZ_TOWER_DEBUG): bit 0 = do not refill weight stages after the first ring pass,
                               // bit 1 = epilogue skips its shared-memory stores, bit 2 = every tap reads the unshifted centre buffer,
                               // bit 4 = no epilogue at all and the MMA warp never waits for activations: the tensor pipe and the
                               // weight ring alone (results are garbage); with bit 4: bit 5 = no weight ring either (producer and
                               // peer forwarder idle, the MMA warp never waits for a stage)
};

constexpr int kStemStages = 3;     // 9 taps x [128 output channels][16 K: 4 planes + zeros] = 4 taps per 16 KB stage
constexpr int kHeadMaxCells = 48;  // the value layer keeps one weight row per thread in registers
constexpr int kHidden = 256;       // Dense(256) of the value head (model.py:129-139) = the 256 epilogue threads
static_assert(kThreads - 64 == kHidden, "one hidden unit of the value head per epilogue thread");

// byte offset of (buffer, row 0, chunk 0) inside the activation area
__device__ __forceinline__ uint32_t buf_row0(int buf) { return (uint32_t)(kPad + buf * kBufRows) * 16u; }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }  // the 8 epilogue warps

// NET = false: the residual tower alone (x -> y, both [cells][128] bf16).
// NET = true : the whole policy/value net of the reference (model/tensorflow/model.py:21-188) for boards whose leaf
//              planes are [cells][4]: stem Conv3x3(4 -> 128) as nine K = 16 MMAs on the same shifted-window scheme,
//              tower, then both heads in the last epilogue (1x1 convolutions from the registers that hold the tower's
//              output, dense layers with one hidden unit per epilogue thread).  Nothing but 8 bytes per cell goes in
//              and A + 1 floats per position come out.
// PAIR = true : two CTAs of a cluster (one SM pair) run as one: cta_group::2 MMAs issued by rank 0 over both CTAs' tiles,
//              each CTA streams half of every weight stage (8 KB, 8-deep ring), the peer's warp 1 forwards "my half has
//              landed" to rank 0, epilogue warps of both CTAs arrive on rank 0's barriers, commits are multicast.
// TREES = true (net mode, 6x7 engines): kTreeWarps more warps per CTA (az_net_forward_trees).  The net kernel leaves 85 % of its
//              issue slots idle and owns its SM (227 KB of shared memory), so nothing can run BESIDE it; these warps run
//              INSIDE it: every tree of P.eng that has no leaf in flight - its last simulations ended in terminal leaves
//              and used up az_step's max_free_sims - goes on simulating (free_sims_tree, az_tree.cuh) instead of waiting
//              for the next az_step.  The serial tail of the tree step moves under the net.
template <bool NET, bool PAIR, bool TREES = false>
__global__ void __launch_bounds__(kThreads + (TREES ? kTreeWarps * 32 : 0), 1) k_tower(TowerParams P) {
    constexpr int kBlockThreads = kThreads + (TREES ? kTreeWarps * 32 : 0);
    __shared__ PathScratch s_path[TREES ? kTreeWarps : 1];
    __shared__ long long s_pub0;  // az_net_tower_timing: clock at which epilogue warp 2 last announced a first channel half
    __shared__ int s_stop;  // TREES: raised by warp 1 when this CTA starts its last tile - the tree warps start nothing new
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kRing = PAIR ? 2 * kStages : kStages;             // ring slots
    constexpr uint32_t kSlotBytes = PAIR ? kStageBytes / 2 : kStageBytes;  // bytes of a weight stage this CTA holds
    constexpr uint32_t kLboW = PAIR ? kLboB / 2 : kLboB;            // 64 or 128 rows x 16 bytes per chunk column
    __shared__ __align__(8) uint64_t s_full[2 * kStages], s_empty[2 * kStages], s_acc, s_act[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (P.timeline && tid == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        atomicMin(P.timeline, t_);
    }
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    // work units: a tile per CTA, or a pair of tiles per CTA pair (rank r takes the r-th tile of the unit)
    const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int kTpu = PAIR ? 2 : 1;
    const uint32_t act = smem_u32(smem), stages = act + kActBytes;
    float* s_bias = reinterpret_cast<float*>(smem + kActBytes + kStages * kStageBytes);  // [stem][depth][2][128]
    float* s_headw = s_bias + (1 + 2 * kMaxDepth) * kC;                                   // [3][128] head convolutions
    const uint32_t bar_acc = smem_u32(&s_acc), bar_act0 = smem_u32(&s_act[0]), bar_act1 = smem_u32(&s_act[1]);
    const int rowstride = P.ppt * P.W, rows_used = P.ppt * P.cells;
    const int stages_per_tile = kStagesPerBlock * P.depth + (NET ? kStemStages : 0);
    if (NET && P.count) {  // gathered batch: every thread reads the same device-side count before anything else
        const int n = min(max(__ldg(P.count), 0), P.n);
        P.n = n;
        P.n_tiles = (n + P.ppt - 1) / P.ppt;
    }

    // one-time: zero the activation area (pads and dead rows stay zero for ever), biases, barriers, TMEM
    for (int i = tid; i < kActBytes / 16; i += kBlockThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < P.depth * 2 * kC; i += kBlockThreads) s_bias[kC + i] = P.bias[i];
    if (NET) {
        for (int i = tid; i < kC; i += kBlockThreads) s_bias[i] = P.stem_bias[i];
        // the existing heads kernel feeds the 1x1 convolutions to the tensor cores as bf16: same rounding here
        for (int i = tid; i < 3 * kC; i += kBlockThreads) s_headw[i] = __bfloat162float(__float2bfloat16(P.heads.conv_w[i]));
    }
    if (tid == 0) {
        s_stop = 0;
        for (int s = 0; s < kRing; ++s) {
            // a stage is there when this CTA's bulk copy has landed; for rank 0 of a pair: ... and the peer's warp 1 has
            // announced its half (one barrier, so the issuing warp polls ONCE per stage)
            mbar_init(smem_u32(&s_full[s]), (PAIR && rank == 0) ? 2 : 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(bar_acc, 1);
        mbar_init(bar_act0, kTpu * (kThreads - 64) / 32);  // one arrival per epilogue warp (of both CTAs of a pair)
        mbar_init(bar_act1, kTpu * (kThreads - 64) / 32);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync();  // the peer's barriers exist and its buffers are zeroed before anything reaches across
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ---------------------------------------------------------------- producer
        // The whole warp runs the loop (the poll is warp-uniform) and one elected lane issues from a predicated block of
        // PTX, so that the loop stays on the uniform datapath.  The first version ran inside `if (lane == 0)`: the compiler
        // then wraps the bulk copy in a broadcast loop (ELECT / R2UR.BROADCAST / BRA.U.ANY) and re-derives the barrier
        // addresses from SR_CgaCtaId every stage - 378 cycles per stage (clock stamps, tools/time_tower.py flag 64), which
        // was what bounded the whole kernel: the tensor pipe needs a stage every 256 cycles.
        const uint32_t lead = elect_one() ? 1u : 0u;
        // (opaque copies: otherwise the compiler re-derives these addresses from SR_CgaCtaId in every iteration)
        const uint32_t full0 = opaque(smem_u32(&s_full[0])), empty0 = opaque(smem_u32(&s_empty[0])), ring0 = opaque(stages);
        uint32_t cnt = 0;
        for (int unit = unit0; unit * kTpu < P.n_tiles && !(P.debug & 32); unit += n_units) {
            const uint8_t* src = P.w_img + rank * kSlotBytes;  // pair: rank r streams output channels 64 r .. 64 r + 63
            for (int s = 0; s < stages_per_tile; ++s, ++cnt, src += kStageBytes) {
                const uint32_t slot = cnt % kRing, k = cnt / kRing;
                mbar_wait(empty0 + slot * 8u, (k & 1) ^ 1);
                if ((P.debug & 64) && P.timing && blockIdx.x == 0 && cnt < 2048 && lane == 0) P.timing[148 * 8 + cnt] = clock64();
                if ((P.debug & 1) && cnt >= (uint32_t)kRing) {
                    if (lead) mbar_arrive(full0 + slot * 8u);
                    __syncwarp();
                    continue;
                }
                asm volatile(
                    "{\n\t.reg .pred pl;\n\tsetp.ne.b32 pl, %4, 0;\n\t"
                    "@pl mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %3;\n\t"
                    "@pl cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2], %3, [%0];\n\t}\n" ::"r"(
                        full0 + slot * 8u),
                    "r"(ring0 + slot * kSlotBytes), "l"(src), "r"(kSlotBytes), "r"(lead)
                    : "memory");
            }
        }
    } else if (warp == 1 && PAIR && rank != 0) {
        // ---------------------------------------------------------------- peer CTA: tell rank 0 that my half of a stage is here
        uint32_t cnt = 0;
        const uint32_t lead = elect_one() ? 1u : 0u;
        const uint32_t full0 = opaque(smem_u32(&s_full[0]));
        uint32_t remote_full0;  // the same barriers in rank 0's shared memory
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote_full0) : "r"(full0), "r"(0u));
        for (int unit = unit0; unit * kTpu < P.n_tiles && !(P.debug & 32); unit += n_units) {
            if (TREES && lane == 0 && (unit + n_units) * kTpu >= P.n_tiles) *(volatile int*)&s_stop = 1;
            for (int s = 0; s < stages_per_tile; ++s, ++cnt) {
                const uint32_t slot = cnt % kRing, k = cnt / kRing;
                mbar_wait(full0 + slot * 8u, k & 1);
                asm volatile(
                    "{\n\t.reg .pred pl;\n\tsetp.ne.b32 pl, %1, 0;\n\t"
                    "@pl mbarrier.arrive.shared::cluster.b64 _, [%0];\n\t}\n" ::"r"(remote_full0 + slot * 8u),
                    "r"(lead)
                    : "memory");
            }
        }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      // ONE elected lane runs the whole role (waits included); the other 31 wait at the end of the kernel.  Inside an
      // elect.sync region every value is trivially warp-uniform, so the compiler keeps descriptors, counters and barrier
      // addresses on the uniform datapath (UIADD3 / UMOV feeding UTCHMMA directly, no R2UR).
      if (elect_one()) {
        uint32_t cnt = 0, act_phase = 0;
        const bool leader = true;
        const uint32_t a_step = 2u * (kLboA >> 4), a_half = 8u * (kLboA >> 4), b_step = 2u * (kLboW >> 4);
        const uint32_t a_buf = (uint32_t)kBufRows;  // descriptor units (16 B) between buffers = rows
        const uint32_t a0 = desc_lo(act + buf_row0(0), kLboA), b0 = desc_lo(stages, kLboW);
        auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
            if (PAIR) mma_bf16_pair(d, a_lo, b_lo, accumulate);
            else mma_bf16(d, a_lo, b_lo, accumulate);
        };
        auto commit = [&](uint32_t bar) {
            if (PAIR) commit_pair(bar);
            else commit_to(bar);
        };
        const bool timed = P.timing != nullptr;
        long long t_act = 0, t_full = 0, t_wake = 0;
        const long long t_begin = clock64();
        auto wait_stage = [&](uint32_t slot, uint32_t k) {
            if (P.debug & 32) return;
            const long long t0 = timed ? clock64() : 0;
            mbar_wait(smem_u32(&s_full[slot]), k & 1);  // pair: completes when BOTH halves of the stage have landed
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (timed) t_full += clock64() - t0;
        };
        auto wait_act = [&](uint32_t bar, uint32_t parity) {
            if (P.debug & 16) return;  // timing experiment: the tensor pipe alone (no epilogue, garbage results)
            const long long t0 = timed ? clock64() : 0;
            mbar_wait(bar, parity);
            if (timed) t_act += clock64() - t0;
        };
        // descriptor of the window a tap reads: left / right masked copy for dx = -1 / +1, shifted by the tap
        auto tap_window = [&](int tap, uint32_t centre) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            return (P.debug & 4) ? centre
                                 : (dx < 0 ? a0 + a_buf : (dx > 0 ? a0 + 2u * a_buf : centre)) + (uint32_t)(dy * rowstride + dx);
        };
        // one weight stage = 64 input channels of one tap: four K = 16 steps.  `ready`: the poll issued inside the previous
        // stage already found this stage's weights (issue_stage); otherwise wait for them the slow way.
        uint32_t ready = 0;
        const uint32_t lead = leader ? 1u : 0u;
        auto stage_mmas = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t first_accumulate) {
            const uint32_t slot = cnt % kRing, k = cnt / kRing;
            if (!ready) wait_stage(slot, k);
            else if (!(P.debug & 32)) asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if ((P.debug & 64) && P.timing && blockIdx.x == 0 && cnt < 2048) {
                P.timing[148 * 8 + 2048 + cnt] = clock64();
                P.timing[148 * 8 + 4096 + cnt] = ready;
            }
            const uint32_t nslot = (cnt + 1) % kRing, nk = (cnt + 1) / kRing;
            ready = issue_stage<PAIR>(d_tmem, a_lo, a_step, b0 + slot * (kSlotBytes >> 4), b_step, first_accumulate,
                                      smem_u32(&s_empty[slot]), smem_u32(&s_full[nslot]), nk & 1, lead);
            ++cnt;
        };
        const bool dbg_centre = (P.debug & 4) != 0;
        const uint32_t row_step = dbg_centre ? 0u : (uint32_t)rowstride;
        for (int unit = unit0; unit * kTpu < P.n_tiles; unit += n_units) {
            if (TREES && (unit + n_units) * kTpu >= P.n_tiles) *(volatile int*)&s_stop = 1;
            if (NET) {
                // stem: the four planes of a cell sit in chunk column 0 (column 1 is zero), one K = 16 MMA per tap; a
                // stage carries the [128][16] weights of four taps
                wait_act(bar_act0, act_phase);
                wait_act(bar_act1, act_phase);
                act_phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                ready = 0;  // the stem's stages wait the plain way
#pragma unroll
                for (int s = 0; s < kStemStages; ++s) {
                    const uint32_t slot = cnt % kRing, k = cnt / kRing;
                    wait_stage(slot, k);
                    if (leader) {
                        const uint32_t b_lo = b0 + slot * (kSlotBytes >> 4);
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt)
                            if (4 * s + tt < 9) mma(tmem, tap_window(4 * s + tt, a0), b_lo + (uint32_t)tt * b_step, (s | tt) ? 1u : 0u);
                        commit(smem_u32(&s_empty[slot]));
                    }
                    ++cnt;
                }
                if (leader) commit(bar_acc);
            }
            for (int b = 0; b < P.depth; ++b) {
#pragma unroll 1
                for (uint32_t half = 0; half < 2; ++half) {  // 0: conv1 on x -> accumulator 0, 1: conv2 on h -> accumulator 1
                    const uint32_t d_tmem = tmem + half * 128u;
                    const uint32_t centre = a0 + half * 3u * a_buf;  // buffer 0 (x) or 3 (h)
                    // channel-half major: the MMAs on input channels 0-63 start as soon as the epilogue (or the tile
                    // load) has written that half of the three copies; it writes channels 64-127 underneath them
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        // weights first: the poll issued a boundary ago usually came too early to see the next stage; asking
                        // again costs nothing here, the warp is about to wait for the epilogue anyway
                        if (!ready) {
                            wait_stage(cnt % kRing, cnt / kRing);
                            ready = 1;
                        }
                        wait_act(kb == 0 ? bar_act0 : bar_act1, act_phase);
                        
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        // the three taps of a filter row per iteration, rows rolled: the fully unrolled version (38 stages of
                        // straight-line code per block) ran at the tensor pipe's rate only on some GPCs - instruction fetch
                        const uint32_t kbo = (uint32_t)kb * a_half;
                        uint32_t row = 0u - row_step;  // dy = -1
#pragma unroll 1
                        for (int dy = 0; dy < 3; ++dy, row += row_step) {  // conv2 accumulates on top of the shortcut
                            stage_mmas(d_tmem, (dbg_centre ? centre : a0 + a_buf - 1u) + row + kbo, (dy == 0 && kb == 0) ? half : 1u);
                            stage_mmas(d_tmem, centre + row + kbo, 1u);
                            stage_mmas(d_tmem, (dbg_centre ? centre : a0 + 2u * a_buf + 1u) + row + kbo, 1u);
                        }
                    }
                    act_phase ^= 1;
                    if (leader) commit(bar_acc);
                    if (half == 0) {
                        // the shortcut needs only x: it runs while the epilogue turns accumulator 0 into h
                        stage_mmas(tmem + 128u, a0, 0u);
                        stage_mmas(tmem + 128u, a0 + a_half, 1u);
                    }
                }
            }
        }
        if ((P.debug & 16) && cnt > 0) {  // nobody else waits for the last MMAs in this mode: do it here, before TMEM is freed
            const uint32_t last = cnt - 1;
            mbar_wait(smem_u32(&s_empty[last % kRing]), (last / kRing) & 1);
        }
        if (timed) {
            P.timing[blockIdx.x * 8 + 0] = clock64() - t_begin;
            P.timing[blockIdx.x * 8 + 1] = t_act;
            P.timing[blockIdx.x * 8 + 2] = t_full;
            P.timing[blockIdx.x * 8 + 7] = t_wake;
        }
      }
    } else if (TREES && warp >= kThreads / 32) {
        // ---------------------------------------------------------------- tree warps: evaluator-free simulations beside the net
        const int tw = warp - kThreads / 32;
        const auto rules = RulesView<C4Rules>::get(P.eng);
        for (int t = (int)blockIdx.x * kTreeWarps + tw; t < P.eng.T && !*(volatile int*)&s_stop; t += (int)gridDim.x * kTreeWarps)
            free_sims_tree<1, 1>(P.eng, rules, t, s_path[tw], lane, P.tree_cap, &s_stop);
    } else {
        // ---------------------------------------------------------------- epilogue / tile load / tile store / heads
        const int e = tid - 64;                       // 0..255
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int sub = (warp - 2) >> 2;              // which 32 columns of each 64-channel half
        const int r = q * 32 + lane;                  // accumulator row = tile row
        // tile row -> (y, position, x): r = y * rowstride + p * W + x
        const int ry = r / rowstride, rrem = r - ry * rowstride, rp = rrem / P.W, rx = rrem - rp * P.W;
        const bool row_live = r < rows_used;
        const bool zero_l = rx == P.W - 1, zero_r = rx == 0;
        const long long row_in_tile = (long long)rp * P.cells + ry * P.W + rx;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 32);
        // the loader's view: row lr = e & 127, chunks sub * 4 .. + 3 of each half (a warp = 32 consecutive rows:
        // conflict-free 512-byte stores)
        const int lr = e & 127, lsub = e >> 7;
        const int ly = lr / rowstride, lrem = lr - ly * rowstride, lp = lrem / P.W, lx = lrem - lp * P.W;
        const bool l_live = lr < rows_used;
        const long long l_row_in_tile = (long long)lp * P.cells + ly * P.W + lx;
        uint32_t acc_phase = 0;
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        // this half of the three copies is complete: every thread orders its stores for the tensor core's proxy, one
        // lane per warp arrives
        auto publish = [&](uint32_t bar) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_remote(bar, 0u);
                else mbar_arrive(bar);
            }
        };
        // net mode: this thread's row of the value head's Dense(256) (hidden unit e) lives in registers for the whole kernel
        float w1row[NET ? kHeadMaxCells : 1], b1u = 0.f, w2u = 0.f;
        if (NET) {
#pragma unroll
            for (int c = 0; c < kHeadMaxCells; ++c) w1row[c] = c < P.cells ? __ldg(P.heads.value1_w + (size_t)e * P.cells + c) : 0.f;
            b1u = __ldg(P.heads.value1_b + e);
            w2u = __ldg(P.heads.value2_w + e);
        }
        // scratch of the heads: the live rows of buffer 3 (dead between the last convolution of a tile and epilogue 1 of the
        // next tile's first block, which rewrites them); one 2016-byte run per chunk column, rows 126-127 and pads untouched
        auto scratch = [&](int col) { return reinterpret_cast<float*>(smem + (uint32_t)col * kLboA + buf_row0(3)); };

        // net mode, tile in: the four planes of a cell -> chunk column 0 of the three copies, chunk column 1 cleared
        auto load_planes = [&](long long pos0) {
            if (l_live) {
                uint4 v4 = zero4;
                if (lsub == 0 && pos0 + lp < P.n) {
                    const long long tree = P.index ? (long long)__ldg(P.index + pos0 + lp) : pos0 + lp;
                    const uint2 pl = __ldg(reinterpret_cast<const uint2*>(P.x) + tree * P.cells + (ly * P.W + lx));
                    v4 = make_uint4(pl.x, pl.y, 0u, 0u);
                }
                const uint32_t off = (uint32_t)lsub * kLboA + (uint32_t)lr * 16u;
                *reinterpret_cast<uint4*>(smem + buf_row0(0) + off) = v4;
                if (lx != P.W - 1) *reinterpret_cast<uint4*>(smem + buf_row0(1) + off) = v4;
                if (lx != 0) *reinterpret_cast<uint4*>(smem + buf_row0(2) + off) = v4;
            }
            publish(bar_act0);
            publish(bar_act1);
        };

        // accumulator `acc` (0 / 1) + bias, ReLU, bf16 -> the three copies (centre buffer `centre`), channel half by
        // channel half; or (last layer) -> global memory / the heads.  next_pos0 >= 0 (net mode, last layer): the next
        // tile's planes are loaded as soon as the accumulator has been read, so that its stem runs under the heads.
        long long t_acc = 0, t_body = 0, t_ld = 0, t_half0 = 0;
        const bool etimed = P.timing != nullptr && warp == 2;
        float hsum0 = 0.f, hsum1 = 0.f, hsum2 = 0.f;  // net mode, last layer: this thread's share of the 1x1 head convolutions
        auto epilogue = [&](int acc, const float* bias, int centre, bool last, long long pos0) {
            const long long te0 = etimed ? clock64() : 0;
            mbar_wait(bar_acc, acc_phase);
            const long long te1 = etimed ? clock64() : 0;
            t_acc += te1 - te0;
            acc_phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            __nv_bfloat16* grow = NET ? nullptr : P.y + (pos0 * P.cells + row_in_tile) * kC + sub * 32;
            const bool store_global = !NET && last && row_live && pos0 + rp < P.n;
            uint32_t v[2][32];  // the first channel half alone is waited for; the second arrives while the first is processed
            tmem_ld32_nowait(taddr + (uint32_t)(acc * 128), v[0]);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            pin32(v[0]);
            const long long te2 = etimed ? clock64() : 0;
            t_ld += te2 - te1;
            tmem_ld32_nowait(taddr + (uint32_t)(acc * 128 + 64), v[1]);
            if (NET && last) hsum0 = hsum1 = hsum2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                if (ch == 1) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    pin32(v[1]);
                }
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float* bb = bias + ch * 64 + sub * 32 + c4 * 8;
                    const float4 b0 = *reinterpret_cast<const float4*>(bb);
                    const float4 b1 = *reinterpret_cast<const float4*>(bb + 4);
                    const uint32_t* w = &v[ch][8 * c4];
                    uint4 o;
                    o.x = pack_relu_bf16(__uint_as_float(w[0]) + b0.x, __uint_as_float(w[1]) + b0.y);
                    o.y = pack_relu_bf16(__uint_as_float(w[2]) + b0.z, __uint_as_float(w[3]) + b0.w);
                    o.z = pack_relu_bf16(__uint_as_float(w[4]) + b1.x, __uint_as_float(w[5]) + b1.y);
                    o.w = pack_relu_bf16(__uint_as_float(w[6]) + b1.z, __uint_as_float(w[7]) + b1.w);
                    const int kc = ch * 8 + sub * 4 + c4;
                    if (last) {
                        if (NET) {
                            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                            const float* cw = s_headw + kc * 8;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {  // the bf16-rounded activations, as the stand-alone heads kernel reads them
                                const float lo = __uint_as_float(ow[j] << 16), hi = __uint_as_float(ow[j] & 0xffff0000u);
                                hsum0 = fmaf(lo, cw[2 * j], fmaf(hi, cw[2 * j + 1], hsum0));
                                hsum1 = fmaf(lo, cw[kC + 2 * j], fmaf(hi, cw[kC + 2 * j + 1], hsum1));
                                hsum2 = fmaf(lo, cw[2 * kC + 2 * j], fmaf(hi, cw[2 * kC + 2 * j + 1], hsum2));
                            }
                        } else if (store_global) {
                            *(reinterpret_cast<uint4*>(grow + ch * 64) + c4) = o;
                        }
                    } else if (row_live && !(P.debug & 2)) {
                        const uint32_t off = (uint32_t)kc * kLboA + (uint32_t)r * 16u;
                        *reinterpret_cast<uint4*>(smem + buf_row0(centre) + off) = o;
                        if (!zero_l) *reinterpret_cast<uint4*>(smem + buf_row0(1) + off) = o;
                        if (!zero_r) *reinterpret_cast<uint4*>(smem + buf_row0(2) + off) = o;
                    }
                }
                if (!last) publish(ch == 0 ? bar_act0 : bar_act1);
                if (etimed && ch == 0) {
                    const long long te3 = clock64();
                    t_half0 += te3 - te2;
                    if (lane == 0) s_pub0 = te3;  // when this warp announced channels 0-63 (read by the MMA warp when it wakes up)
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            if (etimed) t_body += clock64() - te1;
        };
        // ---- heads (model.py:68-149) for the tile's positions, from the 1x1 head convolution shares the last epilogue left in
        // hsum0-2.  Two parts, so that the NEXT tile's stem epilogue can run between them: the tensor core then works on that
        // tile's first convolution while these warps do the dense layers of this one.
        auto heads_partials = [&]() {  // (1) the two column groups of a row meet in scratch
            if (row_live) {
                float* part = scratch(sub) + r * 3;
                part[0] = hsum0;
                part[1] = hsum1;
                part[2] = hsum2;
            }
        };
        auto heads_dense = [&](long long pos0) {
            {
                const int cells = P.cells, A = P.heads.A;
                epi_barrier();
                // (2) + bias, ReLU: policy planes NHWC-flattened [position][cell][2], value plane [position][cell]
                float* hp_s = scratch(2);
                float* hv_s = scratch(3);
                if (sub == 0 && row_live) {
                    const float* p0 = scratch(0) + r * 3;
                    const float* p1 = scratch(1) + r * 3;
                    const int cell = ry * P.W + rx;
                    hp_s[(rp * cells + cell) * 2] = fmaxf(p0[0] + p1[0] + __ldg(P.heads.conv_b), 0.f);
                    hp_s[(rp * cells + cell) * 2 + 1] = fmaxf(p0[1] + p1[1] + __ldg(P.heads.conv_b + 1), 0.f);
                    hv_s[rp * cells + cell] = fmaxf(p0[2] + p1[2] + __ldg(P.heads.conv_b + 2), 0.f);
                }
                epi_barrier();
                // (3) value: hidden unit e of every position of the tile, then the Dense(1) contribution, reduced per warp
                float* red = scratch(4);  // [8 warps][ppt]
                for (int p = 0; p < P.ppt; ++p) {
                    float h = b1u;
#pragma unroll
                    for (int c = 0; c < kHeadMaxCells; ++c)
                        if (c < cells) h = fmaf(hv_s[p * cells + c], w1row[c], h);
                    float contrib = fmaxf(h, 0.f) * w2u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                    if (lane == 0) red[(warp - 2) * P.ppt + p] = contrib;
                }
                // (4) policy: 8 lanes per (position, action) share the 2 * cells products, softmax by the group's first lane
                float* logit_s = scratch(5);  // [ppt][A]
                const int grp = e >> 3, part8 = e & 7;
                {
                    const bool on = grp < P.ppt * A;  // no divergence around the shuffles: idle groups carry zeros
                    const int p = on ? grp / A : 0, a = on ? grp - p * A : 0;
                    float acc2 = 0.f;
                    if (on)
                        for (int i = part8; i < 2 * cells; i += 8)
                            acc2 = fmaf(hp_s[p * 2 * cells + i], __ldg(P.heads.policy_w + (size_t)a * 2 * cells + i), acc2);
                    acc2 += __shfl_xor_sync(0xffffffffu, acc2, 4);
                    acc2 += __shfl_xor_sync(0xffffffffu, acc2, 2);
                    acc2 += __shfl_xor_sync(0xffffffffu, acc2, 1);
                    if (on && part8 == 0) logit_s[p * A + a] = acc2 + __ldg(P.heads.policy_b + a);
                }
                epi_barrier();
                if (e < P.ppt && pos0 + e < P.n) {
                    const long long tree = P.index ? (long long)__ldg(P.index + pos0 + e) : pos0 + e;
                    float vsum = __ldg(P.heads.value2_b);
                    for (int w8 = 0; w8 < 8; ++w8) vsum += red[w8 * P.ppt + e];
                    P.heads.values[tree] = tanhf(vsum);
                    float mx = -INFINITY, ssum = 0.f;
                    for (int a = 0; a < A; ++a) mx = fmaxf(mx, logit_s[e * A + a]);
                    for (int a = 0; a < A; ++a) ssum += expf(logit_s[e * A + a] - mx);
                    for (int a = 0; a < A; ++a) P.heads.priors[tree * A + a] = expf(logit_s[e * A + a] - mx) / ssum;
                }
                epi_barrier();  // scratch (and through it buffer 3) is free again before anybody runs ahead
            }
        };

        // a CTA of a pair whose tile lies beyond the batch still takes part in every barrier: it computes on zeros and
        // stores nothing (pos0 >= n)
        if (NET && unit0 * kTpu < P.n_tiles && !(P.debug & 16)) load_planes((long long)(unit0 * kTpu + (int)rank) * P.ppt);  // later tiles: inside the last epilogue
        bool stem_done = false;
        for (int unit = unit0; unit * kTpu < P.n_tiles && !(P.debug & 16); unit += n_units) {
            const long long pos0 = (long long)(unit * kTpu + (int)rank) * P.ppt;
            const long long next_pos0 = (unit + n_units) * kTpu < P.n_tiles ? (long long)((unit + n_units) * kTpu + (int)rank) * P.ppt : -1;
            if (NET) {
                // stem: accumulator 0 + bias, ReLU -> x and its masked copies (later tiles: done inside the previous tile's tail)
                if (!stem_done) epilogue(0, s_bias, 0, false, pos0);
            } else {
                // ---- tile in: x, x-left-masked, x-right-masked; channels 0-63 first
                uint4 v[8];
                const bool have = l_live && pos0 + lp < P.n;
                const uint4* src = reinterpret_cast<const uint4*>(P.x + (pos0 * P.cells + l_row_in_tile) * kC) + lsub * 4;
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = have ? __ldg(src + (c >> 2) * 8 + (c & 3)) : zero4;
                const bool zl = lx == P.W - 1, zr = lx == 0;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    if (l_live) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t off = (uint32_t)(ch * 8 + lsub * 4 + c) * kLboA + (uint32_t)lr * 16u;
                            *reinterpret_cast<uint4*>(smem + buf_row0(0) + off) = v[ch * 4 + c];
                            if (!zl) *reinterpret_cast<uint4*>(smem + buf_row0(1) + off) = v[ch * 4 + c];
                            if (!zr) *reinterpret_cast<uint4*>(smem + buf_row0(2) + off) = v[ch * 4 + c];
                        }
                    }
                    publish(ch == 0 ? bar_act0 : bar_act1);
                }
            }
            for (int b = 0; b < P.depth; ++b) {
                epilogue(0, s_bias + kC + (b * 2) * kC, 3, false, pos0);                // accumulator 0 -> h
                epilogue(1, s_bias + kC + (b * 2 + 1) * kC, 0, b == P.depth - 1, pos0);  // accumulator 1 -> block output
            }
            if (NET) {
                // buffers 0-2 are dead now: hand the next tile to the tensor core, turn its stem into x as soon as that is
                // there, and only then do this tile's dense heads - under the next tile's first convolution
                if (next_pos0 >= 0) load_planes(next_pos0);
                heads_partials();
                stem_done = next_pos0 >= 0;
                if (stem_done) epilogue(0, s_bias, 0, false, next_pos0);
                heads_dense(pos0);
            }
        }
        if (etimed && lane == 0) {
            P.timing[blockIdx.x * 8 + 3] = t_acc;
            P.timing[blockIdx.x * 8 + 4] = t_body;
            P.timing[blockIdx.x * 8 + 5] = t_ld;
            P.timing[blockIdx.x * 8 + 6] = t_half0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync();  // nobody frees TMEM or exits while the pair's MMAs / remote arrivals may still be in flight
    if (P.timeline && tid == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
        atomicMax(P.timeline + 1, t_);
    }
    if (warp == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

// launches k_tower<NET, PAIR, TREES>: a plain grid of one CTA per SM, or clusters of two CTAs
template <bool NET, bool PAIR, bool TREES = false>
static int launch_tower(const TowerParams& P, int sms, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_tower<NET, PAIR, TREES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
            return az::fail_net(AZ_ERR_CUDA, "az_net_tower / az_net_forward: shared memory request refused");
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    if (PAIR) {
        const int units = (P.n_tiles + 1) / 2, max_units = sms / 2;
        cfg.gridDim = dim3(2u * (unsigned)((units < max_units && !TREES) ? units : max_units));
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3((unsigned)((P.n_tiles < sms && !TREES) ? P.n_tiles : sms));
    }
    cfg.blockDim = dim3(kThreads + (TREES ? kTreeWarps * 32 : 0));
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    if (cudaLaunchKernelEx(&cfg, k_tower<NET, PAIR, TREES>, P) != cudaSuccess) {
        cudaGetLastError();
        return az::fail_net(AZ_ERR_CUDA, "az_net_tower / az_net_forward: launch failed");
    }
    return AZ_OK;
}

static long long* g_timing = nullptr;  // az_net_tower_timing: device buffer [grid][8] for the next launches, or null
static unsigned long long* g_timeline = nullptr;  // az_net_debug_timeline: slots {start, end}, one per launch, round robin
static int g_timeline_slots = 0, g_timeline_next = 0;
static unsigned long long* next_timeline_slot() {
    if (!g_timeline || g_timeline_slots <= 0) return nullptr;
    unsigned long long* tl = g_timeline + 2 * (g_timeline_next % g_timeline_slots);
    g_timeline_next += 1;
    return tl;
}

static int tower_launch_checks(const char* who, int n, int H, int W, int channels, int depth, int* ppt_out) {
    if (n < 0 || H < 1 || W < 1) return az::fail_net(AZ_ERR_ARG, who);
    if (channels != kC) return az::fail_net(AZ_ERR_ARG, "az_net_tower / az_net_forward: built for 128 filters (config.py:71)");
    if (depth < 1 || depth > kMaxDepth) return az::fail_net(AZ_ERR_ARG, "az_net_tower / az_net_forward: depth must be 1..4 (config.py:63 uses 4)");
    const int cells = H * W;
    if (cells > kTileRows) return az::fail_net(AZ_ERR_ARG, "az_net_tower / az_net_forward: a position must fit one 128-row tile (H * W <= 128)");
    const int ppt = kTileRows / cells;
    if (ppt * W + 1 > kPad) return az::fail_net(AZ_ERR_ARG, "az_net_tower / az_net_forward: positions-per-tile * W + 1 must be <= 22 (padding rows)");
    *ppt_out = ppt;
    return AZ_OK;
}

}  // namespace tower
}  // namespace az

extern "C" __attribute__((visibility("default"))) int az_net_tower_timing(void* dev_buffer) {
    az::tower::g_timing = static_cast<long long*>(dev_buffer);
    return AZ_OK;
}

extern "C" __attribute__((visibility("default"))) int az_net_debug_timeline(void* dev_slots, int32_t n_slots) {
    if (n_slots < 0) return az::fail_net(AZ_ERR_ARG, "az_net_debug_timeline: bad argument");
    az::tower::g_timeline = static_cast<unsigned long long*>(dev_slots);
    az::tower::g_timeline_slots = dev_slots ? n_slots : 0;
    az::tower::g_timeline_next = 0;
    return AZ_OK;
}

extern "C" __attribute__((visibility("default"))) int az_net_tower(const void* x, const void* w_img, const float* bias, int32_t n,
                                                                    int32_t H, int32_t W, int32_t channels, int32_t depth,
                                                                    int32_t layout, void* y, void* stream) {
    using namespace az::tower;
    if (n == 0) return AZ_OK;
    if (!x || !w_img || !bias || !y || (layout != 0 && layout != 1)) return az::fail_net(AZ_ERR_ARG, "az_net_tower: bad argument");
    int ppt = 0;
    if (int rc = tower_launch_checks("az_net_tower: bad argument", n, H, W, channels, depth, &ppt)) return rc;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_img) | reinterpret_cast<uintptr_t>(y)) & 15)
        return az::fail_net(AZ_ERR_ARG, "az_net_tower: pointers must be 16-byte aligned");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    TowerParams P{};
    P.x = static_cast<const __nv_bfloat16*>(x);
    P.w_img = static_cast<const uint8_t*>(w_img);
    P.bias = bias;
    P.y = static_cast<__nv_bfloat16*>(y);
    P.n = n, P.W = W, P.cells = H * W, P.ppt = ppt, P.depth = depth, P.n_tiles = (n + ppt - 1) / ppt;
    if (const char* dbg = getenv("AZ_TOWER_DEBUG")) P.debug = atoi(dbg);
    P.timing = g_timing;
    P.timeline = next_timeline_slot();
    return layout ? launch_tower<false, true>(P, sms, static_cast<cudaStream_t>(stream))
                  : launch_tower<false, false>(P, sms, static_cast<cudaStream_t>(stream));
}

static int net_forward_impl(const void* states, const void* w_img, const float* stem_bias, const float* tower_bias,
                            const az_net_head_params* heads, const int32_t* index, const int32_t* count, int32_t n, int32_t H,
                            int32_t W, int32_t channels, int32_t depth, int32_t n_actions, int32_t layout, float* priors,
                            float* values, void* stream, const az_engine* trees = nullptr, int32_t tree_sims = 0) {
    using namespace az::tower;
    if (n == 0) return AZ_OK;
    if (!states || !w_img || !stem_bias || !tower_bias || !heads || !priors || !values || (layout != 0 && layout != 1))
        return az::fail_net(AZ_ERR_ARG, "az_net_forward: bad argument");
    if (!heads->conv_w || !heads->conv_b || !heads->policy_w || !heads->policy_b || !heads->value1_w || !heads->value1_b ||
        !heads->value2_w || !heads->value2_b)
        return az::fail_net(AZ_ERR_ARG, "az_net_forward: null head weights");
    int ppt = 0;
    if (int rc = tower_launch_checks("az_net_forward: bad argument", n, H, W, channels, depth, &ppt)) return rc;
    const int cells = H * W;
    if (cells > kHeadMaxCells) return az::fail_net(AZ_ERR_ARG, "az_net_forward: the fused heads are built for boards of up to 48 cells");
    if (n_actions < 1 || ppt * n_actions * 8 > 256 || ppt * n_actions > 126)
        return az::fail_net(AZ_ERR_ARG, "az_net_forward: positions-per-tile * actions must be <= 32");
    if ((reinterpret_cast<uintptr_t>(states) & 7) || (reinterpret_cast<uintptr_t>(w_img) & 15))
        return az::fail_net(AZ_ERR_ARG, "az_net_forward: misaligned pointer");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return az::fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    TowerParams P{};
    P.x = static_cast<const __nv_bfloat16*>(states);
    P.w_img = static_cast<const uint8_t*>(w_img);
    P.bias = tower_bias;
    P.stem_bias = stem_bias;
    P.heads = HeadParams{heads->conv_w, heads->conv_b, heads->policy_w, heads->policy_b, heads->value1_w, heads->value1_b,
                         heads->value2_w, heads->value2_b, priors, values, n_actions};
    P.index = index;
    P.count = count;
    P.n = n, P.W = W, P.cells = cells, P.ppt = ppt, P.depth = depth, P.n_tiles = (n + ppt - 1) / ppt;
    if (const char* dbg = getenv("AZ_TOWER_DEBUG")) P.debug = atoi(dbg);
    P.timing = g_timing;
    P.timeline = next_timeline_slot();
    if (trees && tree_sims > 0) {
        if (az::engine_view(trees, &P.eng) != 1 || H != 6 || W != 7)
            return az::fail_net(AZ_ERR_ARG, "az_net_forward_trees: the engine must be the plain 6x7 connect-4 configuration (no root noise)");
        P.tree_cap = tree_sims;
        return layout ? launch_tower<true, true, true>(P, sms, static_cast<cudaStream_t>(stream))
                      : launch_tower<true, false, true>(P, sms, static_cast<cudaStream_t>(stream));
    }
    return layout ? launch_tower<true, true>(P, sms, static_cast<cudaStream_t>(stream))
                  : launch_tower<true, false>(P, sms, static_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int az_net_forward(const void* states, const void* w_img, const float* stem_bias,
                                                                      const float* tower_bias, const az_net_head_params* heads,
                                                                      int32_t n, int32_t H, int32_t W, int32_t channels, int32_t depth,
                                                                      int32_t n_actions, int32_t layout, float* priors, float* values,
                                                                      void* stream) {
    return net_forward_impl(states, w_img, stem_bias, tower_bias, heads, nullptr, nullptr, n, H, W, channels, depth, n_actions,
                            layout, priors, values, stream);
}

extern "C" __attribute__((visibility("default"))) int az_net_forward_gathered(
    const void* states, const void* w_img, const float* stem_bias, const float* tower_bias, const az_net_head_params* heads,
    const int32_t* index, const int32_t* count, int32_t n_max, int32_t H, int32_t W, int32_t channels, int32_t depth,
    int32_t n_actions, int32_t layout, float* priors, float* values, void* stream) {
    if (!index || !count) return az::fail_net(AZ_ERR_ARG, "az_net_forward_gathered: null index / count");
    return net_forward_impl(states, w_img, stem_bias, tower_bias, heads, index, count, n_max, H, W, channels, depth, n_actions,
                            layout, priors, values, stream);
}

extern "C" __attribute__((visibility("default"))) int az_net_forward_trees(
    const void* states, const void* w_img, const float* stem_bias, const float* tower_bias, const az_net_head_params* heads,
    const int32_t* index, const int32_t* count, int32_t n_max, int32_t H, int32_t W, int32_t channels, int32_t depth,
    int32_t n_actions, int32_t layout, float* priors, float* values, az_engine* engine, int32_t max_sims, void* stream) {
    if (!index || !count || !engine || max_sims < 1) return az::fail_net(AZ_ERR_ARG, "az_net_forward_trees: bad argument");
    return net_forward_impl(states, w_img, stem_bias, tower_bias, heads, index, count, n_max, H, W, channels, depth, n_actions,
                            layout, priors, values, stream, engine, max_sims);
}
