// az_net.cu - the two memory-bound ends of the policy/value net as hand-written kernels.
//
// The 13 tower convolutions (99.95 % of the FLOPs) run on the tensor cores through cuDNN with fused
// bias/ReLU/residual epilogues (az_b200/net.py).  The stem (K = 36) and the heads (3 output channels,
// two tiny dense stacks) are bandwidth-bound; as library calls they cost ~20 small kernels per
// advance, so they are written here as one kernel each:
//   k_stem   NN input [n][H][W][4] bf16 -> Conv3x3(4 -> C) + folded BN + ReLU -> [n][H][W][C] bf16
//            (reference: model/tensorflow/model.py:36-46, base_layers.py:58-66)
//   k_heads  tower output [n][H*W][C] bf16 -> Conv1x1(C -> 2)+BN+ReLU -> Dense(A) softmax  (model.py:68-103)
//                                          -> Conv1x1(C -> 1)+BN+ReLU -> Dense(256) ReLU -> Dense(1) tanh (:106-149)
//            written straight into the priors / values buffers az_step consumes.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/az_b200.h"
#include "az_mma.cuh"

namespace az {

int fail_net(int code, const char* msg);
constexpr int kMaxDimNet = 11;

// Both kernels put their small GEMM on the legacy tensor path (mma.sync m16n8k16 bf16 -> fp32, az_mma.cuh): the
// stem is 0.4 % and the head convolutions 0.03 % of the net's FLOPs, both are bound by the 44 MB they write / read
// per 4096 positions, and a scalar-FMA version costs 10x the instructions (it measured 69 us against ~10 us of traffic).

// ------------------------------------------------------------------------------------------ stem
// Implicit GEMM per position: D[pixel][cout] = A[pixel][kk] * B[kk][cout], kk = tap*4 + plane (36, padded to 48).
// A is read straight from a zero-bordered copy of the position's planes in shared memory (one 32-bit word =
// two planes of one neighbour cell = one A-fragment register); B (the channel's weights) lives in registers.
// Block = 4 warps, warp w owns output channels [32w, 32w+32); warps are independent (own staging buffer).
constexpr int kStemWarps = 4;
constexpr int kMaxPadCells = (kMaxDimNet + 2) * (kMaxDimNet + 2);

__global__ void __launch_bounds__(kStemWarps * 32) k_stem(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int n, int H, int W,
                                                          __nv_bfloat16* __restrict__ out) {
    __shared__ uint2 s_in[kStemWarps][kMaxPadCells];
    __shared__ int s_base[kMaxDimNet * kMaxDimNet + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int PW = W + 2, cells = H * W, C = 128;
    for (int p = threadIdx.x; p < cells + 16; p += blockDim.x) {
        int q = p < cells ? p : cells - 1;  // rows past the board repeat the last pixel (discarded)
        s_base[p] = (q / W) * PW + q % W;
    }
    for (int i = lane; i < (H + 2) * PW; i += 32) s_in[warp][i] = make_uint2(0u, 0u);
    // B fragments: 3 k-steps x 4 n-tiles
    uint32_t bf[3][4][2];
    float bv[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int co = warp * 32 + nt * 8 + g;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = ks * 16 + t4 * 2 + h * 8;  // pair (kk, kk+1) = planes (ci, ci+1) of one tap
                const int tap = kk >> 2, ci = kk & 3;
                float lo = 0.f, hi = 0.f;
                if (tap < 9) {
                    lo = w[co * 36 + ci * 9 + tap];
                    hi = w[co * 36 + (ci + 1) * 9 + tap];
                }
                bf[ks][nt][h] = pack_bf16(lo, hi);
            }
        bv[nt][0] = bias[warp * 32 + nt * 8 + t4 * 2];
        bv[nt][1] = bias[warp * 32 + nt * 8 + t4 * 2 + 1];
    }
    // per-lane tap offsets into the padded board for the two k-halves of each k-step
    int toff[3][2], wsel[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kk = ks * 16 + t4 * 2 + h * 8;
            int tap = kk >> 2;
            if (tap > 8) tap = 0;  // padding columns: weights are zero, any finite input will do
            toff[ks][h] = (tap / 3) * PW + tap % 3;
            wsel[ks][h] = (kk & 3) >> 1;
        }
    __syncthreads();
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_in[warp]);
    const int mtiles = (cells + 15) >> 4;
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        __syncwarp();
        for (int p = lane; p < cells; p += 32)
            s_in[warp][s_base[p] + PW + 1] = reinterpret_cast<const uint2*>(in)[(size_t)t * cells + p];
        __syncwarp();
        __nv_bfloat16* o = out + (size_t)t * cells * C + warp * 32 + t4 * 2;
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
                for (int nt = 0; nt < 4; ++nt) mma_16816(acc[nt], a, bf[ks][nt]);
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
}

// ------------------------------------------------------------------------------------------ heads
constexpr int kHeadWarps = 16;

// One warp per position, grid-stride; dense weights staged once per block in shared memory.
// 1x1 convolutions: D[pixel][3] = X[pixel][128] * Wc[128][3] with mma.sync, the A fragments loaded straight
// from global memory (every byte of the position is requested exactly once, 96 independent loads per lane).
// Shared layout (floats): policy_w [A][2*cells | 1 pad] | value1_w transposed [cells][256] | per-warp h [3*cells]
// FROM_HD: the 1x1 convolutions were already done by k_head_convs (128-bit loads, float32 accumulation) and `x` is their
// output hd [n][cells][3] float32; the kernel then only runs the dense layers.
template <int C, bool FROM_HD = false>
__global__ void __launch_bounds__(kHeadWarps * 32) k_heads(const __nv_bfloat16* __restrict__ x, HeadParams hp,
                                                           float* __restrict__ priors, float* __restrict__ values) {
    static_assert(C == 128, "8 k-steps of 16 channels");
    extern __shared__ float s_f[];
    const int cells = hp.cells, A = hp.A;
    const int ps = 2 * cells + 1;  // odd row stride: conflict-free across lanes
    float* s_pw = s_f;
    float* s_vw = s_pw + ((A * ps + 3) & ~3);
    float* s_h = s_vw + kHidden * cells;
    // the dense weights arrive already in the padded (odd row stride) layout: two straight 128-bit copies
    {
        const int np4 = (A * ps) >> 2, nv4 = (kHidden * cells) >> 2;
        const float4* gp = reinterpret_cast<const float4*>(hp.policy_w);
        const float4* gv = reinterpret_cast<const float4*>(hp.value1_w);
        for (int i = threadIdx.x; i < np4; i += blockDim.x) reinterpret_cast<float4*>(s_pw)[i] = gp[i];
        for (int i = (np4 << 2) + threadIdx.x; i < A * ps; i += blockDim.x) s_pw[i] = hp.policy_w[i];
        for (int i = threadIdx.x; i < nv4; i += blockDim.x) reinterpret_cast<float4*>(s_vw)[i] = gv[i];
        for (int i = (nv4 << 2) + threadIdx.x; i < kHidden * cells; i += blockDim.x) s_vw[i] = hp.value1_w[i];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    float* h = s_h + warp * 3 * cells;  // [cells][2] policy planes (Keras Flatten of [H][W][2]) then [cells] value plane
    // B fragments of the 1x1 convolutions: column n = g (only n < 3 is non-zero)
    uint32_t bf[C / 16][2];
#pragma unroll
    for (int ks = 0; ks < C / 16; ++ks)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int k = ks * 16 + t4 * 2 + hh * 8;
            bf[ks][hh] = g < 3 ? pack_bf16(hp.conv_w[g * C + k], hp.conv_w[g * C + k + 1]) : 0u;
        }
    const float cb0 = hp.conv_b[0], cb1 = hp.conv_b[1], cb2 = hp.conv_b[2];
    __syncthreads();
    const int mtiles = (cells + 15) >> 4;
    for (int t = blockIdx.x * kHeadWarps + warp; t < hp.n; t += gridDim.x * kHeadWarps) {
        const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + (size_t)t * cells * C);
        if (FROM_HD) {
            const float* hd = reinterpret_cast<const float*>(x) + (size_t)t * cells * 3;
            for (int i = lane; i < 3 * cells; i += 32) {
                const int r = i / 3, pl = i - r * 3;
                h[pl < 2 ? r * 2 + pl : 2 * cells + r] = __ldg(hd + i);
            }
        }
        for (int mt = 0; !FROM_HD && mt < mtiles; ++mt) {
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
            // acc[0], acc[1]: row r0, columns 2*t4, 2*t4+1; acc[2], acc[3]: row r1
            if (t4 == 0) {
                if (r0 < cells) { h[r0 * 2] = fmaxf(acc[0] + cb0, 0.f); h[r0 * 2 + 1] = fmaxf(acc[1] + cb1, 0.f); }
                if (r1 < cells) { h[r1 * 2] = fmaxf(acc[2] + cb0, 0.f); h[r1 * 2 + 1] = fmaxf(acc[3] + cb1, 0.f); }
            } else if (t4 == 1) {
                if (r0 < cells) h[2 * cells + r0] = fmaxf(acc[0] + cb2, 0.f);
                if (r1 < cells) h[2 * cells + r1] = fmaxf(acc[2] + cb2, 0.f);
            }
        }
        __syncwarp();
        // policy: Dense(A) + softmax
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
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float ex[4], sum = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            ex[m] = lane + 32 * m < A ? expf(logit[m] - mx) : 0.f;
            sum += ex[m];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (lane + 32 * m < A) priors[(size_t)t * A + lane + 32 * m] = ex[m] / sum;
        // value: Dense(256) ReLU -> Dense(1) tanh; lane owns hidden units lane, lane+32, ...
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
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) values[t] = tanhf(part + hp.value2_b[0]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ head convolutions alone
// The two 1x1 head convolutions (policy: 2 planes, value: 1 plane; model.py:76-80, :114-118) + ReLU on the tower
// output, for shapes az_net_heads does not cover (chess: 64 cells, 1 880 actions - the dense layers then run through
// cuBLAS).  Bandwidth bound: every activation is read once with 128-bit loads; a half-warp owns one cell (16 lanes x
// 8 channels), three dot products in float32, reduced by shuffles.  x [rows][128] bf16 -> out [rows][3] float32.
__global__ void __launch_bounds__(256) k_head_convs(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ b, long long rows, float* __restrict__ out) {
    __shared__ float sw[3 * 128];
    for (int i = threadIdx.x; i < 3 * 128; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int l16 = threadIdx.x & 15, sub = (threadIdx.x & 31) >> 4;
    float wr[3][8];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int c = 0; c < 8; ++c) wr[o][c] = sw[o * 128 + 8 * l16 + c];
    const float b0 = b[0], b1 = b[1], b2 = b[2];
    const long long warp = (long long)(blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (long long)(gridDim.x * blockDim.x) >> 5;
    constexpr int U = 4;  // cells in flight per half-warp: the kernel is latency bound with one
    for (long long base = 2 * U * warp; base < rows; base += 2 * U * n_warps) {  // 2 * U cells per warp and iteration
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = base + 2 * u + sub;
            v[u] = row < rows ? __ldcs(reinterpret_cast<const uint4*>(x + row * 128) + l16) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = base + 2 * u + sub;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float lo = __uint_as_float(w4[j] << 16), hi = __uint_as_float(w4[j] & 0xffff0000u);
                a0 = fmaf(lo, wr[0][2 * j], a0);
                a0 = fmaf(hi, wr[0][2 * j + 1], a0);
                a1 = fmaf(lo, wr[1][2 * j], a1);
                a1 = fmaf(hi, wr[1][2 * j + 1], a1);
                a2 = fmaf(lo, wr[2][2 * j], a2);
                a2 = fmaf(hi, wr[2][2 * j + 1], a2);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (l16 == 0 && row < rows) {
                out[row * 3 + 0] = fmaxf(a0 + b0, 0.f);
                out[row * 3 + 1] = fmaxf(a1 + b1, 0.f);
                out[row * 3 + 2] = fmaxf(a2 + b2, 0.f);
            }
        }
    }
}


}  // namespace az

using namespace az;
#define AZ_API extern "C" __attribute__((visibility("default")))

AZ_API int az_net_stem(const void* states, const float* w, const float* b, int32_t n, int32_t H, int32_t W, int32_t C,
                       void* out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!states || !w || !b || !out || n < 0 || H < 1 || W < 1) return fail_net(AZ_ERR_ARG, "az_net_stem: bad argument");
    if (C != 128) return fail_net(AZ_ERR_ARG, "az_net_stem: built for 128 filters (config.py:71)");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (H > kMaxDimNet || W > kMaxDimNet) return fail_net(AZ_ERR_ARG, "az_net_stem: board larger than 11x11");
    const int grid = n < sms * 5 ? n : sms * 5;
    k_stem<<<grid, kStemWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(states), w, b, n, H, W, static_cast<__nv_bfloat16*>(out));
    if (cudaGetLastError() != cudaSuccess) return fail_net(AZ_ERR_CUDA, "az_net_stem: launch failed");
    return AZ_OK;
}

static int launch_heads(const void* x, const az_head_weights* hw, int32_t n, int32_t cells, int32_t C, int32_t A, float* priors,
                        float* values, void* stream, bool from_hd);

AZ_API int az_net_heads(const void* x, const az_head_weights* hw, int32_t n, int32_t cells, int32_t C, int32_t A,
                        float* priors, float* values, void* stream) {
    return launch_heads(x, hw, n, cells, C, A, priors, values, stream, false);
}

AZ_API int az_net_heads_dense(const float* hd, const az_head_weights* hw, int32_t n, int32_t cells, int32_t A, float* priors,
                              float* values, void* stream) {
    return launch_heads(hd, hw, n, cells, 128, A, priors, values, stream, true);
}

static int launch_heads(const void* x, const az_head_weights* hw, int32_t n, int32_t cells, int32_t C, int32_t A, float* priors,
                        float* values, void* stream, bool from_hd) {
    if (n == 0) return AZ_OK;
    if (!x || !hw || !priors || !values || n < 0 || cells < 1 || A < 1 || A > AZ_MAX_ACTIONS)
        return fail_net(AZ_ERR_ARG, "az_net_heads: bad argument");
    if (C != 128) return fail_net(AZ_ERR_ARG, "az_net_heads: built for 128 filters (config.py:71)");
    HeadParams hp{hw->conv_w, hw->conv_b, hw->policy_w, hw->policy_b, hw->value1_w, hw->value1_b, hw->value2_w, hw->value2_b,
                  n, cells, A};
    const size_t smem = sizeof(float) * ((((size_t)A * (2 * cells + 1) + 3) & ~(size_t)3) + (size_t)kHidden * cells +
                                         (size_t)kHeadWarps * 3 * cells);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_heads<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_heads<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail_net(AZ_ERR_CUDA, "az_net_heads: shared memory request refused");
        configured = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = (n + kHeadWarps - 1) / kHeadWarps;
    const int per_sm = smem > 100 * 1024 ? 1 : 2;
    if (grid > sms * per_sm) grid = sms * per_sm;
    if (from_hd)
        k_heads<128, true><<<grid, kHeadWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(x), hp, priors, values);
    else
        k_heads<128, false><<<grid, kHeadWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
            static_cast<const __nv_bfloat16*>(x), hp, priors, values);
    if (cudaGetLastError() != cudaSuccess) return fail_net(AZ_ERR_CUDA, "az_net_heads: launch failed");
    return AZ_OK;
}

AZ_API int az_net_head_convs(const void* x, const float* conv_w, const float* conv_b, int32_t n, int32_t cells, int32_t C,
                             float* out, void* stream) {
    if (n == 0) return AZ_OK;
    if (!x || !conv_w || !conv_b || !out || n < 0 || cells < 1) return fail_net(AZ_ERR_ARG, "az_net_head_convs: bad argument");
    if (C != 128) return fail_net(AZ_ERR_ARG, "az_net_head_convs: built for 128 filters (config.py:71)");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail_net(AZ_ERR_NO_DEVICE, "no CUDA device: libaz_b200 has no CPU fallback");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long rows = (long long)n * cells;
    long long grid = (rows + 63) / 64;  // 8 warps x 8 cells per block and iteration
    if (grid > (long long)sms * 8) grid = (long long)sms * 8;
    k_head_convs<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), conv_w,
                                                                               conv_b, rows, out);
    if (cudaGetLastError() != cudaSuccess) return fail_net(AZ_ERR_CUDA, "az_net_head_convs: launch failed");
    return AZ_OK;
}
