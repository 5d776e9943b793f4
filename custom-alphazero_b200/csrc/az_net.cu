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

namespace az {

int fail_net(int code, const char* msg);

// ------------------------------------------------------------------------------------------ stem
// One block = C threads (thread = output channel), grid-stride over boards.  The board is staged in
// shared memory with a zero border so the 3x3 window never branches; each tap is one 128-bit
// broadcast load (the 4 input planes) and 4 FMAs; the 36 weights of the channel live in registers.
template <int C>
__global__ void __launch_bounds__(C) k_stem(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                                            const float* __restrict__ bias, int n, int H, int W,
                                            __nv_bfloat16* __restrict__ out) {
    extern __shared__ float4 s_in[];  // [(H+2)][(W+2)]
    const int co = threadIdx.x, PW = W + 2, cells = H * W;
    float wr[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) wr[i] = w[co * 36 + i];  // OIHW: [co][ci][ky][kx]
    const float b = bias[co];
    for (int i = threadIdx.x; i < (H + 2) * PW; i += C) s_in[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        __syncthreads();
        for (int p = threadIdx.x; p < cells; p += C) {
            const uint2 v = reinterpret_cast<const uint2*>(in)[(size_t)t * cells + p];
            const int y = p / W, x = p - y * W;
            float4 f;
            f.x = __uint_as_float(v.x << 16);
            f.y = __uint_as_float(v.x & 0xffff0000u);
            f.z = __uint_as_float(v.y << 16);
            f.w = __uint_as_float(v.y & 0xffff0000u);
            s_in[(y + 1) * PW + x + 1] = f;
        }
        __syncthreads();
        __nv_bfloat16* o = out + (size_t)t * cells * C + co;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float acc = b;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 f = s_in[(y + ky) * PW + x + kx];
                        acc = fmaf(f.x, wr[0 * 9 + ky * 3 + kx], acc);
                        acc = fmaf(f.y, wr[1 * 9 + ky * 3 + kx], acc);
                        acc = fmaf(f.z, wr[2 * 9 + ky * 3 + kx], acc);
                        acc = fmaf(f.w, wr[3 * 9 + ky * 3 + kx], acc);
                    }
                o[(size_t)(y * W + x) * C] = __float2bfloat16_rn(fmaxf(acc, 0.f));
            }
    }
}

// ------------------------------------------------------------------------------------------ heads
constexpr int kHeadWarps = 8;
constexpr int kHidden = 256;

struct HeadParams {
    const float *conv_w, *conv_b, *policy_w, *policy_b, *value1_w, *value1_b, *value2_w, *value2_b;
    int n, cells, A;
};

// One warp per board, grid-stride; all head weights staged once per block in shared memory.
// Shared layout (floats): conv_w [3][C] | policy_w [A][2*cells | 1 pad] | value1_w [256][cells | 1 pad] |
//                         per-warp h [kHeadWarps][3*cells]
template <int C>
__global__ void __launch_bounds__(kHeadWarps * 32) k_heads(const __nv_bfloat16* __restrict__ x, HeadParams hp,
                                                           float* __restrict__ priors, float* __restrict__ values) {
    extern __shared__ float s_f[];
    const int cells = hp.cells, A = hp.A;
    const int ps = 2 * cells + 1, vs = cells | 1;  // odd row strides: conflict-free across lanes
    float* s_cw = s_f;
    float* s_pw = s_cw + 3 * C;
    float* s_vw = s_pw + A * ps;
    float* s_h = s_vw + kHidden * vs;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_cw[i] = hp.conv_w[i];
    for (int i = threadIdx.x; i < A * 2 * cells; i += blockDim.x) s_pw[(i / (2 * cells)) * ps + i % (2 * cells)] = hp.policy_w[i];
    for (int i = threadIdx.x; i < kHidden * cells; i += blockDim.x) s_vw[(i / cells) * vs + i % cells] = hp.value1_w[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* h = s_h + warp * 3 * cells;  // [cells][2] policy planes then [cells] value plane
    float cw[3][C / 32];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < C / 32; ++j) cw[c][j] = s_cw[c * C + lane * (C / 32) + j];
    const float cb0 = hp.conv_b[0], cb1 = hp.conv_b[1], cb2 = hp.conv_b[2];
    static_assert(C == 128, "one 8-byte load per lane covers C = 128 channels");
    for (int t = blockIdx.x * kHeadWarps + warp; t < hp.n; t += gridDim.x * kHeadWarps) {
        const uint2* row = reinterpret_cast<const uint2*>(x + (size_t)t * cells * C) + lane;
        for (int p0 = 0; p0 < cells; p0 += 4) {  // 4 pixels in flight per lane
            uint2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = p0 + u < cells ? row[(size_t)(p0 + u) * (C / 4)] : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float f0 = __uint_as_float(v[u].x << 16), f1 = __uint_as_float(v[u].x & 0xffff0000u);
                const float f2 = __uint_as_float(v[u].y << 16), f3 = __uint_as_float(v[u].y & 0xffff0000u);
                float a0 = f0 * cw[0][0] + f1 * cw[0][1] + f2 * cw[0][2] + f3 * cw[0][3];
                float a1 = f0 * cw[1][0] + f1 * cw[1][1] + f2 * cw[1][2] + f3 * cw[1][3];
                float a2 = f0 * cw[2][0] + f1 * cw[2][1] + f2 * cw[2][2] + f3 * cw[2][3];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                }
                if (lane == 0 && p0 + u < cells) {
                    h[(p0 + u) * 2 + 0] = fmaxf(a0 + cb0, 0.f);  // Keras Flatten of [H][W][2]
                    h[(p0 + u) * 2 + 1] = fmaxf(a1 + cb1, 0.f);
                    h[2 * cells + p0 + u] = fmaxf(a2 + cb2, 0.f);
                }
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
        // value: Dense(256) ReLU -> Dense(1) tanh
        float part = 0.f;
        const float* hv = h + 2 * cells;
#pragma unroll
        for (int m = 0; m < kHidden / 32; ++m) {
            const int j = lane + 32 * m;
            float acc = hp.value1_b[j];
            const float* wrow = s_vw + j * vs;
            for (int p = 0; p < cells; ++p) acc = fmaf(hv[p], wrow[p], acc);
            part = fmaf(fmaxf(acc, 0.f), hp.value2_w[j], part);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) values[t] = tanhf(part + hp.value2_b[0]);
        __syncwarp();
    }
}

}  // namespace az

using namespace az;
#define AZ_API extern "C" __attribute__((visibility("default")))

AZ_API int az_net_stem(const void* states, const float* w, const float* b, int32_t n, int32_t H, int32_t W, int32_t C,
                       void* out, void* stream) {
    if (!states || !w || !b || !out || n < 0 || H < 1 || W < 1) return fail_net(AZ_ERR_ARG, "az_net_stem: bad argument");
    if (C != 128) return fail_net(AZ_ERR_ARG, "az_net_stem: built for 128 filters (config.py:71)");
    if (n == 0) return AZ_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n < sms * 8 ? n : sms * 8;
    const size_t smem = sizeof(float4) * (size_t)(H + 2) * (W + 2);
    k_stem<128><<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(states), w, b, n, H, W, static_cast<__nv_bfloat16*>(out));
    if (cudaGetLastError() != cudaSuccess) return fail_net(AZ_ERR_CUDA, "az_net_stem: launch failed");
    return AZ_OK;
}

AZ_API int az_net_heads(const void* x, const az_head_weights* hw, int32_t n, int32_t cells, int32_t C, int32_t A,
                        float* priors, float* values, void* stream) {
    if (!x || !hw || !priors || !values || n < 0 || cells < 1 || A < 1 || A > AZ_MAX_ACTIONS)
        return fail_net(AZ_ERR_ARG, "az_net_heads: bad argument");
    if (C != 128) return fail_net(AZ_ERR_ARG, "az_net_heads: built for 128 filters (config.py:71)");
    if (n == 0) return AZ_OK;
    HeadParams hp{hw->conv_w, hw->conv_b, hw->policy_w, hw->policy_b, hw->value1_w, hw->value1_b, hw->value2_w, hw->value2_b,
                  n, cells, A};
    const size_t smem = sizeof(float) * (3 * 128 + (size_t)A * (2 * cells + 1) + (size_t)kHidden * (cells | 1) +
                                         (size_t)kHeadWarps * 3 * cells);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_heads<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail_net(AZ_ERR_CUDA, "az_net_heads: shared memory request refused");
        configured = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = (n + kHeadWarps - 1) / kHeadWarps;
    if (grid > sms * 2) grid = sms * 2;
    k_heads<128><<<grid, kHeadWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), hp, priors, values);
    if (cudaGetLastError() != cudaSuccess) return fail_net(AZ_ERR_CUDA, "az_net_heads: launch failed");
    return AZ_OK;
}
