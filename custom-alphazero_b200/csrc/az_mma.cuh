// az_mma.cuh - legacy tensor-path helpers shared by the stem / heads code (az_net.cu) and the fused
// advance kernel (az_kernels.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace az {

// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// two floats -> one register of two bf16 (lo in the low half): a single cvt.rn.bf16x2.f32
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

constexpr int kHidden = 256;  // ValueHead hidden_dim (model.py:109)

struct HeadParams {
    const float *conv_w, *conv_b, *policy_w, *policy_b, *value1_w, *value1_b, *value2_w, *value2_b;
    int n, cells, A;
};

}  // namespace az
