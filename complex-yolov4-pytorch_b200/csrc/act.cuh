// act.cuh -- activation functions shared by the element-wise passes (elementwise.cu) and the fused
// conv epilogues (conv_tc.cu): Mish (reference src/models/darknet2pytorch.py:22-28), LeakyReLU(0.1) (:265-266).
#pragma once
#include <cuda_runtime.h>

namespace cy4 {

enum { ACT_LINEAR = 0, ACT_LEAKY = 1, ACT_MISH = 2 };

// mish(z) = z * tanh(softplus(z)).  With e = e^z and n = e (e + 2):  tanh(log(1 + e)) = n / (n + 2)
//   mish  = z - 2 z / (n + 2)
//   mish' = 1 - 2/(n+2) + 4 z e (e + 1) / (n + 2)^2
// Both saturate by themselves (e -> inf: 1/(n+2) -> 0 => mish = z, mish' = 1; e -> 0: mish -> 0), which
// is torch's softplus threshold (20) behaviour to fp32 precision, so no select is needed.  The only
// hazard is inf * 0 in the derivative, avoided by clamping z at 40 (e^40 squared is still finite).
__device__ __forceinline__ float mish_f(float z)
{
    const float e = __expf(fminf(z, 40.f));
    const float inv = __fdividef(1.f, fmaf(e, e + 2.f, 2.f));
    return fmaf(-2.f * z, inv, z);
}
__device__ __forceinline__ float mish_grad_f(float z)
{
    const float e = __expf(fminf(z, 40.f));
    const float inv = __fdividef(1.f, fmaf(e, e + 2.f, 2.f));
    const float q = 4.f * z * inv * inv;                  // 4 z / (n+2)^2
    return fmaf(q, fmaf(e, e, e), fmaf(-2.f, inv, 1.f));
}
__device__ __forceinline__ float act_f(float z, int act)
{
    return act == ACT_MISH ? mish_f(z) : (act == ACT_LEAKY ? (z > 0.f ? z : 0.1f * z) : z);
}
__device__ __forceinline__ float act_grad_f(float z, int act)
{
    return act == ACT_MISH ? mish_grad_f(z) : (act == ACT_LEAKY ? (z > 0.f ? 1.f : 0.1f) : 1.f);
}
template <int ACT> __device__ __forceinline__ float act_t(float z) { return ACT == ACT_MISH ? mish_f(z) : (ACT == ACT_LEAKY ? (z > 0.f ? z : 0.1f * z) : z); }
template <int ACT> __device__ __forceinline__ float act_grad_t(float z) { return ACT == ACT_MISH ? mish_grad_f(z) : (ACT == ACT_LEAKY ? (z > 0.f ? 1.f : 0.1f) : 1.f); }

}  // namespace cy4
