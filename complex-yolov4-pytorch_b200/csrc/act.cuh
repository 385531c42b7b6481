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

// ---- packed fp32x2 arithmetic (Blackwell add / mul / fma .f32x2: two fp32 lanes per instruction) ---------------------------
// The element-wise passes are issue-bound on the Mish math (B200 ncu: 70 % issue-active at 0.69 of the copy bandwidth); with
// everything but the two MUFU ops per element done on channel PAIRS the instruction count per element drops by ~40 %.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f32x2_pack(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f32x2_unpack(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 f32x2_add(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f32x2_mul(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 f32x2_fma(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 f32x2_bcast(float a) { return f32x2_pack(a, a); }
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Pair forms of mish_f / mish_grad_f (same formulas; e = 2^(z log2 e) through ex2.approx.ftz, 1/(n+2) through rcp.approx.ftz).
// Forward needs no clamp: e = inf gives 1/(n+2) = 0 and mish = z.  The derivative clamps the exponent at 40 log2(e) like the
// scalar form (e^2 stays finite; 4 z /(n+2)^2 underflows to 0 and the derivative is 1).
__device__ __forceinline__ f32x2 mish_f2(f32x2 z)
{
    float t0, t1;
    f32x2_unpack(f32x2_mul(z, f32x2_bcast(1.4426950408889634f)), t0, t1);
    const f32x2 e = f32x2_pack(ex2_ftz(t0), ex2_ftz(t1));
    float n0, n1;
    f32x2_unpack(f32x2_fma(e, f32x2_add(e, f32x2_bcast(2.f)), f32x2_bcast(2.f)), n0, n1);
    const f32x2 inv = f32x2_pack(rcp_ftz(n0), rcp_ftz(n1));
    return f32x2_fma(f32x2_mul(z, f32x2_bcast(-2.f)), inv, z);
}
__device__ __forceinline__ f32x2 mish_grad_f2(f32x2 z)
{
    float t0, t1;
    f32x2_unpack(f32x2_mul(z, f32x2_bcast(1.4426950408889634f)), t0, t1);
    const f32x2 e = f32x2_pack(ex2_ftz(fminf(t0, 57.70780163555854f)), ex2_ftz(fminf(t1, 57.70780163555854f)));
    float n0, n1;
    f32x2_unpack(f32x2_fma(e, f32x2_add(e, f32x2_bcast(2.f)), f32x2_bcast(2.f)), n0, n1);
    const f32x2 inv = f32x2_pack(rcp_ftz(n0), rcp_ftz(n1));
    const f32x2 q = f32x2_mul(f32x2_mul(f32x2_mul(z, f32x2_bcast(4.f)), inv), inv);          // 4 z / (n+2)^2
    return f32x2_fma(q, f32x2_fma(e, e, e), f32x2_fma(inv, f32x2_bcast(-2.f), f32x2_bcast(1.f)));
}
template <int ACT> __device__ __forceinline__ f32x2 act_t2(f32x2 z)
{
#ifdef CY4_SCALAR_ACT      // A/B side build (tools/gpu_call14.sh): the scalar Mish forms inside the packed passes
    if (ACT == ACT_MISH) { float a, b; f32x2_unpack(z, a, b); return f32x2_pack(mish_f(a), mish_f(b)); }
#endif
    if (ACT == ACT_MISH) return mish_f2(z);
    if (ACT == ACT_LEAKY) { float a, b; f32x2_unpack(z, a, b); return f32x2_pack(fmaxf(a, 0.1f * a), fmaxf(b, 0.1f * b)); }
    return z;
}
template <int ACT> __device__ __forceinline__ f32x2 act_grad_t2(f32x2 z)
{
#ifdef CY4_SCALAR_ACT
    if (ACT == ACT_MISH) { float a, b; f32x2_unpack(z, a, b); return f32x2_pack(mish_grad_f(a), mish_grad_f(b)); }
#endif
    if (ACT == ACT_MISH) return mish_grad_f2(z);
    if (ACT == ACT_LEAKY) { float a, b; f32x2_unpack(z, a, b); return f32x2_pack(a > 0.f ? 1.f : 0.1f, b > 0.f ? 1.f : 0.1f); }
    return f32x2_bcast(1.f);
}

}  // namespace cy4
