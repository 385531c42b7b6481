// conv_epi.cuh -- the fused epilogue transforms shared by conv_tc.cu and conv_pair.cu (ConvKParams::epi_mode, conv_tc.cuh).
// Called by an epilogue thread that holds 32 consecutive fp32 accumulator columns f[0..32) of GEMM row m (output row `orow`,
// first column n0).  On return f holds the values to store; g the second statistics operand of EPI_BWD_DZ (the side value Y);
// accum_in_store tells the caller whether the fp16 read-modify-write of CONV_F_ACCUM is still to be done by the store code.
#pragma once
#include <cuda_fp16.h>

#include "act.cuh"
#include "conv_tc.cuh"

namespace cy4 {

__device__ __forceinline__ void epi_transform(const ConvKParams &p, float (&f)[32], float (&g)[32], bool &accum_in_store, int n0,
                                              int64_t orow, bool row_ok)
{
    if (p.epi_mode == EPI_FWD_ACT) {
        // ---- y = act(acc + shift[n]) (+ residual)
        const float4 *sh4 = (const float4 *)(p.epi_shift + n0);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(sh4 + (i >> 2));
            f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
        }
        if (p.epi_act == ACT_MISH) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = mish_f(f[i]);
        } else if (p.epi_act == ACT_LEAKY) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.f ? f[i] : 0.1f * f[i];
        }
        if (p.side && row_ok) {
            const uint4 *r4 = (const uint4 *)((const __half *)p.side + orow * p.ld_side + n0);
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                const uint4 rv = __ldg(r4 + (i >> 3));
                const __half2 *rh = (const __half2 *)&rv;
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(rh[j]); f[i + 2 * j] += t.x; f[i + 2 * j + 1] += t.y; }
            }
        }
    } else if (p.epi_mode == EPI_BWD_DZ) {
        // ---- dz = (acc (+ old)) * act'(scale[n] * Y + shift[n]); statistics operands (dz, dz * Y)
        if (accum_in_store) {
            accum_in_store = false;
            if (row_ok) {
                const uint4 *o4 = (const uint4 *)((const __half *)p.y + orow * p.ldy + n0);
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    const uint4 ov = o4[i >> 3];
                    const __half2 *oh = (const __half2 *)&ov;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(oh[j]); f[i + 2 * j] += t.x; f[i + 2 * j + 1] += t.y; }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) g[i] = 0.f;
        if (row_ok) {
            const uint4 *y4 = (const uint4 *)((const __half *)p.side + orow * p.ld_side + n0);
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                const uint4 yv = __ldg(y4 + (i >> 3));
                const __half2 *yh = (const __half2 *)&yv;
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(yh[j]); g[i + 2 * j] = t.x; g[i + 2 * j + 1] = t.y; }
            }
        }
        if (p.epi_act != ACT_LINEAR) {
            const float4 *sc4 = (const float4 *)(p.epi_scale + n0), *sh4 = (const float4 *)(p.epi_shift + n0);
            if (p.epi_act == ACT_MISH) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 a = __ldg(sc4 + (i >> 2)), b = __ldg(sh4 + (i >> 2));
                    f[i] *= mish_grad_f(fmaf(g[i], a.x, b.x));
                    f[i + 1] *= mish_grad_f(fmaf(g[i + 1], a.y, b.y));
                    f[i + 2] *= mish_grad_f(fmaf(g[i + 2], a.z, b.z));
                    f[i + 3] *= mish_grad_f(fmaf(g[i + 3], a.w, b.w));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 a = __ldg(sc4 + (i >> 2)), b = __ldg(sh4 + (i >> 2));
                    f[i] *= fmaf(g[i], a.x, b.x) > 0.f ? 1.f : 0.1f;
                    f[i + 1] *= fmaf(g[i + 1], a.y, b.y) > 0.f ? 1.f : 0.1f;
                    f[i + 2] *= fmaf(g[i + 2], a.z, b.z) > 0.f ? 1.f : 0.1f;
                    f[i + 3] *= fmaf(g[i + 3], a.w, b.w) > 0.f ? 1.f : 0.1f;
                }
            }
        }
    }
}

}  // namespace cy4
