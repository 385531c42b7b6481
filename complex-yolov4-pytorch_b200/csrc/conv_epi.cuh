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

// ---- BatchNorm statistics of a staged output slab -------------------------------------------------------------------------
// The warp has just written its [32 rows][32 columns] fp16 slab (64-byte rows, 64B swizzle: 16-byte chunk ^= (row >> 1) & 3)
// for the TMA store.  Statistics are taken from THAT tensor -- the values BatchNorm will normalise -- instead of a
// 31-shuffle reduce-scatter over the fp32 accumulators per statistic: lane (rh = lane >> 4, j = lane & 15) walks the half2
// column pair (2j, 2j+1) down the 16 rows 2t + rh (two rows per wavefront: conflict-free), accumulating sum (y - c) and
// sum (y - c)^2 about the per-channel shift c with packed fp32x2 arithmetic, rows >= nv (beyond M) skipped exactly; one
// xor-16 exchange completes the 32 rows.  ~100 instructions per slab instead of ~400, no dependent shuffle chain.
// On return (every lane): s1a/s1b = sum (y - c) of columns 2j / 2j+1 over the slab's valid rows, s2a/s2b the squares.
__device__ __forceinline__ void slab_stats(const uint8_t *slab, int lane, int nv, float c0, float c1, float &s1a, float &s1b,
                                           float &s2a, float &s2b)
{
    const int rh = lane >> 4, j = lane & 15;
    const uint8_t *base = slab + rh * 64 + (j & 3) * 4;
    const int ch = j >> 2;
    const unsigned long long nc = f32x2_pack(-c0, -c1);
    unsigned long long a1 = f32x2_pack(0.f, 0.f), a2 = a1, b1 = a1, b2 = a1;       // two independent chains
#pragma unroll
    for (int t = 0; t < 16; t += 2) {
        const uint32_t w0 = *(const uint32_t *)(base + t * 128 + ((ch ^ (t & 3)) << 4));
        const uint32_t w1 = *(const uint32_t *)(base + (t + 1) * 128 + ((ch ^ ((t + 1) & 3)) << 4));
        if (2 * t + rh < nv) {
            const float2 y = __half22float2(*(const __half2 *)&w0);
            const unsigned long long d = f32x2_add(f32x2_pack(y.x, y.y), nc);
            a1 = f32x2_add(a1, d); a2 = f32x2_fma(d, d, a2);
        }
        if (2 * t + 2 + rh < nv) {
            const float2 y = __half22float2(*(const __half2 *)&w1);
            const unsigned long long d = f32x2_add(f32x2_pack(y.x, y.y), nc);
            b1 = f32x2_add(b1, d); b2 = f32x2_fma(d, d, b2);
        }
    }
    a1 = f32x2_add(a1, b1); a2 = f32x2_add(a2, b2);
    f32x2_unpack(a1, s1a, s1b); f32x2_unpack(a2, s2a, s2b);
    s1a += __shfl_xor_sync(0xffffffffu, s1a, 16); s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
    s2a += __shfl_xor_sync(0xffffffffu, s2a, 16); s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
}

}  // namespace cy4
