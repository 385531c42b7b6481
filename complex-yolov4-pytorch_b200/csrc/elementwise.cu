// elementwise.cu -- the bandwidth-bound part of the network on NHWC fp16 tensors with an explicit
// channel stride (ld): BatchNorm (training statistics come fused out of the conv epilogue) + Mish /
// LeakyReLU forward and backward, shortcut adds, route (concat / slice) copies, SPP max pooling,
// nearest upsampling, dtype/scale conversions.  All kernels move 16-byte vectors (8 halves).
// Reference ops replaced: nn.BatchNorm2d, Mish, nn.LeakyReLU, torch.cat, shortcut add, nn.MaxPool2d,
// Upsample_expand in src/models/darknet2pytorch.py:22-28,64-79,180-219,256-285.
#include <cuda_fp16.h>

#include <mutex>
#include <set>

#include "common.cuh"
#include "act.cuh"

namespace cy4 {

__device__ __forceinline__ void unpack8(const uint4 &v, float f[8])
{
    const __half2 *h = (const __half2 *)&v;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float f[8])
{
    uint4 v; __half2 *h = (__half2 *)&v;
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// ---- BatchNorm statistics -> per-channel scale / shift ------------------------------------------
__global__ void bn_finalize_kernel(const float *__restrict__ ch_sum, const float *__restrict__ ch_sqsum, float count,
                                   const float *__restrict__ gamma, const float *__restrict__ beta, float *running_mean,
                                   float *running_var, long long *num_batches, float momentum, float eps, int training, int C,
                                   float *__restrict__ scale, float *__restrict__ shift, float *__restrict__ mean_out,
                                   float *__restrict__ rstd_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && num_batches) *num_batches += 1;
    if (c >= C) return;
    float mean, var;
    if (training) {
        const double inv_count = 1.0 / (double)count;
        const double m = (double)ch_sum[c] * inv_count;
        double v = (double)ch_sqsum[c] * inv_count - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        const float unbiased = count > 1.f ? (float)(v * (double)count / ((double)count - 1.0)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    } else {
        mean = running_mean[c]; var = running_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    mean_out[c] = mean;
    rstd_out[c] = rstd;
}

// ---- column-owner mapping for the BN/activation passes -------------------------------------------
// A thread owns ONE 8-channel vector column (its per-channel parameters stay in registers) and walks
// down a CONTIGUOUS chunk of rows owned by its block, kUnroll row groups per iteration with all loads
// issued before any math (memory-level parallelism, sequential DRAM pages per block).  A block covers
// (256 / vectors-per-row) rows per group, so a warp always touches whole 128-byte lines.
// vectors-per-row > 256 (C > 2048) is handled by an outer loop.  Templated on the activation so the
// Mish / LeakyReLU math is branch-free.
constexpr int kUnroll = 4;


struct RowChunk { int64_t begin, end; };
__device__ __forceinline__ RowChunk block_rows(int64_t M, int rpi)
{
    // rows [begin, end) of this block: equal chunks, multiples of rpi * kUnroll
    const int64_t unit = (int64_t)rpi * kUnroll;
    const int64_t units = (M + unit - 1) / unit;
    const int64_t per = (units + gridDim.x - 1) / gridDim.x;
    RowChunk r;
    r.begin = min(M, (int64_t)blockIdx.x * per * unit);
    r.end = min(M, r.begin + per * unit);
    return r;
}

// Training-mode statistics -> scale / shift inside the apply pass (saves the separate cy4_bn_finalize launch per layer).
struct BnFin {
    const float *ch_sum, *ch_sqsum, *gamma, *beta;
    float *running_mean, *running_var; long long *num_batches;
    float *scale_out, *shift_out, *mean_out, *rstd_out;
    const float *stat_shift;      // NULL, or c[ch]: ch_sum / ch_sqsum are sums of (y - c), (y - c)^2 (cy4_conv_fwd_stats)
    float *shift_next;            // NULL, or where the batch mean is published as the NEXT step's c (never the array read above)
    float count, momentum, eps;
    double inv_count;
};

// out = act(y * scale + shift) (+ residual)
// FIN: scale / shift are derived here from the batch sums (every thread for its own 8 channels); block 0 also publishes
// scale / shift / mean / rstd for the backward pass and updates the running statistics exactly like nn.BatchNorm2d.
template <int ACT, bool RES, bool FIN>
__global__ void __launch_bounds__(256, 3)
bn_act_fwd_kernel(const __half *__restrict__ y, int64_t ldy, const float *__restrict__ scale, const float *__restrict__ shift,
                  const __half *__restrict__ res, int64_t ldr, __half *__restrict__ out, int64_t ldo, int64_t M, int C, const BnFin fin)
{
    pdl_trigger();
    pdl_wait();                              // (reads per-channel data the preceding kernel produced right away)
    const int vpr = C >> 3;
    for (int v0 = 0; v0 < vpr; v0 += 256) {
        const int nv = min(256, vpr - v0);
        const int rpi = 256 / nv;
        const int vec = threadIdx.x % nv, rsub = threadIdx.x / nv;
        if (rsub >= rpi) continue;
        const int c0 = (v0 + vec) << 3;
        float sc[8], sh[8];
        if (FIN) {
            const bool publish = blockIdx.x == 0 && rsub == 0;
            if (publish && c0 == 0 && fin.num_batches) *fin.num_batches += 1;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = c0 + k;
                // same arithmetic as bn_finalize_kernel (double mean / variance), with the divisions by the count turned into
                // one multiplication by its double reciprocal: every thread of the grid executes this prologue
                const double ms = (double)fin.ch_sum[c] * fin.inv_count;              // mean of (y - c)
                double v = (double)fin.ch_sqsum[c] * fin.inv_count - ms * ms;         // variance is shift invariant
                if (v < 0.0) v = 0.0;
                const double m = ms + (fin.stat_shift ? (double)fin.stat_shift[c] : 0.0);
                const float mean = (float)m, var = (float)v;
                const float rstd = rsqrtf(var + fin.eps);
                sc[k] = fin.gamma[c] * rstd;
                sh[k] = fin.beta[c] - mean * sc[k];
                if (publish) {
                    const float unbiased = fin.count > 1.f ? (float)(v * (double)fin.count / ((double)fin.count - 1.0)) : var;
                    fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mean;
                    fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * unbiased;
                    fin.scale_out[c] = sc[k]; fin.shift_out[c] = sh[k]; fin.mean_out[c] = mean; fin.rstd_out[c] = rstd;
                    if (fin.shift_next) fin.shift_next[c] = mean;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = scale[c0 + k]; sh[k] = shift[c0 + k]; }
        }
        f32x2 sc2[4], sh2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { sc2[k] = f32x2_pack(sc[2 * k], sc[2 * k + 1]); sh2[k] = f32x2_pack(sh[2 * k], sh[2 * k + 1]); }
        const RowChunk rc = block_rows(M, rpi);
        for (int64_t m = rc.begin + rsub; m < rc.end; m += (int64_t)rpi * kUnroll) {
            uint4 vy[kUnroll], vr[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t mm = m + u * rpi;
                if (mm < rc.end) {
                    vy[u] = *(const uint4 *)(y + mm * ldy + c0);
                    if (RES) vr[u] = *(const uint4 *)(res + mm * ldr + c0);
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t mm = m + u * rpi;
                if (mm < rc.end) {
                    const __half2 *hy = (const __half2 *)&vy[u], *hr = (const __half2 *)&vr[u];
                    uint4 o; __half2 *ho = (__half2 *)&o;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                 // channel pairs: packed fp32x2 math (act.cuh)
                        const float2 yv = __half22float2(hy[k]);
                        f32x2 a = act_t2<ACT>(f32x2_fma(f32x2_pack(yv.x, yv.y), sc2[k], sh2[k]));
                        if (RES) { const float2 rv = __half22float2(hr[k]); a = f32x2_add(a, f32x2_pack(rv.x, rv.y)); }
                        float a0, a1;
                        f32x2_unpack(a, a0, a1);
                        ho[k] = __floats2half2_rn(a0, a1);
                    }
                    *(uint4 *)(out + mm * ldo + c0) = o;
                }
            }
        }
    }
}

// Per-channel sums of dz = dA * act'(z) and dz * xhat  (xhat = (y - mean) * rstd), accumulated as
// sum dz and sum dz*y per thread and combined as rstd * (sum dz*y - mean * sum dz).
template <int ACT>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_reduce_kernel(const __half *__restrict__ y, int64_t ldy, __half *__restrict__ dA, int64_t ldg,
                         const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                         const float *__restrict__ rstd, int64_t M, int C, float *__restrict__ sum_dz,
                         float *__restrict__ sum_dzx)
{
    pdl_trigger();
    pdl_wait();                              // (reads per-channel data the preceding kernel produced right away)
    const int vpr = C >> 3;
    __shared__ float red[2][256][8 + 1];
    for (int v0 = 0; v0 < vpr; v0 += 256) {
        const int nv = min(256, vpr - v0);
        const int rpi = 256 / nv;
        const int vec = threadIdx.x % nv, rsub = threadIdx.x / nv;
        f32x2 a1[4], a2[4];                               // per channel pair: sum dz, sum dz*y
#pragma unroll
        for (int k = 0; k < 4; ++k) { a1[k] = f32x2_bcast(0.f); a2[k] = a1[k]; }
        const int c0 = (v0 + vec) << 3;
        if (rsub < rpi) {
            f32x2 sc2[4], sh2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sc2[k] = f32x2_pack(scale[c0 + 2 * k], scale[c0 + 2 * k + 1]);
                sh2[k] = f32x2_pack(shift[c0 + 2 * k], shift[c0 + 2 * k + 1]);
            }
            const RowChunk rc = block_rows(M, rpi);
            for (int64_t m = rc.begin + rsub; m < rc.end; m += (int64_t)rpi * kUnroll) {
                uint4 vy[kUnroll], vg[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int64_t mm = m + u * rpi;
                    if (mm < rc.end) { vy[u] = *(const uint4 *)(y + mm * ldy + c0); vg[u] = *(const uint4 *)(dA + mm * ldg + c0); }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const int64_t mm = m + u * rpi;
                    if (mm < rc.end) {
                        const __half2 *hy = (const __half2 *)&vy[u], *hg = (const __half2 *)&vg[u];
                        uint4 o; __half2 *ho = (__half2 *)&o;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 yv = __half22float2(hy[k]), gv = __half22float2(hg[k]);
                            const f32x2 yp = f32x2_pack(yv.x, yv.y);
                            const f32x2 dz = f32x2_mul(f32x2_pack(gv.x, gv.y), act_grad_t2<ACT>(f32x2_fma(yp, sc2[k], sh2[k])));
                            a1[k] = f32x2_add(a1[k], dz);
                            a2[k] = f32x2_fma(dz, yp, a2[k]);
                            float d0, d1;
                            f32x2_unpack(dz, d0, d1);
                            ho[k] = __floats2half2_rn(d0, d1);
                        }
                        if (ACT != ACT_LINEAR) *(uint4 *)(dA + mm * ldg + c0) = o;   // dz replaces dA (fp16)
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f32x2_unpack(a1[k], red[0][threadIdx.x][2 * k], red[0][threadIdx.x][2 * k + 1]);
            f32x2_unpack(a2[k], red[1][threadIdx.x][2 * k], red[1][threadIdx.x][2 * k + 1]);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nv * 8; t += 256) {
            const int vv = t >> 3, kk = t & 7;
            float s1 = 0.f, s2 = 0.f;
            for (int r = 0; r < rpi; ++r) { s1 += red[0][r * nv + vv][kk]; s2 += red[1][r * nv + vv][kk]; }
            const int c = ((v0 + vv) << 3) + kk;
            atomicAdd(sum_dz + c, s1);
            atomicAdd(sum_dzx + c, rstd[c] * (s2 - mean[c] * s1));
        }
        __syncthreads();
    }
}

// raw sums of a fused dgrad epilogue (sum dz, sum dz*y) -> sum dz*xhat = rstd * (sum dz*y - mean * sum dz), in place
__global__ void bn_bwd_fixup_kernel(const float *__restrict__ sum_dz, float *__restrict__ sum_dzy, const float *__restrict__ mean,
                                    const float *__restrict__ rstd, int C)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) sum_dzy[c] = rstd[c] * (sum_dzy[c] - mean[c] * sum_dz[c]);
}

// dY = scale * (dz - sum_dz/M - xhat * sum_dzx/M)   (training-mode BN backward; eval: dY = scale*dz)
//    = scale * dz + A * y + B   with  A = -scale*rstd*sum_dzx/M,  B = -scale*sum_dz/M - A*mean
// DZ_READY: the reduce pass already replaced dA by dz, so this pass is three FMAs per element.
template <int ACT, bool DZ_READY>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_kernel(const __half *__restrict__ y, int64_t ldy, const __half *__restrict__ dA, int64_t ldg,
                        const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                        const float *__restrict__ rstd, const float *__restrict__ sum_dz, const float *__restrict__ sum_dzx,
                        float inv_count, int training, __half *__restrict__ dY, int64_t ldd, int64_t M, int C)
{
    pdl_trigger();
    pdl_wait();                              // (reads per-channel data the preceding kernel produced right away)
    const int vpr = C >> 3;
    for (int v0 = 0; v0 < vpr; v0 += 256) {
        const int nv = min(256, vpr - v0);
        const int rpi = 256 / nv;
        const int vec = threadIdx.x % nv, rsub = threadIdx.x / nv;
        if (rsub >= rpi) continue;
        const int c0 = (v0 + vec) << 3;
        float sc[8], sh[8], A[8], Bc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = c0 + k;
            sc[k] = scale[c]; sh[k] = shift[c];
            if (training) {
                A[k] = -sc[k] * rstd[c] * sum_dzx[c] * inv_count;
                Bc[k] = -sc[k] * sum_dz[c] * inv_count - A[k] * mean[c];
            } else { A[k] = 0.f; Bc[k] = 0.f; }
        }
        f32x2 sc2[4], sh2[4], A2[4], B2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sc2[k] = f32x2_pack(sc[2 * k], sc[2 * k + 1]); sh2[k] = f32x2_pack(sh[2 * k], sh[2 * k + 1]);
            A2[k] = f32x2_pack(A[2 * k], A[2 * k + 1]); B2[k] = f32x2_pack(Bc[2 * k], Bc[2 * k + 1]);
        }
        const RowChunk rc = block_rows(M, rpi);
        for (int64_t m = rc.begin + rsub; m < rc.end; m += (int64_t)rpi * kUnroll) {
            uint4 vy[kUnroll], vg[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t mm = m + u * rpi;
                if (mm < rc.end) { vy[u] = *(const uint4 *)(y + mm * ldy + c0); vg[u] = *(const uint4 *)(dA + mm * ldg + c0); }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t mm = m + u * rpi;
                if (mm < rc.end) {
                    const __half2 *hy = (const __half2 *)&vy[u], *hg = (const __half2 *)&vg[u];
                    uint4 o; __half2 *ho = (__half2 *)&o;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 yv = __half22float2(hy[k]), gv = __half22float2(hg[k]);
                        const f32x2 yp = f32x2_pack(yv.x, yv.y);
                        f32x2 dz = f32x2_pack(gv.x, gv.y);
                        if (!DZ_READY) dz = f32x2_mul(dz, act_grad_t2<ACT>(f32x2_fma(yp, sc2[k], sh2[k])));
                        float o0, o1;
                        f32x2_unpack(f32x2_fma(sc2[k], dz, f32x2_fma(A2[k], yp, B2[k])), o0, o1);
                        ho[k] = __floats2half2_rn(o0, o1);
                    }
                    *(uint4 *)(dY + mm * ldd + c0) = o;
                }
            }
        }
    }
}

// out = a (+ b)      (copy / add over channel slices; also gradient accumulation)
__global__ void __launch_bounds__(256)
add_copy_kernel(const __half *__restrict__ a, int64_t lda, const __half *__restrict__ b, int64_t ldb, __half *__restrict__ out,
                int64_t ldo, int64_t M, int C)
{
    pdl_trigger();
    pdl_wait();
    const int vpr = C >> 3;
    const int64_t total = M * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / vpr;
        const int c0 = (int)(i - m * vpr) << 3;
        uint4 va = *(const uint4 *)(a + m * lda + c0);
        if (b) {
            float fa[8], fb[8];
            unpack8(va, fa);
            unpack8(*(const uint4 *)(b + m * ldb + c0), fb);
#pragma unroll
            for (int k = 0; k < 8; ++k) fa[k] += fb[k];
            va = pack8(fa);
        }
        *(uint4 *)(out + m * ldo + c0) = va;
    }
}

// nearest x2 upsample: out[b, 2h+i, 2w+j, c] = in[b, h, w, c]
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const __half *__restrict__ in, int64_t ldi, __half *__restrict__ out, int64_t ldo, int B, int H, int W, int C)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * (2 * H) * (2 * W) * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int ow = (int)(r % (2 * W)); r /= (2 * W);
        const int oh = (int)(r % (2 * H));
        const int b = (int)(r / (2 * H));
        const int64_t src = ((int64_t)b * H + (oh >> 1)) * W + (ow >> 1);
        const int64_t dst = ((int64_t)b * 2 * H + oh) * (2 * W) + ow;
        *(uint4 *)(out + dst * ldo + c0) = *(const uint4 *)(in + src * ldi + c0);
    }
}
// gin[b,h,w,c] (+)= sum of the 2x2 block of gout
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const __half *__restrict__ gout, int64_t ldo, __half *__restrict__ gin, int64_t ldi, int B, int H, int W, int C,
                      int accumulate)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * H * W * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int b = (int)(r / H);
        float acc[8], t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                const int64_t src = ((int64_t)b * 2 * H + 2 * h + dy) * (2 * W) + 2 * w + dx;
                unpack8(*(const uint4 *)(gout + src * ldo + c0), t);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += t[k];
            }
        __half *dst = gin + (((int64_t)b * H + h) * W + w) * ldi + c0;
        if (accumulate) {
            unpack8(*(const uint4 *)dst, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += t[k];
        }
        *(uint4 *)dst = pack8(acc);
    }
}

// max pooling, window k, stride s, symmetric pad (nn.MaxPool2d(k, s, k//2) for s == 1, pad 0 for k == s)
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __half *__restrict__ in, int64_t ldi, __half *__restrict__ out, int64_t ldo, int B, int H, int W, int C,
                   int k, int s, int pad, int Ho, int Wo)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * Ho * Wo * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float best[8], t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) best[q] = -INFINITY;
        for (int dy = 0; dy < k; ++dy) {
            const int h = oh * s - pad + dy;
            if (h < 0 || h >= H) continue;
            for (int dx = 0; dx < k; ++dx) {
                const int w = ow * s - pad + dx;
                if (w < 0 || w >= W) continue;
                unpack8(*(const uint4 *)(in + (((int64_t)b * H + h) * W + w) * ldi + c0), t);
#pragma unroll
                for (int q = 0; q < 8; ++q) best[q] = fmaxf(best[q], t[q]);
            }
        }
        *(uint4 *)(out + (((int64_t)b * Ho + oh) * Wo + ow) * ldo + c0) = pack8(best);
    }
}
// Forward that also records, per output element, WHICH window element was the first maximum (torch's row-major scan order) as
// the window offset dy * k + dx in one byte: the backward pass then routes gradients without re-scanning the k x k windows
// (169 16-byte loads per output vector for the 13 x 13 SPP pool).
// Separable, stride 1: a vertical pass (column maximum + the smallest dy that attains it) into a workspace, then a horizontal
// pass over the column maxima with the tie-break (smaller dy, then smaller dx) -- which IS the first maximum of the row-major scan:
// the minimal dy over all maxima is the minimum of the columns' dy, and among those columns the leftmost wins.  2k loads per
// output instead of k^2.  Rows / columns outside the image are skipped (padding never wins).
__global__ void __launch_bounds__(256)
maxpool_v_kernel(const __half *__restrict__ in, int64_t ldi, __half *__restrict__ vmax, uint8_t *__restrict__ vdy, int B, int H, int W, int C,
                 int k, int pad)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * H * W * vpr;              // one entry per (output row oh, input column w): Ho == H for stride 1
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int w = (int)(r % W); r /= W;
        const int oh = (int)(r % H);
        const int b = (int)(r / H);
        float best[8], t[8]; int bdy[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bdy[q] = -1; }
        for (int dy = 0; dy < k; ++dy) {
            const int h = oh - pad + dy;
            if (h < 0 || h >= H) continue;
            unpack8(*(const uint4 *)(in + (((int64_t)b * H + h) * W + w) * ldi + c0), t);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (t[q] > best[q] || bdy[q] < 0) { best[q] = t[q]; bdy[q] = dy; }
        }
        const int64_t o = ((int64_t)b * H + oh) * W + w;
        *(uint4 *)(vmax + o * C + c0) = pack8(best);
        uint2 pk;
        pk.x = (uint32_t)bdy[0] | ((uint32_t)bdy[1] << 8) | ((uint32_t)bdy[2] << 16) | ((uint32_t)bdy[3] << 24);
        pk.y = (uint32_t)bdy[4] | ((uint32_t)bdy[5] << 8) | ((uint32_t)bdy[6] << 16) | ((uint32_t)bdy[7] << 24);
        *(uint2 *)(vdy + o * C + c0) = pk;
    }
}
__global__ void __launch_bounds__(256)
maxpool_h_kernel(const __half *__restrict__ vmax, const uint8_t *__restrict__ vdy, __half *__restrict__ out, int64_t ldo,
                 uint8_t *__restrict__ argmax, int B, int H, int W, int C, int k, int pad)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * H * W * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int ow = (int)(r % W); r /= W;
        const int oh = (int)(r % H);
        const int b = (int)(r / H);
        float best[8], t[8]; int bdy[8], bdx[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bdy[q] = 1 << 20; bdx[q] = -1; }
        for (int dx = 0; dx < k; ++dx) {
            const int w = ow - pad + dx;
            if (w < 0 || w >= W) continue;
            const int64_t o = ((int64_t)b * H + oh) * W + w;
            unpack8(*(const uint4 *)(vmax + o * C + c0), t);
            const uint2 pk = *(const uint2 *)(vdy + o * C + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int dy = (int)(((q < 4 ? pk.x : pk.y) >> ((q & 3) * 8)) & 255u);
                if (bdx[q] < 0 || t[q] > best[q] || (t[q] == best[q] && dy < bdy[q])) { best[q] = t[q]; bdy[q] = dy; bdx[q] = dx; }
            }
        }
        const int64_t o = ((int64_t)b * H + oh) * W + ow;
        *(uint4 *)(out + o * ldo + c0) = pack8(best);
        int pos[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) pos[q] = bdy[q] * k + bdx[q];
        uint2 pk;
        pk.x = (uint32_t)pos[0] | ((uint32_t)pos[1] << 8) | ((uint32_t)pos[2] << 16) | ((uint32_t)pos[3] << 24);
        pk.y = (uint32_t)pos[4] | ((uint32_t)pos[5] << 8) | ((uint32_t)pos[6] << 16) | ((uint32_t)pos[7] << 24);
        *(uint2 *)(argmax + o * C + c0) = pk;
    }
}
// general (strided) form: one k x k scan per output
__global__ void __launch_bounds__(256)
maxpool_fwd_idx_kernel(const __half *__restrict__ in, int64_t ldi, __half *__restrict__ out, int64_t ldo, uint8_t *__restrict__ argmax,
                       int B, int H, int W, int C, int k, int s, int pad, int Ho, int Wo)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * Ho * Wo * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float best[8], t[8]; int bpos[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bpos[q] = -1; }
        for (int dy = 0; dy < k; ++dy) {
            const int h = oh * s - pad + dy;
            if (h < 0 || h >= H) continue;
            for (int dx = 0; dx < k; ++dx) {
                const int w = ow * s - pad + dx;
                if (w < 0 || w >= W) continue;
                unpack8(*(const uint4 *)(in + (((int64_t)b * H + h) * W + w) * ldi + c0), t);
                const int pos = dy * k + dx;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (t[q] > best[q] || bpos[q] < 0) { best[q] = t[q]; bpos[q] = pos; }
            }
        }
        const int64_t o = ((int64_t)b * Ho + oh) * Wo + ow;
        *(uint4 *)(out + o * ldo + c0) = pack8(best);
        uint2 pk;
        pk.x = (uint32_t)(bpos[0] & 255) | ((uint32_t)(bpos[1] & 255) << 8) | ((uint32_t)(bpos[2] & 255) << 16) | ((uint32_t)(bpos[3] & 255) << 24);
        pk.y = (uint32_t)(bpos[4] & 255) | ((uint32_t)(bpos[5] & 255) << 8) | ((uint32_t)(bpos[6] & 255) << 16) | ((uint32_t)(bpos[7] & 255) << 24);
        *(uint2 *)(argmax + o * C + c0) = pk;
    }
}
__global__ void __launch_bounds__(256)
maxpool_bwd_idx_kernel(const uint8_t *__restrict__ argmax, const __half *__restrict__ gout, int64_t ldo, float *__restrict__ gscratch,
                       int B, int H, int W, int C, int k, int s, int pad, int Ho, int Wo)
{
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * Ho * Wo * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const int64_t o = ((int64_t)b * Ho + oh) * Wo + ow;
        const uint2 pk = *(const uint2 *)(argmax + o * C + c0);
        float g[8];
        unpack8(*(const uint4 *)(gout + o * ldo + c0), g);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pos = (int)(((q < 4 ? pk.x : pk.y) >> ((q & 3) * 8)) & 255u);
            const int dy = pos / k, dx = pos - dy * k;
            const int h = oh * s - pad + dy, w = ow * s - pad + dx;
            atomicAdd(gscratch + (((int64_t)b * H + h) * W + w) * C + c0 + q, g[q]);
        }
    }
}
// gradient: each output pixel routes its gradient to the FIRST maximal element of its window
// (torch's scan order), accumulated in an fp32 scratch [B,H,W,C] with atomics.
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __half *__restrict__ in, int64_t ldi, const __half *__restrict__ gout, int64_t ldo, float *__restrict__ gscratch,
                   int B, int H, int W, int C, int k, int s, int pad, int Ho, int Wo)
{
    // one thread per (output pixel, 8-channel vector): 16-byte window loads, per-channel first-max tracking
    const int vpr = C >> 3;
    const int64_t total = (int64_t)B * Ho * Wo * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr;
        const int c0 = (int)(i - r * vpr) << 3;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float best[8]; int bpos[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bpos[q] = -1; }
        for (int dy = 0; dy < k; ++dy) {
            const int h = oh * s - pad + dy;
            if (h < 0 || h >= H) continue;
            for (int dx = 0; dx < k; ++dx) {
                const int w = ow * s - pad + dx;
                if (w < 0 || w >= W) continue;
                float t[8];
                unpack8(*(const uint4 *)(in + (((int64_t)b * H + h) * W + w) * ldi + c0), t);
                const int pos = h * W + w;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (t[q] > best[q] || bpos[q] < 0) { best[q] = t[q]; bpos[q] = pos; }
            }
        }
        float g[8];
        unpack8(*(const uint4 *)(gout + (((int64_t)b * Ho + oh) * Wo + ow) * ldo + c0), g);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (bpos[q] >= 0) atomicAdd(gscratch + ((int64_t)b * H * W + bpos[q]) * C + c0 + q, g[q]);
    }
}
// gin (+)= fp32 scratch
__global__ void __launch_bounds__(256)
f32_to_f16_accum_kernel(const float *__restrict__ src, int64_t lds, float scale, const float *__restrict__ dscale,
                        __half *__restrict__ dst, int64_t ldd, int64_t M, int C, int accumulate)
{
    if (dscale) scale *= __ldg(dscale);
    const int vpr = C >> 3;
    const int64_t total = M * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / vpr;
        const int c0 = (int)(i - m * vpr) << 3;
        const float4 a = *(const float4 *)(src + m * lds + c0), b = *(const float4 *)(src + m * lds + c0 + 4);
        float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
        __half *d = dst + m * ldd + c0;
        if (accumulate) {
            float t[8];
            unpack8(*(const uint4 *)d, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += t[k];
        }
        *(uint4 *)d = pack8(f);
    }
}

// per-channel column sums of a fp32 [M, ld] matrix (bias gradient of the head convs)
// max |x| over a fp32 buffer -> atomicMax on the bit pattern (non-negative floats order like ints)
__global__ void absmax_f32_kernel(const float *__restrict__ src, int64_t n, float *__restrict__ out)
{
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(src[i]);
        m = (v > m || v != v) ? v : m;             // NaN propagates
    }
    for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, m, o); m = (t > m || t != t) ? t : m; }
    if ((threadIdx.x & 31) == 0) atomicMax((int *)out, __float_as_int(m != m ? INFINITY : m));
}
// scale[0] = 2^k with amax * 2^k ~ target (clamped), scale[1] = 1 / scale[0]; amax == 0 / inf -> 1
__global__ void make_scale_kernel(const float *__restrict__ amax, float target, float *__restrict__ scale)
{
    const float a = amax[0];
    float s = 1.f;
    if (a > 0.f && a < INFINITY) {
        int e = (int)floorf(log2f(target / a));
        e = max(-14, min(24, e));
        s = exp2f((float)e);
    }
    scale[0] = s; scale[1] = 1.f / s;
}

// out[c] += scale * sum_m src[m, c]   (out zeroed by the caller unless accumulating).  Row-major [M, lds] fp32: a warp reads
// one 128-byte row segment (32 columns) per load, 8 warps x 4 independent accumulators walk a contiguous row chunk per block,
// one atomicAdd per column per block.  (The first version ran one block per COLUMN over the whole matrix with 4-byte strided
// reads: 30 blocks, 130-600 us per YOLO head on the backward critical path.)
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float *__restrict__ src, int64_t lds, int64_t M, int C, float scale, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ float red[8][33];
    const int64_t per = (M + gridDim.x - 1) / gridDim.x;
    const int64_t m0 = (int64_t)blockIdx.x * per, m1 = min(M, m0 + per);
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (c < C) {
            int64_t m = m0 + w;
            for (; m + 24 < m1; m += 32) {
                a0 += src[m * lds + c]; a1 += src[(m + 8) * lds + c]; a2 += src[(m + 16) * lds + c]; a3 += src[(m + 24) * lds + c];
            }
            for (; m < m1; m += 8) a0 += src[m * lds + c];
        }
        red[w][lane] = (a0 + a1) + (a2 + a3);
        __syncthreads();
        if (w == 0 && c < C) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k][lane];
            atomicAdd(out + c, t * scale);
        }
        __syncthreads();
    }
}

// Experiment kept as an option ("ew_carveout", default 0, read at a kernel's first launch): ask for the maximum shared-memory
// carve-out so that blocks of these passes can be placed beside a resident weight-gradient CTA (~194 KB of shared memory) of
// the overlapped backward.  Measured on B200 (profiles/r2_overlap_ab.md): 0.7 ms per step SLOWER -- bn_act_bwd_apply alone
// 3.96 -> 4.83 ms: the streaming passes do use their L1 -- and the overlap gains nothing from it.
extern int g_ew_carveout;
static void co_resident(const void *kernel)
{
    static std::mutex mu;
    static std::set<const void *> done;
    std::lock_guard<std::mutex> lock(mu);
    if (!done.insert(kernel).second) return;
    if (g_ew_carveout)
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

static inline int ew_grid(int64_t total)
{
    const int64_t need = (total + 255) / 256;
    return (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)sm_count() * 16));
}

// grid for the column-owner kernels: enough blocks for ~4 resident 256-thread blocks per SM, but
// no more than one block per kUnroll row groups
// Grid cap in blocks per SM: ONE resident wave (3 blocks of the forward pass, 2 of the backward passes fit an SM) -- measured
// on B200 (profiles/r2_overlap_ab.md): 6 blocks per SM (two waves, each block re-deriving its per-channel constants) costs
// 0.2 ms per step in the forward passes and 0.85 ms in the backward ones.  Options "ew_fwd_blocks_per_sm" / "ew_bwd_blocks_per_sm".
extern int g_ew_fwd_bpsm, g_ew_bwd_bpsm;
static inline int col_grid(int64_t M, int C, int blocks_per_sm)
{
    const int vpr = C / 8;
    const int rpi = std::max(1, 256 / std::min(vpr, 256));
    const int64_t groups = (M + (int64_t)rpi * kUnroll - 1) / ((int64_t)rpi * kUnroll);
    return (int)std::max<int64_t>(1, std::min<int64_t>(groups, (int64_t)sm_count() * blocks_per_sm));
}

}  // namespace cy4

using namespace cy4;

#define EW_CHECK_C(C, who) CY4_CHECK_ARG((C) > 0 && ((C) % 8) == 0, who ": channels must be a positive multiple of 8")

extern "C" {

int cy4_bn_finalize(const float *ch_sum, const float *ch_sqsum, float count, const float *gamma, const float *beta,
                    float *running_mean, float *running_var, int64_t *num_batches_tracked, float momentum, float eps,
                    int training, int C, float *scale, float *shift, float *mean, float *rstd, void *stream)
{
    CY4_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && mean && rstd && C > 0 &&
                  (!training || (ch_sum && ch_sqsum && count > 0)), "cy4_bn_finalize: bad argument");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ch_sum, ch_sqsum, count, gamma, beta, running_mean, running_var,
                                                                         (long long *)num_batches_tracked, momentum, eps, training, C,
                                                                         scale, shift, mean, rstd);
    return cy4_launch_status("cy4_bn_finalize");
}

static int bn_act_fwd_launch(const void *y, int64_t ldy, const float *scale, const float *shift, int act, const void *residual, int64_t ldr,
                             void *out, int64_t ldo, int64_t M, int C, const BnFin *fin, void *stream)
{
    BnFin f;
    memset(&f, 0, sizeof(f));
    if (fin) f = *fin;
#define CY4_FWD(ACT, RES, FIN)                                                                                              \
    do {                                                                                                                    \
        co_resident((const void *)bn_act_fwd_kernel<ACT, RES, FIN>);                                                        \
        launch_pdl(bn_act_fwd_kernel<ACT, RES, FIN>, col_grid(M, C, g_ew_fwd_bpsm), 256, (cudaStream_t)stream, (const __half *)y, ldy, scale, shift, \
                   (const __half *)residual, ldr, (__half *)out, ldo, M, C, f);                                                 \
    } while (0)
#define CY4_FWD_A(RES, FIN)                                                                     \
    do {                                                                                        \
        if (act == ACT_MISH) CY4_FWD(ACT_MISH, RES, FIN);                                       \
        else if (act == ACT_LEAKY) CY4_FWD(ACT_LEAKY, RES, FIN);                                \
        else CY4_FWD(ACT_LINEAR, RES, FIN);                                                     \
    } while (0)
    if (fin) { if (residual) CY4_FWD_A(true, true); else CY4_FWD_A(false, true); }
    else { if (residual) CY4_FWD_A(true, false); else CY4_FWD_A(false, false); }
#undef CY4_FWD_A
#undef CY4_FWD
    return cy4_launch_status("cy4_bn_act_fwd");
}

int cy4_bn_act_fwd(const void *y, int64_t ldy, const float *scale, const float *shift, int act, const void *residual, int64_t ldr,
                   void *out, int64_t ldo, int64_t M, int C, void *stream)
{
    EW_CHECK_C(C, "cy4_bn_act_fwd");
    CY4_CHECK_ARG(y && scale && shift && out && M >= 0 && (ldy % 8) == 0 && (ldo % 8) == 0 && (ldr % 8) == 0, "cy4_bn_act_fwd: bad argument");
    if (M == 0) return 0;
    return bn_act_fwd_launch(y, ldy, scale, shift, act, residual, ldr, out, ldo, M, C, nullptr, stream);
}

int cy4_bn_train_act_fwd(const void *y, int64_t ldy, const float *ch_sum, const float *ch_sqsum, float count, const float *gamma,
                         const float *beta, float *running_mean, float *running_var, int64_t *num_batches_tracked, float momentum,
                         float eps, float *scale, float *shift, float *mean, float *rstd, int act, const void *residual, int64_t ldr,
                         void *out, int64_t ldo, int64_t M, int C, const float *stat_shift, float *shift_next, void *stream)
{
    EW_CHECK_C(C, "cy4_bn_train_act_fwd");
    CY4_CHECK_ARG(!shift_next || shift_next != stat_shift, "cy4_bn_train_act_fwd: shift_next must not alias stat_shift (every block reads it)");
    CY4_CHECK_ARG(y && ch_sum && ch_sqsum && gamma && beta && running_mean && running_var && scale && shift && mean && rstd && out &&
                  count > 0 && M > 0 && (ldy % 8) == 0 && (ldo % 8) == 0 && (ldr % 8) == 0, "cy4_bn_train_act_fwd: bad argument");
    BnFin f;
    f.ch_sum = ch_sum; f.ch_sqsum = ch_sqsum; f.gamma = gamma; f.beta = beta;
    f.running_mean = running_mean; f.running_var = running_var; f.num_batches = (long long *)num_batches_tracked;
    f.scale_out = scale; f.shift_out = shift; f.mean_out = mean; f.rstd_out = rstd;
    f.count = count; f.momentum = momentum; f.eps = eps; f.inv_count = 1.0 / (double)count;
    f.stat_shift = stat_shift; f.shift_next = shift_next;
    return bn_act_fwd_launch(y, ldy, nullptr, nullptr, act, residual, ldr, out, ldo, M, C, &f, stream);
}

int cy4_bn_act_bwd_reduce(const void *y, int64_t ldy, void *dA, int64_t ldg, const float *scale, const float *shift,
                          const float *mean, const float *rstd, int act, int64_t M, int C, float *sum_dz, float *sum_dzx, void *stream)
{
    EW_CHECK_C(C, "cy4_bn_act_bwd_reduce");
    CY4_CHECK_ARG(y && dA && scale && shift && mean && rstd && sum_dz && sum_dzx && M >= 0, "cy4_bn_act_bwd_reduce: bad argument");
    if (M == 0) return 0;
#define CY4_RED(ACT)                                                                                                        \
    do {                                                                                                                    \
        co_resident((const void *)bn_act_bwd_reduce_kernel<ACT>);                                                           \
        launch_pdl(bn_act_bwd_reduce_kernel<ACT>, col_grid(M, C, g_ew_bwd_bpsm), 256, (cudaStream_t)stream, (const __half *)y, ldy, (__half *)dA, ldg, scale, \
                   shift, mean, rstd, M, C, sum_dz, sum_dzx);                                                                   \
    } while (0)
    if (act == ACT_MISH) CY4_RED(ACT_MISH); else if (act == ACT_LEAKY) CY4_RED(ACT_LEAKY); else CY4_RED(ACT_LINEAR);
#undef CY4_RED
    return cy4_launch_status("cy4_bn_act_bwd_reduce");
}

int cy4_bn_bwd_fixup(const float *sum_dz, float *sum_dzy_inout, const float *mean, const float *rstd, int C, void *stream)
{
    CY4_CHECK_ARG(sum_dz && sum_dzy_inout && mean && rstd && C > 0, "cy4_bn_bwd_fixup: bad argument");
    bn_bwd_fixup_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum_dz, sum_dzy_inout, mean, rstd, C);
    return cy4_launch_status("cy4_bn_bwd_fixup");
}

int cy4_bn_act_bwd_apply(const void *y, int64_t ldy, const void *dA, int64_t ldg, const float *scale, const float *shift,
                         const float *mean, const float *rstd, const float *sum_dz, const float *sum_dzx, float inv_count,
                         int training, int act, int dz_ready, void *dY, int64_t ldd, int64_t M, int C, void *stream)
{
    EW_CHECK_C(C, "cy4_bn_act_bwd_apply");
    CY4_CHECK_ARG(y && dA && scale && shift && mean && rstd && sum_dz && sum_dzx && dY && M >= 0, "cy4_bn_act_bwd_apply: bad argument");
    if (M == 0) return 0;
#define CY4_APP(ACT, RDY)                                                                                                   \
    do {                                                                                                                    \
        co_resident((const void *)bn_act_bwd_apply_kernel<ACT, RDY>);                                                       \
        launch_pdl(bn_act_bwd_apply_kernel<ACT, RDY>, col_grid(M, C, g_ew_bwd_bpsm), 256, (cudaStream_t)stream, (const __half *)y, ldy, (const __half *)dA, ldg, \
                   scale, shift, mean, rstd, sum_dz, sum_dzx, inv_count, training, (__half *)dY, ldd, M, C);                    \
    } while (0)
    // dz_ready: cy4_bn_act_bwd_reduce ran on the same dA buffer before (it leaves dz = dA*act'(z) there)
    if (dz_ready || act == ACT_LINEAR) CY4_APP(ACT_LINEAR, true);
    else if (act == ACT_MISH) CY4_APP(ACT_MISH, false);
    else CY4_APP(ACT_LEAKY, false);
#undef CY4_APP
    return cy4_launch_status("cy4_bn_act_bwd_apply");
}

int cy4_add_copy(const void *a, int64_t lda, const void *b, int64_t ldb, void *out, int64_t ldo, int64_t M, int C, void *stream)
{
    EW_CHECK_C(C, "cy4_add_copy");
    CY4_CHECK_ARG(a && out && M >= 0 && (lda % 8) == 0 && (ldo % 8) == 0 && (ldb % 8) == 0, "cy4_add_copy: bad argument");
    if (M == 0) return 0;
    co_resident((const void *)add_copy_kernel);
    launch_pdl(add_copy_kernel, ew_grid(M * (C / 8)), 256, (cudaStream_t)stream, (const __half *)a, lda, (const __half *)b, ldb, (__half *)out, ldo, M, C);
    return cy4_launch_status("cy4_add_copy");
}

int cy4_upsample2x_fwd(const void *in, int64_t ldi, void *out, int64_t ldo, int B, int H, int W, int C, void *stream)
{
    EW_CHECK_C(C, "cy4_upsample2x_fwd");
    CY4_CHECK_ARG(in && out, "cy4_upsample2x_fwd: null pointer");
    upsample2x_fwd_kernel<<<ew_grid((int64_t)B * 4 * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const __half *)in, ldi, (__half *)out, ldo, B, H, W, C);
    return cy4_launch_status("cy4_upsample2x_fwd");
}

int cy4_upsample2x_bwd(const void *gout, int64_t ldo, void *gin, int64_t ldi, int B, int H, int W, int C, int accumulate, void *stream)
{
    EW_CHECK_C(C, "cy4_upsample2x_bwd");
    CY4_CHECK_ARG(gout && gin, "cy4_upsample2x_bwd: null pointer");
    upsample2x_bwd_kernel<<<ew_grid((int64_t)B * H * W * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const __half *)gout, ldo, (__half *)gin, ldi, B, H, W, C, accumulate);
    return cy4_launch_status("cy4_upsample2x_bwd");
}

int cy4_maxpool_fwd(const void *in, int64_t ldi, void *out, int64_t ldo, int B, int H, int W, int C, int k, int stride, int pad, void *stream)
{
    EW_CHECK_C(C, "cy4_maxpool_fwd");
    CY4_CHECK_ARG(in && out && k > 0 && stride > 0, "cy4_maxpool_fwd: bad argument");
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    maxpool_fwd_kernel<<<ew_grid((int64_t)B * Ho * Wo * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const __half *)in, ldi, (__half *)out, ldo, B, H, W, C, k, stride, pad, Ho, Wo);
    return cy4_launch_status("cy4_maxpool_fwd");
}

int cy4_maxpool_bwd(const void *in, int64_t ldi, const void *gout, int64_t ldo, float *gscratch, int B, int H, int W, int C, int k,
                    int stride, int pad, void *stream)
{
    EW_CHECK_C(C, "cy4_maxpool_bwd");
    CY4_CHECK_ARG(in && gout && gscratch && k > 0 && stride > 0, "cy4_maxpool_bwd: bad argument");
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    maxpool_bwd_kernel<<<ew_grid((int64_t)B * Ho * Wo * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const __half *)in, ldi, (const __half *)gout, ldo, gscratch,
                                                                                            B, H, W, C, k, stride, pad, Ho, Wo);
    return cy4_launch_status("cy4_maxpool_bwd");
}

int cy4_maxpool_fwd_idx(const void *in, int64_t ldi, void *out, int64_t ldo, void *argmax, void *workspace, int B, int H, int W, int C, int k,
                        int stride, int pad, void *stream)
{
    EW_CHECK_C(C, "cy4_maxpool_fwd_idx");
    CY4_CHECK_ARG(in && out && argmax && k > 0 && k * k <= 255 && stride > 0, "cy4_maxpool_fwd_idx: bad argument (window of at most 255 elements)");
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    const int64_t n = (int64_t)B * H * W * C;
    if (workspace && stride == 1 && Ho == H && Wo == W) {
        // separable: column maxima (+ first dy) into the workspace [n halves | n bytes], then the row pass
        __half *vmax = (__half *)workspace;
        uint8_t *vdy = (uint8_t *)workspace + 2 * n;
        maxpool_v_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const __half *)in, ldi, vmax, vdy, B, H, W, C, k, pad);
        maxpool_h_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>(vmax, vdy, (__half *)out, ldo, (uint8_t *)argmax, B, H, W, C, k, pad);
        return cy4_launch_status("cy4_maxpool_fwd_idx", 2);
    }
    maxpool_fwd_idx_kernel<<<ew_grid((int64_t)B * Ho * Wo * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const __half *)in, ldi, (__half *)out, ldo,
                                                                                                (uint8_t *)argmax, B, H, W, C, k, stride, pad, Ho, Wo);
    return cy4_launch_status("cy4_maxpool_fwd_idx");
}

int cy4_maxpool_bwd_idx(const void *argmax, const void *gout, int64_t ldo, float *gscratch, int B, int H, int W, int C, int k, int stride,
                        int pad, void *stream)
{
    EW_CHECK_C(C, "cy4_maxpool_bwd_idx");
    CY4_CHECK_ARG(argmax && gout && gscratch && k > 0 && k * k <= 255 && stride > 0, "cy4_maxpool_bwd_idx: bad argument");
    const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    maxpool_bwd_idx_kernel<<<ew_grid((int64_t)B * Ho * Wo * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)argmax, (const __half *)gout, ldo,
                                                                                                gscratch, B, H, W, C, k, stride, pad, Ho, Wo);
    return cy4_launch_status("cy4_maxpool_bwd_idx");
}

int cy4_f32_to_f16(const float *src, int64_t lds, float scale, const float *dscale, void *dst, int64_t ldd, int64_t M, int C,
                   int accumulate, void *stream)
{
    EW_CHECK_C(C, "cy4_f32_to_f16");
    CY4_CHECK_ARG(src && dst && M >= 0 && (lds % 4) == 0 && (ldd % 8) == 0, "cy4_f32_to_f16: bad argument");
    if (M == 0) return 0;
    f32_to_f16_accum_kernel<<<ew_grid(M * (C / 8)), 256, 0, (cudaStream_t)stream>>>(src, lds, scale, dscale, (__half *)dst, ldd, M, C, accumulate);
    return cy4_launch_status("cy4_f32_to_f16");
}

int cy4_absmax_f32(const float *src, int64_t n, float *amax, void *stream)
{
    CY4_CHECK_ARG(src && amax && n >= 0, "cy4_absmax_f32: bad argument");
    if (n == 0) return 0;
    absmax_f32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(src, n, amax);
    return cy4_launch_status("cy4_absmax_f32");
}

int cy4_make_scale(const float *amax, float target, float *scale2, void *stream)
{
    CY4_CHECK_ARG(amax && scale2 && target > 0.f, "cy4_make_scale: bad argument");
    make_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(amax, target, scale2);
    return cy4_launch_status("cy4_make_scale");
}

int cy4_colsum_f32(const float *src, int64_t lds, int64_t M, int C, float scale, float *out, int accumulate, void *stream)
{
    CY4_CHECK_ARG(src && out && M >= 0 && C > 0, "cy4_colsum_f32: bad argument");
    if (!accumulate) CY4_CUDA(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), (cudaStream_t)stream));
    if (M == 0) return 0;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((M + 63) / 64, (int64_t)sm_count() * 4));
    colsum_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, M, C, scale, out);
    return cy4_launch_status("cy4_colsum_f32");
}

}  // extern "C"
