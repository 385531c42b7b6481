// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   D[m, n] = sum_{tap, c} X[pixel(m) + tap, c] * W[n, tap, c]       m = output pixel, n = out channel
//
// One persistent CTA per SM, 6 warps, three roles:
//   warp 0      TMA producer   im2col-mode TMA gathers a 128-pixel x kchunk-channel slab of the NHWC
//                              activation per (tap, channel chunk) -- padding / stride / image borders
//                              are resolved by the TMA unit (OOB zero fill); a tiled TMA brings the
//                              matching [block_n x kchunk] slab of the K-major packed weights.
//   warp 1      MMA issuer     one lane issues tcgen05.mma (M=128, N=block_n, K=16) on the 128B/64B
//                              swizzled smem slabs; fp32 accumulators live in TMEM, double buffered
//                              (2 x block_n columns) so the epilogue of tile i overlaps tile i+1.
//   warps 2..9  epilogue       two groups of 4 warps, group g drains accumulator stage g (tiles alternate):
//                              tcgen05.ld the accumulator rows, fused per-channel sum / sum^2 for the
//                              training-mode BatchNorm that follows (warp reduce-scatter into per-CTA smem
//                              accumulators), bias / fp16 or fp32 conversion, swizzled smem slab + TMA store.
// The same kernel serves fprop, stride-1 dgrad (flipped/transposed weights) and the four parity
// classes of stride-2 dgrad (generic tap table + strided output-row mapping).
// Reference ops replaced: nn.Conv2d inside src/models/darknet2pytorch.py:247-278 (cuDNN via ATen).
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_tc.cuh"
#include "conv_epi.cuh"

namespace cy4 {
using namespace sm100;

constexpr int kBlockM = 128;
constexpr int kStages = 4;                         // stage REGION = 4 x 48 KB; split into p.stages (<= kMaxStages) slots
constexpr int kMaxStages = 12;
constexpr int kThreads = 320;                      // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kAStageBytes = kBlockM * 128;        // 16 KB (kchunk 64) ; 8 KB used when kchunk 32
constexpr int kBStageBytes = 256 * 128;            // 32 KB (block_n 256, kchunk 64)
constexpr int kMaxStatCh = 1024;                   // per-CTA shared accumulators for the BN statistics
constexpr int kOutStageBytes = 8 * 32 * 64;        // per epilogue warp: 32 rows x 64 B (32 fp16 columns) slab for the TMA store
constexpr int kCtlOffset = kStages * (kAStageBytes + kBStageBytes) + kOutStageBytes;
constexpr int kCtlBytes = 512;                     // barriers, TMEM base, slab table
constexpr int kSmemBytes = kCtlOffset + 1024 /*align*/ + kCtlBytes + 2 * kMaxStatCh * 4;
constexpr uint32_t kTmemCols = 512;

struct SmemCtl {
    uint64_t full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

// Cycle probes of the two single-thread loops (compile with -DCY4_PROBE; tools/probe_pipeline.py). Off in the product build.
#ifdef CY4_PROBE
__device__ unsigned long long g_probe[16];
#define PROBE_DECL unsigned long long pr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_ = 0
#define PROBE_T0 pt_ = clock64()
#define PROBE_ACC(i) do { long long n_ = clock64(); pr_[i] += (unsigned long long)(n_ - pt_); pt_ = n_; } while (0)
#define PROBE_CNT(i) pr_[i] += 1
#define PROBE_FLUSH(base) do { if (lane == 0 && blockIdx.x == 0) for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_probe[(base) + i_], pr_[i_]); } while (0)
#define PPROBE_DECL PROBE_DECL
#define PPROBE_T0 PROBE_T0
#define PPROBE_ACC(i) PROBE_ACC(i)
#define PPROBE_CNT(i) PROBE_CNT(i)
#define PPROBE_FLUSH(base) PROBE_FLUSH(base)
#else
#define PROBE_DECL
#define PROBE_T0
#define PROBE_ACC(i)
#define PROBE_CNT(i)
#define PROBE_FLUSH(base)
#define PPROBE_DECL
#define PPROBE_T0
#define PPROBE_ACC(i)
#define PPROBE_CNT(i)
#define PPROBE_FLUSH(base)
#endif

// the bottleneck-experiment switches (cy4_set_option "debug") only exist in -DCY4_PROBE builds
#ifdef CY4_PROBE
#define CY4_DBG (p.debug)
#else
#define CY4_DBG 0
#endif

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const ConvKParams p)
{
    pdl_trigger();                           // the next kernel of the stream may start its own set-up (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // The TMA -> MMA round trip is ~1.5-2 us; a slot only holds 12..48 KB, so narrow / small-K layers need
    // many more slots in flight than the 4 that fit for the 128x256x64 tile (measured).
    const int nst = p.stages, kps = p.kps;
    uint8_t *sA = smem;
    uint8_t *sB = smem + nst * kps * p.a_stage;
    uint8_t *sOut = smem + kStages * (kAStageBytes + kBStageBytes);
    SmemCtl *ctl = (SmemCtl *)(smem + kCtlOffset);
    float *sstat = (float *)(smem + kCtlOffset + kCtlBytes);     // [2][kMaxStatCh]
    static_assert(sizeof(SmemCtl) <= kCtlBytes, "control block does not fit");
    const bool smem_stats = (p.flags & CONV_F_STATS) && p.tiles_n * p.block_n <= kMaxStatCh;
    if (smem_stats)
        for (int i = threadIdx.x; i < 2 * kMaxStatCh; i += kThreads) sstat[i] = 0.f;

    // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps
    // the single-thread loops' operands in uniform registers (no R2UR broadcast loops around TMA / MMA issue)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // Work units: (group of `cluster` consecutive m tiles) x n tile.  The CTAs of a cluster take the m
    // tiles of one group, walk the same k loop and share every weight slab through TMA multicast.
    const int cs = p.cluster;
    const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
    const int unit0 = cs > 1 ? (int)cluster_id_x() : (int)blockIdx.x;
    const int unit_step = cs > 1 ? (int)ncluster_x() : (int)gridDim.x;
    const int cls_units = ((p.tiles_m + cs - 1) / cs) * p.tiles_n;          // units of one tap class
    const int ncls = p.ncls > 1 ? p.ncls : 1;
    const int units = cls_units * ncls;
    const uint16_t cmask = (uint16_t)((1u << cs) - 1);
    const uint32_t a_bytes = kBlockM * p.kchunk * 2, b_bytes = p.block_n * p.kchunk * 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        if (p.flags & CONV_F_TMA_OUT) prefetch_tmap(&tmC);
        for (int s = 0; s < nst; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], cs); }
        for (int s = 0; s < 2; ++s) { mbar_init(&ctl->tmem_full[s], 1); mbar_init(&ctl->tmem_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(&ctl->tmem_base);
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();          // peers' barriers are initialised before any multicast targets them
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
    pdl_wait();                              // set-up done; from here on global memory written by the preceding kernels is touched

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            PPROBE_DECL;
            int stage = 0; uint32_t phase = 0;
            const int b_rows = p.block_n / cs;                 // weight rows this CTA fetches (and multicasts)
            for (int t = unit0; t < units; t += unit_step) {
                int cls, tt;
                unit_decode(t, cls_units, ncls, p.cls_interleave, unit_step, cls, tt);
                const int num_kb = (ncls > 1 ? p.cls_ntap[cls] : p.ntaps) * p.cin_chunks;
                const int n_blk = tt % p.tiles_n, m_blk = (tt / p.tiles_n) * cs + crank;
                const int m0 = m_blk * kBlockM;
                // base pixel of the tile in the im2col "base pixel" space
                const int img = m0 / (p.Po * p.Qo);
                const int rem = m0 - img * (p.Po * p.Qo);
                const int pi = rem / p.Qo, qi = rem - pi * p.Qo;
                const int bw = qi * p.tstride + p.lower_w, bh = pi * p.tstride + p.lower_h;
                int tap = ncls > 1 ? p.cls_tap0[cls] : 0, cc = 0;
                for (int g0 = 0; g0 < num_kb; g0 += kps) {
                    const int cnt = min(kps, num_kb - g0);         // k-blocks of this slot
                    PPROBE_T0;
                    mbar_wait(&ctl->empty[stage], phase ^ 1);
                    PPROBE_ACC(0);
                    if (CY4_DBG == 2 || CY4_DBG == 3 || CY4_DBG == 6) { mbar_expect_tx(&ctl->full[stage], 0); if (++stage == nst) { stage = 0; phase ^= 1; } continue; }
                    mbar_expect_tx(&ctl->full[stage], (uint32_t)cnt * (a_bytes + b_bytes));
                    for (int j = 0; j < cnt; ++j) {
                        const int slot = stage * kps + j;
                        if (p.a_mode == 1)
                            tma_load_im2col_4d(&tmA, &ctl->full[stage], sA + slot * p.a_stage, cc * p.kchunk, bw, bh, img,
                                               (uint16_t)p.tap_ow[tap], (uint16_t)p.tap_oh[tap]);
                        else
                            tma_load_2d(&tmA, &ctl->full[stage], sA + slot * p.a_stage, cc * p.kchunk, m0);
                        if (cs > 1)
                            tma_load_2d_mc(&tmB, &ctl->full[stage], sB + slot * p.b_stage + crank * b_rows * p.kchunk * 2,
                                           p.tap_kofs[tap] + cc * p.kchunk, n_blk * p.block_n + crank * b_rows, cmask);
                        else
                            tma_load_2d(&tmB, &ctl->full[stage], sB + slot * p.b_stage, p.tap_kofs[tap] + cc * p.kchunk,
                                        n_blk * p.block_n);
                        if (++cc == p.cin_chunks) { cc = 0; ++tap; }
                    }
                    PPROBE_ACC(1);
                    PPROBE_CNT(2);
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
            }
            PPROBE_FLUSH(8);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_f16(kBlockM, p.block_n, p.ab_fmt, 0, 0);
        const uint32_t dhi = smem_desc_hi(p.kchunk == 64 ? 1024 : 512, p.kchunk == 64 ? SW_128B : SW_64B);
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA), 16), b_lo0 = smem_desc_lo(smem_u32(sB), 16);
        const uint32_t a_step = (uint32_t)p.a_stage >> 4, b_step = (uint32_t)p.b_stage >> 4;
        const bool k64 = p.kchunk == 64;
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        PROBE_DECL;
        for (int t = unit0; t < units; t += unit_step) {
            PROBE_T0;
            mbar_wait(&ctl->tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            PROBE_ACC(0);
            const uint32_t d_tmem = tmem_base + acc * p.block_n;
            int cls, tt_;
            unit_decode(t, cls_units, ncls, p.cls_interleave, unit_step, cls, tt_);
            const int num_kb = (ncls > 1 ? p.cls_ntap[cls] : p.ntaps) * p.cin_chunks;
            for (int g0 = 0; g0 < num_kb; g0 += kps) {
                const int cnt = min(kps, num_kb - g0);
                PROBE_T0;
                mbar_wait(&ctl->full[stage], phase);
                PROBE_ACC(1);
                tc_fence_after();
                PROBE_ACC(2);
                if (elect_one()) {
                    // descriptor low words advance by whole slots (addresses are 1024-aligned and < 256 KB, so the
                    // 14-bit address field never carries): one multiply-add per operand per k-block
                    uint32_t a_lo = a_lo0 + (uint32_t)(stage * kps) * a_step, b_lo = b_lo0 + (uint32_t)(stage * kps) * b_step;
                    for (int j = 0; j < cnt; ++j, a_lo += a_step, b_lo += b_step) {
                        if (CY4_DBG != 1 && CY4_DBG != 3 && CY4_DBG != 6) {
                            umma_f16_lohi(d_tmem, a_lo, dhi, b_lo, dhi, idesc, (g0 + j) != 0);
                            umma_f16_lohi(d_tmem, a_lo + 2, dhi, b_lo + 2, dhi, idesc, 1);       // +32 bytes of K per MMA
                            if (k64) {
                                umma_f16_lohi(d_tmem, a_lo + 4, dhi, b_lo + 4, dhi, idesc, 1);
                                umma_f16_lohi(d_tmem, a_lo + 6, dhi, b_lo + 6, dhi, idesc, 1);
                            }
                        }
                    }
                    PROBE_ACC(3);
                    if (CY4_DBG == 6) mbar_arrive(&ctl->empty[stage]);             // experiment: plain arrive instead of the commit
                    else if (cs > 1) umma_commit_mc(&ctl->empty[stage], cmask);    // the slot is free once EVERY CTA has consumed it
                    else umma_commit(&ctl->empty[stage]);
                    if (g0 + cnt == num_kb) umma_commit(&ctl->tmem_full[acc]);
                    PROBE_ACC(4);
                }
                PROBE_ACC(5);                  // (no __syncwarp: the next elect.sync is the warp's convergence point)
                PROBE_CNT(6);
                if (++stage == nst) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        PROBE_FLUSH(0);
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are this warp's
        const int group = (warp - 2) >> 2;            // group g owns accumulator stage g
        const int acc = group; uint32_t acc_phase = 0;
        int seq = 0;
        const bool slab_st = (p.flags & (CONV_F_STATS | CONV_F_TMA_OUT | CONV_F_ACC_STATS)) == (CONV_F_STATS | CONV_F_TMA_OUT) && p.epi_mode != EPI_BWD_DZ;
        int slab_i = 0;                               // rotating output slab of this warp
        for (int t = unit0; t < units; t += unit_step, ++seq) {
            if ((seq & 1) != group) continue;
            int cls, tt;
                unit_decode(t, cls_units, ncls, p.cls_interleave, unit_step, cls, tt);
            const int n_blk = tt % p.tiles_n, m_blk = (tt / p.tiles_n) * cs + crank;
            const int m = m_blk * kBlockM + quarter * 32 + lane;          // this thread's GEMM row
            const bool row_ok = m < p.M;
            // output row address (dense, or a strided parity class of a larger image)
            int64_t orow = m;
            if (p.omap) {
                const int img = m / (p.Po * p.Qo);
                const int rem = m - img * (p.Po * p.Qo);
                const int pi = rem / p.Qo, qi = rem - pi * p.Qo;
                const int oh0 = ncls > 1 ? p.cls_oh0[cls] : p.oh0, ow0 = ncls > 1 ? p.cls_ow0[cls] : p.ow0;
                orow = ((int64_t)img * p.OH + (pi * p.ostep + oh0)) * p.OW + (qi * p.ostep + ow0);
            }
            mbar_wait(&ctl->tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * p.block_n;
            for (int c = 0; c < p.block_n / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c * 32, v);
                tmem_ld_wait();
                const int n0 = n_blk * p.block_n + c * 32;
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                float g[32];                              // second statistics operand (EPI_BWD_DZ: the side value Y)
                bool accum_in_store = (p.flags & CONV_F_ACCUM) != 0;
                epi_transform(p, f, g, accum_in_store, n0, orow, row_ok);
                if (CY4_DBG == 4 || CY4_DBG == 5) {
                    if (CY4_DBG == 5 && f[0] == 12345.678f) ((float *)p.y)[0] = f[1];      // keep the TMEM load alive, store nothing
                } else if (p.flags & CONV_F_TMA_OUT) {
                    // fp16 slab [32 rows][cw columns] of this warp in swizzled smem, then one TMA store
                    // (coalesced, clipped to M rows by the tensor map).  cw = 64 (128B swizzle) or 32 (64B).
                    // (32 columns = 64-byte rows, 64B swizzle: 16-byte chunk index ^= (row >> 1) & 3)
                    // Narrow layers are bound by the TMA-store round trip of this slab, so they rotate through
                    // up to 4 slabs per warp (extra slabs live at the tail of the 192 KB stage region).
                    uint8_t *slab = (slab_i == 0 ? sOut : smem + kStages * (kAStageBytes + kBStageBytes) - slab_i * kOutStageBytes) + (warp - 2) * (32 * 64);
                    const int cw = 32;
                    const int sub = 0;
                    if (lane == 0) tma_store_wait_read_n(p.slab_bufs - 1);   // the store that used this slab has read it
                    __syncwarp();
                    if (++slab_i == p.slab_bufs) slab_i = 0;
                    const int rowbytes = cw * 2;
                    const int xr = (lane >> 1) & 3;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 o; __half2 *ph = (__half2 *)&o;
#pragma unroll
                        for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(f[i + 2 * j], f[i + 2 * j + 1]);
                        const int chunk = (sub + i) >> 3;            // 16-byte chunk index inside the row
                        *(uint4 *)(slab + lane * rowbytes + ((chunk ^ xr) << 4)) = o;
                    }
                    if (sub + 32 == cw) {                            // slab complete
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            if (p.flags & CONV_F_ACCUM) tma_reduce_add_2d(&tmC, slab, n_blk * p.block_n + c * 32 - sub, m_blk * kBlockM + quarter * 32);
                            else tma_store_2d(&tmC, slab, n_blk * p.block_n + c * 32 - sub, m_blk * kBlockM + quarter * 32);
                            tma_store_commit();
                        }
                    }
                    if (slab_st) {          // BatchNorm statistics of the staged slab (conv_epi.cuh slab_stats)
                        float c0 = 0.f, c1 = 0.f;
                        if (p.stat_shift) { const float2 cv = __ldg((const float2 *)(p.stat_shift + n0) + (lane & 15)); c0 = cv.x; c1 = cv.y; }
                        float t1a, t1b, t2a, t2b;
                        slab_stats(slab, lane, p.M - (m_blk * kBlockM + quarter * 32), c0, c1, t1a, t1b, t2a, t2b);
                        const int col = n0 + 2 * (lane & 15), sq = lane >> 4;           // lanes 0..15 add the sums, 16..31 the squares
                        float *dst = smem_stats ? sstat + sq * kMaxStatCh + col : (sq ? p.ch_sqsum : p.ch_sum) + col;
                        if (smem_stats || col < p.N) atomicAdd(dst, sq ? t2a : t1a);
                        if (smem_stats || col + 1 < p.N) atomicAdd(dst + 1, sq ? t2b : t1b);
                    }
                } else if (p.flags & CONV_F_OUT_F32) {
                    if (row_ok) {
                        float *dst = (float *)p.y + orow * p.ldy + n0;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 o;
                            o.x = f[i] + (p.bias ? __ldg(p.bias + n0 + i) : 0.f);
                            o.y = f[i + 1] + (p.bias ? __ldg(p.bias + n0 + i + 1) : 0.f);
                            o.z = f[i + 2] + (p.bias ? __ldg(p.bias + n0 + i + 2) : 0.f);
                            o.w = f[i + 3] + (p.bias ? __ldg(p.bias + n0 + i + 3) : 0.f);
                            *(float4 *)(dst + i) = o;
                        }
                    }
                } else {
                    if (row_ok) {
                        __half *dst = (__half *)p.y + orow * p.ldy + n0;
                        if (accum_in_store) {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                uint4 old = *(const uint4 *)(dst + i);
                                const __half2 *oh = (const __half2 *)&old;
                                uint4 o; __half2 *ph = (__half2 *)&o;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 of = __half22float2(oh[j]);
                                    ph[j] = __floats2half2_rn(f[i + 2 * j] + of.x, f[i + 2 * j + 1] + of.y);
                                }
                                *(uint4 *)(dst + i) = o;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                uint4 o; __half2 *ph = (__half2 *)&o;
#pragma unroll
                                for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(f[i + 2 * j], f[i + 2 * j + 1]);
                                *(uint4 *)(dst + i) = o;
                            }
                        }
                    }
                }
                if ((p.flags & CONV_F_STATS) && !slab_st) {      // (no staged slab, or EPI_BWD_DZ: reduce-scatter over the accumulators)
                    // Per-channel sum and sum of squares over this warp's 32 rows (rows >= M are exact
                    // zeros: their im2col pixels are out of bounds).  Reduce-scatter: lane L ends up
                    // with the totals of column L.
                    float s1[32], s2[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { s1[i] = f[i]; s2[i] = f[i] * (p.epi_mode == EPI_BWD_DZ ? g[i] : f[i]); }
                    if (p.stat_shift) {
                        // shifted sums: sum (y - c), sum (y - c)^2 with c ~ the channel mean (last step's): the batch variance
                        // s2/n - (s1/n)^2 then has no cancellation however large |mean| / sigma is.  Rows beyond M are exact
                        // zeros of the GEMM, not samples: they must not contribute -c.
                        const float4 *c4 = (const float4 *)(p.stat_shift + n0);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 cv = __ldg(c4 + (i >> 2));
                            const float d0 = row_ok ? f[i] - cv.x : 0.f, d1 = row_ok ? f[i + 1] - cv.y : 0.f;
                            const float d2 = row_ok ? f[i + 2] - cv.z : 0.f, d3 = row_ok ? f[i + 3] - cv.w : 0.f;
                            s1[i] = d0; s1[i + 1] = d1; s1[i + 2] = d2; s1[i + 3] = d3;
                            s2[i] = d0 * d0; s2[i + 1] = d1 * d1; s2[i + 2] = d2 * d2; s2[i + 3] = d3 * d3;
                        }
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool hi = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float send1 = hi ? s1[i] : s1[i + off];
                            const float send2 = hi ? s2[i] : s2[i + off];
                            const float r1 = __shfl_xor_sync(0xffffffffu, send1, off);
                            const float r2 = __shfl_xor_sync(0xffffffffu, send2, off);
                            s1[i] = (hi ? s1[i + off] : s1[i]) + r1;
                            s2[i] = (hi ? s2[i + off] : s2[i]) + r2;
                        }
                    }
                    if (smem_stats) {            // one global atomic per channel per CTA, at the end
                        atomicAdd(sstat + n0 + lane, s1[0]);
                        atomicAdd(sstat + kMaxStatCh + n0 + lane, s2[0]);
                    } else if (n0 + lane < p.N) {
                        atomicAdd(p.ch_sum + n0 + lane, s1[0]);
                        atomicAdd(p.ch_sqsum + n0 + lane, s2[0]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->tmem_empty[acc]);
            acc_phase ^= 1;
        }
    }
    if (warp >= 2 && lane == 0 && (p.flags & CONV_F_TMA_OUT)) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it
    if (warp == 1) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
    if (smem_stats)
        for (int c = threadIdx.x; c < p.N; c += kThreads) {
            atomicAdd(p.ch_sum + c, sstat[c]);
            atomicAdd(p.ch_sqsum + c, sstat[kMaxStatCh + c]);
        }
}

// --------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                     const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                     CUtensorMapFloatOOBfill);

static PFN_encodeTiled g_encodeTiled = nullptr;
static PFN_encodeIm2col g_encodeIm2col = nullptr;
static int g_driver_version = 0;

static int load_driver_entry_points()
{
    if (g_encodeTiled && g_encodeIm2col) return 0;
    cudaDriverEntryPointQueryResult qr;
    void *fn = nullptr;
    CY4_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available in this driver"); return -2; }
    g_encodeTiled = (PFN_encodeTiled)fn;
    fn = nullptr;
    CY4_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeIm2col not available in this driver"); return -2; }
    g_encodeIm2col = (PFN_encodeIm2col)fn;
    cudaDriverGetVersion(&g_driver_version);
    return 0;
}

int make_tmap_2d(CUtensorMap *tm, const void *base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes, int dtype_bf16)
{
    if (load_driver_entry_points()) return -2;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encodeTiled(tm, dtype_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                               const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu stride %llu box %u x %u", (int)r,
                                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner, box_outer); return -2; }
    return 0;
}

// NHWC activation viewed as (C, W, H, N); ld = channel stride in elements.
int make_tmap_im2col(CUtensorMap *tm, const void *base, int C, int W, int H, int N, int64_t ld, int lower_w, int lower_h,
                     int upper_w, int upper_h, int chan_per_pixel, int pixels_per_col, int tstride, int swizzle_bytes,
                     int dtype_bf16)
{
    if (load_driver_entry_points()) return -2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
    int lower[2] = {lower_w, lower_h};
    int upper[2] = {upper_w, upper_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)tstride, (cuuint32_t)tstride, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encodeIm2col(tm, dtype_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                                const_cast<void *>(base), dims, strides, lower, upper, (cuuint32_t)chan_per_pixel,
                                (cuuint32_t)pixels_per_col, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed (%d): C %d W %d H %d N %d ld %lld lower %d,%d upper %d,%d cpp %d ppc %d",
                                       (int)r, C, W, H, N, (long long)ld, lower_w, lower_h, upper_w, upper_h, chan_per_pixel, pixels_per_col); return -2; }
    // Driver workaround carried by CUTLASS (cute/atom/copy_traits_sm90_im2col.hpp): for drivers
    // <= 13.1 and tensors smaller than 128 KiB, bit 21 of the second descriptor word must be cleared.
    if (g_driver_version <= 13010 && (uint64_t)ld * 2 * W * H * N < 131072)
        reinterpret_cast<uint64_t *>(tm)[1] &= ~(1llu << 21);
    return 0;
}

int launch_conv_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const ConvKParams &p, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        CY4_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    const int cs = p.cluster;
    const int units = ((p.tiles_m + cs - 1) / cs) * p.tiles_n * (p.ncls > 1 ? p.ncls : 1);
    const int grid = std::min(units, sm_count() / cs) * cs;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_launch_attr(attr, cs > 1 ? 1 : 0);
    CY4_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, tmA, tmB, tmC, p));
    return cy4_launch_status("conv_tc_kernel");
}

}  // namespace cy4

#ifdef CY4_PROBE
extern "C" __attribute__((visibility("default"))) int cy4_probe_read(unsigned long long *out16)
{
    cudaDeviceSynchronize();
    unsigned long long z[16] = {0};
    if (cudaMemcpyFromSymbol(out16, cy4::g_probe, sizeof(z)) != cudaSuccess) return -1;
    return cudaMemcpyToSymbol(cy4::g_probe, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif
