// conv_pair.cu -- CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM convolution, used for every eligible launch
// (kchunk 64, N tile 128 / 256, >= 2 m tiles; cy4_set_option("conv_pair", 0) falls back to conv_tc.cu).  Measured on B200
// (profiles/r2_conv_shape_bench.md): never slower than the 1-CTA kernel, -27 % on the 128-channel 3x3 layers.
//
// conv_tc.cu's 128 x block_n tiles are bound by the L2 -> shared-memory ingest of their operand slabs
// (48 KB per 512 tensor cycles for N = 256).  Here two CTAs on the SMs of one TPC work
// on a 256 x block_n tile: each CTA loads the activation slab of ITS 128 pixels and HALF of the weight slab
// (block_n / 2 rows), the leader CTA issues one M = 256 `tcgen05.mma.cta_group::2` per K = 16 step, and each
// SM's tensor core reads the other half of the weights from its peer's shared memory.  Per CTA: 32 KB instead
// of 48 KB per k-block.  Accumulator rows [128 r, 128 r + 128) live in the TMEM of CTA r; each CTA runs its own
// epilogue (identical to conv_tc.cu).
//
// Protocol (after CUTLASS' 2-SM collectives):
//   * TMA loads of both CTAs signal the LEADER's full barrier (`.cta_group::2`, barrier address with the peer
//     bit cleared); the leader's producer posts the expected byte count of both CTAs;
//   * the leader's MMA warp waits on its full barrier, issues the MMAs and commits with
//     `tcgen05.commit.cta_group::2 ... .multicast::cluster` to the empty barrier (and, at the end of a tile, the
//     tmem_full barrier) of BOTH CTAs;
//   * the epilogue warps of both CTAs release an accumulator stage by arriving on the LEADER's tmem_empty
//     barrier (count 8);
//   * TMEM is allocated / freed with the cta_group::2 forms by warp 1 of both CTAs.
// Restrictions: kchunk = 64 (Cin % 64 == 0), block_n in {128, 256}, one k-block per slot.
#include <cuda_fp16.h>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_tc.cuh"
#include "conv_epi.cuh"
#include "sm100_pair.cuh"

namespace cy4 {
using namespace sm100;

constexpr int kPM = 128;                            // rows per CTA (256 per pair)
constexpr int kPThreads = 320;                      // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kPRegion = 4 * 49152;                 // operand slots (+ extra output slabs at its tail)
constexpr int kPMaxStages = 12;
constexpr int kPMaxStatCh = 1024;
constexpr int kPOutStage = 8 * 32 * 64;
constexpr int kPCtlOffset = kPRegion + kPOutStage;
constexpr int kPCtlBytes = 512;
constexpr int kPSmem = kPCtlOffset + 1024 + kPCtlBytes + 2 * kPMaxStatCh * 4;
struct PairCtl {
    uint64_t full[kPMaxStages], empty[kPMaxStages], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kPThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const ConvKParams p)
{
    pdl_trigger();                           // the next kernel of the stream may start its own set-up (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nst = p.stages;
    uint8_t *sA = smem;
    uint8_t *sB = smem + nst * p.a_stage;                       // this CTA's HALF of the weight slab per slot
    uint8_t *sOut = smem + kPRegion;
    PairCtl *ctl = (PairCtl *)(smem + kPCtlOffset);
    float *sstat = (float *)(smem + kPCtlOffset + kPCtlBytes);
    static_assert(sizeof(PairCtl) <= kPCtlBytes, "control block does not fit");
    const bool smem_stats = (p.flags & CONV_F_STATS) && p.tiles_n * p.block_n <= kPMaxStatCh;
    if (smem_stats)
        for (int i = threadIdx.x; i < 2 * kPMaxStatCh; i += kPThreads) sstat[i] = 0.f;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();                   // 0 = leader (issues the MMAs), 1 = peer
    const int unit0 = (int)cluster_id_x(), unit_step = (int)ncluster_x();
    const int cls_units = ((p.tiles_m + 1) / 2) * p.tiles_n;    // pairs of m tiles x n tiles (of one tap class)
    const int ncls = p.ncls > 1 ? p.ncls : 1;
    const int units = cls_units * ncls;
    const int half_n = p.block_n / 2;
    const uint32_t a_bytes = kPM * 64 * 2, b_half_bytes = (uint32_t)half_n * 64 * 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        if (p.flags & CONV_F_TMA_OUT) prefetch_tmap(&tmC);
        for (int s = 0; s < nst; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&ctl->tmem_full[s], 1); mbar_init(&ctl->tmem_empty[s], 8); }   // 4 warps x 2 CTAs
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();                      // both CTAs' barriers exist before anything remote targets them
    if (warp == 1) tmem_alloc_pair<512>(&ctl->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
    pdl_wait();                              // set-up done; from here on global memory written by the preceding kernels is touched

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = unit0; t < units; t += unit_step) {
                int cls, tt;
                unit_decode(t, cls_units, ncls, p.cls_interleave, unit_step, cls, tt);
                const int num_kb = (ncls > 1 ? p.cls_ntap[cls] : p.ntaps) * p.cin_chunks;
                const int n_blk = tt % p.tiles_n, m_blk = (tt / p.tiles_n) * 2 + crank;
                const int m0 = m_blk * kPM;
                const int img = m0 / (p.Po * p.Qo);
                const int rem = m0 - img * (p.Po * p.Qo);
                const int pi = rem / p.Qo, qi = rem - pi * p.Qo;
                const int bw = qi * p.tstride + p.lower_w, bh = pi * p.tstride + p.lower_h;
                int tap = ncls > 1 ? p.cls_tap0[cls] : 0, cc = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&ctl->empty[stage], phase ^ 1);                  // my own slot was consumed (multicast commit)
                    if (crank == 0) mbar_expect_tx(&ctl->full[stage], 2 * (a_bytes + b_half_bytes));   // bytes of BOTH CTAs
                    if (p.a_mode == 1)
                        tma_load_im2col_4d_pair(&tmA, &ctl->full[stage], sA + stage * p.a_stage, cc * 64, bw, bh, img,
                                                (uint16_t)p.tap_ow[tap], (uint16_t)p.tap_oh[tap]);
                    else
                        tma_load_2d_pair(&tmA, &ctl->full[stage], sA + stage * p.a_stage, cc * 64, m0);
                    tma_load_2d_pair(&tmB, &ctl->full[stage], sB + stage * p.b_stage, p.tap_kofs[tap] + cc * 64,
                                     n_blk * p.block_n + crank * half_n);
                    if (++cc == p.cin_chunks) { cc = 0; ++tap; }
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (crank == 0) {
            const uint32_t idesc = make_idesc_f16(2 * kPM, p.block_n, p.ab_fmt, 0, 0);       // M = 256 across the pair
            const uint32_t dhi = smem_desc_hi(1024, SW_128B);
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA), 16), b_lo0 = smem_desc_lo(smem_u32(sB), 16);
            const uint32_t a_step = (uint32_t)p.a_stage >> 4, b_step = (uint32_t)p.b_stage >> 4;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int t = unit0; t < units; t += unit_step) {
                mbar_wait(&ctl->tmem_empty[acc], acc_phase ^ 1);         // both CTAs' epilogues have drained this stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.block_n;
                int cls, tt_;
            unit_decode(t, cls_units, ncls, p.cls_interleave, unit_step, cls, tt_);
            const int num_kb = (ncls > 1 ? p.cls_ntap[cls] : p.ntaps) * p.cin_chunks;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&ctl->full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_lo = a_lo0 + (uint32_t)stage * a_step, b_lo = b_lo0 + (uint32_t)stage * b_step;
                        umma_f16_lohi_pair(d_tmem, a_lo, dhi, b_lo, dhi, idesc, kb != 0);
                        umma_f16_lohi_pair(d_tmem, a_lo + 2, dhi, b_lo + 2, dhi, idesc, 1);
                        umma_f16_lohi_pair(d_tmem, a_lo + 4, dhi, b_lo + 4, dhi, idesc, 1);
                        umma_f16_lohi_pair(d_tmem, a_lo + 6, dhi, b_lo + 6, dhi, idesc, 1);
                        umma_commit_pair(&ctl->empty[stage]);
                        if (kb == num_kb - 1) umma_commit_pair(&ctl->tmem_full[acc]);
                    }
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        const int quarter = warp & 3;
        const int group = (warp - 2) >> 2;
        const int acc = group; uint32_t acc_phase = 0;
        int seq = 0;
        const bool slab_st = (p.flags & (CONV_F_STATS | CONV_F_TMA_OUT | CONV_F_ACC_STATS)) == (CONV_F_STATS | CONV_F_TMA_OUT) && p.epi_mode != EPI_BWD_DZ;
        int slab_i = 0;
        for (int t = unit0; t < units; t += unit_step, ++seq) {
            if ((seq & 1) != group) continue;
            int cls, tt;
                unit_decode(t, cls_units, ncls, p.cls_interleave, unit_step, cls, tt);
            const int n_blk = tt % p.tiles_n, m_blk = (tt / p.tiles_n) * 2 + crank;
            const int m = m_blk * kPM + quarter * 32 + lane;
            const bool row_ok = m < p.M;
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
                float g[32];
                bool accum_in_store = (p.flags & CONV_F_ACCUM) != 0;
                epi_transform(p, f, g, accum_in_store, n0, orow, row_ok);
                if (p.flags & CONV_F_TMA_OUT) {
                    uint8_t *slab = (slab_i == 0 ? sOut : smem + kPRegion - slab_i * kPOutStage) + (warp - 2) * (32 * 64);
                    if (lane == 0) tma_store_wait_read_n(p.slab_bufs - 1);
                    __syncwarp();
                    if (++slab_i == p.slab_bufs) slab_i = 0;
                    const int xr = (lane >> 1) & 3;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 o; __half2 *ph = (__half2 *)&o;
#pragma unroll
                        for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(f[i + 2 * j], f[i + 2 * j + 1]);
                        *(uint4 *)(slab + lane * 64 + (((i >> 3) ^ xr) << 4)) = o;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.flags & CONV_F_ACCUM) tma_reduce_add_2d(&tmC, slab, n_blk * p.block_n + c * 32, m_blk * kPM + quarter * 32);
                        else tma_store_2d(&tmC, slab, n_blk * p.block_n + c * 32, m_blk * kPM + quarter * 32);
                        tma_store_commit();
                    }
                    if (slab_st) {          // BatchNorm statistics of the staged slab (conv_epi.cuh slab_stats)
                        float c0 = 0.f, c1 = 0.f;
                        if (p.stat_shift) { const float2 cv = __ldg((const float2 *)(p.stat_shift + n0) + (lane & 15)); c0 = cv.x; c1 = cv.y; }
                        float t1a, t1b, t2a, t2b;
                        slab_stats(slab, lane, p.M - (m_blk * kPM + quarter * 32), c0, c1, t1a, t1b, t2a, t2b);
                        const int col = n0 + 2 * (lane & 15), sq = lane >> 4;           // lanes 0..15 add the sums, 16..31 the squares
                        float *dst = smem_stats ? sstat + sq * kPMaxStatCh + col : (sq ? p.ch_sqsum : p.ch_sum) + col;
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
                } else if (row_ok) {
                    __half *dst = (__half *)p.y + orow * p.ldy + n0;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 o; __half2 *ph = (__half2 *)&o;
                        if (accum_in_store) {
                            const uint4 old = *(const uint4 *)(dst + i);
                            const __half2 *oh = (const __half2 *)&old;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 of = __half22float2(oh[j]);
                                ph[j] = __floats2half2_rn(f[i + 2 * j] + of.x, f[i + 2 * j + 1] + of.y);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(f[i + 2 * j], f[i + 2 * j + 1]);
                        }
                        *(uint4 *)(dst + i) = o;
                    }
                }
                if ((p.flags & CONV_F_STATS) && !slab_st) {      // (no staged slab, or EPI_BWD_DZ: reduce-scatter over the accumulators)
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
                    if (smem_stats) {
                        atomicAdd(sstat + n0 + lane, s1[0]);
                        atomicAdd(sstat + kPMaxStatCh + n0 + lane, s2[0]);
                    } else if (n0 + lane < p.N) {
                        atomicAdd(p.ch_sum + n0 + lane, s1[0]);
                        atomicAdd(p.ch_sqsum + n0 + lane, s2[0]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&ctl->tmem_empty[acc]);     // the leader's MMA warp owns the accumulator hand-over
            acc_phase ^= 1;
        }
    }
    if (warp >= 2 && lane == 0 && (p.flags & CONV_F_TMA_OUT)) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // neither CTA frees TMEM / leaves while its peer may still touch it
    if (warp == 1) { tc_fence_after(); tmem_dealloc_pair<512>(tmem_base); }
    if (smem_stats)
        for (int c = threadIdx.x; c < p.N; c += kPThreads) {
            atomicAdd(p.ch_sum + c, sstat[c]);
            atomicAdd(p.ch_sqsum + c, sstat[kPMaxStatCh + c]);
        }
}

// Can this launch use the pair kernel?  (kchunk 64, N tile of 128 or 256, at least one full pair of m tiles)
bool conv_pair_eligible(const ConvKParams &p)
{
    return p.kchunk == 64 && (p.block_n == 256 || p.block_n == 128) && p.tiles_m >= 2 && sm_count() >= 2;
}

// tmB must have been encoded with a box of block_n / 2 rows (run_generic does that when it selects this path).
int launch_conv_pair(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const ConvKParams &p_in, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        CY4_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
        attr_set = true;
    }
    ConvKParams p = p_in;
    p.kps = 1;
    p.slab_bufs = p.block_n <= 128 ? 2 : 1;
    p.a_stage = kPM * 64 * 2;                                    // 16 KB
    p.b_stage = (p.block_n / 2) * 64 * 2;                        // 16 KB (N = 256) or 8 KB (N = 128): this CTA's half
    p.stages = std::max(2, std::min(kPMaxStages, (kPRegion - (p.slab_bufs - 1) * kPOutStage) / (p.a_stage + p.b_stage)));
    const int units = ((p.tiles_m + 1) / 2) * p.tiles_n * (p.ncls > 1 ? p.ncls : 1);
    const int grid = std::min(units, sm_count() / 2) * 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPThreads);
    cfg.dynamicSmemBytes = kPSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_launch_attr(attr, 1);
    CY4_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel, tmA, tmB, tmC, p));
    return cy4_launch_status("conv_pair_kernel");
}

}  // namespace cy4
