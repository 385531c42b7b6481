// conv_wgrad.cu -- weight gradient of the convolution on tcgen05:
//
//   dW[co, tap, ci] = sum_{pixels m} dY[m, co] * X[pixel(m) + tap, ci]
//
// a GEMM whose reduction dimension is the pixel index, so BOTH operands are "MN-major" in shared
// memory (channels contiguous, pixels strided) -- the UMMA instruction descriptor's a_major/b_major
// bits select that, and TMA delivers the slabs in exactly the canonical MN-major 128B-swizzle
// layout: 128-pixel x 64-channel boxes, 8-pixel groups 1024 B apart (SBO), 64-channel column blocks
// one box (16 KB) apart (LBO).  X goes through the same im2col-mode tensor map as fprop.
//
// Work item = (128 output channels) x (<= 256 input channels) x (one tap) x (one slice of the
// pixel range); split-K slices add their fp32 partial tile into dw_acc with red.global.add.
// Reference op replaced: the autograd backward of nn.Conv2d (weight), src/train.py:212.
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"
#include "sm100_pair.cuh"
#include "conv_tc.cuh"

namespace cy4 {
using namespace sm100;
extern int g_wgrad_cluster;      // conv_api.cu (cy4_set_option)
extern int g_debug;              // 1: skip the MMAs, 2: skip the TMA loads (bottleneck experiments only)
extern int g_wgrad_wide32;       // conv_api.cu ("wgrad_wide32"): one N = 32*taps MMA per K step for 32-channel X instead of one per tap
extern int g_wgrad_pair;         // conv_api.cu (cy4_set_option "wgrad_pair"): CTA-pair kernel for the eligible launches

// The TMA unit sustains roughly one bulk-tensor instruction per ~350 cycles per SM regardless of the
// box size (measured: tools/bottleneck.py), so the boxes are made as large as the
// 128-byte swizzle span allows: 128 pixels x 64 channels (16 KB).  Two stages of 96 KB.
constexpr int kWStages = 2;
constexpr int kWThreads = 192;
constexpr int kPixBlk = 128;                        // pixels (GEMM K) per pipeline stage
constexpr int kWAStage = 2 * kPixBlk * 128;         // dY: two 64-channel boxes  = 32 KB
constexpr int kWBStage = 4 * kPixBlk * 128;         // X : up to four boxes      = 64 KB
constexpr int kWSmem = kWStages * (kWAStage + kWBStage) + 1024 + 256;

struct WgradParams {
    int Mpix;                    // total pixels (B*Ho*Wo)
    int Cout, Cin;               // real sizes
    int m_tiles, n_tiles, block_n, ntaps, ksplit, kblocks;   // kblocks = ceil(Mpix/64)
    int tpc, tap_groups;         // taps handled by one CTA (accumulators tpc * block_n TMEM columns <= 256)
    int cluster;                 // CTAs per cluster (consecutive m tiles) sharing every X slab through TMA multicast
    int pair;                    // 1: conv_wgrad_pair_kernel (cta_group::2)
    int wide32;                  // 1: 32-channel X taps as one wide-N B operand
    int debug;
    int b_boxes;                 // block_n / 64 (or 1 when the 64B-swizzle N=32 path is used)
    int b_sw64;                  // 1: X has 32 channels, single [64 px x 32 ch] box, 64B swizzle
    int a_matrix;                // 1: X is a plain matrix (tiled TMA), only with ntaps == 1
    int Po, Qo, tstride, lower_w, lower_h;
    uint8_t tap_ow[kMaxTaps], tap_oh[kMaxTaps];
    float *dw; int64_t dw_row;   // dw_acc[co * dw_row + tap * cin_pad + ci]
    int cin_pad;
};

struct WCtl {
    uint64_t full[kWStages], empty[kWStages], tmem_full;
    uint32_t tmem_base;
};

#ifdef CY4_PROBE
#define CY4_WDBG (p.debug)
#else
#define CY4_WDBG 0
#endif

__global__ void __launch_bounds__(kWThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgradParams p)
{
    pdl_trigger();                           // the next kernel of the stream may start its own set-up (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + kWStages * kWAStage;
    WCtl *ctl = (WCtl *)(smem + kWStages * (kWAStage + kWBStage));
    // (shuffled warp index / TMEM base: keeps the issue loops' operands in uniform registers, see conv_tc.cu)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    // decode the work item; the CTAs of a cluster differ only in the m tile
    const int cs = p.cluster;
    const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << cs) - 1);
    int item = blockIdx.x / cs;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int tg = item % p.tap_groups; item /= p.tap_groups;
    const int tap0 = tg * p.tpc, ntap = min(p.tpc, p.ntaps - tap0);
    const int n_blk = item % p.n_tiles;
    const int m_blk = (item / p.n_tiles) * cs + crank;
    const int kb_per = (p.kblocks + p.ksplit - 1) / p.ksplit;
    const int kb0 = ks * kb_per, kb1 = min(p.kblocks, kb0 + kb_per);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDy); prefetch_tmap(&tmX);
        for (int s = 0; s < kWStages; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], cs); }
        mbar_init(&ctl->tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(&ctl->tmem_base);
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
    pdl_wait();                              // set-up done; from here on global memory written by the preceding kernels is touched

    if (nkb > 0) {
        if (warp == 0) {
            if (elect_one()) {
                const int a_boxes = (m_blk * 128 + 64 < p.Cout) ? 2 : 1;       // skip a fully out-of-range box
                const uint32_t b_tap_bytes = p.b_sw64 ? kPixBlk * 64 : p.b_boxes * kPixBlk * 128;
                const uint32_t bytes = a_boxes * kPixBlk * 128 + ntap * b_tap_bytes;
                int stage = 0; uint32_t phase = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int m0 = kb * kPixBlk;
                    mbar_wait(&ctl->empty[stage], phase ^ 1);
                    if (CY4_WDBG == 2) { mbar_expect_tx(&ctl->full[stage], 0); if (++stage == kWStages) { stage = 0; phase ^= 1; } continue; }
                    mbar_expect_tx(&ctl->full[stage], bytes);
                    for (int bx = 0; bx < a_boxes; ++bx)
                        tma_load_2d(&tmDy, &ctl->full[stage], sA + stage * kWAStage + bx * (kPixBlk * 128), m_blk * 128 + bx * 64, m0);
                    if (p.a_matrix) {
                        for (int bx = 0; bx < p.b_boxes; ++bx)
                            tma_load_2d(&tmX, &ctl->full[stage], sB + stage * kWBStage + bx * (kPixBlk * 128),
                                        n_blk * p.block_n + bx * 64, m0);
                    } else {
                        const int img = m0 / (p.Po * p.Qo);
                        const int rem = m0 - img * (p.Po * p.Qo);
                        const int pi = rem / p.Qo, qi = rem - pi * p.Qo;
                        const int bw = qi * p.tstride + p.lower_w, bh = pi * p.tstride + p.lower_h;
                        for (int t = 0; t < ntap; ++t)
                            for (int bx = 0; bx < p.b_boxes; ++bx) {
                                uint8_t *dst = sB + stage * kWBStage + t * b_tap_bytes + bx * (kPixBlk * 128);
                                if (cs == 1)
                                    tma_load_im2col_4d(&tmX, &ctl->full[stage], dst, n_blk * p.block_n + bx * 64, bw, bh, img,
                                                       (uint16_t)p.tap_ow[tap0 + t], (uint16_t)p.tap_oh[tap0 + t]);
                                else if ((t * p.b_boxes + bx) % cs == crank)     // this CTA's share, multicast to the cluster
                                    tma_load_im2col_4d_mc(&tmX, &ctl->full[stage], dst, n_blk * p.block_n + bx * 64, bw, bh, img,
                                                          (uint16_t)p.tap_ow[tap0 + t], (uint16_t)p.tap_oh[tap0 + t], cmask);
                            }
                    }
                    if (++stage == kWStages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            const uint32_t idesc = make_idesc_f16(128, p.block_n, 0, 1, 1);       // both operands MN-major
            const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
            const uint32_t a_hi = smem_desc_hi(1024, SW_128B);
            const uint32_t b_hi = p.b_sw64 ? smem_desc_hi(512, SW_64B) : smem_desc_hi(1024, SW_128B);
            const uint32_t b_kstep = p.b_sw64 ? (16 * 64 / 16) : (16 * 128 / 16);          // 16 pixel rows, in 16-byte units
            const uint32_t b_tap_bytes = p.b_sw64 ? kPixBlk * 64 : p.b_boxes * kPixBlk * 128;
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&ctl->full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    // descriptors as (lo, hi) halves: hi is a loop constant, lo advances by 16 pixel rows per MMA
                    const uint32_t a_lo0 = smem_desc_lo(a_base + stage * kWAStage, kPixBlk * 128);
                    if (p.b_sw64 && ntap > 1 && p.wide32) {
                        // 32-channel X (64B swizzle): the taps' [128 px x 32 ch] slabs are consecutive 32-column blocks of ONE
                        // MN-major B operand (LBO = slab size), so a single N = 32 * ntap MMA per K step replaces ntap MMAs of
                        // N = 32 -- those are bound by the MMA dispatch rate (an N = 32 instruction costs as much issue time
                        // as an N = 64 one), not by bytes.  The accumulator columns are the same: tap t at column 32 t.
                        const uint32_t idesc_w = make_idesc_f16(128, 32 * ntap, 0, 1, 1);
                        const uint32_t b_lo0 = smem_desc_lo(b_base + stage * kWBStage, b_tap_bytes);
                        if (CY4_WDBG != 1) {
#pragma unroll
                            for (int k = 0; k < kPixBlk / 16; ++k)
                                umma_f16_lohi(tmem_base, a_lo0 + k * (16 * 128 / 16), a_hi, b_lo0 + k * b_kstep, b_hi, idesc_w, (kb | k) != 0);
                        }
                    } else
                    for (int t = 0; t < ntap; ++t) {
                        const uint32_t b_addr = b_base + stage * kWBStage + t * b_tap_bytes;
                        const uint32_t b_lo0 = p.b_sw64 ? smem_desc_lo(b_addr, 0) : smem_desc_lo(b_addr, kPixBlk * 128);
                        const uint32_t d_t = tmem_base + t * p.block_n;
                        if (CY4_WDBG != 1) {
#pragma unroll
                            for (int k = 0; k < kPixBlk / 16; ++k)
                                umma_f16_lohi(d_t, a_lo0 + k * (16 * 128 / 16), a_hi, b_lo0 + k * b_kstep, b_hi, idesc, (kb | k) != 0);
                        }
                    }
                    if (cs > 1) umma_commit_mc(&ctl->empty[stage], cmask);
                    else umma_commit(&ctl->empty[stage]);
                    if (kb == nkb - 1) umma_commit(&ctl->tmem_full);
                }
                __syncwarp();
                if (++stage == kWStages) { stage = 0; phase ^= 1; }
            }
        } else {
            const int quarter = warp & 3;
            const int co = m_blk * 128 + quarter * 32 + lane;
            mbar_wait(&ctl->tmem_full, 0);
            tc_fence_after();
            for (int t = 0; t < ntap; ++t) {
                float *row = p.dw + (int64_t)co * p.dw_row + (int64_t)(tap0 + t) * p.cin_pad;
                for (int c = 0; c < p.block_n / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + t * p.block_n + c * 32, v);
                    tmem_ld_wait();
                    const int ci0 = n_blk * p.block_n + c * 32;
                    if (co < p.Cout) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)       // Cin is a multiple of 4: 16-byte vector reductions
                            if (ci0 + i < p.Cin)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + ci0 + i), "f"(__uint_as_float(v[i])),
                                             "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                                             : "memory");
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(tmem_base); }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair (tcgen05 cta_group::2) form for layers with Cout % 256 == 0 and a 128- or 256-channel X tile.
// Like fprop, the 1-CTA kernel is bound by the L2 -> shared-memory ingest of its operand slabs (96 KB per 1024 tensor cycles
// for the widest tile: 94 B/clk against ~60 B/clk that an SM sustains when all of them pull).  Two CTAs on consecutive
// 128-channel m tiles work on one (256 x N) tile: each loads the dY slab of ITS 128 output channels and HALF of the X slab
// (N / 2 input channels of every tap); the leader issues M = 256 MMAs whose B operand is read from both CTAs' shared
// memory.  64 KB per CTA per 128-pixel k-block instead of 96 KB, and three pipeline stages instead of two.
constexpr int kWPStages = 3;
constexpr int kWPAStage = 2 * kPixBlk * 128;        // dY: two 64-channel boxes  = 32 KB
constexpr int kWPBStage = 2 * kPixBlk * 128;        // X : this CTA's half       = 32 KB
constexpr int kWPSmem = kWPStages * (kWPAStage + kWPBStage) + 1024 + 256;

struct WPCtl {
    uint64_t full[kWPStages], empty[kWPStages], tmem_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWThreads, 1)
conv_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WgradParams p)
{
    pdl_trigger();                           // the next kernel of the stream may start its own set-up (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + kWPStages * kWPAStage;
    WPCtl *ctl = (WPCtl *)(smem + kWPStages * (kWPAStage + kWPBStage));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();                   // 0 = leader (issues the MMAs)
    int item = blockIdx.x >> 1;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int tg = item % p.tap_groups; item /= p.tap_groups;
    const int tap0 = tg * p.tpc, ntap = min(p.tpc, p.ntaps - tap0);
    const int n_blk = item % p.n_tiles;
    const int m_blk = (item / p.n_tiles) * 2 + crank;
    const int kb_per = (p.kblocks + p.ksplit - 1) / p.ksplit;
    const int kb0 = ks * kb_per, kb1 = min(p.kblocks, kb0 + kb_per);
    const int nkb = kb1 - kb0;                                  // identical in both CTAs of the pair
    const int hb = p.b_boxes >> 1;                              // X boxes (64 channels each) per tap that THIS CTA loads
    const uint32_t b_tap_half = (uint32_t)hb * kPixBlk * 128;   // bytes of one tap's half slab

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDy); prefetch_tmap(&tmX);
        for (int s = 0; s < kWPStages; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
        mbar_init(&ctl->tmem_full, 1);
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();                      // both CTAs' barriers exist before anything remote targets them
    if (warp == 1) tmem_alloc_pair<256>(&ctl->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
    pdl_wait();                              // set-up done; from here on global memory written by the preceding kernels is touched

    if (nkb > 0) {
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t my_bytes = 2 * kPixBlk * 128 + (uint32_t)ntap * b_tap_half;
                int stage = 0; uint32_t phase = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int m0 = kb * kPixBlk;
                    mbar_wait(&ctl->empty[stage], phase ^ 1);                  // my own slot was consumed (multicast commit)
                    if (crank == 0) mbar_expect_tx(&ctl->full[stage], 2 * my_bytes);         // bytes of BOTH CTAs
                    for (int bx = 0; bx < 2; ++bx)
                        tma_load_2d_pair(&tmDy, &ctl->full[stage], sA + stage * kWPAStage + bx * (kPixBlk * 128), m_blk * 128 + bx * 64, m0);
                    const int img = m0 / (p.Po * p.Qo);
                    const int rem = m0 - img * (p.Po * p.Qo);
                    const int pi = rem / p.Qo, qi = rem - pi * p.Qo;
                    const int bw = qi * p.tstride + p.lower_w, bh = pi * p.tstride + p.lower_h;
                    for (int t = 0; t < ntap; ++t)
                        for (int j = 0; j < hb; ++j)
                            tma_load_im2col_4d_pair(&tmX, &ctl->full[stage], sB + stage * kWPBStage + t * b_tap_half + j * (kPixBlk * 128),
                                                    n_blk * p.block_n + (crank * hb + j) * 64, bw, bh, img,
                                                    (uint16_t)p.tap_ow[tap0 + t], (uint16_t)p.tap_oh[tap0 + t]);
                    if (++stage == kWPStages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (crank == 0) {
                const uint32_t idesc = make_idesc_f16(256, p.block_n, 0, 1, 1);       // M = 256 across the pair, both operands MN-major
                const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
                const uint32_t hi = smem_desc_hi(1024, SW_128B);
                int stage = 0; uint32_t phase = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&ctl->full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_lo0 = smem_desc_lo(a_base + stage * kWPAStage, kPixBlk * 128);
                        for (int t = 0; t < ntap; ++t) {
                            const uint32_t b_lo0 = smem_desc_lo(b_base + stage * kWPBStage + t * b_tap_half, kPixBlk * 128);
                            const uint32_t d_t = tmem_base + t * p.block_n;
#pragma unroll
                            for (int k = 0; k < kPixBlk / 16; ++k)
                                umma_f16_lohi_pair(d_t, a_lo0 + k * (16 * 128 / 16), hi, b_lo0 + k * (16 * 128 / 16), hi, idesc, (kb | k) != 0);
                        }
                        umma_commit_pair(&ctl->empty[stage]);
                        if (kb == nkb - 1) umma_commit_pair(&ctl->tmem_full);
                    }
                    __syncwarp();
                    if (++stage == kWPStages) { stage = 0; phase ^= 1; }
                }
            }
        } else {
            const int quarter = warp & 3;
            const int co = m_blk * 128 + quarter * 32 + lane;
            mbar_wait(&ctl->tmem_full, 0);
            tc_fence_after();
            for (int t = 0; t < ntap; ++t) {
                float *row = p.dw + (int64_t)co * p.dw_row + (int64_t)(tap0 + t) * p.cin_pad;
                for (int c = 0; c < p.block_n / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + t * p.block_n + c * 32, v);
                    tmem_ld_wait();
                    const int ci0 = n_blk * p.block_n + c * 32;
                    if (co < p.Cout) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            if (ci0 + i < p.Cin)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + ci0 + i), "f"(__uint_as_float(v[i])),
                                             "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                                             : "memory");
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // neither CTA frees TMEM / leaves while its peer may still touch it
    if (warp == 1) { tc_fence_after(); tmem_dealloc_pair<256>(tmem_base); }
}

}  // namespace cy4

using namespace cy4;

// Tile geometry and split-K factor of one weight-gradient launch (shared by cy4_conv_wgrad and cy4_conv_wgrad_plan).
static void wgrad_tiling(const cy4_conv_desc *d, WgradParams &p)
{
    const int k = d->ksize;
    const bool sw64 = d->Cin == 32;
    const int cin64 = sw64 ? 32 : (d->Cin + 63) / 64 * 64;
    p.Mpix = d->B * d->Ho * d->Wo;
    p.Cout = d->Cout; p.Cin = d->Cin;
    p.m_tiles = (d->Cout + 127) / 128;
    p.block_n = sw64 ? 32 : std::min(cin64, 256);
    p.n_tiles = (cin64 + p.block_n - 1) / p.block_n;
    p.b_boxes = sw64 ? 1 : p.block_n / 64;
    p.b_sw64 = sw64 ? 1 : 0;
    p.ntaps = k * k;
    p.kblocks = (p.Mpix + kPixBlk - 1) / kPixBlk;
    // several taps per CTA while their accumulators (tpc * block_n columns) and X slabs (<= 32 KB per
    // stage) fit: dY is then loaded once per tap group instead of once per tap
    p.tpc = std::max(1, std::min(p.ntaps, 256 / p.block_n));
    if (p.ntaps == 9) p.tpc = p.tpc >= 5 ? 5 : (p.tpc >= 3 ? 3 : p.tpc);      // balanced groups: 5+4, 3+3+3, 2+2+2+2+1
    p.tap_groups = (p.ntaps + p.tpc - 1) / p.tpc;
    // pairs of CTAs on consecutive m tiles share (multicast) the X slabs; not for the matrix (stem) mode
    p.debug = g_debug;
    p.wide32 = g_wgrad_wide32;
    p.cluster = (g_wgrad_cluster >= 2 && !(d->flags & CY4_CONV_A_MATRIX) && p.m_tiles % 2 == 0) ? 2 : 1;
    // CTA pairs (conv_wgrad_pair_kernel): both m tiles of a pair full, an even number of 64-channel X boxes per tap
    p.pair = (g_wgrad_pair && !(d->flags & CY4_CONV_A_MATRIX) && !sw64 && d->Cout % 256 == 0 && p.block_n >= 128 && p.cluster == 1) ? 1 : 0;
    const int items = (p.pair ? p.m_tiles / 2 : p.m_tiles) * p.n_tiles * p.tap_groups;       // work items (CTAs, or CTA pairs)
    // Split-K factor.  One CTA per SM fits (193 KB of smem), so the kernel runs in waves of sm_count() CTAs and a grid of
    // 300 CTAs costs three CTA durations, not 2.03: pick the split (up to ~2 waves of CTAs) that minimises
    // waves x (k-blocks per CTA + the fixed prologue / TMEM drain / red.add epilogue, ~6 k-blocks' worth of time).
    {
        const int sms = p.pair ? sm_count() / 2 : sm_count();      // concurrent work items per wave
        const int ks_max = std::max(1, std::min(p.kblocks, (2 * sms + items - 1) / items));
        double best = 1e30;
        p.ksplit = 1;
        for (int ks = 1; ks <= ks_max; ++ks) {
            const int waves = (items * ks + sms - 1) / sms;
            const int kb_per = (p.kblocks + ks - 1) / ks;
            const double cost = (double)waves * (kb_per + 6.0);
            if (cost < best - 1e-9) { best = cost; p.ksplit = ks; }
        }
    }
}

extern "C" int cy4_conv_wgrad(const cy4_conv_desc *d, const void *x, const void *dy, float *dw_acc, void *stream)
{
    CY4_CHECK_ARG(d && x && dy && dw_acc, "cy4_conv_wgrad: null pointer");
    CY4_CHECK_ARG(d->ksize >= 1 && d->ksize <= 3 && d->stride >= 1 && d->stride <= 2, "cy4_conv_wgrad: bad geometry");
    const int k = d->ksize;
    const int cout64 = (d->Cout + 63) / 64 * 64;
    CY4_CHECK_ARG(d->ldy >= cout64 && d->ldy % 8 == 0, "cy4_conv_wgrad: dy must be allocated with ld >= Cout rounded up to 64");
    const bool sw64 = d->Cin == 32;     // one [64 px x 32 ch] box, 64B swizzle: never reads past the 32 channels
    const int cin64 = sw64 ? 32 : (d->Cin + 63) / 64 * 64;
    CY4_CHECK_ARG(d->Cin % 32 == 0 && d->ldx >= (sw64 ? 32 : cin64) && d->ldx % 8 == 0, "cy4_conv_wgrad: x must be allocated with ld >= Cin rounded up to 64 (or Cin == 32)");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    wgrad_tiling(d, p);
    const int items = (p.pair ? p.m_tiles / 2 : p.m_tiles) * p.n_tiles * p.tap_groups;
    p.a_matrix = (d->flags & CY4_CONV_A_MATRIX) ? 1 : 0;
    if (p.a_matrix) CY4_CHECK_ARG(k == 1 && d->stride == 1 && d->pad == 0, "cy4_conv_wgrad: matrix mode needs a 1x1/s1/p0 conv");
    p.Po = d->Ho; p.Qo = d->Wo; p.tstride = d->stride; p.lower_w = p.lower_h = -d->pad;
    for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) { p.tap_ow[r * k + s] = (uint8_t)s; p.tap_oh[r * k + s] = (uint8_t)r; }
    p.cin_pad = d->Cin;
    p.dw = dw_acc; p.dw_row = (int64_t)k * k * d->Cin;
    if (p.Mpix <= 0) return 0;

    alignas(64) CUtensorMap tmDy, tmX;
    int rc = make_tmap_2d(&tmDy, dy, (uint64_t)cout64, (uint64_t)p.Mpix, (uint64_t)d->ldy * 2, 64, kPixBlk, 128, 0);
    if (rc) return rc;
    if (p.a_matrix)
        rc = make_tmap_2d(&tmX, x, (uint64_t)cin64, (uint64_t)p.Mpix, (uint64_t)d->ldx * 2, sw64 ? 32 : 64, kPixBlk, sw64 ? 64 : 128, 0);
    else
        rc = make_tmap_im2col(&tmX, x, cin64, d->Wi, d->Hi, d->B, d->ldx, -d->pad, -d->pad, d->pad - (k - 1), d->pad - (k - 1),
                              sw64 ? 32 : 64, kPixBlk, d->stride, sw64 ? 64 : 128, 0);
    if (rc) return rc;
    if (d->flags & CY4_CONV_ZERO_ACC)     // zero right before use: the red.global.add then hit lines that are hot in L2
        CY4_CUDA(cudaMemsetAsync(dw_acc, 0, (size_t)((d->Cout + 31) / 32 * 32) * k * k * d->Cin * sizeof(float), (cudaStream_t)stream));
    static bool attr_set = false;
    if (!attr_set) {
        CY4_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem));
        attr_set = true;
    }
    if (p.pair) {
        static bool pattr_set = false;
        if (!pattr_set) {
            CY4_CUDA(cudaFuncSetAttribute(conv_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWPSmem));
            pattr_set = true;
        }
        cudaLaunchConfig_t pc;
        memset(&pc, 0, sizeof(pc));
        pc.gridDim = dim3((p.m_tiles / 2) * p.n_tiles * p.tap_groups * p.ksplit * 2);
        pc.blockDim = dim3(kWThreads);
        pc.dynamicSmemBytes = kWPSmem;
        pc.stream = (cudaStream_t)stream;
        cudaLaunchAttribute pa[2];
        pa[0].id = cudaLaunchAttributeClusterDimension;
        pa[0].val.clusterDim.x = 2; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
        pc.attrs = pa; pc.numAttrs = pdl_launch_attr(pa, 1);
        CY4_CUDA(cudaLaunchKernelEx(&pc, conv_wgrad_pair_kernel, tmDy, tmX, p));
        return cy4_launch_status("cy4_conv_wgrad (pair)");
    }
    const int grid = items * p.ksplit;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kWThreads);
    cfg.dynamicSmemBytes = kWSmem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_launch_attr(attr, p.cluster > 1 ? 1 : 0);
    CY4_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel, tmDy, tmX, p));
    return cy4_launch_status("cy4_conv_wgrad");
}

extern "C" int cy4_conv_wgrad_plan(const cy4_conv_desc *d, int32_t *out8)
{
    CY4_CHECK_ARG(d && out8, "cy4_conv_wgrad_plan: null pointer");
    CY4_CHECK_ARG(d->ksize >= 1 && d->ksize <= 3 && d->Cin % 32 == 0 && d->Cout > 0 && d->B > 0, "cy4_conv_wgrad_plan: bad geometry");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    wgrad_tiling(d, p);
    out8[0] = p.m_tiles; out8[1] = p.n_tiles; out8[2] = p.block_n; out8[3] = p.tpc; out8[4] = p.tap_groups;
    out8[5] = p.kblocks; out8[6] = p.ksplit; out8[7] = p.m_tiles * p.n_tiles * p.tap_groups * p.ksplit;      // CTAs (a pair counts as two)
    return 0;
}
