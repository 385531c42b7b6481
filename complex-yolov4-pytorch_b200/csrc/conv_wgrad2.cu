// conv_wgrad2.cu -- EXPERIMENTAL persistent variant of the weight-gradient kernel (cy4_set_option("wgrad_variant", 2);
// default off, not yet measured on hardware -- DESIGN.md section 4.3, lever 2).
//
// Same operand layout and MMA sequence as conv_wgrad.cu (MN-major dY / im2col(X) slabs of 128 pixels, split-K over
// pixels, red.global.add of the fp32 partial tiles).  What changes is the control structure, borrowed from the fprop
// kernel (conv_tc.cu):
//   * one CTA per SM walks a list of work items (m tile, n tile, tap group, k slice) instead of one CTA per item: the
//     prologue (barrier init, TMEM allocation, tensor-map prefetch) is paid once, items can be small (good balance, no
//     wave quantisation) and
//   * two accumulator stages of 256 TMEM columns: the TMEM drain + red.add epilogue of item i overlaps the TMA / MMA
//     main loop of item i+1; the smem ring keeps running across item boundaries.
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_tc.cuh"

namespace cy4 {
using namespace sm100;

constexpr int kW2Stages = 2;
constexpr int kW2Threads = 192;                      // TMA warp, MMA warp, 4 epilogue warps
constexpr int kW2Pix = 128;                          // pixels (GEMM K) per pipeline stage
constexpr int kW2AStage = 2 * kW2Pix * 128;          // dY: two 64-channel boxes  = 32 KB
constexpr int kW2BStage = 4 * kW2Pix * 128;          // X : up to four boxes      = 64 KB
constexpr int kW2Smem = kW2Stages * (kW2AStage + kW2BStage) + 1024 + 256;

struct Wgrad2Params {
    int Mpix, Cout, Cin;
    int m_tiles, n_tiles, block_n, ntaps, ksplit, kblocks, tpc, tap_groups, n_work;
    int b_boxes, b_sw64, a_matrix;
    int Po, Qo, tstride, lower_w, lower_h;
    uint8_t tap_ow[kMaxTaps], tap_oh[kMaxTaps];
    float *dw; int64_t dw_row; int cin_pad;
};

struct W2Ctl {
    uint64_t full[kW2Stages], empty[kW2Stages], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

struct W2Item { int m_blk, n_blk, tap0, ntap, kb0, nkb; };

__device__ __forceinline__ W2Item w2_decode(const Wgrad2Params &p, int w)
{
    W2Item it;
    const int ks = w % p.ksplit; w /= p.ksplit;
    const int tg = w % p.tap_groups; w /= p.tap_groups;
    it.tap0 = tg * p.tpc; it.ntap = min(p.tpc, p.ntaps - it.tap0);
    it.n_blk = w % p.n_tiles; it.m_blk = w / p.n_tiles;
    const int kb_per = (p.kblocks + p.ksplit - 1) / p.ksplit;
    it.kb0 = ks * kb_per;
    it.nkb = min(p.kblocks, it.kb0 + kb_per) - it.kb0;
    return it;
}

__global__ void __launch_bounds__(kW2Threads, 1)
conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const Wgrad2Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + kW2Stages * kW2AStage;
    W2Ctl *ctl = (W2Ctl *)(smem + kW2Stages * (kW2AStage + kW2BStage));
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmDy); prefetch_tmap(&tmX);
        for (int s = 0; s < kW2Stages; ++s) { mbar_init(&ctl->full[s], 1); mbar_init(&ctl->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&ctl->tmem_full[a], 1); mbar_init(&ctl->tmem_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&ctl->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
    const uint32_t b_tap_bytes = p.b_sw64 ? kW2Pix * 64 : p.b_boxes * kW2Pix * 128;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
                const W2Item it = w2_decode(p, w);
                if (it.nkb <= 0) continue;
                const int a_boxes = (it.m_blk * 128 + 64 < p.Cout) ? 2 : 1;       // skip a fully out-of-range box
                const uint32_t bytes = a_boxes * kW2Pix * 128 + it.ntap * b_tap_bytes;
                for (int kb = it.kb0; kb < it.kb0 + it.nkb; ++kb) {
                    const int m0 = kb * kW2Pix;
                    mbar_wait(&ctl->empty[stage], phase ^ 1);
                    mbar_expect_tx(&ctl->full[stage], bytes);
                    for (int bx = 0; bx < a_boxes; ++bx)
                        tma_load_2d(&tmDy, &ctl->full[stage], sA + stage * kW2AStage + bx * (kW2Pix * 128), it.m_blk * 128 + bx * 64, m0);
                    if (p.a_matrix) {
                        for (int bx = 0; bx < p.b_boxes; ++bx)
                            tma_load_2d(&tmX, &ctl->full[stage], sB + stage * kW2BStage + bx * (kW2Pix * 128), it.n_blk * p.block_n + bx * 64, m0);
                    } else {
                        const int img = m0 / (p.Po * p.Qo);
                        const int rem = m0 - img * (p.Po * p.Qo);
                        const int pi = rem / p.Qo, qi = rem - pi * p.Qo;
                        const int bw = qi * p.tstride + p.lower_w, bh = pi * p.tstride + p.lower_h;
                        for (int t = 0; t < it.ntap; ++t)
                            for (int bx = 0; bx < p.b_boxes; ++bx)
                                tma_load_im2col_4d(&tmX, &ctl->full[stage], sB + stage * kW2BStage + t * b_tap_bytes + bx * (kW2Pix * 128),
                                                   it.n_blk * p.block_n + bx * 64, bw, bh, img, (uint16_t)p.tap_ow[it.tap0 + t],
                                                   (uint16_t)p.tap_oh[it.tap0 + t]);
                    }
                    if (++stage == kW2Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_f16(128, p.block_n, 0, 1, 1);       // both operands MN-major
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        const uint32_t a_hi = smem_desc_hi(1024, SW_128B);
        const uint32_t b_hi = p.b_sw64 ? smem_desc_hi(512, SW_64B) : smem_desc_hi(1024, SW_128B);
        const uint32_t b_kstep = p.b_sw64 ? (16 * 64 / 16) : (16 * 128 / 16);          // 16 pixel rows, in 16-byte units
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            const W2Item it = w2_decode(p, w);
            if (it.nkb <= 0) continue;
            mbar_wait(&ctl->tmem_empty[acc], acc_phase ^ 1);          // the epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t d_acc = tmem_base + (uint32_t)acc * 256u;
            for (int kb = 0; kb < it.nkb; ++kb) {
                mbar_wait(&ctl->full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo0 = smem_desc_lo(a_base + stage * kW2AStage, kW2Pix * 128);
                    for (int t = 0; t < it.ntap; ++t) {
                        const uint32_t b_addr = b_base + stage * kW2BStage + t * b_tap_bytes;
                        const uint32_t b_lo0 = p.b_sw64 ? smem_desc_lo(b_addr, 0) : smem_desc_lo(b_addr, kW2Pix * 128);
                        const uint32_t d_t = d_acc + t * p.block_n;
#pragma unroll
                        for (int k = 0; k < kW2Pix / 16; ++k)
                            umma_f16_lohi(d_t, a_lo0 + k * (16 * 128 / 16), a_hi, b_lo0 + k * b_kstep, b_hi, idesc, (kb | k) != 0);
                    }
                    umma_commit(&ctl->empty[stage]);
                    if (kb == it.nkb - 1) umma_commit(&ctl->tmem_full[acc]);
                }
                if (++stage == kW2Stages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are this warp's
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
            const W2Item it = w2_decode(p, w);
            if (it.nkb <= 0) continue;
            const int co = it.m_blk * 128 + quarter * 32 + lane;
            mbar_wait(&ctl->tmem_full[acc], acc_phase);
            tc_fence_after();
            for (int t = 0; t < it.ntap; ++t) {
                float *row = p.dw + (int64_t)co * p.dw_row + (int64_t)(it.tap0 + t) * p.cin_pad;
                for (int c = 0; c < p.block_n / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)acc * 256u + t * p.block_n + c * 32, v);
                    tmem_ld_wait();
                    const int ci0 = it.n_blk * p.block_n + c * 32;
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
            tc_fence_before();                        // the TMEM reads are complete before the MMA warp may overwrite the stage
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

// Host side: same tiling as conv_wgrad.cu; the split-K factor aims at ~4 items per CTA of at least 8 k-blocks each.
int conv_wgrad2_launch(const cy4_conv_desc *d, const void *x, const void *dy, float *dw_acc, void *stream)
{
    const int k = d->ksize;
    const int cout64 = (d->Cout + 63) / 64 * 64;
    const bool sw64 = d->Cin == 32;
    const int cin64 = sw64 ? 32 : (d->Cin + 63) / 64 * 64;
    Wgrad2Params p;
    memset(&p, 0, sizeof(p));
    p.Mpix = d->B * d->Ho * d->Wo;
    p.Cout = d->Cout; p.Cin = d->Cin;
    p.m_tiles = (d->Cout + 127) / 128;
    p.block_n = sw64 ? 32 : std::min(cin64, 256);
    p.n_tiles = (cin64 + p.block_n - 1) / p.block_n;
    p.b_boxes = sw64 ? 1 : p.block_n / 64;
    p.b_sw64 = sw64 ? 1 : 0;
    p.ntaps = k * k;
    p.kblocks = (p.Mpix + kW2Pix - 1) / kW2Pix;
    p.tpc = std::max(1, std::min(p.ntaps, 256 / p.block_n));
    if (p.ntaps == 9) p.tpc = p.tpc >= 5 ? 5 : (p.tpc >= 3 ? 3 : p.tpc);
    p.tap_groups = (p.ntaps + p.tpc - 1) / p.tpc;
    const int items = p.m_tiles * p.n_tiles * p.tap_groups;
    const int sms = sm_count();
    p.ksplit = std::max(1, std::min(std::max(1, p.kblocks / 8), (4 * sms + items - 1) / items));
    p.n_work = items * p.ksplit;
    p.a_matrix = (d->flags & CY4_CONV_A_MATRIX) ? 1 : 0;
    if (p.a_matrix) CY4_CHECK_ARG(k == 1 && d->stride == 1 && d->pad == 0, "cy4_conv_wgrad: matrix mode needs a 1x1/s1/p0 conv");
    p.Po = d->Ho; p.Qo = d->Wo; p.tstride = d->stride; p.lower_w = p.lower_h = -d->pad;
    for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) { p.tap_ow[r * k + s] = (uint8_t)s; p.tap_oh[r * k + s] = (uint8_t)r; }
    p.cin_pad = d->Cin;
    p.dw = dw_acc; p.dw_row = (int64_t)k * k * d->Cin;
    if (p.Mpix <= 0) return 0;

    alignas(64) CUtensorMap tmDy, tmX;
    int rc = make_tmap_2d(&tmDy, dy, (uint64_t)cout64, (uint64_t)p.Mpix, (uint64_t)d->ldy * 2, 64, kW2Pix, 128, 0);
    if (rc) return rc;
    if (p.a_matrix)
        rc = make_tmap_2d(&tmX, x, (uint64_t)cin64, (uint64_t)p.Mpix, (uint64_t)d->ldx * 2, sw64 ? 32 : 64, kW2Pix, sw64 ? 64 : 128, 0);
    else
        rc = make_tmap_im2col(&tmX, x, cin64, d->Wi, d->Hi, d->B, d->ldx, -d->pad, -d->pad, d->pad - (k - 1), d->pad - (k - 1),
                              sw64 ? 32 : 64, kW2Pix, d->stride, sw64 ? 64 : 128, 0);
    if (rc) return rc;
    if (d->flags & CY4_CONV_ZERO_ACC)
        CY4_CUDA(cudaMemsetAsync(dw_acc, 0, (size_t)((d->Cout + 31) / 32 * 32) * k * k * d->Cin * sizeof(float), (cudaStream_t)stream));
    static bool attr_set = false;
    if (!attr_set) {
        CY4_CUDA(cudaFuncSetAttribute(conv_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2Smem));
        attr_set = true;
    }
    const int grid = std::min(p.n_work, sms);
    conv_wgrad2_kernel<<<grid, kW2Threads, kW2Smem, (cudaStream_t)stream>>>(tmDy, tmX, p);
    return cy4_launch_status("cy4_conv_wgrad (persistent variant)");
}

}  // namespace cy4
