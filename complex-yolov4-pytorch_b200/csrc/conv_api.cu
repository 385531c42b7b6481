// conv_api.cu -- C-ABI of the convolution stack: fprop / dgrad launch set-up (tensor maps, tap
// tables, tile shapes), weight packing, stem im2col.  Kernels live in conv_tc.cu / conv_wgrad.cu.
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "conv_tc.cuh"

namespace cy4 {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int pick_block_n(int cout_pad)
{
    for (int bn : {256, 128, 64, 32})
        if (cout_pad % bn == 0) return bn;
    return 32;
}

// Tunables (cy4_set_option).  Weight-slab multicast across 2/4-CTA clusters is implemented and
// parity-tested, but measured neutral on B200 for these layer shapes, so it is off by default.
int g_disable_tma_out = getenv("CY4_NO_TMA_OUT") != nullptr;
int g_cluster = getenv("CY4_CLUSTER") ? atoi(getenv("CY4_CLUSTER")) : 1;
int g_debug = 0;
int g_kps_max = 4;         // k-blocks per pipeline slot (upper bound; 1 disables the packing)
int g_conv_pair = 1;        // CTA-pair (cta_group::2) conv kernel for the eligible launches (conv_pair.cu); 0: always the 1-CTA kernel
int g_wgrad_cluster = getenv("CY4_WGRAD_CLUSTER") ? atoi(getenv("CY4_WGRAD_CLUSTER")) : 1;
int g_wgrad_pair = 1;       // CTA-pair weight-gradient kernel for the eligible launches (conv_wgrad.cu conv_wgrad_pair_kernel)
int g_wgrad_wide32 = 1;     // weight gradient of 32-channel inputs: the taps of a CTA as ONE N = 32*taps MMA operand
int g_accum_tma = 1;        // accumulate-mode outputs through TMA reduce-add stores (0: per-thread read-modify-write)
int g_pdl = getenv("CY4_PDL") ? atoi(getenv("CY4_PDL")) : 0;   // programmatic dependent launch of the hot kernels (common.cuh)
int g_ew_fwd_bpsm = 3, g_ew_bwd_bpsm = 2;   // grid caps of the BN / activation passes in blocks per SM (one resident wave each)
int g_ew_carveout = 0;      // 1: BN / activation passes ask for the max shared-memory carve-out (measured: -0.7 ms/step WORSE, the passes want their L1)
int g_dgrad_interleave = 1; // merged stride-2 dgrad: the four parity classes of a tile in neighbouring work units (dY re-reads hit L2)
int g_slab_stats = 1;       // BatchNorm statistics read off the staged fp16 output slab (0: reduce-scatter over the fp32 accumulators)
int g_conv1x1_matrix = 0;   // 1: 1x1 / stride-1 convs (fprop and dgrad) read their activation through a plain 2-D tiled TMA instead of im2col mode

// Generic launch: `a` is an NHWC tensor (C=Ca channels, ld lda) convolved with the tap table.
struct GenericConv {
    const void *a; int B, Ha, Wa, Ca; int64_t lda;
    int lower_w, lower_h, upper_w, upper_h, tstride;   // im2col box
    int Po, Qo;                                        // base-pixel grid per image
    int ntaps; uint8_t ow[kMaxTaps], oh[kMaxTaps]; int kofs[kMaxTaps];
    const void *w; int w_rows_pad; int64_t w_ktot;     // packed weights [rows_pad][ktot]
    int N;                                             // real output channels
    void *y; int64_t ldy; uint32_t flags;
    int omap, OH, OW, ostep, oh0, ow0;
    const float *bias; float *ch_sum, *ch_sqsum; const float *stat_shift;
    int a_matrix;
    int epi_mode, epi_act; const float *epi_scale, *epi_shift; const void *side; int64_t ld_side;   // fused epilogue (conv_tc.cuh)
    int ncls; uint8_t cls_tap0[4], cls_ntap[4], cls_oh0[4], cls_ow0[4];                              // tap classes merged into one launch
};

static int run_generic(const GenericConv &g, cudaStream_t st)
{
    ConvKParams p;
    memset(&p, 0, sizeof(p));
    p.kchunk = (g.Ca % 64 == 0) ? 64 : 32;
    if (g.Ca % 32 != 0) { set_error("conv: input channels (%d) must be a multiple of 32", g.Ca); return -1; }
    p.cin_chunks = g.Ca / p.kchunk;
    p.M = g.B * g.Po * g.Qo;
    p.N = g.N;
    p.block_n = pick_block_n(g.w_rows_pad);
    p.tiles_m = (p.M + 127) / 128;
    p.tiles_n = g.w_rows_pad / p.block_n;
    p.ntaps = g.ntaps;
    p.a_mode = g.a_matrix ? 0 : 1;
    p.debug = g_debug;
    p.ab_fmt = 0;
    p.Po = g.Po; p.Qo = g.Qo; p.tstride = g.tstride; p.lower_w = g.lower_w; p.lower_h = g.lower_h;
    for (int t = 0; t < g.ntaps; ++t) { p.tap_ow[t] = g.ow[t]; p.tap_oh[t] = g.oh[t]; p.tap_kofs[t] = g.kofs[t]; }
    p.y = g.y; p.ldy = g.ldy; p.flags = g.flags;
    p.omap = g.omap; p.OH = g.OH; p.OW = g.OW; p.ostep = g.ostep; p.oh0 = g.oh0; p.ow0 = g.ow0;
    p.bias = g.bias; p.ch_sum = g.ch_sum; p.ch_sqsum = g.ch_sqsum; p.stat_shift = g.stat_shift;
    p.epi_mode = g.epi_mode; p.epi_act = g.epi_act; p.epi_scale = g.epi_scale; p.epi_shift = g.epi_shift;
    p.side = g.side; p.ld_side = g.ld_side;
    p.ncls = g.ncls;
    // (small layers: dY stays in the 126 MB L2 between the class passes, and their few units per CTA balance better class-major)
    p.cls_interleave = g_dgrad_interleave && g.ncls > 1 && (int64_t)g.B * g.Ha * g.Wa * g.Ca * 2 >= (64ll << 20);
    int min_taps = g.ntaps;
    for (int c = 0; c < g.ncls && c < 4; ++c) {
        p.cls_tap0[c] = g.cls_tap0[c]; p.cls_ntap[c] = g.cls_ntap[c]; p.cls_oh0[c] = g.cls_oh0[c]; p.cls_ow0[c] = g.cls_ow0[c];
        min_taps = std::min(min_taps, (int)g.cls_ntap[c]);
    }
    if (p.M <= 0) return 0;
    const int swz = p.kchunk * 2;
    alignas(64) CUtensorMap tmA, tmB;
    int rc;
    if (g.a_matrix)
        rc = make_tmap_2d(&tmA, g.a, (uint64_t)g.Ca, (uint64_t)p.M, (uint64_t)g.lda * 2, p.kchunk, 128, swz, 0);
    else
        rc = make_tmap_im2col(&tmA, g.a, g.Ca, g.Wa, g.Ha, g.B, g.lda, g.lower_w, g.lower_h, g.upper_w, g.upper_h, p.kchunk,
                              128, g.tstride, swz, 0);
    if (rc) return rc;
    // clusters of 2 CTAs share each weight slab through TMA multicast (halves the L2 -> SM weight traffic)
    p.cluster = (p.tiles_m >= 2 && g_cluster >= 2) ? 2 : 1;
    if (g_cluster >= 4 && p.tiles_m >= 8 && p.block_n >= 64) p.cluster = 4;
    const bool pair = g_conv_pair && conv_pair_eligible(p);
    if (pair) p.cluster = 2;             // weight box of block_n / 2 rows: each CTA of the pair loads its half
    rc = make_tmap_2d(&tmB, g.w, (uint64_t)g.w_ktot, (uint64_t)g.w_rows_pad, (uint64_t)g.w_ktot * 2, p.kchunk, p.block_n / p.cluster, swz, 0);
    if (rc) return rc;
    // output slabs per epilogue warp: 4 for narrow tiles (epilogue-bound, TMA-store latency), else 1
    p.slab_bufs = p.block_n <= 64 ? 4 : (p.block_n <= 128 ? 2 : 1);
    const int slab_extra = (p.slab_bufs - 1) * 8 * 32 * 64;
    // split the 192 KB stage region into as many pipeline slots as fit (at most 12)
    p.a_stage = 128 * p.kchunk * 2;
    p.b_stage = (p.block_n * p.kchunk * 2 + 1023) / 1024 * 1024;
    // Every slot costs the issue threads a barrier round trip and a commit (~250 cycles, tools/probe_pipeline.py),
    // more than the MMAs of one narrow k-block (128 cycles at N = 64): pack several k-blocks per slot while a
    // slot stays <= 48 KB.
    const int num_kb = min_taps * p.cin_chunks;          // (of the shortest tap class)
    p.kps = (p.cluster == 1 && g_kps_max > 1) ? std::max(1, std::min(std::min(g_kps_max, num_kb), 49152 / (p.a_stage + p.b_stage))) : 1;
    p.stages = std::max(2, std::min(12, (4 * 49152 - slab_extra) / (p.kps * (p.a_stage + p.b_stage))));
    alignas(64) CUtensorMap tmC = tmB;
    // Accumulating launches (y += result: gradients of tensors with several consumers) keep the slab path too: the slab is
    // ADDED to global memory by a TMA reduce (cp.reduce.async.bulk.tensor ... .add, fp16) instead of a per-thread strided
    // read-modify-write.  Not for EPI_BWD_DZ, which needs the total in registers before it applies act'.
    const bool accum_tma = (p.flags & CONV_F_ACCUM) && p.epi_mode != EPI_BWD_DZ && g_accum_tma;
    if (!(p.flags & CONV_F_OUT_F32) && (!(p.flags & CONV_F_ACCUM) || accum_tma) && !p.omap && !g_disable_tma_out) {
        // dense fp16 output: epilogue stages 32-row slabs in swizzled smem and TMA-stores them
        const int cw = 32;
        rc = make_tmap_2d(&tmC, g.y, (uint64_t)g.w_rows_pad, (uint64_t)p.M, (uint64_t)g.ldy * 2, cw, 32, cw * 2, 0);
        if (rc) return rc;
        p.flags |= CONV_F_TMA_OUT;
        if (!g_slab_stats) p.flags |= CONV_F_ACC_STATS;
    }
    return pair ? launch_conv_pair(tmA, tmB, tmC, p, st) : launch_conv_tc(tmA, tmB, tmC, p, st);
}

static int check_conv_desc(const cy4_conv_desc *d, const char *who)
{
    if (!d || d->B <= 0 || d->Hi <= 0 || d->Wi <= 0 || d->Cin <= 0 || d->Cout <= 0 || d->ksize <= 0 || d->ksize > 3 ||
        d->stride <= 0 || d->stride > 2 || d->pad < 0) {
        set_error("%s: bad conv descriptor", who);
        return -1;
    }
    const int ho = (d->Hi + 2 * d->pad - d->ksize) / d->stride + 1, wo = (d->Wi + 2 * d->pad - d->ksize) / d->stride + 1;
    if (ho != d->Ho || wo != d->Wo) { set_error("%s: Ho/Wo (%d,%d) inconsistent with the conv geometry (%d,%d)", who, d->Ho, d->Wo, ho, wo); return -1; }
    if ((d->ldx % 8) || d->ldx < d->Cin) { set_error("%s: ldx must be a multiple of 8 and >= Cin", who); return -1; }
    return 0;
}

// ---- packing kernels -----------------------------------------------------------------------------
__global__ void pack_fprop_kernel(const float *__restrict__ w, int Cout, int Cin, int k, int cin_pad, int cout_pad, __half *__restrict__ out)
{
    // out[co][r][s][ci]  (ci < cin_pad), zero padded
    const int64_t total = (int64_t)cout_pad * k * k * cin_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_pad);
        int64_t r = i / cin_pad;
        const int s = (int)(r % k); r /= k;
        const int rr = (int)(r % k);
        const int co = (int)(r / k);
        float v = 0.f;
        if (co < Cout && ci < Cin) v = w[(((int64_t)co * Cin + ci) * k + rr) * k + s];
        out[i] = __float2half_rn(v);
    }
}

__global__ void pack_dgrad_kernel(const float *__restrict__ w, int Cout, int Cin, int k, int cin_rows_pad, __half *__restrict__ out)
{
    // out[ci][r][s][co]
    const int64_t total = (int64_t)cin_rows_pad * k * k * Cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        int64_t r = i / Cout;
        const int s = (int)(r % k); r /= k;
        const int rr = (int)(r % k);
        const int ci = (int)(r / k);
        float v = 0.f;
        if (ci < Cin) v = w[(((int64_t)co * Cin + ci) * k + rr) * k + s];
        out[i] = __float2half_rn(v);
    }
}

__global__ void unpack_wgrad_kernel(const float *__restrict__ acc, int Cout, int Cin, int k, int cin_pad, float scale,
                                    const float *__restrict__ dscale, int accumulate, float *__restrict__ gw)
{
    if (dscale) scale *= __ldg(dscale);
    // gw[co][ci][r][s] (+)= scale * acc[co][r*k+s][ci]
    const int64_t total = (int64_t)Cout * Cin * k * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i % k);
        int64_t r = i / k;
        const int rr = (int)(r % k); r /= k;
        const int ci = (int)(r % Cin);
        const int co = (int)(r / Cin);
        const float v = scale * acc[((int64_t)co * k * k + rr * k + s) * cin_pad + ci];
        gw[i] = accumulate ? gw[i] + v : v;
    }
}

// ---- batched variants: one launch for every conv layer of the network ------------------------------
// Both directions are (co, ci, tap) <-> (co, tap, ci) / (ci, tap, co) re-layouts.  A block moves one 32 x 32 (co x ci) tile with
// all its taps through shared memory so that global reads AND writes are contiguous runs (the direct form reads the fp32
// parameters with a stride of k*k or Cin*k*k elements and ran at ~1/8 of the copy bandwidth).
constexpr int kPkT = 32;                                  // tile edge
constexpr int kPkTaps = 9;                                // k <= 3
constexpr int kPkLd = kPkT * kPkTaps + 1;                 // +1: conflict-free column reads

// The tiles of ALL items form one list (item.tile_begin = prefix sum of the per-item tile counts, filled by the caller), walked by a
// grid of a few blocks per SM: the dozen 4.7 M-parameter layers that dominate the bytes are spread over the whole GPU instead of
// over the blocks of one grid row.
template <typename Item>
__device__ __forceinline__ int find_item(const Item *__restrict__ items, int n, int tile)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {                                     // last item with tile_begin <= tile
        const int mid = (lo + hi + 1) >> 1;
        if (items[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
pack_batched_kernel(const cy4_pack_item *__restrict__ items, int n)
{
    __shared__ float t[kPkT][kPkLd];
    const cy4_pack_item last = items[n - 1];
    const int total = last.tile_begin + ((last.cin_pad + kPkT - 1) / kPkT) * ((last.cout_pad + kPkT - 1) / kPkT);
    for (int gt = items[0].tile_begin + blockIdx.x; gt < total; gt += gridDim.x) {      // (a sub-range of a table starts at its own tile_begin)
        const cy4_pack_item it = items[find_item(items, n, gt)];
        const float *__restrict__ w = it.w_oihw;
        const int kk = it.ksize * it.ksize, Cin = it.Cin, Cout = it.Cout;
        const int run = kPkT * kk;                        // floats of one co row inside the tile
        const int tiles_ci = (it.cin_pad + kPkT - 1) / kPkT;
        __half *of = (__half *)it.w_fprop, *od = (__half *)it.w_dgrad;
        const int tile = gt - it.tile_begin;
        const int co0 = (tile / tiles_ci) * kPkT, ci0 = (tile % tiles_ci) * kPkT;
        // (index arithmetic without divisions: a warp owns whole rows, lanes run along the contiguous dimension; Cin is a
        // multiple of 32, so a ci tile is always full)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < kPkT; r += 8) {
            const int co = co0 + r;
            const float *src = w + (int64_t)co * Cin * kk + (int64_t)ci0 * kk;
            for (int c = lane; c < run; c += 32) t[r][c] = co < Cout ? src[c] : 0.f;
        }
        __syncthreads();
        if (of) {                                         // [cout_pad][tap][Cin]
            for (int r = warp; r < kPkT; r += 8) {
                const int co = co0 + r;
                if (co >= it.cout_pad) break;
                const float fs = (it.fold_scale && co < Cout) ? __ldg(it.fold_scale + co) : 1.f;
                for (int tap = 0; tap < kk; ++tap)
                    of[((int64_t)co * kk + tap) * Cin + ci0 + lane] = __float2half_rn(t[r][lane * kk + tap] * fs);
            }
        }
        if (od && co0 + lane < it.cout_pad) {             // [cin_pad][tap][cout_pad]
            for (int cil = warp; cil < kPkT; cil += 8)
                for (int tap = 0; tap < kk; ++tap)
                    od[((int64_t)(ci0 + cil) * kk + tap) * it.cout_pad + co0 + lane] = __float2half_rn(t[lane][cil * kk + tap]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
unpack_batched_kernel(const cy4_unpack_item *__restrict__ items, int n, const float *__restrict__ dscale)
{
    __shared__ float t[kPkT][kPkLd];
    const float scale = dscale ? __ldg(dscale) : 1.f;
    const cy4_unpack_item last = items[n - 1];
    const int total = last.tile_begin + ((last.Cin + kPkT - 1) / kPkT) * ((last.Cout + kPkT - 1) / kPkT);
    for (int gt = items[0].tile_begin + blockIdx.x; gt < total; gt += gridDim.x) {      // (a sub-range of a table starts at its own tile_begin)
        const cy4_unpack_item it = items[find_item(items, n, gt)];
        const int kk = it.ksize * it.ksize, Cin = it.Cin, Cout = it.Cout;
        const int run = kPkT * kk;
        const int tiles_ci = (Cin + kPkT - 1) / kPkT;
        const int tile = gt - it.tile_begin;
        const int co0 = (tile / tiles_ci) * kPkT, ci0 = (tile % tiles_ci) * kPkT;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < kPkT; r += 8) {                      // read acc[co][tap][ci0 .. ci0+32): 32 contiguous floats
            const int co = co0 + r;
            for (int tap = 0; tap < kk; ++tap)
                t[r][lane * kk + tap] = co < Cout ? it.dw_acc[((int64_t)co * kk + tap) * Cin + ci0 + lane] : 0.f;
        }
        __syncthreads();
        for (int r = warp; r < kPkT; r += 8) {                      // write gw[co][ci0*kk .. (ci0+32)*kk): one contiguous run
            const int co = co0 + r;
            if (co >= Cout) break;
            float *dst = it.gw_oihw + (int64_t)co * Cin * kk + (int64_t)ci0 * kk;
            for (int c = lane; c < run; c += 32) dst[c] = scale * t[r][c];
        }
        __syncthreads();
    }
}

__global__ void stem_im2col_kernel(const float *__restrict__ x, int B, int C, int H, int W, int k, int stride, int pad, int Ho, int Wo,
                                   __half *__restrict__ cols)
{
    // one thread per output pixel: 32 halves (64 B) per row, (r, s, c) order then zeros
    const int64_t M = (int64_t)B * Ho * Wo;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int q = (int)(m % Wo);
    const int pp = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    __align__(16) __half row[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) row[i] = __float2half_rn(0.f);
    int idx = 0;
    for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) {
            const int h = pp * stride - pad + r, w = q * stride - pad + s;
            const bool in = h >= 0 && h < H && w >= 0 && w < W;
            for (int c = 0; c < C; ++c, ++idx)
                if (idx < 32 && in) row[idx] = __float2half_rn(__ldg(x + (((int64_t)b * C + c) * H + h) * W + w));
        }
    uint4 *dst = (uint4 *)(cols + m * 32);
    const uint4 *src = (const uint4 *)row;
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = src[i];
}

// The shape every complex-yolov4 cfg uses (3 channels, 3 x 3, pad 1): fully unrolled, so the 27 taps stay in registers (the generic
// kernel's dynamically indexed row[] lives in local memory) and each (c, r, s) load is one coalesced 128-byte line per warp.
template <int STRIDE>
__global__ void __launch_bounds__(256)
stem_im2col_c3k3_kernel(const float *__restrict__ x, int B, int H, int W, int Ho, int Wo, __half *__restrict__ cols)
{
    const int64_t M = (int64_t)B * Ho * Wo;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int q = (int)(m % Wo);
    const int pp = (int)((m / Wo) % Ho);
    const int b = (int)(m / ((int64_t)Wo * Ho));
    const float *xb = x + (int64_t)b * 3 * H * W;
    float v[32];
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int h = pp * STRIDE - 1 + r;
        const bool hin = h >= 0 && h < H;
#pragma unroll
        for (int s_ = 0; s_ < 3; ++s_) {
            const int w = q * STRIDE - 1 + s_;
            const bool in = hin && w >= 0 && w < W;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[(r * 3 + s_) * 3 + c] = in ? __ldg(xb + ((int64_t)c * H + h) * W + w) : 0.f;
        }
    }
    uint4 *dst = (uint4 *)(cols + m * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 o; __half2 *ph = (__half2 *)&o;
#pragma unroll
        for (int j = 0; j < 4; ++j) ph[j] = __floats2half2_rn(v[i * 8 + 2 * j], v[i * 8 + 2 * j + 1]);
        dst[i] = o;
    }
}

}  // namespace cy4

using namespace cy4;

extern "C" {

int cy4_set_option(const char *name, int value)
{
    CY4_CHECK_ARG(name, "cy4_set_option: null name");
    if (!strcmp(name, "conv_cluster")) { CY4_CHECK_ARG(value == 1 || value == 2 || value == 4, "conv_cluster must be 1, 2 or 4"); g_cluster = value; return 0; }
    if (!strcmp(name, "wgrad_cluster")) { CY4_CHECK_ARG(value == 1 || value == 2, "wgrad_cluster must be 1 or 2"); g_wgrad_cluster = value; return 0; }
    if (!strcmp(name, "tma_store")) { g_disable_tma_out = value ? 0 : 1; return 0; }
    if (!strcmp(name, "kblocks_per_slot")) { g_kps_max = value < 1 ? 1 : (value > 8 ? 8 : value); return 0; }
    if (!strcmp(name, "wgrad_wide32")) { g_wgrad_wide32 = value ? 1 : 0; return 0; }
    if (!strcmp(name, "accum_tma")) { g_accum_tma = value ? 1 : 0; return 0; }
    if (!strcmp(name, "wgrad_pair")) { g_wgrad_pair = value ? 1 : 0; return 0; }
    if (!strcmp(name, "conv1x1_matrix")) { g_conv1x1_matrix = value ? 1 : 0; return 0; }
    if (!strcmp(name, "pdl")) { g_pdl = value ? 1 : 0; return 0; }
    if (!strcmp(name, "ew_fwd_blocks_per_sm")) { CY4_CHECK_ARG(value >= 1 && value <= 32, "ew_fwd_blocks_per_sm must be in 1..32"); g_ew_fwd_bpsm = value; return 0; }
    if (!strcmp(name, "ew_bwd_blocks_per_sm")) { CY4_CHECK_ARG(value >= 1 && value <= 32, "ew_bwd_blocks_per_sm must be in 1..32"); g_ew_bwd_bpsm = value; return 0; }
    if (!strcmp(name, "ew_carveout")) { g_ew_carveout = value ? 1 : 0; return 0; }
    if (!strcmp(name, "dgrad_interleave")) { g_dgrad_interleave = value ? 1 : 0; return 0; }
    if (!strcmp(name, "slab_stats")) { g_slab_stats = value ? 1 : 0; return 0; }
    if (!strcmp(name, "conv_pair")) { g_conv_pair = value ? 1 : 0; return 0; }
    if (!strcmp(name, "debug")) { g_debug = value; return 0; }      // bottleneck experiments: results are garbage
    set_error("cy4_set_option: unknown option %s", name);
    return -1;
}

struct EpiArgs { int mode, act; const float *scale, *shift; const void *side; int64_t ld_side; const float *stat_shift; };

static int conv_fwd_impl(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, const float *bias, float *ch_sum,
                         float *ch_sqsum, const EpiArgs *epi, void *stream)
{
    if (check_conv_desc(d, "cy4_conv_fwd")) return -1;
    CY4_CHECK_ARG(x && w_fprop && y, "cy4_conv_fwd: null pointer");
    CY4_CHECK_ARG(!(d->flags & CY4_CONV_STATS) || (ch_sum && ch_sqsum), "cy4_conv_fwd: STATS needs ch_sum / ch_sqsum");
    const int cout_pad = round_up(d->Cout, 32);
    CY4_CHECK_ARG(d->ldy >= cout_pad, "cy4_conv_fwd: ldy must be >= Cout rounded up to 32");
    CY4_CHECK_ARG((d->flags & CY4_CONV_OUT_F32) ? (d->ldy % 4 == 0) : (d->ldy % 8 == 0), "cy4_conv_fwd: ldy alignment");
    GenericConv g;
    memset(&g, 0, sizeof(g));
    const int k = d->ksize;
    g.a = x; g.B = d->B; g.Ha = d->Hi; g.Wa = d->Wi; g.Ca = d->Cin; g.lda = d->ldx;
    g.lower_w = g.lower_h = -d->pad;
    g.upper_w = g.upper_h = d->pad - (k - 1);
    g.tstride = d->stride;
    g.Po = d->Ho; g.Qo = d->Wo;
    g.ntaps = k * k;
    for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) { g.ow[r * k + s] = (uint8_t)s; g.oh[r * k + s] = (uint8_t)r; g.kofs[r * k + s] = (r * k + s) * d->Cin; }
    g.w = w_fprop; g.w_rows_pad = cout_pad; g.w_ktot = (int64_t)k * k * d->Cin;
    g.N = d->Cout;
    g.y = y; g.ldy = d->ldy;
    g.flags = ((d->flags & CY4_CONV_OUT_F32) ? CONV_F_OUT_F32 : 0) | ((d->flags & CY4_CONV_STATS) ? CONV_F_STATS : 0) |
              ((d->flags & CY4_CONV_ACCUM) ? CONV_F_ACCUM : 0);
    g.bias = bias; g.ch_sum = ch_sum; g.ch_sqsum = ch_sqsum;
    g.a_matrix = (d->flags & CY4_CONV_A_MATRIX) ? 1 : 0;
    if (g.a_matrix) CY4_CHECK_ARG(k == 1 && d->stride == 1 && d->pad == 0, "cy4_conv_fwd: matrix mode needs a 1x1/s1/p0 conv");
    if (g_conv1x1_matrix && k == 1 && d->stride == 1 && d->pad == 0) g.a_matrix = 1;
    if (epi) { g.epi_mode = epi->mode; g.epi_act = epi->act; g.epi_scale = epi->scale; g.epi_shift = epi->shift; g.side = epi->side; g.ld_side = epi->ld_side;
               g.stat_shift = epi->stat_shift; }
    return run_generic(g, (cudaStream_t)stream);
}

int cy4_conv_fwd_stats(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, float *ch_sum, float *ch_sqsum,
                       const float *stat_shift, void *stream)
{
    CY4_CHECK_ARG(d && (d->flags & CY4_CONV_STATS) && ch_sum && ch_sqsum, "cy4_conv_fwd_stats: needs CY4_CONV_STATS and the two sum arrays");
    CY4_CHECK_ARG(d->Cout % 32 == 0, "cy4_conv_fwd_stats: Cout must be a multiple of 32");
    const EpiArgs e = {EPI_NONE, 0, nullptr, nullptr, nullptr, 0, stat_shift};
    return conv_fwd_impl(d, x, w_fprop, y, nullptr, ch_sum, ch_sqsum, &e, stream);
}

int cy4_conv_fwd(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, const float *bias, float *ch_sum,
                 float *ch_sqsum, void *stream)
{
    return conv_fwd_impl(d, x, w_fprop, y, bias, ch_sum, ch_sqsum, nullptr, stream);
}

int cy4_conv_fwd_fused(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, const float *shift, int act,
                       const void *residual, int64_t ldr, void *stream)
{
    CY4_CHECK_ARG(d && shift && act >= 0 && act <= 2, "cy4_conv_fwd_fused: bad argument");
    CY4_CHECK_ARG(!(d->flags & (CY4_CONV_OUT_F32 | CY4_CONV_STATS | CY4_CONV_ACCUM)), "cy4_conv_fwd_fused: fp16 output without statistics / accumulation only");
    CY4_CHECK_ARG(d->Cout % 32 == 0, "cy4_conv_fwd_fused: Cout must be a multiple of 32");
    CY4_CHECK_ARG(!residual || (ldr % 8 == 0 && ldr >= d->Cout), "cy4_conv_fwd_fused: residual ld");
    const EpiArgs e = {EPI_FWD_ACT, act, nullptr, shift, residual, ldr, nullptr};
    return conv_fwd_impl(d, x, w_fprop, y, nullptr, nullptr, nullptr, &e, stream);
}

static int conv_dgrad_impl(const cy4_conv_desc *d, const void *dy, const void *w_dgrad, void *dx, const EpiArgs *epi, float *s_dz,
                           float *s_dzy, void *stream);

int cy4_conv_dgrad(const cy4_conv_desc *d, const void *dy, const void *w_dgrad, void *dx, void *stream)
{
    return conv_dgrad_impl(d, dy, w_dgrad, dx, nullptr, nullptr, nullptr, stream);
}

int cy4_conv_dgrad_fused(const cy4_conv_desc *d, const void *dy, const void *w_dgrad, void *dx, const void *y_producer, int64_t ldyp,
                         const float *scale, const float *shift, int act, float *sum_dz, float *sum_dzy, void *stream)
{
    CY4_CHECK_ARG(d && y_producer && scale && shift && sum_dz && sum_dzy && act >= 0 && act <= 2, "cy4_conv_dgrad_fused: bad argument");
    CY4_CHECK_ARG(ldyp % 8 == 0 && ldyp >= d->Cin, "cy4_conv_dgrad_fused: producer ld");
    const EpiArgs e = {EPI_BWD_DZ, act, scale, shift, y_producer, ldyp, nullptr};
    return conv_dgrad_impl(d, dy, w_dgrad, dx, &e, sum_dz, sum_dzy, stream);
}

static int conv_dgrad_impl(const cy4_conv_desc *d, const void *dy, const void *w_dgrad, void *dx, const EpiArgs *epi, float *s_dz,
                           float *s_dzy, void *stream)
{
    if (check_conv_desc(d, "cy4_conv_dgrad")) return -1;
    CY4_CHECK_ARG(dy && w_dgrad && dx, "cy4_conv_dgrad: null pointer");
    CY4_CHECK_ARG(d->Cout % 32 == 0, "cy4_conv_dgrad: Cout must be a multiple of 32 (pad dy)");
    CY4_CHECK_ARG(d->Cin % 32 == 0 && d->ldx % 8 == 0 && d->ldy % 8 == 0, "cy4_conv_dgrad: channel alignment");
    const int k = d->ksize;
    GenericConv g;
    memset(&g, 0, sizeof(g));
    g.a = dy; g.B = d->B; g.Ha = d->Ho; g.Wa = d->Wo; g.Ca = d->Cout; g.lda = d->ldy;
    g.w = w_dgrad; g.w_rows_pad = round_up(d->Cin, 32); g.w_ktot = (int64_t)k * k * d->Cout;
    g.N = d->Cin;
    g.y = dx; g.ldy = d->ldx;
    g.flags = (d->flags & CY4_CONV_ACCUM) ? CONV_F_ACCUM : 0;
    if (epi) {
        g.epi_mode = epi->mode; g.epi_act = epi->act; g.epi_scale = epi->scale; g.epi_shift = epi->shift; g.side = epi->side; g.ld_side = epi->ld_side;
        g.flags |= CONV_F_STATS; g.ch_sum = s_dz; g.ch_sqsum = s_dzy;
    }
    if (d->stride == 1) {
        CY4_CHECK_ARG(d->pad == k / 2 && (k & 1), "cy4_conv_dgrad: stride 1 needs odd k and pad = k/2");
        // dx[h] = sum_r dy[h + pad - r] w[r]: base = h + lower, offset o = k-1-r
        g.lower_w = g.lower_h = d->pad - (k - 1);
        g.upper_w = g.upper_h = g.lower_w;               // Ho == Hi: the box spans Hi base pixels
        g.tstride = 1;
        g.Po = d->Hi; g.Qo = d->Wi;
        if (g_conv1x1_matrix && k == 1) g.a_matrix = 1;
        g.ntaps = k * k;
        for (int o_r = 0; o_r < k; ++o_r)
            for (int o_s = 0; o_s < k; ++o_s) {
                const int t = o_r * k + o_s;
                g.ow[t] = (uint8_t)o_s; g.oh[t] = (uint8_t)o_r;
                g.kofs[t] = ((k - 1 - o_r) * k + (k - 1 - o_s)) * d->Cout;
            }
        return run_generic(g, (cudaStream_t)stream);
    }
    CY4_CHECK_ARG(d->stride == 2 && k == 3 && d->pad == 1 && (d->Hi % 2 == 0) && (d->Wi % 2 == 0),
                  "cy4_conv_dgrad: stride 2 is implemented for k=3, pad=1, even input size");
    // hi = 2*ho - 1 + r.  Parity class ph: ph=0 -> (r=1, o=0);  ph=1 -> (r=0, o=1), (r=2, o=0), where o = ho - i.
    g.lower_w = g.lower_h = 0; g.upper_w = g.upper_h = 0; g.tstride = 1;
    g.Po = d->Hi / 2; g.Qo = d->Wi / 2;
    g.omap = 1; g.OH = d->Hi; g.OW = d->Wi; g.ostep = 2;
    static const int cls_n[2] = {1, 2};
    static const int cls_r[2][2] = {{1, 0}, {0, 2}};
    static const int cls_o[2][2] = {{0, 0}, {1, 0}};
    if (pick_block_n(g.w_rows_pad) <= 32) {
        // Narrow input (Cin = 32): these tiles are bound by the issue threads and pack several k-blocks per pipeline slot; the
        // packing factor is per launch, and the 1-tap class would force 1 -- one launch per parity class (measured faster here).
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                g.oh0 = ph; g.ow0 = pw;
                g.ntaps = 0;
                for (int a = 0; a < cls_n[ph]; ++a)
                    for (int b = 0; b < cls_n[pw]; ++b) {
                        const int t = g.ntaps++;
                        g.oh[t] = (uint8_t)cls_o[ph][a]; g.ow[t] = (uint8_t)cls_o[pw][b];
                        g.kofs[t] = (cls_r[ph][a] * 3 + cls_r[pw][b]) * d->Cout;
                    }
                const int rc = run_generic(g, (cudaStream_t)stream);
                if (rc) return rc;
            }
        return 0;
    }
    // All four parity classes in ONE launch (work unit -> class): a single grid instead of four part-filled ones
    // (-20...45 % on the wide layers, profiles/r2_conv_shape_bench.md).
    g.ntaps = 0; g.ncls = 0;
    for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
            const int c = g.ncls++;
            g.cls_oh0[c] = (uint8_t)ph; g.cls_ow0[c] = (uint8_t)pw;
            g.cls_tap0[c] = (uint8_t)g.ntaps;
            for (int a = 0; a < cls_n[ph]; ++a)
                for (int b = 0; b < cls_n[pw]; ++b) {
                    const int t = g.ntaps++;
                    g.oh[t] = (uint8_t)cls_o[ph][a]; g.ow[t] = (uint8_t)cls_o[pw][b];
                    g.kofs[t] = (cls_r[ph][a] * 3 + cls_r[pw][b]) * d->Cout;
                }
            g.cls_ntap[c] = (uint8_t)(g.ntaps - g.cls_tap0[c]);
        }
    return run_generic(g, (cudaStream_t)stream);
}

int cy4_pack_weight_fprop(const float *w_oihw, int Cout, int Cin, int ksize, int cin_pad, void *w_packed, void *stream)
{
    CY4_CHECK_ARG(w_oihw && w_packed && Cout > 0 && Cin > 0 && ksize > 0 && cin_pad >= Cin, "cy4_pack_weight_fprop: bad argument");
    const int cout_pad = round_up(Cout, 32);
    const int64_t total = (int64_t)cout_pad * ksize * ksize * cin_pad;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 4096);
    pack_fprop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, ksize, cin_pad, cout_pad, (__half *)w_packed);
    return cy4_launch_status("cy4_pack_weight_fprop");
}

int cy4_pack_weight_dgrad(const float *w_oihw, int Cout, int Cin, int ksize, void *w_packed, void *stream)
{
    CY4_CHECK_ARG(w_oihw && w_packed && Cout > 0 && Cin > 0 && ksize > 0, "cy4_pack_weight_dgrad: bad argument");
    const int rows = round_up(Cin, 32);
    const int64_t total = (int64_t)rows * ksize * ksize * Cout;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 4096);
    pack_dgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, ksize, rows, (__half *)w_packed);
    return cy4_launch_status("cy4_pack_weight_dgrad");
}

int cy4_unpack_wgrad(const float *dw_acc, int Cout, int Cin, int ksize, int cin_pad, float scale, const float *dscale, int accumulate,
                     float *gw_oihw, void *stream)
{
    CY4_CHECK_ARG(dw_acc && gw_oihw && Cout > 0 && Cin > 0 && ksize > 0 && cin_pad >= Cin, "cy4_unpack_wgrad: bad argument");
    const int64_t total = (int64_t)Cout * Cin * ksize * ksize;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 4096);
    unpack_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dw_acc, Cout, Cin, ksize, cin_pad, scale, dscale, accumulate, gw_oihw);
    return cy4_launch_status("cy4_unpack_wgrad");
}

int cy4_pack_weights_batched(const cy4_pack_item *items_dev, int n, void *stream)
{
    CY4_CHECK_ARG(items_dev && n > 0, "cy4_pack_weights_batched: bad argument");
    pack_batched_kernel<<<sm_count() * 6, 256, 0, (cudaStream_t)stream>>>(items_dev, n);
    return cy4_launch_status("cy4_pack_weights_batched");
}

int cy4_unpack_wgrad_batched(const cy4_unpack_item *items_dev, int n, const float *dscale, void *stream)
{
    CY4_CHECK_ARG(items_dev && n > 0, "cy4_unpack_wgrad_batched: bad argument");
    unpack_batched_kernel<<<sm_count() * 6, 256, 0, (cudaStream_t)stream>>>(items_dev, n, dscale);
    return cy4_launch_status("cy4_unpack_wgrad_batched");
}

int cy4_stem_im2col(const float *x_nchw, int B, int C, int H, int W, int ksize, int stride, int pad, void *cols, void *stream)
{
    CY4_CHECK_ARG(x_nchw && cols && B > 0 && C > 0 && C * ksize * ksize <= 32, "cy4_stem_im2col: needs C*k*k <= 32");
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    const int64_t M = (int64_t)B * Ho * Wo;
    if (C == 3 && ksize == 3 && pad == 1 && (stride == 1 || stride == 2)) {
        if (stride == 1) stem_im2col_c3k3_kernel<1><<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, B, H, W, Ho, Wo, (__half *)cols);
        else stem_im2col_c3k3_kernel<2><<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, B, H, W, Ho, Wo, (__half *)cols);
        return cy4_launch_status("cy4_stem_im2col");
    }
    stem_im2col_kernel<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x_nchw, B, C, H, W, ksize, stride, pad, Ho, Wo, (__half *)cols);
    return cy4_launch_status("cy4_stem_im2col");
}

}  // extern "C"
