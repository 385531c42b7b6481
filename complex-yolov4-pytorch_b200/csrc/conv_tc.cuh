// conv_tc.cuh -- kernel parameter block and host helpers shared by the tensor-core conv kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cy4 {

enum : uint32_t {
    CONV_F_OUT_F32 = 1u,   // fp32 output (+ optional bias) instead of fp16
    CONV_F_STATS = 2u,     // accumulate per-channel sum / sum^2 of the fp32 accumulators
    CONV_F_ACCUM = 4u,     // y += result (fp16 read-modify-write), used by dgrad into shared grads
    CONV_F_TMA_OUT = 16u,  // internal: fp16 tile staged in swizzled smem and written with TMA stores
    CONV_F_ACC_STATS = 32u,// internal (option slab_stats=0): statistics by reduce-scatter over the fp32 accumulators even when a slab is staged
};

// Fused epilogues (ConvKParams::epi_mode).  Both apply per-output-channel parameters and read a "side" fp16 tensor with
// the output's row mapping.
enum : int {
    EPI_NONE = 0,
    // y = act(acc + shift[n]) (+ side[m, n]): eval-mode inference with BatchNorm folded into the weights (scale) and this
    // shift; `side` is the residual of a fused shortcut.  Reference: darknet2pytorch.py:247-278 with model.eval().
    EPI_FWD_ACT = 1,
    // Input gradient fused with the FIRST pass of the producer's BatchNorm/activation backward: v = acc (+ old gradient
    // when CONV_F_ACCUM), z = scale[n] * side[m, n] + shift[n] (side = the producer's raw conv output Y),
    // dz = v * act'(z) is stored instead of v, and sum_m dz / sum_m dz * Y are accumulated into ch_sum / ch_sqsum
    // (what cy4_bn_act_bwd_reduce computes in a separate pass over the tensor).
    EPI_BWD_DZ = 2,
};

constexpr int kMaxTaps = 16;

struct ConvKParams {
    int M, N;                    // GEMM rows (output pixels of this launch), real output channels
    int tiles_m, tiles_n, block_n;
    int kchunk, cin_chunks, ntaps;
    int a_mode;                  // 0: A is a plain [M, K] matrix (tiled TMA) ; 1: im2col TMA
    int cluster;                 // CTAs per cluster sharing (multicasting) the weight slabs: 1, 2 or 4
    int debug;                   // 1: skip MMAs, 2: skip TMA loads (bottleneck experiments)
    int stages, a_stage, b_stage;   // pipeline depth and per-stage bytes (the 192 KB stage region is split to fit)
    int kps;                     // k-blocks per pipeline slot (one barrier round trip per kps k-blocks)
    int slab_bufs;               // output slabs per epilogue warp (1, 2 or 4): TMA-store latency hiding for narrow layers
    int ab_fmt;                  // 0 fp16, 1 bf16
    // im2col base-pixel space: pixel m -> (img, pi, qi) over Po x Qo ; TMA base = (qi*tstride + lower_w, ...)
    int Po, Qo, tstride, lower_w, lower_h;
    uint8_t tap_ow[kMaxTaps], tap_oh[kMaxTaps];
    int tap_kofs[kMaxTaps];      // k offset of each tap inside a packed weight row
    // output
    void *y; int64_t ldy;
    int omap, OH, OW, ostep, oh0, ow0;   // strided output-row mapping (stride-2 dgrad parity classes)
    const float *bias; float *ch_sum, *ch_sqsum;
    const float *stat_shift;     // CONV_F_STATS: NULL, or per-channel c[n]: the sums are of (y - c) and (y - c)^2 (rows >= M excluded)
    uint32_t flags;
    // Several tap classes in ONE launch (stride-2 dgrad: the four output-parity classes, each with its own taps and output
    // offsets, all over the same Po x Qo base-pixel grid).  Work unit t -> class t / cls_units.  ncls <= 1: one class made of
    // taps [0, ntaps) and (oh0, ow0) above.
    int ncls, cls_units;
    int cls_interleave;          // 1: unit -> (tile t / ncls, class rotated per tile), see unit_decode
    uint8_t cls_tap0[4], cls_ntap[4], cls_oh0[4], cls_ow0[4];
    // fused epilogue
    int epi_mode, epi_act;
    const float *epi_scale, *epi_shift;  // per output channel
    const void *side; int64_t ld_side;   // fp16 [rows, ld_side], same row mapping as y
};

// Work unit t of a launch with several tap classes -> (class, tile of that class).  Class-major order (t / cls_units) walks the
// whole dY tensor once PER CLASS: 4 DRAM passes over dY for the stride-2 input gradients (ncu: 1.1 GB of DRAM traffic for 0.57 GB
// of algorithmic bytes on the 64->128 layer).  Interleaved order puts the four classes of one tile into units 4*tile .. 4*tile+3,
// which neighbouring CTAs process at the same time, so three of the four dY reads hit L2.  The class of a unit is rotated by the
// tile index: a function of the tile only (each tile still gets every class exactly once) that makes every CTA of the static
// round-robin cycle through all classes (they carry 1 / 2 / 2 / 4 taps) for grids of 148 CTAs and of 74 CTA pairs alike (max / mean
// load 1.01-1.03 on the layers it is used for; the host enables it only when dY exceeds what L2 keeps, conv_api.cu).
__device__ __forceinline__ void unit_decode(int t, int cls_units, int ncls, int interleave, int grid_units, int &cls, int &tt)
{
    if (ncls <= 1) { cls = 0; tt = t; return; }
    if (!interleave) { cls = t / cls_units; tt = t - cls * cls_units; return; }
    tt = t / ncls;
    cls = (t - tt * ncls + tt) % ncls;
}

int make_tmap_2d(CUtensorMap *tm, const void *base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int swizzle_bytes, int dtype_bf16);
int make_tmap_im2col(CUtensorMap *tm, const void *base, int C, int W, int H, int N, int64_t ld, int lower_w, int lower_h,
                     int upper_w, int upper_h, int chan_per_pixel, int pixels_per_col, int tstride, int swizzle_bytes,
                     int dtype_bf16);
int launch_conv_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const ConvKParams &p, cudaStream_t st);
// experimental CTA-pair (cta_group::2) variant, conv_pair.cu; tmB must be encoded with a box of block_n / 2 rows
bool conv_pair_eligible(const ConvKParams &p);
int launch_conv_pair(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const ConvKParams &p, cudaStream_t st);

}  // namespace cy4
