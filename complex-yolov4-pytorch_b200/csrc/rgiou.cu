// rgiou.cu -- rotated-box IoU / GIoU kernels (pairs, anchors, small geometry helpers).
// Compiled with --fmad=false (see rbox.cuh).  Reference: src/utils/iou_rotated_boxes_utils.py,
// src/utils/cal_intersection_rotated_boxes.py.
#include "common.cuh"
#include "rbox.cuh"

namespace cy4 {

constexpr int kPairBlock = 128;

// One thread per pair.  The [n,6] rows are staged through shared memory so that global loads and
// the gradient stores are fully coalesced (a warp reads 32*6 consecutive floats).
template <bool GRAD>
__global__ void __launch_bounds__(kPairBlock)
rgiou_pairs_kernel(const float *__restrict__ pred6, const float *__restrict__ tgt6, int64_t n, int giou,
                   float *__restrict__ iou_out, float *__restrict__ term_out,
                   const float *__restrict__ gterm, float *__restrict__ gpred6)
{
    __shared__ PolySmem<kPairBlock> sm;
    __shared__ float stage[2][kPairBlock * 6];
    const int tid = threadIdx.x;
    for (int64_t base = (int64_t)blockIdx.x * kPairBlock; base < n; base += (int64_t)gridDim.x * kPairBlock) {
        const int64_t cnt = min((int64_t)kPairBlock, n - base);
        const int nflt = (int)cnt * 6;
        for (int i = tid; i < nflt; i += kPairBlock) {
            stage[0][i] = __ldg(pred6 + base * 6 + i);
            stage[1][i] = __ldg(tgt6 + base * 6 + i);
        }
        __syncthreads();
        float P[6], T[6], g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float iou = 0.f, term = 0.f;
        const bool live = tid < cnt;
        if (live) {
#pragma unroll
            for (int c = 0; c < 6; ++c) { P[c] = stage[0][tid * 6 + c]; T[c] = stage[1][tid * 6 + c]; }
            rgiou_pair<kPairBlock, GRAD>(P, T, giou != 0, sm, tid, iou, term, g);
            iou_out[base + tid] = iou;
            term_out[base + tid] = term;
        }
        if (GRAD) {
            __syncthreads();
            if (live) {
                const float go = gterm ? gterm[base + tid] : 1.0f;
#pragma unroll
                for (int c = 0; c < 6; ++c) stage[0][tid * 6 + c] = g[c] * go;
            }
            __syncthreads();
            for (int i = tid; i < nflt; i += kPairBlock) gpred6[base * 6 + i] = stage[0][i];
        }
        __syncthreads();
    }
}

// Reference accumulation `giou_loss += term` (iou_rotated_boxes_utils.py:133) is sequential fp32;
// kept exactly for n <= 4096 (one thread), a fixed-shape block tree beyond that.
__global__ void sum_f32_kernel(const float *__restrict__ v, int64_t n, float *__restrict__ out)
{
    __shared__ float red[1024];
    if (n <= 4096) {
        if (threadIdx.x == 0) {
            float a = 0.f;
            for (int64_t i = 0; i < n; ++i) a = a + v[i];
            out[0] = a;
        }
        return;
    }
    float a = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += v[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

__global__ void corners_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ w,
                               const float *__restrict__ l, const float *__restrict__ yaw, int64_t n,
                               float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float cx[4], cy[4], cs, sn;
    box_corners(x[i], y[i], w[i], l[i], yaw[i], cx, cy, cs, sn);
#pragma unroll
    for (int k = 0; k < 4; ++k) { out[i * 8 + 2 * k] = cx[k]; out[i * 8 + 2 * k + 1] = cy[k]; }
}

__global__ void __launch_bounds__(kPairBlock)
quad_inter_kernel(const float *__restrict__ r1, const float *__restrict__ r2, int64_t n, float *__restrict__ area)
{
    __shared__ PolySmem<kPairBlock> sm;
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * kPairBlock + tid;
    if (i >= n) return;
    float px[4], py[4], tx[4], ty[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        px[k] = r1[i * 8 + 2 * k]; py[k] = r1[i * 8 + 2 * k + 1];
        tx[k] = r2[i * 8 + 2 * k]; ty[k] = r2[i * 8 + 2 * k + 1];
    }
    int buf;
    const int m = clip_ref<kPairBlock>(px, py, tx, ty, sm, tid, buf);
    float a = 0.f;
    if (m > 2) {
        const float s = torch_small_sum(m, [&](int k) {
            const int j = (k + 1 == m) ? 0 : k + 1;
            return sm.x[buf][k][tid] * sm.y[buf][j][tid] - sm.y[buf][k][tid] * sm.x[buf][j][tid];
        });
        a = fabsf(s) * 0.5f;
    }
    area[i] = a;
}

__global__ void poly_area_kernel(const float *__restrict__ pts, int k, float *__restrict__ area)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float s;
    if (k <= 8) {
        s = torch_small_sum(k, [&](int i) {
            const int j = (i + 1 == k) ? 0 : i + 1;
            return pts[2 * i] * pts[2 * j + 1] - pts[2 * i + 1] * pts[2 * j];
        });
    } else {
        s = 0.f;
        for (int i = 0; i < k; ++i) {
            const int j = (i + 1 == k) ? 0 : i + 1;
            s = s + (pts[2 * i] * pts[2 * j + 1] - pts[2 * i + 1] * pts[2 * j]);
        }
    }
    area[0] = fabsf(s) * 0.5f;
}

__global__ void anchor_iou_kernel(const float *__restrict__ anchors4, int nA, const float *__restrict__ tgt4,
                                  int64_t nT, float *__restrict__ ious)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nT * nA) return;
    const int a = (int)(i / nT);
    const int64_t t = i % nT;
    float A[4], T[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { A[k] = anchors4[a * 4 + k]; T[k] = tgt4[t * 4 + k]; }
    ious[i] = anchor_target_iou(A, T);
}

}  // namespace cy4

using namespace cy4;

extern "C" {

int cy4_rgiou_pairs(const float *pred6, const float *tgt6, int64_t n, uint32_t flags, float *iou, float *term,
                    const float *gterm, float *gpred6, void *stream)
{
    CY4_CHECK_ARG(n >= 0, "cy4_rgiou_pairs: n < 0");
    if (n == 0) return 0;
    CY4_CHECK_ARG(pred6 && tgt6 && iou && term, "cy4_rgiou_pairs: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t blocks_needed = (n + kPairBlock - 1) / kPairBlock;
    const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)sm_count() * 16);
    const int giou = (flags & CY4_F_GIOU) ? 1 : 0;
    if (gpred6)
        rgiou_pairs_kernel<true><<<grid, kPairBlock, 0, st>>>(pred6, tgt6, n, giou, iou, term, gterm, gpred6);
    else
        rgiou_pairs_kernel<false><<<grid, kPairBlock, 0, st>>>(pred6, tgt6, n, giou, iou, term, nullptr, nullptr);
    return cy4_launch_status("cy4_rgiou_pairs");
}

int cy4_sum_f32_seq(const float *term, int64_t n, float *out, void *stream)
{
    CY4_CHECK_ARG(n >= 0 && out && (term || n == 0), "cy4_sum_f32_seq: bad argument");
    sum_f32_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(term, n, out);
    return cy4_launch_status("cy4_sum_f32_seq");
}

int cy4_corners(const float *x, const float *y, const float *w, const float *l, const float *yaw, int64_t n,
                float *corners, void *stream)
{
    CY4_CHECK_ARG(n >= 0, "cy4_corners: n < 0");
    if (n == 0) return 0;
    CY4_CHECK_ARG(x && y && w && l && yaw && corners, "cy4_corners: null pointer");
    corners_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, w, l, yaw, n, corners);
    return cy4_launch_status("cy4_corners");
}

int cy4_quad_intersection_area(const float *rect1, const float *rect2, int64_t n, float *area, void *stream)
{
    CY4_CHECK_ARG(n >= 0, "cy4_quad_intersection_area: n < 0");
    if (n == 0) return 0;
    CY4_CHECK_ARG(rect1 && rect2 && area, "cy4_quad_intersection_area: null pointer");
    quad_inter_kernel<<<(unsigned)((n + kPairBlock - 1) / kPairBlock), kPairBlock, 0, (cudaStream_t)stream>>>(rect1, rect2, n, area);
    return cy4_launch_status("cy4_quad_intersection_area");
}

int cy4_poly_area(const float *pts, int k, float *area, void *stream)
{
    CY4_CHECK_ARG(pts && area && k >= 0 && k <= 16, "cy4_poly_area: bad argument (k <= 16)");
    poly_area_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pts, k, area);
    return cy4_launch_status("cy4_poly_area");
}

int cy4_anchor_iou(const float *anchors4, int nA, const float *tgt4, int64_t nT, float *ious, void *stream)
{
    CY4_CHECK_ARG(nA >= 0 && nT >= 0, "cy4_anchor_iou: negative size");
    if (nA == 0 || nT == 0) return 0;
    CY4_CHECK_ARG(anchors4 && tgt4 && ious, "cy4_anchor_iou: null pointer");
    const int64_t tot = nT * nA;
    anchor_iou_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, (cudaStream_t)stream>>>(anchors4, nA, tgt4, nT, ious);
    return cy4_launch_status("cy4_anchor_iou");
}

}  // extern "C"
