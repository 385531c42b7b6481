// nms.cu -- SURVEY section 8 (f1): rotated IoU one-vs-many, post_processing_v2 (confidence filter + sort +
// rotated NMS with confidence-weighted merging) and the true-positive matching of evaluate.py, on the device.
// Reference: src/utils/evaluation_utils.py:152-210 and :322-357.  Compiled with --fmad=false: the reference
// evaluates every product / sum of the box arithmetic as a separate fp32 torch op.
//
// Geometry = the fp64 convex clipper of rbox.cuh on fp32 corners (shapely/GEOS in the reference), then
//   iou = reciprocal((s_area + t_area) - inter + 1e-16) * inter      (Tensor.__rtruediv__, all fp32)
#include <cfloat>
#include <climits>

#include "common.cuh"
#include "rbox.cuh"

namespace cy4 {

struct Quad { float x[4], y[4], area; };

__device__ __forceinline__ void make_quad(const float b[6], Quad &q)
{
    float cs, sn;
    box_corners(b[0], b[1], b[2], b[3], atan2_cr(b[4], b[5]), q.x, q.y, cs, sn);
    q.area = b[2] * b[3];
}

// evaluation_utils.py:203-207
__device__ __forceinline__ float eval_iou(const Quad &s, const Quad &m)
{
    const float inter = (float)convex_inter64(s.x, s.y, m.x, m.y);
    const float den = ((s.area + m.area) - inter) + 1e-16f;
    return (1.0f / den) * inter;
}

// ious[i, j] = IoU(a[i], b[j]);  n == 1 is iou_rotated_single_vs_multi_boxes_cpu
__global__ void __launch_bounds__(128)
rbox_iou_matrix_kernel(const float *__restrict__ a6, int64_t n, const float *__restrict__ b6, int64_t m, float *__restrict__ ious)
{
    const int64_t total = n * m;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = k / m, j = k - i * m;
        float A[6], Bx[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { A[c] = __ldg(a6 + i * 6 + c); Bx[c] = __ldg(b6 + j * 6 + c); }
        Quad qa, qb;
        make_quad(A, qa); make_quad(Bx, qb);
        ious[k] = eval_iou(qa, qb);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Anchor k-means distance matrix: utils/find_anchors.py:53-59 compute_iou for every (box, cluster)
// pair.  Boxes and clusters are (w, l, yaw) centred at the origin.  Corners: kitti_bev_utils.get_corners evaluates them in
// float64 and stores float32 (:96-120); polygons, areas and the intersection are float64 (shapely);
// iou = inter / (box_area + cluster_area - inter + 1e-12) in float64, returned as float32 (np.array(..., dtype=np.float32)).
__device__ __forceinline__ void kmeans_corners(double w, double l, double yaw, float cx[4], float cy[4])
{
    double sn, cs;
    sincos(yaw, &sn, &cs);
    cx[0] = (float)(0.0 - w / 2 * cs - l / 2 * sn);  cy[0] = (float)(0.0 - w / 2 * sn + l / 2 * cs);   // front left
    cx[1] = (float)(0.0 - w / 2 * cs + l / 2 * sn);  cy[1] = (float)(0.0 - w / 2 * sn - l / 2 * cs);   // rear left
    cx[2] = (float)(0.0 + w / 2 * cs + l / 2 * sn);  cy[2] = (float)(0.0 + w / 2 * sn - l / 2 * cs);   // rear right
    cx[3] = (float)(0.0 + w / 2 * cs - l / 2 * sn);  cy[3] = (float)(0.0 + w / 2 * sn + l / 2 * cs);   // front right
}
__device__ __forceinline__ double quad_area64(const float x[4], const float y[4])
{
    double s = 0.0;
    for (int i = 0; i < 4; ++i) { const int j = (i + 1) & 3; s += (double)x[i] * y[j] - (double)y[i] * x[j]; }
    return fabs(s) * 0.5;
}
__global__ void __launch_bounds__(128)
kmeans_iou_kernel(const double *__restrict__ boxes3, int64_t n, const double *__restrict__ clusters3, int k, float *__restrict__ ious)
{
    const int64_t total = n * k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / k; const int j = (int)(t - i * k);
        float ax[4], ay[4], bx[4], by[4];
        kmeans_corners(boxes3[3 * i], boxes3[3 * i + 1], boxes3[3 * i + 2], ax, ay);
        kmeans_corners(clusters3[3 * j], clusters3[3 * j + 1], clusters3[3 * j + 2], bx, by);
        const double inter = convex_inter64(ax, ay, bx, by);
        ious[t] = (float)(inter / (quad_area64(ax, ay) + quad_area64(bx, by) - inter + 1e-12));
    }
}

// ---------------------------------------------------------------------------------------------------------
// post_processing_v2: one CTA per image.
//   1. rows with conf >= conf_thresh are compacted (any order), score = conf * max(cls)         (:333-339)
//   2. sorted by score, descending; equal scores keep the lower row index first (the reference's
//      unstable argsort leaves ties undefined)                                                   (:341)
//   3. greedy loop over the sorted list: IoU(head, every live box) in parallel, `invalid` = IoU > nms_thresh
//      and same class; head box <- sum(conf * box) / sum(conf) over the invalid set, summed in list order
//      in fp32; the invalid set (and always the head) leaves the list                            (:346-355)
constexpr int kNmsThreads = 512;
constexpr int kNmsMaxCand = 4096;

struct NmsWs {                 // per image, in global memory (L2 resident)
    float row[kNmsMaxCand][9];  // x y w l im re conf cls_conf cls_pred, sorted order
    Quad quad[kNmsMaxCand];
};

__global__ void __launch_bounds__(kNmsThreads)
nms_v2_kernel(const float *__restrict__ pred, int N, int nC, float conf_thresh, float nms_thresh, int max_cand,
              float *__restrict__ out9, int32_t *__restrict__ counts, int32_t *__restrict__ found, NmsWs *__restrict__ ws_all)
{
    __shared__ float s_score[kNmsMaxCand];
    __shared__ int s_idx[kNmsMaxCand];
    __shared__ uint8_t s_alive[kNmsMaxCand], s_inv[kNmsMaxCand];
    __shared__ float s_sum[7];
    __shared__ int s_cnt, s_head, s_keep;
    const int tid = threadIdx.x, img = blockIdx.x, row_len = 7 + nC;
    const float *P = pred + (int64_t)img * N * row_len;
    NmsWs &ws = ws_all[img];
    float *out = out9 + (int64_t)img * max_cand * 9;
    if (tid == 0) s_cnt = 0;
    __syncthreads();

    // 1. filter + score
    for (int i = tid; i < N; i += kNmsThreads) {
        const float conf = __ldg(P + (int64_t)i * row_len + 6);
        if (conf >= conf_thresh) {
            const int pos = atomicAdd(&s_cnt, 1);
            if (pos < max_cand) {
                float best = __ldg(P + (int64_t)i * row_len + 7);
                for (int c = 1; c < nC; ++c) best = fmaxf(best, __ldg(P + (int64_t)i * row_len + 7 + c));
                s_score[pos] = conf * best;
                s_idx[pos] = i;
            }
        }
    }
    __syncthreads();
    const int total = s_cnt;
    const int cnt = min(total, max_cand);
    if (tid == 0) found[img] = total;          // > max_cand: the host wrapper raises (the list was truncated)
    if (cnt == 0) { if (tid == 0) counts[img] = 0; return; }

    // 2. bitonic sort, (score desc, row asc)
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + tid; i < n2; i += kNmsThreads) { s_score[i] = -FLT_MAX; s_idx[i] = INT_MAX; }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += kNmsThreads) {
                const int l = i ^ j;
                if (l > i) {
                    const float sa = s_score[i], sb = s_score[l];
                    const int ia = s_idx[i], ib = s_idx[l];
                    const bool a_first = sa > sb || (sa == sb && ia < ib);      // a belongs before b in the final order
                    const bool up = (i & k) == 0;
                    if (up ? !a_first : a_first) { s_score[i] = sb; s_score[l] = sa; s_idx[i] = ib; s_idx[l] = ia; }
                }
            }
            __syncthreads();
        }

    // gather the sorted rows, class argmax (first maximum), corners
    for (int c = tid; c < cnt; c += kNmsThreads) {
        const float *r = P + (int64_t)s_idx[c] * row_len;
        float b[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { b[k] = __ldg(r + k); ws.row[c][k] = b[k]; }
        ws.row[c][6] = __ldg(r + 6);
        float best = __ldg(r + 7); int arg = 0;
        for (int k = 1; k < nC; ++k) { const float v = __ldg(r + 7 + k); if (v > best) { best = v; arg = k; } }
        ws.row[c][7] = best; ws.row[c][8] = (float)arg;
        make_quad(b, ws.quad[c]);
        s_alive[c] = 1;
    }
    if (tid == 0) { s_head = 0; s_keep = 0; }
    __syncthreads();

    // 3. greedy NMS with merging
    while (true) {
        const int head = s_head;
        if (head >= cnt) break;
        const Quad hq = ws.quad[head];
        const float hcls = ws.row[head][8];
        for (int j = head + tid; j < cnt; j += kNmsThreads) {
            uint8_t inv = 0;
            if (s_alive[j]) inv = (eval_iou(hq, ws.quad[j]) > nms_thresh) && (ws.row[j][8] == hcls);
            s_inv[j] = inv;
        }
        __syncthreads();
        if (tid < 7) {            // column sums in list order: (weights * boxes).sum(0) and weights.sum()
            float acc = 0.f;
            for (int j = head; j < cnt; ++j)
                if (s_inv[j]) { const float w = ws.row[j][6]; acc = acc + (tid < 6 ? w * ws.row[j][tid] : w); }
            s_sum[tid] = acc;
        }
        __syncthreads();
        const int keep = s_keep;
        if (tid < 6) out[keep * 9 + tid] = s_sum[tid] / s_sum[6];
        else if (tid < 9) out[keep * 9 + tid] = ws.row[head][tid];
        for (int j = head + tid; j < cnt; j += kNmsThreads)
            if (s_inv[j]) s_alive[j] = 0;
        __syncthreads();
        if (tid == 0) {
            s_alive[head] = 0;     // (the reference would spin forever on a head that does not suppress itself)
            int h = head + 1;
            while (h < cnt && !s_alive[h]) ++h;
            s_head = h; s_keep = keep + 1;
        }
        __syncthreads();
    }
    if (tid == 0) counts[img] = s_keep;
}

// ---------------------------------------------------------------------------------------------------------
// get_batch_statistics_rotated_bbox (:152-183): one warp per image, detections in list order.
//   annotations = targets[targets[:,0] == img][:, 1:]  (cls, x, y, w, l, im, re; x..l already in pixels)
//   for each detection: stop when every annotation is matched; skip if its class is not among the annotation
//   classes; IoU vs all annotations, first maximum; true positive if IoU >= thresh and that annotation is free.
constexpr int kMaxAnn = 256;

__global__ void __launch_bounds__(32)
eval_match_kernel(const float *__restrict__ dets9, const int32_t *__restrict__ counts, int max_det,
                  const float *__restrict__ targets8, int64_t nT, float iou_thresh, uint8_t *__restrict__ tp, int32_t *__restrict__ status)
{
    __shared__ int s_ann[kMaxAnn];
    __shared__ uint8_t s_used[kMaxAnn];
    const int img = blockIdx.x, lane = threadIdx.x;
    const int nd = counts[img];
    const float *D = dets9 + (int64_t)img * max_det * 9;
    uint8_t *TP = tp + (int64_t)img * max_det;
    for (int i = lane; i < max_det; i += 32) TP[i] = 0;
    // ordered list of this image's annotations
    int na = 0;
    for (int64_t base = 0; base < nT; base += 32) {
        const int64_t t = base + lane;
        const bool mine = t < nT && targets8[t * 8] == (float)img;
        const unsigned bal = __ballot_sync(0xffffffffu, mine);
        if (mine) { const int pos = na + __popc(bal & ((1u << lane) - 1)); if (pos < kMaxAnn) s_ann[pos] = (int)t; }
        na += __popc(bal);
    }
    if (lane == 0) status[img] = na;             // > kMaxAnn: the host wrapper raises
    na = min(na, kMaxAnn);
    for (int i = lane; i < na; i += 32) s_used[i] = 0;
    __syncwarp();
    int matched = 0;
    for (int p = 0; p < nd && na > 0; ++p) {
        if (matched == na) break;
        float b[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) b[c] = D[p * 9 + c];
        const float label = D[p * 9 + 8];
        bool any = false;
        float best = -1.f; int arg = INT_MAX;
        Quad pq; make_quad(b, pq);
        for (int a = lane; a < na; a += 32) {
            const float *T = targets8 + (int64_t)s_ann[a] * 8;
            any |= (T[1] == label);
            float tb[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) tb[c] = T[2 + c];
            Quad tq; make_quad(tb, tq);
            const float v = eval_iou(pq, tq);
            if (v > best) { best = v; arg = a; }          // ascending a per lane: keeps the first maximum
        }
        any = __any_sync(0xffffffffu, any);
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        bool hit = any && best >= iou_thresh && arg < na && !s_used[arg];
        hit = __shfl_sync(0xffffffffu, hit, 0);           // every lane has read s_used before lane 0 updates it
        if (hit) {
            if (lane == 0) { TP[p] = 1; s_used[arg] = 1; }
            ++matched;
        }
        __syncwarp();
    }
}

}  // namespace cy4

using namespace cy4;

extern "C" {

int cy4_rbox_iou_matrix(const float *a6, int64_t n, const float *b6, int64_t m, float *ious, void *stream)
{
    CY4_CHECK_ARG(n >= 0 && m >= 0, "cy4_rbox_iou_matrix: negative size");
    if (n == 0 || m == 0) return 0;
    CY4_CHECK_ARG(a6 && b6 && ious, "cy4_rbox_iou_matrix: null pointer");
    const int64_t total = n * m;
    const int grid = (int)std::min<int64_t>((total + 127) / 128, (int64_t)sm_count() * 16);
    rbox_iou_matrix_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a6, n, b6, m, ious);
    return cy4_launch_status("cy4_rbox_iou_matrix");
}

int cy4_kmeans_iou(const double *boxes3, int64_t n, const double *clusters3, int k, float *ious, void *stream)
{
    CY4_CHECK_ARG(n >= 0 && k >= 0, "cy4_kmeans_iou: negative size");
    if (n == 0 || k == 0) return 0;
    CY4_CHECK_ARG(boxes3 && clusters3 && ious, "cy4_kmeans_iou: null pointer");
    const int64_t total = n * k;
    const int grid = (int)std::min<int64_t>((total + 127) / 128, (int64_t)sm_count() * 16);
    kmeans_iou_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(boxes3, n, clusters3, k, ious);
    return cy4_launch_status("cy4_kmeans_iou");
}

int cy4_nms_max_candidates(void) { return kNmsMaxCand; }

size_t cy4_nms_workspace_bytes(int B) { return B > 0 ? (size_t)B * sizeof(NmsWs) : 0; }

int cy4_nms_rotated_v2(const float *pred, int B, int N, int nC, float conf_thresh, float nms_thresh, float *out9,
                       int32_t *counts, int32_t *found, void *workspace, void *stream)
{
    CY4_CHECK_ARG(B >= 0 && N >= 0 && nC >= 1, "cy4_nms_rotated_v2: bad sizes");
    if (B == 0) return 0;
    CY4_CHECK_ARG(pred && out9 && counts && found && workspace, "cy4_nms_rotated_v2: null pointer");
    nms_v2_kernel<<<B, kNmsThreads, 0, (cudaStream_t)stream>>>(pred, N, nC, conf_thresh, nms_thresh, kNmsMaxCand, out9, counts, found,
                                                               (NmsWs *)workspace);
    return cy4_launch_status("cy4_nms_rotated_v2");
}

int cy4_eval_match(const float *dets9, const int32_t *counts, int B, int max_det, const float *targets8, int64_t nT,
                   float iou_thresh, uint8_t *tp, int32_t *n_ann, void *stream)
{
    CY4_CHECK_ARG(B >= 0 && max_det >= 0 && nT >= 0, "cy4_eval_match: bad sizes");
    if (B == 0) return 0;
    CY4_CHECK_ARG(dets9 && counts && tp && n_ann && (targets8 || nT == 0), "cy4_eval_match: null pointer");
    eval_match_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(dets9, counts, max_det, targets8, nT, iou_thresh, tp, n_ann);
    return cy4_launch_status("cy4_eval_match");
}

int cy4_eval_max_annotations(void) { return kMaxAnn; }

}  // extern "C"
