// yolo_head.cu -- fused YOLO head for the training step: decode, rotated target assignment,
// the 9 loss terms, the 18 metrics and the gradient w.r.t. the raw head tensor, without a single
// host synchronisation.  Reference: src/models/yolo_layer.py:53-67 (grid offsets), :69-142
// (build_targets), :144-253 (forward).  Compiled with --fmad=false (fp32 op order of the reference).
//
// Data flow (all on one stream):
//   memset(cellmap, clsbits)
//   yolo_targets_kernel   1 thread / target : anchor IoU (fp64 clip), best anchor, cell, regression
//                                             targets, GIoU(+grad) of the cell's pred box; scatters
//                                             cellmap (atomicMax = last writer wins) and
//                                             the multi-hot class bits (atomicOr)
//   yolo_dense_kernel     1 thread / cell   : decode -> output row; 19 partial sums per block
//   yolo_finalize_kernel  1 block           : fixed-order reduction, loss, metrics, normalisers
//   yolo_dense_bwd_kernel 1 thread / cell   : d loss / d raw head (fully overwrites dpred)
//   yolo_targets_bwd_kernel 1 thread/target : adds the GIoU gradient of every target (duplicates too)
#include "common.cuh"
#include "rbox.cuh"

namespace cy4 {

constexpr int kTgtBlock = 128;
constexpr int kDenseBlock = 256;
constexpr int kNAcc = 19;
constexpr int kMaxDenseGrid = 1184;   // 148 SMs x 8 resident 256-thread blocks

enum Acc { A_NOBJ, A_NNOOBJ, A_SX, A_SY, A_SW, A_SH, A_SIM, A_SRE, A_SIMRE, A_BCE_OBJ, A_BCE_NOOBJ, A_BCE_CLS,
           A_CLSACC, A_CONF_OBJ, A_CONF_NOOBJ, A_CONF50, A_IOU50, A_IOU75, A_IOUSUM };

struct TargetRec {
    int32_t b, a, gj, gi, label, valid;
    float tx, ty, tw, th, tim, tre;      // regression targets (yolo_layer.py:122-129)
    float iou, term, clsmatch;           // pred<->target IoU, GIoU summand, argmax(pred_cls)==label
    float gbox[6];                       // d term / d pred_box (x, y, w, l, im, re)
    float gx, gy, gw, gh, gim, gre;      // target box in grid units
};

struct Final {
    float nobj, nnoobj, nT, pad;
    float loss, giou_loss;
};

struct Workspace {
    int32_t *cellmap;      // [B*nA*G*G]  0 plain noobj, 1 ignored, t+2 = obj owned by target t
    uint32_t *clsbits;     // [B*nA*G*G]  multi-hot class bits of obj cells
    TargetRec *rec;        // [nT]
    float *partials;       // [kMaxDenseGrid * kNAcc]
    Final *fin;
    float *anchor_ious;    // [nA * nT]
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t carve(const cy4_yolo_desc *d, int64_t nT, void *base, Workspace *ws)
{
    const size_t cells = (size_t)d->B * d->nA * d->G * d->G;
    size_t off = 0;
    char *p = (char *)base;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return p ? p + o : nullptr; };
    void *a0 = take(cells * 4), *a1 = take(cells * 4), *a2 = take((size_t)std::max<int64_t>(nT, 1) * sizeof(TargetRec));
    void *a3 = take((size_t)kMaxDenseGrid * kNAcc * 4), *a4 = take(sizeof(Final));
    void *a5 = take((size_t)std::max<int64_t>(nT, 1) * d->nA * 4);
    if (ws) { ws->cellmap = (int32_t *)a0; ws->clsbits = (uint32_t *)a1; ws->rec = (TargetRec *)a2;
              ws->partials = (float *)a3; ws->fin = (Final *)a4; ws->anchor_ious = (float *)a5; }
    return off;
}

struct HeadView {
    const float *p; int64_t sB, sC, sH, sW; int nA, nC, G;
    __device__ __forceinline__ float at(int b, int a, int c, int gj, int gi) const
    { return __ldg(p + b * sB + (int64_t)(a * (nC + 7) + c) * sC + gj * sH + gi * sW); }
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float bce_log(float p) { return fmaxf(logf(p), -100.0f); }   // torch clamps log at -100

// ------------------------------------------------------------------------------------------------
// 1 thread per target.  FROM_RAW: pred box decoded from the raw head tensor; otherwise read from
// the caller's pred_boxes [B,nA,G,G,6] / pred_cls [B,nA,G,G,nC] (build_targets API).
template <bool FROM_RAW>
__global__ void __launch_bounds__(kTgtBlock)
yolo_targets_kernel(cy4_yolo_desc d, HeadView hv, const float *__restrict__ pred_boxes, const float *__restrict__ pred_cls,
                    const float *__restrict__ anchors4, const float *__restrict__ targets8, int64_t nT,
                    Workspace ws, int32_t *__restrict__ status)
{
    __shared__ PolySmem<kTgtBlock> sm;
    const int tid = threadIdx.x;
    const int64_t t = (int64_t)blockIdx.x * kTgtBlock + tid;
    if (t >= nT) return;
    const float *tg = targets8 + t * 8;
    const float Gf = (float)d.G;
    TargetRec r;
    r.b = (int32_t)(int64_t)tg[0];                 // .long() truncation (yolo_layer.py:96)
    r.label = (int32_t)(int64_t)tg[1];
    r.gx = tg[2] * Gf; r.gy = tg[3] * Gf; r.gw = tg[4] * Gf; r.gh = tg[5] * Gf;   // :97
    r.gim = tg[6]; r.gre = tg[7];
    r.gi = (int32_t)(int64_t)r.gx; r.gj = (int32_t)(int64_t)r.gy;               // :112
    // anchors with best IoU (:103-107); ties -> first index
    float best = -1.0f; int best_n = 0;
    const float T4[4] = {r.gw, r.gh, r.gim, r.gre};
    for (int a = 0; a < d.nA; ++a) {
        const float A4[4] = {anchors4[a * 4], anchors4[a * 4 + 1], anchors4[a * 4 + 2], anchors4[a * 4 + 3]};
        const float v = anchor_target_iou(A4, T4);
        ws.anchor_ious[(int64_t)a * nT + t] = v;
        if (v > best) { best = v; best_n = a; }
    }
    r.a = best_n;
    r.valid = (r.b >= 0 && r.b < d.B && r.gi >= 0 && r.gi < d.G && r.gj >= 0 && r.gj < d.G &&
               r.label >= 0 && r.label < d.nC) ? 1 : 0;
    r.tx = r.gx - floorf(r.gx); r.ty = r.gy - floorf(r.gy);                      // :122-123
    r.tw = logf(r.gw / anchors4[best_n * 4] + 1e-16f);                          // :125-126
    r.th = logf(r.gh / anchors4[best_n * 4 + 1] + 1e-16f);
    r.tim = r.gim; r.tre = r.gre;
    r.iou = 0.f; r.term = 0.f; r.clsmatch = 0.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) r.gbox[c] = 0.f;
    if (!r.valid) {
        atomicOr(status, 1);
        ws.rec[t] = r;
        return;
    }
    const int64_t cell = (((int64_t)r.b * d.nA + r.a) * d.G + r.gj) * d.G + r.gi;
    atomicMax(ws.cellmap + cell, (int32_t)t + 2);                                // :114-115, last writer wins
    atomicOr(ws.clsbits + cell, 1u << r.label);                                  // :132 (multi-hot on duplicates)
    for (int a = 0; a < d.nA; ++a)                                               // :118-119 strict >
        if (ws.anchor_ious[(int64_t)a * nT + t] > d.ignore_thresh)
            atomicMax(ws.cellmap + ((((int64_t)r.b * d.nA + a) * d.G + r.gj) * d.G + r.gi), 1);
    // pred box of the assigned cell (:134) and class argmax (:133)
    float P[6];
    int amax = 0;
    if (FROM_RAW) {
        P[0] = sigmoidf_(hv.at(r.b, r.a, 0, r.gj, r.gi)) + (float)r.gi;
        P[1] = sigmoidf_(hv.at(r.b, r.a, 1, r.gj, r.gi)) + (float)r.gj;
        P[2] = fminf(expf(hv.at(r.b, r.a, 2, r.gj, r.gi)), 1e3f) * anchors4[r.a * 4];
        P[3] = fminf(expf(hv.at(r.b, r.a, 3, r.gj, r.gi)), 1e3f) * anchors4[r.a * 4 + 1];
        P[4] = hv.at(r.b, r.a, 4, r.gj, r.gi);
        P[5] = hv.at(r.b, r.a, 5, r.gj, r.gi);
        float bestc = sigmoidf_(hv.at(r.b, r.a, 7, r.gj, r.gi));
        for (int c = 1; c < d.nC; ++c) {
            const float v = sigmoidf_(hv.at(r.b, r.a, 7 + c, r.gj, r.gi));
            if (v > bestc) { bestc = v; amax = c; }
        }
    } else {
#pragma unroll
        for (int c = 0; c < 6; ++c) P[c] = pred_boxes[cell * 6 + c];
        float bestc = pred_cls[cell * d.nC];
        for (int c = 1; c < d.nC; ++c) {
            const float v = pred_cls[cell * d.nC + c];
            if (v > bestc) { bestc = v; amax = c; }
        }
    }
    r.clsmatch = (amax == r.label) ? 1.0f : 0.0f;
    const float Tb[6] = {r.gx, r.gy, r.gw, r.gh, r.gim, r.gre};
    rgiou_pair<kTgtBlock, true>(P, Tb, d.use_giou != 0, sm, tid, r.iou, r.term, r.gbox);
    ws.rec[t] = r;
}

// cell id -> (b, a, gj, gi).  a_fastest: (b, gj, gi, a) order for channel-last heads.
__device__ __forceinline__ void cell_coords(int64_t i, const cy4_yolo_desc &d, bool a_fastest, int &b, int &a, int &gj, int &gi)
{
    if (a_fastest) { a = (int)(i % d.nA); i /= d.nA; gi = (int)(i % d.G); i /= d.G; gj = (int)(i % d.G); b = (int)(i / d.G); }
    else { gi = (int)(i % d.G); i /= d.G; gj = (int)(i % d.G); i /= d.G; a = (int)(i % d.nA); b = (int)(i / d.nA); }
}

template <bool TRAIN>
__global__ void __launch_bounds__(kDenseBlock)
yolo_dense_kernel(cy4_yolo_desc d, HeadView hv, const float *__restrict__ anchors4, Workspace ws,
                  float *__restrict__ output, int a_fastest)
{
    const int64_t cells = (int64_t)d.B * d.nA * d.G * d.G;
    const float stride = d.img_size / (float)d.G;
    const int nO = 7 + d.nC;
    float acc[kNAcc];
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) acc[k] = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * kDenseBlock + threadIdx.x; i < cells; i += (int64_t)gridDim.x * kDenseBlock) {
        int b, a, gj, gi;
        cell_coords(i, d, a_fastest != 0, b, a, gj, gi);
        const float vx = hv.at(b, a, 0, gj, gi), vy = hv.at(b, a, 1, gj, gi), vw = hv.at(b, a, 2, gj, gi);
        const float vh = hv.at(b, a, 3, gj, gi), vim = hv.at(b, a, 4, gj, gi), vre = hv.at(b, a, 5, gj, gi);
        const float px = sigmoidf_(vx), py = sigmoidf_(vy), conf = sigmoidf_(hv.at(b, a, 6, gj, gi));
        if (output) {
            float *o = output + ((int64_t)b * d.nA * d.G * d.G + ((int64_t)a * d.G + gj) * d.G + gi) * nO;
            o[0] = (px + (float)gi) * stride;
            o[1] = (py + (float)gj) * stride;
            o[2] = (fminf(expf(vw), 1e3f) * anchors4[a * 4]) * stride;
            o[3] = (fminf(expf(vh), 1e3f) * anchors4[a * 4 + 1]) * stride;
            o[4] = vim; o[5] = vre; o[6] = conf;
            for (int c = 0; c < d.nC; ++c) o[7 + c] = sigmoidf_(hv.at(b, a, 7 + c, gj, gi));
        }
        if (TRAIN) {
            const int64_t cell = (((int64_t)b * d.nA + a) * d.G + gj) * d.G + gi;
            const int32_t st = ws.cellmap[cell];
            if (conf > 0.5f) acc[A_CONF50] += 1.f;
            if (st == 0) {
                acc[A_NNOOBJ] += 1.f;
                acc[A_BCE_NOOBJ] += -bce_log(1.0f - conf);
                acc[A_CONF_NOOBJ] += conf;
            } else if (st >= 2) {
                const TargetRec r = ws.rec[st - 2];
                const uint32_t bits = ws.clsbits[cell];
                acc[A_NOBJ] += 1.f;
                float e;
                e = px - r.tx; acc[A_SX] += e * e;
                e = py - r.ty; acc[A_SY] += e * e;
                e = vw - r.tw; acc[A_SW] += e * e;
                e = vh - r.th; acc[A_SH] += e * e;
                e = vim - r.tim; acc[A_SIM] += e * e;
                e = vre - r.tre; acc[A_SRE] += e * e;
                e = 1.0f - sqrtf(vim * vim + vre * vre); acc[A_SIMRE] += e * e;
                acc[A_BCE_OBJ] += -bce_log(conf);
                for (int c = 0; c < d.nC; ++c) {
                    const float pc = sigmoidf_(hv.at(b, a, 7 + c, gj, gi));
                    acc[A_BCE_CLS] += ((bits >> c) & 1u) ? -bce_log(pc) : -bce_log(1.0f - pc);
                }
                acc[A_CLSACC] += r.clsmatch;
                acc[A_CONF_OBJ] += conf;
                acc[A_IOUSUM] += r.iou;
                const float det = (conf > 0.5f ? 1.f : 0.f) * r.clsmatch;
                if (r.iou > 0.5f) acc[A_IOU50] += det;
                if (r.iou > 0.75f) acc[A_IOU75] += det;
            }
        }
    }
    if (TRAIN) {
        __shared__ float red[kDenseBlock / 32][kNAcc];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < kNAcc; ++k) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < kNAcc) {
            float v = 0.f;
            for (int w = 0; w < kDenseBlock / 32; ++w) v += red[w][threadIdx.x];
            ws.partials[(int64_t)blockIdx.x * kNAcc + threadIdx.x] = v;
        }
    }
}

__global__ void yolo_finalize_kernel(cy4_yolo_desc d, Workspace ws, int nblocks, int64_t nT,
                                     float *__restrict__ loss, float *__restrict__ metrics)
{
    __shared__ double S[kNAcc];
    __shared__ float gsum;
    if (threadIdx.x < kNAcc) {
        double v = 0.0;
        for (int b = 0; b < nblocks; ++b) v += (double)ws.partials[(int64_t)b * kNAcc + threadIdx.x];
        S[threadIdx.x] = v;
    }
    if (threadIdx.x == 32) {      // giou_loss += term over all targets, sequential fp32 (:133), then / nT (:137-138)
        float a = 0.f;
        for (int64_t t = 0; t < nT; ++t) if (ws.rec[t].valid) a = a + ws.rec[t].term;
        gsum = a;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const float nobj = (float)S[A_NOBJ], nnoobj = (float)S[A_NNOOBJ];
    const float giou_loss = nT > 0 ? gsum / (float)nT : 0.f;
    const float loss_x = (float)S[A_SX] / nobj, loss_y = (float)S[A_SY] / nobj;
    const float loss_w = (float)S[A_SW] / nobj, loss_h = (float)S[A_SH] / nobj;
    const float loss_im = (float)S[A_SIM] / nobj, loss_re = (float)S[A_SRE] / nobj;
    const float loss_imre = (float)S[A_SIMRE] / nobj;
    const float loss_eular = (loss_im + loss_re) + loss_imre;                       // :207
    const float l_conf_obj = (float)S[A_BCE_OBJ] / nobj;
    const float l_conf_noobj = (float)S[A_BCE_NOOBJ] / nnoobj;
    const float l_cls = (float)S[A_BCE_CLS] / (nobj * (float)d.nC);
    float loss_obj, total;
    if (d.use_giou) {                                                                // :213-215
        loss_obj = l_conf_obj + l_conf_noobj;
        total = ((giou_loss * 3.54f + loss_eular * 3.54f) + loss_obj * 64.3f) + l_cls * 37.4f;
    } else {                                                                         // :216-218
        loss_obj = 1.0f * l_conf_obj + 100.0f * l_conf_noobj;
        total = (((((loss_x + loss_y) + loss_w) + loss_h) + loss_eular) + loss_obj) + l_cls;
    }
    loss[0] = total;
    Final f; f.nobj = nobj; f.nnoobj = nnoobj; f.nT = (float)nT; f.pad = 0.f; f.loss = total; f.giou_loss = giou_loss;
    *ws.fin = f;
    if (metrics) {                                                                   // :232-251 order
        metrics[0] = total;
        metrics[1] = (float)S[A_IOUSUM] / nobj;
        metrics[2] = giou_loss;
        metrics[3] = loss_x; metrics[4] = loss_y; metrics[5] = loss_w; metrics[6] = loss_h;
        metrics[7] = loss_eular; metrics[8] = loss_im; metrics[9] = loss_re;
        metrics[10] = loss_obj; metrics[11] = l_cls;
        metrics[12] = 100.0f * ((float)S[A_CLSACC] / nobj);
        metrics[13] = (float)S[A_IOU50] / (nobj + 1e-16f);                           // recall50
        metrics[14] = (float)S[A_IOU75] / (nobj + 1e-16f);                           // recall75
        metrics[15] = (float)S[A_IOU50] / ((float)S[A_CONF50] + 1e-16f);             // precision
        metrics[16] = (float)S[A_CONF_OBJ] / nobj;
        metrics[17] = (float)S[A_CONF_NOOBJ] / nnoobj;
    }
}

// torch's binary_cross_entropy backward: (p - t) / max((1 - p) * p, 1e-12), then through the sigmoid.
__device__ __forceinline__ float bce_grad_raw(float p, float t)
{
    return ((p - t) / fmaxf((1.0f - p) * p, 1e-12f)) * (p * (1.0f - p));
}

__global__ void __launch_bounds__(kDenseBlock)
yolo_dense_bwd_kernel(cy4_yolo_desc d, HeadView hv, Workspace ws, const float *__restrict__ gloss,
                      float *__restrict__ dpred, int64_t dsB, int64_t dsC, int64_t dsH, int64_t dsW, int a_fastest)
{
    const int64_t cells = (int64_t)d.B * d.nA * d.G * d.G;
    const float go = gloss[0];
    const Final f = *ws.fin;
    const int nO = 7 + d.nC;
    const float w_xy = d.use_giou ? 0.f : 1.f;
    const float w_e = d.use_giou ? 3.54f : 1.f;
    const float w_obj = d.use_giou ? 64.3f : 1.f;
    const float w_noobj = d.use_giou ? 64.3f : 100.f;
    const float w_cls = d.use_giou ? 37.4f : 1.f;
    for (int64_t i = (int64_t)blockIdx.x * kDenseBlock + threadIdx.x; i < cells; i += (int64_t)gridDim.x * kDenseBlock) {
        int b, a, gj, gi;
        cell_coords(i, d, a_fastest != 0, b, a, gj, gi);
        const int64_t cell = (((int64_t)b * d.nA + a) * d.G + gj) * d.G + gi;
        const int32_t st = ws.cellmap[cell];
        float *o = dpred + b * dsB + (int64_t)(a * nO) * dsC + gj * dsH + gi * dsW;
        if (st == 0) {
            const float conf = sigmoidf_(hv.at(b, a, 6, gj, gi));
            for (int c = 0; c < nO; ++c) o[c * dsC] = 0.f;
            o[6 * dsC] = go * (w_noobj / f.nnoobj) * bce_grad_raw(conf, 0.f);
        } else if (st == 1) {
            for (int c = 0; c < nO; ++c) o[c * dsC] = 0.f;
        } else {
            const TargetRec r = ws.rec[st - 2];
            const uint32_t bits = ws.clsbits[cell];
            const float vx = hv.at(b, a, 0, gj, gi), vy = hv.at(b, a, 1, gj, gi), vw = hv.at(b, a, 2, gj, gi);
            const float vh = hv.at(b, a, 3, gj, gi), vim = hv.at(b, a, 4, gj, gi), vre = hv.at(b, a, 5, gj, gi);
            const float px = sigmoidf_(vx), py = sigmoidf_(vy), conf = sigmoidf_(hv.at(b, a, 6, gj, gi));
            const float inv = 1.0f / f.nobj;
            o[0 * dsC] = go * w_xy * 2.0f * (px - r.tx) * inv * (px * (1.0f - px));
            o[1 * dsC] = go * w_xy * 2.0f * (py - r.ty) * inv * (py * (1.0f - py));
            o[2 * dsC] = go * w_xy * 2.0f * (vw - r.tw) * inv;
            o[3 * dsC] = go * w_xy * 2.0f * (vh - r.th) * inv;
            const float rr = sqrtf(vim * vim + vre * vre);
            const float k = -2.0f * (1.0f - rr) / rr;        // d (1-r)^2 / d im = k * im
            o[4 * dsC] = go * w_e * inv * (2.0f * (vim - r.tim) + k * vim);
            o[5 * dsC] = go * w_e * inv * (2.0f * (vre - r.tre) + k * vre);
            o[6 * dsC] = go * (w_obj * inv) * bce_grad_raw(conf, 1.0f);
            const float invc = inv / (float)d.nC;
            for (int c = 0; c < d.nC; ++c) {
                const float pc = sigmoidf_(hv.at(b, a, 7 + c, gj, gi));
                o[(7 + c) * dsC] = go * (w_cls * invc) * bce_grad_raw(pc, ((bits >> c) & 1u) ? 1.0f : 0.0f);
            }
        }
    }
}

__global__ void yolo_targets_bwd_kernel(cy4_yolo_desc d, HeadView hv, const float *__restrict__ anchors4, Workspace ws,
                                        int64_t nT, const float *__restrict__ gloss, float *__restrict__ dpred,
                                        int64_t dsB, int64_t dsC, int64_t dsH, int64_t dsW)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nT) return;
    const TargetRec r = ws.rec[t];
    if (!r.valid) return;
    const float s = gloss[0] * 3.54f / (float)nT;
    const int nO = 7 + d.nC;
    float *o = dpred + r.b * dsB + (int64_t)(r.a * nO) * dsC + r.gj * dsH + r.gi * dsW;
    const float px = sigmoidf_(hv.at(r.b, r.a, 0, r.gj, r.gi)), py = sigmoidf_(hv.at(r.b, r.a, 1, r.gj, r.gi));
    const float ew = expf(hv.at(r.b, r.a, 2, r.gj, r.gi)), eh = expf(hv.at(r.b, r.a, 3, r.gj, r.gi));
    atomicAdd(o + 0 * dsC, s * r.gbox[0] * (px * (1.0f - px)));
    atomicAdd(o + 1 * dsC, s * r.gbox[1] * (py * (1.0f - py)));
    atomicAdd(o + 2 * dsC, ew <= 1e3f ? s * r.gbox[2] * anchors4[r.a * 4] * ew : 0.f);       // clamp(max=1e3) gate
    atomicAdd(o + 3 * dsC, eh <= 1e3f ? s * r.gbox[3] * anchors4[r.a * 4 + 1] * eh : 0.f);
    atomicAdd(o + 4 * dsC, s * r.gbox[4]);
    atomicAdd(o + 5 * dsC, s * r.gbox[5]);
}

// Dense materialisation of build_targets' 13 outputs from cellmap / clsbits / records.
__global__ void build_targets_dense_kernel(cy4_yolo_desc d, Workspace ws, float *iou_scores, float *class_mask,
                                           uint8_t *obj_mask, uint8_t *noobj_mask, float *tx, float *ty, float *tw,
                                           float *th, float *tim, float *tre, float *tcls, float *tconf)
{
    const int64_t cells = (int64_t)d.B * d.nA * d.G * d.G;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const int32_t st = ws.cellmap[i];
    const bool isobj = st >= 2;
    TargetRec r;
    if (isobj) r = ws.rec[st - 2];
    obj_mask[i] = isobj ? 1 : 0;
    noobj_mask[i] = st == 0 ? 1 : 0;
    tconf[i] = isobj ? 1.f : 0.f;
    iou_scores[i] = isobj ? r.iou : 0.f;
    class_mask[i] = isobj ? r.clsmatch : 0.f;
    tx[i] = isobj ? r.tx : 0.f; ty[i] = isobj ? r.ty : 0.f; tw[i] = isobj ? r.tw : 0.f; th[i] = isobj ? r.th : 0.f;
    tim[i] = isobj ? r.tim : 0.f; tre[i] = isobj ? r.tre : 0.f;
    const uint32_t bits = isobj ? ws.clsbits[i] : 0u;
    for (int c = 0; c < d.nC; ++c) tcls[i * d.nC + c] = ((bits >> c) & 1u) ? 1.f : 0.f;
}

__global__ void build_targets_misc_kernel(Workspace ws, int64_t nT, float *giou_loss, int64_t *idx)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float a = 0.f;
        for (int64_t t = 0; t < nT; ++t) if (ws.rec[t].valid) a = a + ws.rec[t].term;
        giou_loss[0] = nT > 0 ? a / (float)nT : 0.f;
    }
    if (idx) {
        const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t < nT) {
            const TargetRec r = ws.rec[t];
            idx[0 * nT + t] = r.b; idx[1 * nT + t] = r.a; idx[2 * nT + t] = r.gj; idx[3 * nT + t] = r.gi; idx[4 * nT + t] = r.label;
        }
    }
}

static int check_desc(const cy4_yolo_desc *d, const char *who)
{
    if (!d || d->B <= 0 || d->G <= 0 || d->nA <= 0 || d->nC <= 0 || d->nC > 32) {
        set_error("%s: bad descriptor (need B,G,nA > 0 and 0 < nC <= 32)", who);
        return -1;
    }
    return 0;
}

static HeadView make_view(const cy4_yolo_desc *d, const float *pred)
{
    HeadView hv; hv.p = pred; hv.sB = d->sB; hv.sC = d->sC; hv.sH = d->sH; hv.sW = d->sW; hv.nA = d->nA; hv.nC = d->nC; hv.G = d->G;
    return hv;
}

static int dense_grid(int64_t cells)
{
    const int64_t need = (cells + kDenseBlock - 1) / kDenseBlock;
    return (int)std::min<int64_t>(need, std::min<int64_t>(kMaxDenseGrid, (int64_t)sm_count() * 8));
}

}  // namespace cy4

using namespace cy4;

extern "C" {

size_t cy4_yolo_workspace_bytes(const cy4_yolo_desc *d, int64_t nT)
{
    if (check_desc(d, "cy4_yolo_workspace_bytes") || nT < 0) return 0;
    return carve(d, nT, nullptr, nullptr);
}

int cy4_yolo_decode(const cy4_yolo_desc *d, const float *pred, const float *anchors4, float *output, void *stream)
{
    if (check_desc(d, "cy4_yolo_decode")) return -1;
    CY4_CHECK_ARG(pred && anchors4 && output, "cy4_yolo_decode: null pointer");
    const int64_t cells = (int64_t)d->B * d->nA * d->G * d->G;
    Workspace ws = {};
    const int a_fast = d->sC < d->sW ? 1 : 0;
    yolo_dense_kernel<false><<<dense_grid(cells), kDenseBlock, 0, (cudaStream_t)stream>>>(*d, make_view(d, pred), anchors4, ws, output, a_fast);
    return cy4_launch_status("cy4_yolo_decode");
}

int cy4_yolo_loss_fwd(const cy4_yolo_desc *d, const float *pred, const float *anchors4, const float *targets8, int64_t nT,
                      float *output, float *loss, float *metrics, int32_t *status, void *workspace, void *stream)
{
    if (check_desc(d, "cy4_yolo_loss_fwd")) return -1;
    CY4_CHECK_ARG(pred && anchors4 && loss && status && workspace && nT >= 0 && (targets8 || nT == 0),
                  "cy4_yolo_loss_fwd: null pointer / negative nT");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws;
    carve(d, nT, workspace, &ws);
    const int64_t cells = (int64_t)d->B * d->nA * d->G * d->G;
    CY4_CUDA(cudaMemsetAsync(ws.cellmap, 0, (size_t)cells * 4, st));
    CY4_CUDA(cudaMemsetAsync(ws.clsbits, 0, (size_t)cells * 4, st));
    CY4_CUDA(cudaMemsetAsync(status, 0, 4, st));
    HeadView hv = make_view(d, pred);
    if (nT > 0)
        yolo_targets_kernel<true><<<(unsigned)((nT + kTgtBlock - 1) / kTgtBlock), kTgtBlock, 0, st>>>(
            *d, hv, nullptr, nullptr, anchors4, targets8, nT, ws, status);
    const int grid = dense_grid(cells);
    const int a_fast = d->sC < d->sW ? 1 : 0;
    yolo_dense_kernel<true><<<grid, kDenseBlock, 0, st>>>(*d, hv, anchors4, ws, output, a_fast);
    yolo_finalize_kernel<<<1, 64, 0, st>>>(*d, ws, grid, nT, loss, metrics);
    return cy4_launch_status("cy4_yolo_loss_fwd", nT > 0 ? 3 : 2);
}

int cy4_yolo_loss_bwd(const cy4_yolo_desc *d, const float *pred, const float *anchors4, const float *targets8, int64_t nT,
                      const float *gloss, const void *workspace, float *dpred, int64_t dsB, int64_t dsC, int64_t dsH,
                      int64_t dsW, void *stream)
{
    (void)targets8;
    if (check_desc(d, "cy4_yolo_loss_bwd")) return -1;
    CY4_CHECK_ARG(pred && anchors4 && gloss && workspace && dpred && nT >= 0, "cy4_yolo_loss_bwd: null pointer / negative nT");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws;
    carve(d, nT, const_cast<void *>(workspace), &ws);
    const int64_t cells = (int64_t)d->B * d->nA * d->G * d->G;
    HeadView hv = make_view(d, pred);
    const int a_fast = dsC < dsW ? 1 : 0;
    yolo_dense_bwd_kernel<<<dense_grid(cells), kDenseBlock, 0, st>>>(*d, hv, ws, gloss, dpred, dsB, dsC, dsH, dsW, a_fast);
    if (nT > 0 && d->use_giou)
        yolo_targets_bwd_kernel<<<(unsigned)((nT + 127) / 128), 128, 0, st>>>(*d, hv, anchors4, ws, nT, gloss, dpred, dsB, dsC, dsH, dsW);
    return cy4_launch_status("cy4_yolo_loss_bwd", (nT > 0 && d->use_giou) ? 2 : 1);
}

int cy4_build_targets(const cy4_yolo_desc *d, const float *pred_boxes, const float *pred_cls, const float *targets8,
                      int64_t nT, const float *anchors4, float *iou_scores, float *giou_loss, float *class_mask,
                      uint8_t *obj_mask, uint8_t *noobj_mask, float *tx, float *ty, float *tw, float *th, float *tim,
                      float *tre, float *tcls, float *tconf, int64_t *idx, int32_t *status, void *workspace, void *stream)
{
    if (check_desc(d, "cy4_build_targets")) return -1;
    CY4_CHECK_ARG(pred_boxes && pred_cls && anchors4 && iou_scores && giou_loss && class_mask && obj_mask && noobj_mask &&
                  tx && ty && tw && th && tim && tre && tcls && tconf && status && workspace && nT >= 0 && (targets8 || nT == 0),
                  "cy4_build_targets: null pointer / negative nT");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace ws;
    carve(d, nT, workspace, &ws);
    const int64_t cells = (int64_t)d->B * d->nA * d->G * d->G;
    CY4_CUDA(cudaMemsetAsync(ws.cellmap, 0, (size_t)cells * 4, st));
    CY4_CUDA(cudaMemsetAsync(ws.clsbits, 0, (size_t)cells * 4, st));
    CY4_CUDA(cudaMemsetAsync(status, 0, 4, st));
    HeadView hv = {};
    if (nT > 0)
        yolo_targets_kernel<false><<<(unsigned)((nT + kTgtBlock - 1) / kTgtBlock), kTgtBlock, 0, st>>>(
            *d, hv, pred_boxes, pred_cls, anchors4, targets8, nT, ws, status);
    build_targets_dense_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(*d, ws, iou_scores, class_mask, obj_mask, noobj_mask,
                                                                              tx, ty, tw, th, tim, tre, tcls, tconf);
    build_targets_misc_kernel<<<(unsigned)std::max<int64_t>(1, (nT + 127) / 128), 128, 0, st>>>(ws, nT, giou_loss, idx);
    return cy4_launch_status("cy4_build_targets", nT > 0 ? 3 : 2);
}

}  // extern "C"
