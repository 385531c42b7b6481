// rbox.cuh -- device-side rotated-box geometry for sm_100a.
//
// One thread owns one (pred, target) box pair.  The clip polygon (<= 8 live vertices) lives in
// shared memory, laid out [buffer][vertex][thread] so that every dynamic vertex index is a
// conflict-free LDS/STS; the 8 hull candidates stay in registers (all indices are compile-time
// after unrolling).  Arithmetic mirrors the reference's fp32 operation order -- this translation
// unit is compiled with --fmad=false and without fast-math so every product and sum rounds
// separately, as torch's CPU ops do:
//   corners        src/utils/iou_rotated_boxes_utils.py:34-61
//   clip / area    src/utils/cal_intersection_rotated_boxes.py:16-96   (quirks F5/F6 kept)
//   iou / giou     src/utils/iou_rotated_boxes_utils.py:98-142
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cy4 {

constexpr int kMaxV = 10;   // a convex quad clipped by 4 half-planes has <= 8 vertices; 2 spare

// Per-thread view of the shared-memory polygon storage.
template <int BLOCK>
struct PolySmem {
    float x[2][kMaxV][BLOCK];
    float y[2][kMaxV][BLOCK];
    signed char s[2][kMaxV][BLOCK];
};

// atan2 evaluated in fp64 and rounded once (see box_corners).
__device__ __forceinline__ float atan2_cr(float y, float x) { return (float)atan2((double)y, (double)x); }

__device__ __forceinline__ void box_corners(float x, float y, float w, float l, float yaw,
                                            float cx[4], float cy[4], float &cs, float &sn)
{
    // fp64 sincos rounded once to fp32: correctly rounded like glibc / (almost always) Sleef, so the
    // corner coordinates are bit-identical to the reference's for the vast majority of boxes.
    double sd, cd;
    sincos((double)yaw, &sd, &cd);
    cs = (float)cd; sn = (float)sd;
    const float hw = w / 2.0f, hl = l / 2.0f;
    const float a = hw * cs, b = hl * sn, c = hw * sn, d = hl * cs;
    cx[0] = (x - a) - b;  cy[0] = (y - c) + d;   // front left
    cx[1] = (x - a) + b;  cy[1] = (y - c) - d;   // rear left
    cx[2] = (x + a) + b;  cy[2] = (y + c) - d;   // rear right
    cx[3] = (x + a) - b;  cy[3] = (y + c) + d;   // front right
}

// torch's CPU fp32 .sum() of n <= 8 contiguous values (see DESIGN.md "summation order"):
// n <= 4 or n >= 8: left to right;  5..7: ((v0 + v4 + .. + v[n-1]) + v1) + v2) + v3.
template <typename F>
__device__ __forceinline__ float torch_small_sum(int n, F term)
{
    float a;
    if (n <= 4 || n >= 8) {
        a = 0.0f;
        for (int i = 0; i < n; ++i) a = a + term(i);
    } else {
        a = term(0);
        for (int i = 4; i < n; ++i) a = a + term(i);
        a = a + term(1); a = a + term(2); a = a + term(3);
    }
    return a;
}

// Sutherland-Hodgman clip of rect1 (px,py) by the 4 edges of rect2 (tx,ty), reference order and
// predicates.  Result: vertex count, polygon in sm.{x,y,s}[buf][..][tid]; s = rect1 corner id or -1.
template <int BLOCK>
__device__ __forceinline__ int clip_ref(const float px[4], const float py[4], const float tx[4], const float ty[4],
                                        PolySmem<BLOCK> &sm, int tid, int &buf_out)
{
    int cur = 0, n = 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sm.x[0][i][tid] = px[i]; sm.y[0][i][tid] = py[i]; sm.s[0][i][tid] = (signed char)i; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (n <= 2) break;
        const float p0 = tx[e], p1 = ty[e], q0 = tx[(e + 1) & 3], q1 = ty[(e + 1) & 3];
        const float a = q1 - p1;
        const float b = p0 - q0;
        const float c = q0 * p1 - q1 * p0;
        const int nxt = cur ^ 1;
        int m = 0;
        float sx = sm.x[cur][0][tid], sy = sm.y[cur][0][tid];
        signed char ss = sm.s[cur][0][tid];
        float sv = (a * sx + b * sy) + c;
        const float v0x = sx, v0y = sy, v0v = sv;
        for (int i = 0; i < n; ++i) {
            float ex, ey, ev; signed char es = 0;
            if (i + 1 < n) {
                ex = sm.x[cur][i + 1][tid]; ey = sm.y[cur][i + 1][tid]; es = sm.s[cur][i + 1][tid];
                ev = (a * ex + b * ey) + c;
            } else { ex = v0x; ey = v0y; ev = v0v; }
            if (sv <= 0.0f && m < kMaxV) {
                sm.x[nxt][m][tid] = sx; sm.y[nxt][m][tid] = sy; sm.s[nxt][m][tid] = ss; ++m;
            }
            if (sv * ev < 0.0f && m < kMaxV) {
                const float a2 = ey - sy;
                const float b2 = sx - ex;
                const float c2 = ex * sy - ey * sx;
                const float w = a * b2 - b * a2;
                sm.x[nxt][m][tid] = (b * c2 - c * b2) / w;
                sm.y[nxt][m][tid] = (c * a2 - a * c2) / w;
                sm.s[nxt][m][tid] = (signed char)-1; ++m;
            }
            sx = ex; sy = ey; sv = ev; ss = es;
        }
        if (m > 0) { n = m; cur = nxt; }
        else break;                          // F5: keep the previous polygon
    }
    buf_out = cur;
    return n;
}

// Gradient accumulator: d term / d (x, y, w, l, yaw) through the pred-box corner formulas.
struct CornerGrad {
    float gx = 0.f, gy = 0.f, gw = 0.f, gl = 0.f, gyaw = 0.f;
    float cs, sn, hw, hl;
    __device__ __forceinline__ void add(int corner, float dx, float dy)
    {
        const float SW = corner >= 2 ? 1.f : -1.f;
        const float SLX = (corner == 1 || corner == 2) ? 1.f : -1.f;
        const float SLY = (corner == 0 || corner == 3) ? 1.f : -1.f;
        gx += dx; gy += dy;
        gw += 0.5f * SW * (dx * cs + dy * sn);
        gl += 0.5f * (dx * SLX * sn + dy * SLY * cs);
        gyaw += dx * (-SW * hw * sn + SLX * hl * cs) + dy * (SW * hw * cs - SLY * hl * sn);
    }
};

// Exact convex quad/quad intersection area in fp64 on the fp32 corners (stands in for
// shapely/GEOS at iou_rotated_boxes_utils.py:91,118-120).  Local-memory arrays: this path is cold
// (3 x nT anchor pairs per layer; GIoU=False metrics).
__device__ inline double convex_inter64(const float ax[4], const float ay[4], const float bx[4], const float by[4])
{
    double X[2][kMaxV], Y[2][kMaxV];
    int cur = 0, n = 4;
    for (int i = 0; i < 4; ++i) { X[0][i] = ax[i]; Y[0][i] = ay[i]; }
    double o2 = 0.0;
    for (int i = 0; i < 4; ++i) { int j = (i + 1) & 3; o2 += (double)bx[i] * by[j] - (double)by[i] * bx[j]; }
    const double sg = o2 >= 0.0 ? 1.0 : -1.0;
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double px = bx[e], py = by[e], qx = bx[(e + 1) & 3], qy = by[(e + 1) & 3];
        const int nxt = cur ^ 1;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double ds = sg * ((qx - px) * (Y[cur][i] - py) - (qy - py) * (X[cur][i] - px));
            const double dt = sg * ((qx - px) * (Y[cur][j] - py) - (qy - py) * (X[cur][j] - px));
            if (ds >= 0.0 && m < kMaxV) { X[nxt][m] = X[cur][i]; Y[nxt][m] = Y[cur][i]; ++m; }
            if (((ds > 0.0 && dt < 0.0) || (ds < 0.0 && dt > 0.0)) && m < kMaxV) {
                const double t = ds / (ds - dt);
                X[nxt][m] = X[cur][i] + t * (X[cur][j] - X[cur][i]);
                Y[nxt][m] = Y[cur][i] + t * (Y[cur][j] - Y[cur][i]);
                ++m;
            }
        }
        n = m; cur = nxt;
    }
    if (n < 3) return 0.0;
    double s = 0.0;
    for (int i = 0; i < n; ++i) { const int j = (i + 1 == n) ? 0 : i + 1; s += X[cur][i] * Y[cur][j] - Y[cur][i] * X[cur][j]; }
    return fabs(s) * 0.5;
}

// One (pred, target) pair: iou, giou term and (GRAD) d term / d pred6.
// P/T = (x, y, w, l, im, re).  giou: reference clipper + hull; else exact intersection, term = 1 - iou.
template <int BLOCK, bool GRAD>
__device__ __forceinline__ void rgiou_pair(const float P[6], const float T[6], bool giou, PolySmem<BLOCK> &sm, int tid,
                                           float &iou, float &term, float g[6])
{
    float px[4], py[4], tx[4], ty[4], tcs, tsn;
    CornerGrad cg;
    const float tyaw = atan2_cr(T[4], T[5]);
    box_corners(T[0], T[1], T[2], T[3], tyaw, tx, ty, tcs, tsn);
    const float pyaw = atan2_cr(P[4], P[5]);
    box_corners(P[0], P[1], P[2], P[3], pyaw, px, py, cg.cs, cg.sn);
    cg.hw = 0.5f * P[2]; cg.hl = 0.5f * P[3];
    const float t_area = T[2] * T[3];
    const float p_area = P[2] * P[3];
    float g_parea = 0.f;

    if (giou) {
        int buf;
        const int m = clip_ref<BLOCK>(px, py, tx, ty, sm, tid, buf);
        const bool inter_is_tensor = m > 2;
        float ssum = 0.f, inter = 0.f;
        if (inter_is_tensor) {
            ssum = torch_small_sum(m, [&](int i) {
                const int j = (i + 1 == m) ? 0 : i + 1;
                return sm.x[buf][i][tid] * sm.y[buf][j][tid] - sm.y[buf][i][tid] * sm.x[buf][j][tid];
            });
            inter = fabsf(ssum) * 0.5f;
        }
        const float uni = (p_area + t_area) - inter;
        iou = inter_is_tensor ? inter / (uni + 1e-16f) : (1.0f / (uni + 1e-16f)) * inter;

        // ---- convex hull of the 8 corners: Jarvis march, counter-clockwise, from the
        // lexicographically smallest point, collinear points skipped; predicates in fp64.
        float qx[8], qy[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qx[i] = px[i]; qy[i] = py[i]; qx[4 + i] = tx[i]; qy[4 + i] = ty[i]; }
        float hx[8], hy[8]; int hid[8];
        float stx = qx[0], sty = qy[0]; int sid = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j)
            if (qx[j] < stx || (qx[j] == stx && qy[j] < sty)) { stx = qx[j]; sty = qy[j]; sid = j; }
        hx[0] = stx; hy[0] = sty; hid[0] = sid;
        int hn = 1;
        float cx = stx, cy = sty;
        bool done = false;
#pragma unroll
        for (int step = 1; step < 8; ++step) {
            hx[step] = 0.f; hy[step] = 0.f; hid[step] = -1;
            if (!done) {
                int nid = -1; float nx = 0.f, ny = 0.f; double nd = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool distinct = (qx[j] != cx) || (qy[j] != cy);
                    if (distinct) {
                        const double dxj = (double)qx[j] - cx, dyj = (double)qy[j] - cy;
                        const double dj = dxj * dxj + dyj * dyj;
                        if (nid < 0) { nid = j; nx = qx[j]; ny = qy[j]; nd = dj; }
                        else {
                            const double cr = ((double)nx - cx) * dyj - ((double)ny - cy) * dxj;
                            if (cr < 0.0 || (cr == 0.0 && dj > nd)) { nid = j; nx = qx[j]; ny = qy[j]; nd = dj; }
                        }
                    }
                }
                if (nid < 0 || (nx == stx && ny == sty)) done = true;
                else { hx[step] = nx; hy[step] = ny; hid[step] = nid; hn = step + 1; cx = nx; cy = ny; }
            }
        }
        // (cx, cy) is now the last hull vertex
        float hsum = 0.f, carea = 0.f;
        if (hn >= 3) {
            // hull terms with compile-time indices: t_i = x_i*y_{i+1} - y_i*x_{i+1}
            float ht[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float nxv = (i + 1 < hn) ? hx[(i + 1) & 7] : hx[0];
                const float nyv = (i + 1 < hn) ? hy[(i + 1) & 7] : hy[0];
                ht[i] = hx[i] * nyv - hy[i] * nxv;
            }
            if (hn <= 4 || hn >= 8) {
                hsum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) if (i < hn) hsum = hsum + ht[i];
            } else {
                hsum = ht[0];
#pragma unroll
                for (int i = 4; i < 8; ++i) if (i < hn) hsum = hsum + ht[i];
                hsum = hsum + ht[1]; hsum = hsum + ht[2]; hsum = hsum + ht[3];
            }
            carea = fabsf(hsum) * 0.5f;
        }
        term = 1.0f - (iou - (carea - uni) / (carea + 1e-16f));

        if (GRAD) {
            const float Ue = uni + 1e-16f, Ce = carea + 1e-16f;
            float dI = inter_is_tensor ? (-(1.0f / Ue) - inter / (Ue * Ue) + 1.0f / Ce) : 0.f;
            g_parea = inter / (Ue * Ue) - 1.0f / Ce;
            const float dC = Ue / (Ce * Ce);
            if (inter_is_tensor) {
                const float sgn = ssum > 0.f ? 0.5f : (ssum < 0.f ? -0.5f : 0.f);
                dI *= sgn;
                for (int i = 0; i < m; ++i) {
                    const int src = sm.s[buf][i][tid];
                    if (src >= 0) {
                        const int nx_ = (i + 1 == m) ? 0 : i + 1, pv = (i == 0) ? m - 1 : i - 1;
                        cg.add(src, dI * (sm.y[buf][nx_][tid] - sm.y[buf][pv][tid]),
                               dI * (sm.x[buf][pv][tid] - sm.x[buf][nx_][tid]));
                    }
                }
            }
            if (hn >= 3) {
                const float sgn = hsum > 0.f ? 0.5f : (hsum < 0.f ? -0.5f : 0.f);
                const float dCs = dC * sgn;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < hn && hid[i] >= 0 && hid[i] < 4) {
                        const float nxv = (i + 1 < hn) ? hx[(i + 1) & 7] : hx[0];
                        const float nyv = (i + 1 < hn) ? hy[(i + 1) & 7] : hy[0];
                        const float pxv = (i == 0) ? cx : hx[(i + 7) & 7];
                        const float pyv = (i == 0) ? cy : hy[(i + 7) & 7];
                        cg.add(hid[i], dCs * (nyv - pyv), dCs * (pxv - nxv));
                    }
                }
            }
        }
    } else {
        const float inter = (float)convex_inter64(px, py, tx, ty);
        const float uni = (p_area + t_area) - inter;
        iou = (1.0f / (uni + 1e-16f)) * inter;       // python float / tensor == reciprocal * float
        term = 1.0f - iou;
        if (GRAD) { const float Ue = uni + 1e-16f; g_parea = inter / (Ue * Ue); }
    }

    if (GRAD) {
        const float gw = cg.gw + g_parea * P[3];
        const float gl = cg.gl + g_parea * P[2];
        const float r2 = P[4] * P[4] + P[5] * P[5];
        g[0] = cg.gx; g[1] = cg.gy; g[2] = gw; g[3] = gl;
        g[4] = cg.gyaw * (P[5] / r2);
        g[5] = cg.gyaw * (-P[4] / r2);
    }
}

// IoU of an anchor and a target box (w, l, im, re), both centred at (100,100):
// iou_rotated_boxes_utils.py:64-95.  fp32 corners, fp64 intersection rounded once to fp32,
// then reciprocal-multiply (Tensor.__rtruediv__).
__device__ inline float anchor_target_iou(const float A[4], const float T[4])
{
    float ax[4], ay[4], bx[4], by[4], cs, sn;
    box_corners(100.0f, 100.0f, A[0], A[1], atan2_cr(A[2], A[3]), ax, ay, cs, sn);
    box_corners(100.0f, 100.0f, T[0], T[1], atan2_cr(T[2], T[3]), bx, by, cs, sn);
    const float aa = A[0] * A[1], ta = T[0] * T[1];
    const float inter = (float)convex_inter64(ax, ay, bx, by);
    const float den = ((aa + ta) - inter) + 1e-16f;
    return (1.0f / den) * inter;
}

}  // namespace cy4
