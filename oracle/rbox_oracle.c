/*
 * oracle/rbox_oracle.c -- CPU restatement of the reference's rotated-box geometry.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (complex-yolov4-pytorch_b200/)
 * may include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity pinning: the reference ships no tests or golden vectors, and its
 * third-party arithmetic (shapely/GEOS, scipy/Qhull, torch/Sleef) is unpinned.  This
 * restatement is pinned against outputs of the reference itself, generated in the build
 * container by oracle/gen_golden.py (real reference code imported from /root/reference/src
 * with a convex-quad stand-in for the missing shapely) and committed under tests/golden/.
 * The GEOS path (anchor<->target IoU, GIoU=False intersection) is therefore
 * "parity unpinned" with respect to real GEOS; it is cross-checked against
 * cv2.intersectConvexConvex instead.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Compile with -ffp-contract=off: the reference rounds to fp32
 * after every op.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAXV 24

/* utils/iou_rotated_boxes_utils.py:34-61  get_corners_vectorize.
 * Python precedence: x - w / 2 * cos - l / 2 * sin  ==  (x - ((w/2)*cos)) - ((l/2)*sin). */
void orc_corners(float x, float y, float w, float l, float yaw, float c[4][2])
{
    float cs = cosf(yaw), sn = sinf(yaw);
    float hw = w / 2.0f, hl = l / 2.0f;
    c[0][0] = (x - hw * cs) - hl * sn;   c[0][1] = (y - hw * sn) + hl * cs;   /* front left  */
    c[1][0] = (x - hw * cs) + hl * sn;   c[1][1] = (y - hw * sn) - hl * cs;   /* rear left   */
    c[2][0] = (x + hw * cs) + hl * sn;   c[2][1] = (y + hw * sn) - hl * cs;   /* rear right  */
    c[3][0] = (x + hw * cs) - hl * sn;   c[3][1] = (y + hw * sn) + hl * cs;   /* front right */
}

/* torch (2.11, CPU, AVX2/AVX512 dispatch) evaluates a contiguous fp32 .sum() of n<=8
 * elements as: n<=4 or n==8 sequential; 5<=n<=7  ((v0+v4+..+v[n-1]) + v1) + v2) + v3.
 * Probed in the build container (see DESIGN.md "summation order"). */
static float torch_small_sum(const float *v, int n)
{
    float a;
    int i;
    if (n <= 4 || n >= 8) {
        a = 0.0f;
        for (i = 0; i < n; ++i) a = a + v[i];
        return a;
    }
    a = v[0];
    for (i = 4; i < n; ++i) a = a + v[i];
    a = a + v[1]; a = a + v[2]; a = a + v[3];
    return a;
}

/* utils/cal_intersection_rotated_boxes.py:93-96  PolyArea2D (fp32). Returns the signed sum
 * through *signed_sum (needed for the backward) and |sum|*0.5 as the value. */
static float poly_area32(const float p[][2], int n, float *signed_sum)
{
    float terms[MAXV];
    int i;
    for (i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        terms[i] = p[i][0] * p[j][1] - p[i][1] * p[j][0];
    }
    float s = torch_small_sum(terms, n);
    if (signed_sum) *signed_sum = s;
    return fabsf(s) * 0.5f;
}

float orc_poly_area(const float *pts, int n)
{
    return poly_area32((const float (*)[2])pts, n, 0);
}

/* utils/cal_intersection_rotated_boxes.py:42-90  intersection_area, fp32, with both
 * reference quirks kept:
 *   F5 (:81-84)  a clip stage that removes every vertex breaks WITHOUT clearing the polygon;
 *   F6 (:38-39)  intersection points are built by torch.tensor([...]) => constants for autograd.
 * src[i] = index 0..3 of the rect1 corner vertex i still is, or -1 for a (detached)
 * intersection point.  Returns the vertex count. */
static int clip_ref(const float r1[4][2], const float r2[4][2], float poly[MAXV][2], int src[MAXV])
{
    int n = 4, e, i;
    for (i = 0; i < 4; ++i) { poly[i][0] = r1[i][0]; poly[i][1] = r1[i][1]; src[i] = i; }
    for (e = 0; e < 4; ++e) {
        if (n <= 2) break;                                   /* :59 */
        const float *p = r2[e], *q = r2[(e + 1) % 4];
        float a = q[1] - p[1];                               /* :25 */
        float b = p[0] - q[0];                               /* :26 */
        float c = q[0] * p[1] - q[1] * p[0];                 /* :27 */
        float val[MAXV];
        for (i = 0; i < n; ++i) val[i] = (a * poly[i][0] + b * poly[i][1]) + c;   /* :30 */
        float np_[MAXV][2]; int ns[MAXV]; int m = 0;
        for (i = 0; i < n; ++i) {
            int j = (i + 1) % n;
            float sv = val[i], tv = val[j];
            if (sv <= 0.0f && m < MAXV) {                     /* :73 */
                np_[m][0] = poly[i][0]; np_[m][1] = poly[i][1]; ns[m] = src[i]; ++m;
            }
            if (sv * tv < 0.0f && m < MAXV) {                 /* :75 */
                float a2 = poly[j][1] - poly[i][1];
                float b2 = poly[i][0] - poly[j][0];
                float c2 = poly[j][0] * poly[i][1] - poly[j][1] * poly[i][0];
                float w = a * b2 - b * a2;                   /* :37 */
                np_[m][0] = (b * c2 - c * b2) / w;           /* :38 */
                np_[m][1] = (c * a2 - a * c2) / w;
                ns[m] = -1; ++m;
            }
        }
        if (m > 0) {                                         /* :81 */
            n = m;
            for (i = 0; i < m; ++i) { poly[i][0] = np_[i][0]; poly[i][1] = np_[i][1]; src[i] = ns[i]; }
        } else {
            break;                                           /* :84  (F5) */
        }
    }
    return n;
}

float orc_intersection_area(const float *rect1, const float *rect2)
{
    float poly[MAXV][2]; int src[MAXV];
    int n = clip_ref((const float (*)[2])rect1, (const float (*)[2])rect2, poly, src);
    if (n <= 2) return 0.0f;                                 /* :87-88 */
    return poly_area32(poly, n, 0);
}

/* Convex hull of the 8 corners, replacing scipy.spatial.ConvexHull
 * (utils/iou_rotated_boxes_utils.py:129-131: indices only, counter-clockwise in 2-D).
 * Andrew monotone chain in fp64 on the fp32 coordinates, collinear points dropped.
 * Qhull's starting vertex is internal and has no simple rule (probed); ours starts at the
 * lexicographically smallest (x, y) point.  The start only rotates the fp32 summation order
 * of PolyArea2D. Returns the number of hull vertices, indices in idx[]. */
static int hull8(const float pts[8][2], int idx[8])
{
    int order[8], i, j, k = 0, h[16];
    for (i = 0; i < 8; ++i) order[i] = i;
    for (i = 1; i < 8; ++i) {          /* insertion sort by (x, y), stable */
        int o = order[i];
        for (j = i - 1; j >= 0; --j) {
            int q = order[j];
            if (pts[q][0] > pts[o][0] || (pts[q][0] == pts[o][0] && pts[q][1] > pts[o][1])) order[j + 1] = q;
            else break;
        }
        order[j + 1] = o;
    }
#define CROSS(o, a, b) (((double)pts[a][0] - pts[o][0]) * ((double)pts[b][1] - pts[o][1]) - \
                        ((double)pts[a][1] - pts[o][1]) * ((double)pts[b][0] - pts[o][0]))
    for (i = 0; i < 8; ++i) {
        while (k >= 2 && CROSS(h[k - 2], h[k - 1], order[i]) <= 0.0) --k;
        h[k++] = order[i];
    }
    int lo = k + 1;
    for (i = 6; i >= 0; --i) {
        while (k >= lo && CROSS(h[k - 2], h[k - 1], order[i]) <= 0.0) --k;
        h[k++] = order[i];
    }
#undef CROSS
    k -= 1;                            /* last point == first point */
    if (k < 1) k = 1;
    /* duplicates of identical points can survive the chain; drop exact repeats */
    int m = 0;
    for (i = 0; i < k; ++i) {
        int dup = 0;
        for (j = 0; j < m; ++j)
            if (pts[idx[j]][0] == pts[h[i]][0] && pts[idx[j]][1] == pts[h[i]][1]) { dup = 1; break; }
        if (!dup) idx[m++] = h[i];
    }
    return m;
}

/* Exact (fp64) convex quad/quad intersection area, standing in for shapely/GEOS
 * (utils/iou_rotated_boxes_utils.py:91 and :118-120).  Sutherland-Hodgman in fp64 on the
 * fp32 corner coordinates; orientation-independent; empty => 0. */
static double clip_area64(const float r1[4][2], const float r2[4][2])
{
    double poly[MAXV][2], np_[MAXV][2];
    int n = 4, e, i;
    for (i = 0; i < 4; ++i) { poly[i][0] = r1[i][0]; poly[i][1] = r1[i][1]; }
    /* orientation of the clip polygon */
    double o2 = 0.0;
    for (i = 0; i < 4; ++i) {
        int j = (i + 1) % 4;
        o2 += (double)r2[i][0] * r2[j][1] - (double)r2[i][1] * r2[j][0];
    }
    double sgn = o2 >= 0.0 ? 1.0 : -1.0;
    for (e = 0; e < 4 && n > 0; ++e) {
        double px = r2[e][0], py = r2[e][1], qx = r2[(e + 1) % 4][0], qy = r2[(e + 1) % 4][1];
        int m = 0;
        for (i = 0; i < n; ++i) {
            int j = (i + 1) % n;
            double ds = sgn * ((qx - px) * (poly[i][1] - py) - (qy - py) * (poly[i][0] - px));
            double dt = sgn * ((qx - px) * (poly[j][1] - py) - (qy - py) * (poly[j][0] - px));
            if (ds >= 0.0) { np_[m][0] = poly[i][0]; np_[m][1] = poly[i][1]; ++m; }
            if ((ds > 0.0 && dt < 0.0) || (ds < 0.0 && dt > 0.0)) {
                double t = ds / (ds - dt);
                np_[m][0] = poly[i][0] + t * (poly[j][0] - poly[i][0]);
                np_[m][1] = poly[i][1] + t * (poly[j][1] - poly[i][1]);
                ++m;
            }
        }
        n = m;
        for (i = 0; i < m; ++i) { poly[i][0] = np_[i][0]; poly[i][1] = np_[i][1]; }
    }
    if (n < 3) return 0.0;
    double s = 0.0;
    for (i = 0; i < n; ++i) {
        int j = (i + 1) % n;
        s += poly[i][0] * poly[j][1] - poly[i][1] * poly[j][0];
    }
    return fabs(s) * 0.5;
}

double orc_convex_inter64(const float *rect1, const float *rect2)
{
    return clip_area64((const float (*)[2])rect1, (const float (*)[2])rect2);
}

/* utils/iou_rotated_boxes_utils.py:64-79 + :82-95:  IoU of every anchor with every target,
 * both placed at fix_xy = (100, 100).  boxes are (w, l, im, re).
 * Op order (:91-93): intersection is a Python float => cast to fp32; the division is
 * Tensor.__rtruediv__ == reciprocal(denominator) * intersection, all fp32. */
void orc_anchor_iou(const float *anchors4, int nA, const float *tgt4, int64_t nT, float *ious /* [nA,nT] */)
{
    int a; int64_t t;
    for (a = 0; a < nA; ++a) {
        float ac[4][2];
        const float *A = anchors4 + 4 * a;
        orc_corners(100.0f, 100.0f, A[0], A[1], atan2f(A[2], A[3]), ac);
        float aa = A[0] * A[1];
        for (t = 0; t < nT; ++t) {
            const float *T = tgt4 + 4 * t;
            float tc[4][2];
            orc_corners(100.0f, 100.0f, T[0], T[1], atan2f(T[2], T[3]), tc);
            float ta = T[0] * T[1];
            float inter = (float)clip_area64(ac, tc);
            float den = ((aa + ta) - inter) + 1e-16f;
            ious[(int64_t)a * nT + t] = (1.0f / den) * inter;
        }
    }
}

/* utils/evaluation_utils.py:186-210  iou_rotated_single_vs_multi_boxes_cpu, for every pair (a[i], b[j]).
 * rows (x, y, w, l, im, re).  Corners in fp32 (bev_utils.get_corners / get_corners_vectorize :213-239, same
 * left-to-right order as orc_corners), intersection exact in fp64 (shapely), then fp32:
 * iou = reciprocal((s_area + m_area) - inter + 1e-16) * inter   (python float / tensor). */
void orc_iou_matrix(const float *a6, int64_t n, const float *b6, int64_t m, float *ious)
{
    int64_t i, j;
    for (i = 0; i < n; ++i) {
        const float *A = a6 + 6 * i;
        float ac[4][2];
        orc_corners(A[0], A[1], A[2], A[3], atan2f(A[4], A[5]), ac);
        float aa = A[2] * A[3];
        for (j = 0; j < m; ++j) {
            const float *B = b6 + 6 * j;
            float bc[4][2];
            orc_corners(B[0], B[1], B[2], B[3], atan2f(B[4], B[5]), bc);
            float ba = B[2] * B[3];
            float inter = (float)clip_area64(ac, bc);
            float den = ((aa + ba) - inter) + 1e-16f;
            ious[i * m + j] = (1.0f / den) * inter;
        }
    }
}

/* utils/find_anchors.py:53-59 compute_iou for every (box, cluster) pair.  Rows (w, l, yaw) float64,
 * boxes centred at the origin.  Corners: data_process/kitti_bev_utils.py:96-120 get_corners -- float64 arithmetic stored into a
 * float32 array; polygon areas and the intersection in float64 (shapely); iou = inter / (a1 + a2 - inter + 1e-12) in float64,
 * stored as float32 (np.array(iou, dtype=np.float32)). */
static void kmeans_corners(double w, double l, double yaw, float c[4][2])
{
    double cs = cos(yaw), sn = sin(yaw);
    c[0][0] = (float)(0.0 - w / 2 * cs - l / 2 * sn);  c[0][1] = (float)(0.0 - w / 2 * sn + l / 2 * cs);
    c[1][0] = (float)(0.0 - w / 2 * cs + l / 2 * sn);  c[1][1] = (float)(0.0 - w / 2 * sn - l / 2 * cs);
    c[2][0] = (float)(0.0 + w / 2 * cs + l / 2 * sn);  c[2][1] = (float)(0.0 + w / 2 * sn - l / 2 * cs);
    c[3][0] = (float)(0.0 + w / 2 * cs - l / 2 * sn);  c[3][1] = (float)(0.0 + w / 2 * sn + l / 2 * cs);
}
static double quad_area64(const float c[4][2])
{
    double s = 0.0;
    int i;
    for (i = 0; i < 4; ++i) { int j = (i + 1) & 3; s += (double)c[i][0] * c[j][1] - (double)c[i][1] * c[j][0]; }
    return fabs(s) * 0.5;
}
void orc_kmeans_iou(const double *boxes3, int64_t n, const double *clusters3, int k, float *ious)
{
    int64_t i;
    int j;
    for (i = 0; i < n; ++i) {
        float bc[4][2];
        kmeans_corners(boxes3[3 * i], boxes3[3 * i + 1], boxes3[3 * i + 2], bc);
        double ba = quad_area64(bc);
        for (j = 0; j < k; ++j) {
            float cc[4][2];
            kmeans_corners(clusters3[3 * j], clusters3[3 * j + 1], clusters3[3 * j + 2], cc);
            double inter = clip_area64(bc, cc);
            ious[i * k + j] = (float)(inter / (ba + quad_area64(cc) - inter + 1e-12));
        }
    }
}

/* utils/iou_rotated_boxes_utils.py:98-142  iou_pred_vs_target_boxes, element-wise pairs.
 * flags bit0: GIoU (reference clipper :122 + hull term :128-133); otherwise the shapely path
 * (:118-120, exact intersection) with term = 1 - iou (:135).
 * Outputs per pair: iou, term (the summand of giou_loss), and if grad6 != NULL the
 * gradient d term / d pred(x,y,w,l,im,re) under the reference's autograd semantics (F6:
 * intersection points are constants; hull vertices contribute only where they are pred
 * corners; in the shapely path only p_w*p_l carries gradient). Gradients are evaluated in
 * fp64 from the fp32 forward quantities. */
void orc_rgiou_pairs(const float *pred6, const float *tgt6, int64_t n, uint32_t flags,
                     float *iou_out, float *term_out, float *grad6)
{
    int64_t k;
    const int giou = (flags & 1u) != 0;
    for (k = 0; k < n; ++k) {
        const float *P = pred6 + 6 * k, *T = tgt6 + 6 * k;
        float pc[4][2], tc[4][2];
        float pyaw = atan2f(P[4], P[5]), tyaw = atan2f(T[4], T[5]);
        orc_corners(T[0], T[1], T[2], T[3], tyaw, tc);
        orc_corners(P[0], P[1], P[2], P[3], pyaw, pc);
        float t_area = T[2] * T[3];
        float p_area = P[2] * P[3];
        /* d term / d corner (x,y) of the 4 pred corners, and d term / d p_area */
        double gc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        double g_parea = 0.0;
        float inter, uni, iou, term;

        if (giou) {
            float poly[MAXV][2]; int src[MAXV];
            int m = clip_ref(pc, tc, poly, src);
            float ssum = 0.0f;
            int inter_is_tensor = m > 2;
            inter = inter_is_tensor ? poly_area32(poly, m, &ssum) : 0.0f;
            uni = (p_area + t_area) - inter;                           /* :124 */
            if (inter_is_tensor) iou = inter / (uni + 1e-16f);         /* :125 tensor / tensor */
            else iou = (1.0f / (uni + 1e-16f)) * inter;                /* python 0. / tensor => reciprocal * 0. */
            float all8[8][2]; int hid[8], i;
            for (i = 0; i < 4; ++i) { all8[i][0] = pc[i][0]; all8[i][1] = pc[i][1]; all8[4 + i][0] = tc[i][0]; all8[4 + i][1] = tc[i][1]; }
            int hn = hull8(all8, hid);
            float hp[8][2]; float hsum = 0.0f;
            for (i = 0; i < hn; ++i) { hp[i][0] = all8[hid[i]][0]; hp[i][1] = all8[hid[i]][1]; }
            float carea = hn >= 3 ? poly_area32(hp, hn, &hsum) : 0.0f;
            term = 1.0f - (iou - (carea - uni) / (carea + 1e-16f));    /* :133 */

            if (grad6) {
                /* term = 1 - I/(U+e) + (C-U)/(C+e),  U = pa + ta - I */
                double I = inter, U = uni, C = carea, e = 1e-16;
                double dterm_dI = -(1.0 / (U + e)) - I / ((U + e) * (U + e)) /* via U=..-I: d(-I/(U+e))/dI = -1/(U+e) - I/(U+e)^2 */
                                  + 1.0 / (C + e);                        /* d((C-U)/(C+e))/dI = +1/(C+e) */
                double dterm_dpa = I / ((U + e) * (U + e)) - 1.0 / (C + e);
                double dterm_dC = (U + e) / ((C + e) * (C + e));          /* d((C-U)/(C+e))/dC = (U+e)/(C+e)^2 */
                if (!inter_is_tensor) dterm_dI = 0.0;                      /* python float: no graph */
                g_parea = dterm_dpa;
                if (inter_is_tensor) {
                    double sg = ssum > 0.0f ? 0.5 : (ssum < 0.0f ? -0.5 : 0.0);
                    for (i = 0; i < m; ++i) if (src[i] >= 0) {
                        int nx = (i + 1) % m, pv = (i + m - 1) % m;
                        gc[src[i]][0] += dterm_dI * sg * ((double)poly[nx][1] - poly[pv][1]);
                        gc[src[i]][1] += dterm_dI * sg * ((double)poly[pv][0] - poly[nx][0]);
                    }
                }
                if (hn >= 3) {
                    double sg = hsum > 0.0f ? 0.5 : (hsum < 0.0f ? -0.5 : 0.0);
                    for (i = 0; i < hn; ++i) if (hid[i] < 4) {
                        int nx = (i + 1) % hn, pv = (i + hn - 1) % hn;
                        gc[hid[i]][0] += dterm_dC * sg * ((double)hp[nx][1] - hp[pv][1]);
                        gc[hid[i]][1] += dterm_dC * sg * ((double)hp[pv][0] - hp[nx][0]);
                    }
                }
            }
        } else {
            inter = (float)clip_area64(pc, tc);                        /* shapely .area, python float */
            uni = (p_area + t_area) - inter;
            iou = (1.0f / (uni + 1e-16f)) * inter;                     /* __rtruediv__ */
            term = 1.0f - iou;                                         /* :135 */
            if (grad6) {
                double I = inter, U = uni, e = 1e-16;
                g_parea = I / ((U + e) * (U + e));
            }
        }
        iou_out[k] = iou;
        term_out[k] = term;

        if (grad6) {
            /* corners -> (x, y, w, l, yaw) -> (im, re);  utils/iou_rotated_boxes_utils.py:46-59 */
            double cs = cos((double)pyaw), sn = sin((double)pyaw);
            double w = P[2], l = P[3];
            /* signs (sw, sl): x_i = x + sw*(w/2)cs + slx*(l/2)sn ; y_i = y + sw*(w/2)sn + sly*(l/2)cs */
            static const double SW[4] = {-1, -1, 1, 1};
            static const double SLX[4] = {-1, 1, 1, -1};
            static const double SLY[4] = {1, -1, -1, 1};
            double gx = 0, gy = 0, gw = 0, gl = 0, gyaw = 0;
            int i;
            for (i = 0; i < 4; ++i) {
                double dx = gc[i][0], dy = gc[i][1];
                gx += dx; gy += dy;
                gw += dx * SW[i] * 0.5 * cs + dy * SW[i] * 0.5 * sn;
                gl += dx * SLX[i] * 0.5 * sn + dy * SLY[i] * 0.5 * cs;
                gyaw += dx * (-SW[i] * 0.5 * w * sn + SLX[i] * 0.5 * l * cs)
                      + dy * (SW[i] * 0.5 * w * cs - SLY[i] * 0.5 * l * sn);
            }
            gw += g_parea * l;
            gl += g_parea * w;
            double im = P[4], re = P[5], r2 = im * im + re * re;
            double gim = gyaw * (re / r2), gre = gyaw * (-im / r2);
            float *G = grad6 + 6 * k;
            G[0] = (float)gx; G[1] = (float)gy; G[2] = (float)gw; G[3] = (float)gl; G[4] = (float)gim; G[5] = (float)gre;
        }
    }
}

/* fp64 "truth" for reporting (not reference-compatible): exact intersection, exact hull. */
void orc_rgiou_pairs_exact64(const float *pred6, const float *tgt6, int64_t n, double *iou_out, double *term_out)
{
    int64_t k;
    for (k = 0; k < n; ++k) {
        const float *P = pred6 + 6 * k, *T = tgt6 + 6 * k;
        float pc[4][2], tc[4][2];
        orc_corners(T[0], T[1], T[2], T[3], atan2f(T[4], T[5]), tc);
        orc_corners(P[0], P[1], P[2], P[3], atan2f(P[4], P[5]), pc);
        double I = clip_area64(pc, tc);
        double U = (double)P[2] * P[3] + (double)T[2] * T[3] - I;
        float all8[8][2]; int hid[8], i;
        for (i = 0; i < 4; ++i) { all8[i][0] = pc[i][0]; all8[i][1] = pc[i][1]; all8[4 + i][0] = tc[i][0]; all8[4 + i][1] = tc[i][1]; }
        int hn = hull8(all8, hid);
        double s = 0.0;
        for (i = 0; i < hn; ++i) {
            int j = (i + 1) % hn;
            s += (double)all8[hid[i]][0] * all8[hid[j]][1] - (double)all8[hid[i]][1] * all8[hid[j]][0];
        }
        double C = fabs(s) * 0.5;
        iou_out[k] = I / (U + 1e-16);
        term_out[k] = 1.0 - (iou_out[k] - (C - U) / (C + 1e-16));
    }
}
