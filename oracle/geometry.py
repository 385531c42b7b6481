"""ctypes binding of oracle/rbox_oracle.c (the C restatement of the reference geometry).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Reference being restated:
src/utils/iou_rotated_boxes_utils.py:34-142 and src/utils/cal_intersection_rotated_boxes.py:16-96.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "rbox_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_corners.argtypes = [ctypes.c_float] * 5 + [fp]
        L.orc_poly_area.argtypes = [fp, ctypes.c_int]
        L.orc_poly_area.restype = ctypes.c_float
        L.orc_intersection_area.argtypes = [fp, fp]
        L.orc_intersection_area.restype = ctypes.c_float
        L.orc_convex_inter64.argtypes = [fp, fp]
        L.orc_convex_inter64.restype = ctypes.c_double
        L.orc_anchor_iou.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int64, fp]
        L.orc_rgiou_pairs.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_uint32, fp, fp, fp]
        L.orc_rgiou_pairs_exact64.argtypes = [fp, fp, ctypes.c_int64, dp, dp]
        L.orc_iou_matrix.argtypes = [fp, ctypes.c_int64, fp, ctypes.c_int64, fp]
        L.orc_kmeans_iou.argtypes = [dp, ctypes.c_int64, dp, ctypes.c_int, fp]
        L.orc_kmeans_iou.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def corners(x, y, w, l, yaw):
    out = np.zeros((4, 2), np.float32)
    lib().orc_corners(float(x), float(y), float(w), float(l), float(yaw), _p(out))
    return out


def poly_area(pts):
    pts = _f32(pts)
    return float(lib().orc_poly_area(_p(pts), pts.shape[0]))


def intersection_area(rect1, rect2):
    r1, r2 = _f32(rect1), _f32(rect2)
    return float(lib().orc_intersection_area(_p(r1), _p(r2)))


def convex_inter64(rect1, rect2):
    r1, r2 = _f32(rect1), _f32(rect2)
    return float(lib().orc_convex_inter64(_p(r1), _p(r2)))


def anchor_iou(anchors4, targets4):
    a, t = _f32(anchors4).reshape(-1, 4), _f32(targets4).reshape(-1, 4)
    out = np.zeros((a.shape[0], t.shape[0]), np.float32)
    lib().orc_anchor_iou(_p(a), a.shape[0], _p(t), t.shape[0], _p(out))
    return out


def rgiou_pairs(pred6, tgt6, giou=True, with_grad=False):
    """Returns (iou[n], term[n]) or (iou, term, grad[n,6])."""
    p, t = _f32(pred6).reshape(-1, 6), _f32(tgt6).reshape(-1, 6)
    assert p.shape == t.shape
    n = p.shape[0]
    iou = np.zeros(n, np.float32)
    term = np.zeros(n, np.float32)
    grad = np.zeros((n, 6), np.float32) if with_grad else None
    lib().orc_rgiou_pairs(_p(p), _p(t), n, 1 if giou else 0, _p(iou), _p(term),
                          _p(grad) if with_grad else None)
    return (iou, term, grad) if with_grad else (iou, term)


def rgiou_pairs_exact64(pred6, tgt6):
    p, t = _f32(pred6).reshape(-1, 6), _f32(tgt6).reshape(-1, 6)
    n = p.shape[0]
    iou = np.zeros(n, np.float64)
    term = np.zeros(n, np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().orc_rgiou_pairs_exact64(_p(p), _p(t), n, iou.ctypes.data_as(dp), term.ctypes.data_as(dp))
    return iou, term


def iou_matrix(a6, b6):
    """IoU of every box of a6 [n,6] with every box of b6 [m,6] -> [n,m] fp32 (evaluation_utils.py:186-210)."""
    a, b = _f32(a6).reshape(-1, 6), _f32(b6).reshape(-1, 6)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if a.shape[0] and b.shape[0]:
        lib().orc_iou_matrix(_p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def kmeans_iou(boxes3, clusters3):
    """find_anchors.py:53-59 for all pairs: [n,3] x [k,3] float64 (w, l, yaw) -> float32 [n,k]."""
    b = np.ascontiguousarray(np.asarray(boxes3, np.float64)).reshape(-1, 3)
    c = np.ascontiguousarray(np.asarray(clusters3, np.float64)).reshape(-1, 3)
    out = np.zeros((b.shape[0], c.shape[0]), np.float32)
    dp = ctypes.POINTER(ctypes.c_double)
    if b.shape[0] and c.shape[0]:
        lib().orc_kmeans_iou(b.ctypes.data_as(dp), b.shape[0], c.ctypes.data_as(dp), c.shape[0], _p(out))
    return out
