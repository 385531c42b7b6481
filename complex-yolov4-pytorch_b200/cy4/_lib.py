"""ctypes loader for csrc/libcy4.so (the C-ABI declared in include/cy4.h).

There is NO fallback: if the library is missing or no sm_100 device is usable, every op raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CY4_LIB_NAME selects a side build of the same sources (tools/probe_pipeline.py); the default is the product library
SO_PATH = os.path.normpath(os.path.join(_HERE, "..", "csrc", os.environ.get("CY4_LIB_NAME", "libcy4.so")))

c_f = ctypes.c_void_p        # device pointers travel as void*
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_u32 = ctypes.c_uint32
c_vp = ctypes.c_void_p


class YoloDesc(ctypes.Structure):
    _fields_ = [("B", c_i32), ("G", c_i32), ("nA", c_i32), ("nC", c_i32),
                ("sB", c_i64), ("sC", c_i64), ("sH", c_i64), ("sW", c_i64),
                ("img_size", ctypes.c_float), ("ignore_thresh", ctypes.c_float),
                ("use_giou", c_u32), ("reserved", c_u32)]


class BevDesc(ctypes.Structure):
    _fields_ = [("minX", ctypes.c_float), ("maxX", ctypes.c_float), ("minY", ctypes.c_float), ("maxY", ctypes.c_float),
                ("minZ", ctypes.c_float), ("maxZ", ctypes.c_float), ("discretization", ctypes.c_float), ("max_height", ctypes.c_float),
                ("H", c_i32), ("W", c_i32), ("apply_filter", c_i32), ("reserved", c_i32)]


_SIGS = {
    "cy4_version": (ctypes.c_int, []),
    "cy4_last_error": (ctypes.c_char_p, []),
    "cy4_device_ok": (ctypes.c_int, []),
    "cy4_kernel_launches": (ctypes.c_longlong, [ctypes.c_int]),
    "cy4_note_graph_replay": (ctypes.c_int, [ctypes.c_int]),
    "cy4_rgiou_pairs": (ctypes.c_int, [c_f, c_f, c_i64, c_u32, c_f, c_f, c_f, c_f, c_vp]),
    "cy4_sum_f32_seq": (ctypes.c_int, [c_f, c_i64, c_f, c_vp]),
    "cy4_corners": (ctypes.c_int, [c_f, c_f, c_f, c_f, c_f, c_i64, c_f, c_vp]),
    "cy4_quad_intersection_area": (ctypes.c_int, [c_f, c_f, c_i64, c_f, c_vp]),
    "cy4_poly_area": (ctypes.c_int, [c_f, ctypes.c_int, c_f, c_vp]),
    "cy4_anchor_iou": (ctypes.c_int, [c_f, ctypes.c_int, c_f, c_i64, c_f, c_vp]),
    "cy4_yolo_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(YoloDesc), c_i64]),
    "cy4_yolo_decode": (ctypes.c_int, [ctypes.POINTER(YoloDesc), c_f, c_f, c_f, c_vp]),
    "cy4_yolo_loss_fwd": (ctypes.c_int, [ctypes.POINTER(YoloDesc), c_f, c_f, c_f, c_i64, c_f, c_f, c_f, c_f, c_vp, c_vp]),
    "cy4_yolo_loss_bwd": (ctypes.c_int, [ctypes.POINTER(YoloDesc), c_f, c_f, c_f, c_i64, c_f, c_vp, c_f,
                                        c_i64, c_i64, c_i64, c_i64, c_vp]),
    # evaluation
    "cy4_rbox_iou_matrix": (ctypes.c_int, [c_f, c_i64, c_f, c_i64, c_f, c_vp]),
    "cy4_kmeans_iou": (ctypes.c_int, [c_f, c_i64, c_f, ctypes.c_int, c_f, c_vp]),
    "cy4_nms_max_candidates": (ctypes.c_int, []),
    "cy4_nms_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "cy4_nms_rotated_v2": (ctypes.c_int, [c_f, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, c_f,
                                         c_vp, c_vp, c_vp, c_vp]),
    "cy4_eval_max_annotations": (ctypes.c_int, []),
    "cy4_eval_match": (ctypes.c_int, [c_f, c_vp, ctypes.c_int, ctypes.c_int, c_f, c_i64, ctypes.c_float, c_vp, c_vp, c_vp]),
    # LiDAR -> BEV rasteriser
    "cy4_bev_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "cy4_bev_rasterize": (ctypes.c_int, [c_f, c_vp, ctypes.c_int, ctypes.POINTER(BevDesc), c_f, c_vp, c_vp, c_vp]),
    "cy4_build_targets": (ctypes.c_int, [ctypes.POINTER(YoloDesc), c_f, c_f, c_f, c_i64, c_f] + [c_f] * 13 +
                          [c_f, c_f, c_vp, c_vp]),
}

_lib = None


def _bind(L, table):
    for name, (res, args) in table.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    """Returns the loaded library or raises -- never falls back to another implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("cy4: %s is missing -- build it with `python __graft_entry__.py` "
                               "(there is no CPU / PyTorch fallback for this path)" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        _bind(L, _SIGS)
        from . import _sigs_engine
        _bind(L, _sigs_engine.SIGS)
        _lib = L
    return _lib


def check(rc, what=""):
    if rc < 0:
        raise RuntimeError("cy4 %s failed (%d): %s" % (what, rc, lib().cy4_last_error().decode()))


def require_device():
    if not torch.cuda.is_available():
        raise RuntimeError("cy4: no CUDA device -- the hot path has no CPU fallback")
    L = lib()
    check(L.cy4_device_ok(), "device check")


def stream():
    return torch.cuda.current_stream().cuda_stream


def p(t):
    return None if t is None else t.data_ptr()
