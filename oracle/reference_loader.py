"""Import the UNMODIFIED reference modules: from /root/reference/src in the build container, else from the
byte-identical copy oracle/_ref/src that build() makes there (oracle/make_ref.py; git-ignored, travels to the
GPU box with the snapshot).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py, the live cross-checks in tests/ (skipped
when no reference tree is available), bench.py --impl reference and the GIoU micro-benchmark's CPU leg.
"""
import importlib
import os
import sys

from . import make_ref

REF_SRC = make_ref.ref_src() or make_ref.SRC


def available():
    return os.path.isdir(REF_SRC)


def load():
    """Returns a dict of the reference modules on the hot path."""
    from . import shapely_standin
    kind = shapely_standin.install()
    # The reference imports its own top-level packages `models`, `utils`.  Make sure ours
    # (complex-yolov4-pytorch_b200/) do not shadow them while loading, then restore.
    saved_path = list(sys.path)
    pk = ("models", "utils", "data_process", "config")
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k in pk or k.split(".")[0] in pk}
    sys.path = [REF_SRC] + [p for p in sys.path if "complex-yolov4-pytorch_b200" not in p]
    try:
        mods = {
            "clip": importlib.import_module("utils.cal_intersection_rotated_boxes"),
            "iou": importlib.import_module("utils.iou_rotated_boxes_utils"),
            "yolo": importlib.import_module("models.yolo_layer"),
            "darknet": importlib.import_module("models.darknet2pytorch"),
            "eval": importlib.import_module("utils.evaluation_utils"),        # section 8 row f1 (pulls in data_process.kitti_bev_utils, row f3)
        }
    finally:
        for k in list(sys.modules):
            if k in pk or k.split(".")[0] in pk:
                sys.modules["_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved_mods)
        sys.path = saved_path
    mods["shapely_kind"] = kind
    return mods
