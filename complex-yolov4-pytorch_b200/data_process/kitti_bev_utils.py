"""Drop-in for the reference's src/data_process/kitti_bev_utils.py: removePoints,
makeBVFeature and build_yolo_target with the reference's signatures; the rasterisation runs on the GPU
(cy4/bevops.py -> csrc/bev.cu).  Every other public name of the reference module (label reading, drawing helpers)
is re-exported from the reference's own file when that tree is importable, so kitti_dataset.py keeps working."""
import importlib.util
import math
import os
import sys

import numpy as np

from cy4 import bevops


def _load_reference_module():
    here = os.path.dirname(os.path.abspath(__file__))
    for p in list(sys.path):
        cand = os.path.join(p, "data_process", "kitti_bev_utils.py")
        if os.path.exists(cand) and os.path.dirname(os.path.abspath(cand)) != here:
            spec = importlib.util.spec_from_file_location("_cy4_ref_kitti_bev_utils", cand)
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except Exception:            # its own imports (cv2, config) are not our concern
                return None
            return mod
    return None


_ref = _load_reference_module()
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("_")})


def removePoints(PointCloud, BoundaryCond):
    """Points inside the inclusive boundary, z shifted by -minZ (reference :18-36).  Host-side and cheap; the batched
    device path (cy4.bevops.rasterize(apply_filter=True)) fuses it into the scatter kernel."""
    c = np.asarray(PointCloud)
    b = BoundaryCond
    keep = ((c[:, 0] >= b["minX"]) & (c[:, 0] <= b["maxX"]) & (c[:, 1] >= b["minY"]) & (c[:, 1] <= b["maxY"]) &
            (c[:, 2] >= b["minZ"]) & (c[:, 2] <= b["maxZ"]))
    out = c[keep]
    out[:, 2] = out[:, 2] - b["minZ"]
    return out


def makeBVFeature(PointCloud_, Discretization, bc):
    """[n,4] cropped + shifted points -> float64 [3, 608, 608] (intensity, height, density), reference :39-76.
    The density channel carries float32 precision (the reference rounds it to float32 right after, kitti_dataset.py:113)."""
    import torch.utils.data as tud
    if tud.get_worker_info() is not None:
        # KittiDataset.__getitem__ runs in forked DataLoader workers (train.py defaults to num_workers=4) and a forked
        # child cannot use the parent's CUDA context.  This drop-in therefore does NOT replace the per-sample call inside
        # workers: it hands it back to the reference's own function.  The GPU rasteriser is for main-process / batched use
        # (num_workers=0, or cy4.bevops.rasterize(list_of_scans) on the whole batch).
        if _ref is None:
            raise RuntimeError("data_process.kitti_bev_utils.makeBVFeature was called inside a DataLoader worker process: the GPU "
                               "rasteriser cannot run in a forked worker and the reference module is not importable to take the "
                               "call -- use num_workers=0 or rasterise the batch in the main process (cy4.bevops.rasterize)")
        return _ref.makeBVFeature(PointCloud_, Discretization, bc)
    rgb = bevops.rasterize([np.asarray(PointCloud_, np.float32)], bc, Discretization, bevops.BEV_HEIGHT, bevops.BEV_WIDTH, apply_filter=False)
    return rgb[0].cpu().numpy().astype(np.float64)


def build_yolo_target(labels):
    """[n,8] (cls, x, y, z, h, w, l, yaw) -> [k,7] float32 (cls, y, x, w, l, im, re) normalised to the BEV boundary; boxes
    whose centre is outside are dropped; +0.3 m on w and l (reference :122-138)."""
    bc = bevops.BOUNDARY
    sx, sy = bc["maxX"] - bc["minX"], bc["maxY"] - bc["minY"]
    rows = []
    for cl, x, y, _z, _h, w, l, yaw in np.asarray(labels):
        if not (bc["minX"] < x < bc["maxX"] and bc["minY"] < y < bc["maxY"]):
            continue
        ang = float(np.pi * 2 - yaw)
        rows.append([cl, (y - bc["minY"]) / sy, (x - bc["minX"]) / sx, (w + 0.3) / sy, (l + 0.3) / sx, math.sin(ang), math.cos(ang)])
    return np.array(rows, dtype=np.float32)
