"""LiDAR point clouds -> [B,3,608,608] BEV maps on the device: the step right before the
training hot path.  Replaces the numpy lexsort / unique pipeline of the reference's
src/data_process/kitti_bev_utils.py:18-76 with two kernels (csrc/bev.cu).  No CPU implementation."""
import numpy as np
import torch

from . import _lib

# config/kitti_config.py:13-36 of the reference
BOUNDARY = {"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}
BEV_WIDTH = BEV_HEIGHT = 608
DISCRETIZATION = (BOUNDARY["maxX"] - BOUNDARY["minX"]) / BEV_HEIGHT


def _desc(bc, discretization, H, W, apply_filter):
    return _lib.BevDesc(float(bc["minX"]), float(bc["maxX"]), float(bc["minY"]), float(bc["maxY"]), float(bc["minZ"]), float(bc["maxZ"]),
                        float(discretization), float(np.abs(bc["maxZ"] - bc["minZ"])), int(H), int(W), 1 if apply_filter else 0, 0)


def rasterize(clouds, bc=BOUNDARY, discretization=DISCRETIZATION, H=BEV_HEIGHT, W=BEV_WIDTH, apply_filter=True, check=False):
    """clouds: list of [n_i,4] fp32 arrays / tensors (x, y, z, intensity), raw scans when apply_filter (removePoints is
    fused) or already cropped + z-shifted ones otherwise.  Returns a cuda tensor [B,3,H,W] fp32 (intensity, height,
    density) -- the layout Darknet.forward takes.  check=True synchronises and raises if a point fell outside the
    reference's scratch map (numpy would wrap around or raise there)."""
    _lib.require_device()
    L = _lib.lib()
    ts = [torch.as_tensor(c, dtype=torch.float32).reshape(-1, 4) for c in clouds]
    B = len(ts)
    sizes = [t.shape[0] for t in ts]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    if B and sum(sizes):
        pts = torch.cat([t if t.is_cuda else t.pin_memory().cuda(non_blocking=True) for t in ts]).contiguous()
    else:
        pts = torch.zeros(1, 4, device="cuda")
    out = torch.empty(B, 3, H, W, device="cuda", dtype=torch.float32)
    if B == 0:
        return out
    offs_d = offs.cuda()
    dropped = torch.empty(B, device="cuda", dtype=torch.int32)
    ws = torch.empty(L.cy4_bev_workspace_bytes(B, H, W), device="cuda", dtype=torch.uint8)
    d = _desc(bc, discretization, H, W, apply_filter)
    _lib.check(L.cy4_bev_rasterize(pts.data_ptr(), offs_d.data_ptr(), B, d, out.data_ptr(), dropped.data_ptr(), ws.data_ptr(), _lib.stream()),
               "bev_rasterize")
    if check and int(dropped.sum().item()):
        raise IndexError("cy4 bev rasterize: %d points outside the map (apply removePoints first)" % int(dropped.sum().item()))
    return out
