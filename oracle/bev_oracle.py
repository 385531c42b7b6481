"""CPU restatement of the reference's LiDAR -> BEV rasteriser and target builder.
TEST INFRASTRUCTURE ONLY.  Follows /root/reference/src/data_process/kitti_bev_utils.py:
  removePoints :18-36, makeBVFeature :39-76, build_yolo_target :122-138
with config/kitti_config.py:13-36 (boundary, 608 x 608 cells, discretization 50/608).
Pinned against the unmodified reference by oracle/gen_golden.py (tests/golden/bev_raster.npz)."""
import math

import numpy as np

BOUNDARY = {"minX": 0, "maxX": 50, "minY": -25, "maxY": 25, "minZ": -2.73, "maxZ": 1.27}       # kitti_config.py:14-21
BEV_H = BEV_W = 608                                                                             # :33-34
DISCRETIZATION = (BOUNDARY["maxX"] - BOUNDARY["minX"]) / BEV_H                                 # :36


def remove_points(cloud, bc=BOUNDARY):
    """inclusive crop, then z -= minZ in float32 (:28-34)."""
    c = np.asarray(cloud, np.float32)
    keep = ((c[:, 0] >= bc["minX"]) & (c[:, 0] <= bc["maxX"]) & (c[:, 1] >= bc["minY"]) & (c[:, 1] <= bc["maxY"]) &
            (c[:, 2] >= bc["minZ"]) & (c[:, 2] <= bc["maxZ"]))
    c = c[keep].copy()
    c[:, 2] = c[:, 2] - bc["minZ"]
    return c


def make_bv_feature(cloud, disc=DISCRETIZATION, bc=BOUNDARY, H=BEV_H, W=BEV_W):
    """[n,4] float32 (x, y, z-shifted, intensity) -> [3, H, W] float64 (intensity, height, density).
    Per cell: the point with the greatest z, the earliest one among equals (lexsort is stable, :49-50,:55);
    height = z / float(|maxZ - minZ|) in float32 (:58-59); density = min(1, ln(count + 1) / ln 64) (:67)."""
    c = np.asarray(cloud, np.float32)
    ix = np.floor(c[:, 0] / np.float32(disc)).astype(np.int64)                                   # :45
    iy = (np.floor(c[:, 1] / np.float32(disc)) + np.float32((W + 1) / 2)).astype(np.int64)       # :46 (astype truncates)
    inten = np.zeros((H + 1, W + 1)); height = np.zeros((H + 1, W + 1)); dens = np.zeros((H + 1, W + 1))
    cell = ix * (W + 1) + iy
    order = np.lexsort((np.arange(len(c)), -c[:, 2], cell))        # by cell, highest z first, file order among ties
    cs = cell[order]
    first = np.ones(len(cs), bool); first[1:] = cs[1:] != cs[:-1]
    top = order[first]
    counts = np.diff(np.append(np.nonzero(first)[0], len(cs)))
    max_height = float(np.abs(bc["maxZ"] - bc["minZ"]))
    height[ix[top], iy[top]] = c[top, 2] / np.float32(max_height)
    inten[ix[top], iy[top]] = c[top, 3]
    dens[ix[top], iy[top]] = np.minimum(1.0, np.log(counts + 1) / np.log(64))
    out = np.zeros((3, H, W))
    out[0], out[1], out[2] = inten[:H, :W], height[:H, :W], dens[:H, :W]
    return out


def build_yolo_target(labels, bc=BOUNDARY):
    """[n,8] (cls, x, y, z, h, w, l, yaw) lidar-frame labels -> [k,7] float32 (cls, y1, x1, w1, l1, im, re) (:122-138)."""
    rows = []
    for cl, x, y, z, h, w, l, yaw in np.asarray(labels):
        l, w = l + 0.3, w + 0.3
        yaw = np.pi * 2 - yaw
        if bc["minX"] < x < bc["maxX"] and bc["minY"] < y < bc["maxY"]:
            sy, sx = bc["maxY"] - bc["minY"], bc["maxX"] - bc["minX"]
            rows.append([cl, (y - bc["minY"]) / sy, (x - bc["minX"]) / sx, w / sy, l / sx, math.sin(float(yaw)), math.cos(float(yaw))])
    return np.array(rows, dtype=np.float32)
