"""Anchor k-means on rotated boxes: the algorithm of the reference's
src/utils/find_anchors.py:53-105 (Find_Anchors.compute_iou / avg_iou / kmeans) with the box-vs-cluster IoU matrix -- the
reference's per-pair shapely loop, num_boxes x num_anchors polygon intersections per iteration -- computed by one kernel
launch (cy4_kmeans_iou, csrc/nms.cu).  Assignment (first arg-min), the per-cluster median update, the yaw-0 clusters, the
seeded initial choice and the stopping rule are the reference's, so the same boxes give the same anchors.  No CPU path."""
import numpy as np
import torch

from . import _lib


def iou_matrix(boxes_wh_yaw, clusters_wh_yaw):
    """[n,3] x [k,3] (w, l, yaw; float64) -> float32 [n,k] on the device (find_anchors.py:53-59 for all boxes)."""
    _lib.require_device()
    L = _lib.lib()
    def dev64(a):
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
        return a.to("cuda", torch.float64).reshape(-1, 3).contiguous()
    b, c = dev64(boxes_wh_yaw), dev64(clusters_wh_yaw)
    out = torch.empty(b.shape[0], c.shape[0], device=b.device, dtype=torch.float32)
    _lib.check(L.cy4_kmeans_iou(b.data_ptr(), b.shape[0], c.data_ptr(), c.shape[0], out.data_ptr(), _lib.stream()), "kmeans_iou")
    return out


def avg_iou(boxes_wh_yaw, clusters_wh_yaw):
    """find_anchors.py:61-62: mean over boxes of the best IoU with any cluster (float32 IoUs, float64 mean like np.mean)."""
    return float(iou_matrix(boxes_wh_yaw, clusters_wh_yaw).max(1).values.double().mean())


def kmeans(boxes_wh_yaw, num_anchors, seed=0, max_iter=10000, verbose=False):
    """find_anchors.py:64-105.  boxes_wh_yaw: float64 [n,3] rows (int(w px), int(l px), yaw).  Returns (clusters [k,3]
    float64 in the reference's order, number of iterations).  The initial centres are boxes drawn with
    np.random.seed(seed); np.random.choice(n, k, replace=False); cluster yaw is forced to 0 (:77,:101)."""
    boxes = np.asarray(boxes_wh_yaw, dtype=np.float64)
    n = boxes.shape[0]
    last = np.zeros((n,))
    np.random.seed(seed)
    cluster = boxes[np.random.choice(n, num_anchors, replace=False)].copy()
    cluster[:, 2] = 0
    loops = 0
    boxes_dev = torch.as_tensor(boxes).cuda()
    while loops < max_iter:
        loops += 1
        if verbose:
            print("iteration %d:" % loops, ", ".join("[%d, %d, %.0f]" % (int(w), int(h), y) for w, h, y in cluster))
        dist = 1 - iou_matrix(boxes_dev, cluster).cpu().numpy().astype(np.float64)     # `distance` is a float64 array (:66)
        near = np.argmin(dist, axis=1)
        if (last == near).all():
            break
        for j in range(num_anchors):
            cluster[j] = np.median(boxes[near == j], axis=0)
        cluster[:, 2] = 0
        last = near
    return cluster, loops
