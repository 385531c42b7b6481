"""CPU restatement of the reference's anchor k-means (src/utils/find_anchors.py:53-105) -- TEST INFRASTRUCTURE ONLY.
compute_iou / avg_iou / kmeans with the (box, cluster) IoU matrix from the C oracle (oracle/rbox_oracle.c orc_kmeans_iou);
everything else is the reference's numpy code path: seeded initial choice (:73-75), yaw-0 clusters (:77,:101), float64
distance matrix, first arg-min assignment (:93), stop when the assignment repeats (:95), per-cluster median update (:99-100)."""
import numpy as np

from . import geometry as og


def avg_iou(boxes_wh, cluster):
    return float(np.mean(np.max(og.kmeans_iou(boxes_wh, cluster), axis=1)))


def kmeans(boxes_wh, num_anchors, seed=0, max_iter=10000):
    boxes = np.asarray(boxes_wh, dtype=np.float64)
    n = boxes.shape[0]
    distance = np.empty((n, num_anchors))
    last_clu = np.zeros((n,))
    np.random.seed(seed)
    cluster = boxes[np.random.choice(n, num_anchors, replace=False)].copy()
    cluster[:, 2] = 0
    loops = 0
    while loops < max_iter:
        loops += 1
        distance[:] = 1 - og.kmeans_iou(boxes, cluster)
        near = np.argmin(distance, axis=1)
        if (last_clu == near).all():
            break
        for j in range(num_anchors):
            cluster[j] = np.median(boxes[near == j], axis=0)
        cluster[:, 2] = 0
        last_clu = near
    return cluster, loops
