"""Seeded synthetic inputs for the hot path.  numpy PCG64 streams are
stable across machines, so tests, golden fixtures and bench.py all see the same data."""
import math

import numpy as np


def make_pairs(n, seed=7, grid=76.0, disjoint_frac=0.0):
    """n (pred, target) rotated-box pairs in grid units, as YoloLayer.build_targets feeds
    iou_pred_vs_target_boxes (reference src/models/yolo_layer.py:134).  Returns two [n,6] fp32
    arrays (x, y, w, l, im, re)."""
    rng = np.random.default_rng(seed)
    tx = rng.uniform(0, grid, n); ty = rng.uniform(0, grid, n)
    tw = rng.uniform(1.25, 3.75, n); tl = rng.uniform(1.9, 8.1, n)
    tyaw = rng.uniform(-math.pi, math.pi, n)
    tgt = np.stack([tx, ty, tw, tl, np.sin(tyaw), np.cos(tyaw)], 1)
    px = tx + rng.uniform(-0.5, 0.5, n); py = ty + rng.uniform(-0.5, 0.5, n)
    pw = tw * np.exp(rng.normal(0, 0.3, n)); pl = tl * np.exp(rng.normal(0, 0.3, n))
    pim = np.sin(tyaw) + rng.normal(0, 0.3, n); pre = np.cos(tyaw) + rng.normal(0, 0.3, n)
    if disjoint_frac > 0:
        far = rng.uniform(0, 1, n) < disjoint_frac
        px = np.where(far, px + rng.choice([-1, 1], n) * rng.uniform(12, 30, n), px)
        py = np.where(far, py + rng.choice([-1, 1], n) * rng.uniform(12, 30, n), py)
    pred = np.stack([px, py, pw, pl, pim, pre], 1)
    return pred.astype(np.float32), tgt.astype(np.float32)


def make_targets(batch, per_image=5, seed=4321, img_size=608, strides=(8, 16, 32), total=None):
    """[nT,8] fp32 targets (image, class, x, y, w, l, im, re) with x..l normalised to [0,1), as
    collate_fn emits (reference src/data_process/kitti_dataset.py:216-233).  (image, cell) is
    distinct at every stride so target assignment has no duplicate cells.  total pins nT."""
    rng = np.random.default_rng(seed)
    n = total if total is not None else batch * per_image
    rows, used = [], set()
    k = 0
    while len(rows) < n:
        b = (len(rows) // per_image) % batch if total is None else int(rng.integers(0, batch))
        cls = int(rng.integers(0, 3))
        x = rng.uniform(0.05, 0.95); y = rng.uniform(0.05, 0.95)
        cells = [(b, s, int(np.float32(x) * np.float32(img_size // s)), int(np.float32(y) * np.float32(img_size // s)))
                 for s in strides]
        k += 1
        if any(c in used for c in cells) and k < 100000:
            continue
        used.update(cells)
        w = rng.uniform(10, 30) / img_size; l = rng.uniform(15, 65) / img_size
        yaw = rng.uniform(-math.pi, math.pi)
        rows.append([b, cls, x, y, w, l, math.sin(yaw), math.cos(yaw)])
    t = np.asarray(rows, dtype=np.float32)
    return t[np.argsort(t[:, 0], kind="stable")]


def make_bev(batch, seed=1234, img_size=608, channels=3):
    """[B,3,608,608] fp32 in [0,1) -- the BEV stand-in BASELINE.json names ("synthetic BEV")."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.rand(batch, channels, img_size, img_size, generator=g, dtype=torch.float32)


def make_detections(batch, targets, n_rows=22743, n_classes=3, dup=6, clutter=40, seed=99, img_size=608,
                    conf_lo=0.5, margin=None):
    """Raw network output [B, n_rows, 7+nC] fp32 (x, y, w, l, im, re, conf, cls...) in pixels, as Darknet.forward
    returns it in eval mode (reference darknet2pytorch.py:228), for the post-processing / evaluation path:
    `dup` jittered high-confidence copies of every target (so NMS has clusters to merge, most of them true
    positives), `clutter` random confident boxes per image (false positives), the rest background with
    conf < 0.3.  Scores are distinct.  targets: make_targets() output ([nT,8], x..l normalised).
    margin=(nms_thresh, iou_thresh, eps): hook for tests that need decisions away from the thresholds (unused
    by default; tests filter with the oracle instead)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((batch, n_rows, 7 + n_classes), np.float32)
    out[:, :, 0:2] = rng.uniform(0, img_size, (batch, n_rows, 2))
    out[:, :, 2] = rng.uniform(8, 30, (batch, n_rows)); out[:, :, 3] = rng.uniform(12, 70, (batch, n_rows))
    yaw = rng.uniform(-math.pi, math.pi, (batch, n_rows))
    out[:, :, 4] = np.sin(yaw); out[:, :, 5] = np.cos(yaw)
    out[:, :, 6] = rng.uniform(0.0, 0.3, (batch, n_rows))
    out[:, :, 7:] = rng.uniform(0.0, 1.0, (batch, n_rows, n_classes))
    for b in range(batch):
        tg = targets[targets[:, 0] == b]
        rows = rng.choice(n_rows, size=len(tg) * dup + clutter, replace=False)
        confs = rng.permutation(np.linspace(conf_lo + 0.01, 0.999, len(rows))).astype(np.float32)     # distinct
        k = 0
        for t in tg:
            cls = int(t[1])
            x, y, w, l = t[2] * img_size, t[3] * img_size, t[4] * img_size, t[5] * img_size
            tyaw = math.atan2(t[6], t[7])
            for _ in range(dup):
                r = rows[k]
                jy = tyaw + rng.normal(0, 0.06)
                amp = rng.uniform(0.9, 1.1)                      # im/re are not normalised by the network
                out[b, r, :6] = [x + rng.normal(0, 1.5), y + rng.normal(0, 1.5), w * math.exp(rng.normal(0, 0.06)),
                                 l * math.exp(rng.normal(0, 0.06)), amp * math.sin(jy), amp * math.cos(jy)]
                out[b, r, 6] = confs[k]
                pc = rng.uniform(0.0, 0.3, n_classes); pc[cls if rng.uniform() < 0.9 else (cls + 1) % n_classes] = rng.uniform(0.7, 1.0)
                out[b, r, 7:] = pc
                k += 1
        for _ in range(clutter):
            r = rows[k]
            out[b, r, 6] = confs[k]
            pc = rng.uniform(0.0, 0.3, n_classes); pc[int(rng.integers(0, n_classes))] = rng.uniform(0.7, 1.0)
            out[b, r, 7:] = pc
            k += 1
    return out


def make_point_cloud(n=120000, seed=5, ties=True):
    """[n,4] fp32 (x, y, z, intensity) LiDAR-like frame for the BEV rasteriser: ~60 % of the points inside the front
    KITTI boundary (0..50 m, -25..25 m, -2.73..1.27 m) with a 1/r density, ground plane + object clusters, the rest
    outside (to exercise removePoints); `ties` adds exact duplicates of heights within a cell and points exactly on
    the inclusive bounds."""
    rng = np.random.default_rng(seed)
    r = 3.0 + 70.0 * rng.uniform(0, 1, n) ** 2
    th = rng.uniform(-math.pi, math.pi, n)
    x, y = r * np.cos(th), r * np.sin(th)
    z = -1.7 + 0.05 * rng.normal(size=n)
    obj = rng.uniform(size=n) < 0.25
    z[obj] = rng.uniform(-1.7, 1.5, obj.sum())
    out_z = rng.uniform(size=n) < 0.02
    z[out_z] = rng.uniform(-4.0, 3.0, out_z.sum())
    pts = np.stack([x, y, z, rng.uniform(0, 1, n)], 1).astype(np.float32)
    if ties:
        k = min(2000, n // 10)
        src = rng.choice(n, k, replace=False); dst = rng.choice(n, k, replace=False)
        pts[dst, :3] = pts[src, :3]                                  # same cell, same height, different intensity
        edge = rng.choice(n, 64, replace=False)
        pts[edge[:16], 0] = 0.0; pts[edge[16:32], 0] = 50.0; pts[edge[32:40], 1] = -25.0; pts[edge[40:48], 1] = 25.0
        pts[edge[48:56], 2] = np.float32(-2.73); pts[edge[56:], 2] = np.float32(1.27)
    return pts
