"""GPU: anchor k-means -- cy4.anchors / the drop-in utils.find_anchors.Find_Anchors on the device
against the reference golden (tests/golden/anchors_kmeans.npz, produced by the unmodified src/utils/find_anchors.py) and
against the C oracle on a larger set.  The IoU matrix is float64 arithmetic rounded to float32 on both sides: bit-exact
except where the fp64 sincos / clip of device and host libm differ in the last bit (<= 1 float32 ulp allowed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_kmeans_matches_reference_golden(golden):
    from cy4 import anchors as an
    g = golden("anchors_kmeans.npz")
    boxes, k = g["boxes"], int(g["num_anchors"])
    iou = an.iou_matrix(boxes[:60], g["cluster"]).cpu().numpy()
    assert np.abs(iou - g["iou_first60"]).max() <= 1.2e-7
    cluster, loops = an.kmeans(boxes, k)
    assert loops == int(g["loops"]) and np.array_equal(cluster, g["cluster"])
    assert abs(an.avg_iou(boxes, cluster) - float(g["avg_iou"])) <= 1e-6


def test_dropin_find_anchors_class(golden):
    import sys, os
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
    from utils.find_anchors import Find_Anchors
    g = golden("anchors_kmeans.npz")
    s = Find_Anchors.from_boxes(g["boxes"])
    s.kmeans(int(g["num_anchors"]))
    assert s.loop_cnt == int(g["loops"]) and np.array_equal(s.cluster, g["cluster"])
    assert np.abs(s.compute_iou(7) - g["iou_first60"][7]).max() <= 1.2e-7
    assert abs(s.avg_iou() - float(g["avg_iou"])) <= 1e-6


def test_iou_matrix_vs_oracle_large():
    from cy4 import anchors as an
    from oracle import geometry as og
    rng = np.random.RandomState(11)
    n = 20000
    boxes = np.stack([rng.randint(6, 40, n).astype(float), rng.randint(8, 90, n).astype(float), rng.uniform(-np.pi, np.pi, n)], 1)
    clusters = np.stack([rng.randint(6, 40, 9).astype(float), rng.randint(8, 90, 9).astype(float), np.zeros(9)], 1)
    got = an.iou_matrix(boxes, clusters).cpu().numpy()
    ref = og.kmeans_iou(boxes, clusters)
    d = np.abs(got - ref)
    assert d.max() <= 1.2e-7 and (d > 0).mean() < 0.02
    assert (got.argmax(1) == ref.argmax(1)).mean() > 0.9999
