"""CPU: the anchor k-means oracle (oracle/anchors_oracle.py + orc_kmeans_iou) against a golden run
of the UNMODIFIED reference Find_Anchors (src/utils/find_anchors.py:53-105, shapely stand-in) on 300 seeded KITTI-like
boxes, 6 anchors: same IoUs (float32), same assignment history (iteration count), identical final clusters."""
import numpy as np


def test_kmeans_oracle_matches_reference_golden(golden):
    from oracle import anchors_oracle as ao, geometry as og
    g = golden("anchors_kmeans.npz")
    boxes, k = g["boxes"], int(g["num_anchors"])
    iou = og.kmeans_iou(boxes[:60], g["cluster"])
    assert np.abs(iou - g["iou_first60"]).max() <= 1.2e-7          # float64 ratio rounded to float32: at most 1 ulp
    cluster, loops = ao.kmeans(boxes, k)
    assert loops == int(g["loops"]) and np.array_equal(cluster, g["cluster"])
    assert abs(ao.avg_iou(boxes, cluster) - float(g["avg_iou"])) <= 1e-6
    assert (cluster[:, 2] == 0).all()                               # anchors keep yaw 0 (find_anchors.py:77,101)


def test_kmeans_iou_known_answers():
    from oracle import geometry as og
    b = np.array([[4.0, 8.0, 0.0], [4.0, 8.0, np.pi / 2], [10.0, 20.0, 0.3]])
    m = og.kmeans_iou(b, b)
    assert np.allclose(np.diag(m), 1.0, atol=1e-6)
    assert abs(m[0, 1] - 16.0 / (32 + 32 - 16)) <= 1e-6            # 4x8 against its 90-degree rotation: 4x4 overlap
    assert np.allclose(m, m.T, atol=1e-7)
