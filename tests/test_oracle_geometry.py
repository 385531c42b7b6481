"""CPU: the oracle (oracle/rbox_oracle.c) against the golden vectors generated from the reference
itself (oracle/gen_golden.py) and against the known answers quoted in SURVEY.md section 4 / 8c."""
import numpy as np
import pytest

from oracle import geometry as og

# The reference's own fp32 noise on these quantities is ~1e-4 (torch/Sleef trig vs libm
# differ in the last ulp for ~5-15% of inputs, the hull start vertex of Qhull is unspecified).
TOL_IOU = 1e-4
TOL_TERM = 1e-4


def test_known_answers(golden):
    g = golden("rgiou_pairs.npz")
    iou, term = og.rgiou_pairs(g["ka_pred"], g["ka_tgt"], giou=True)
    np.testing.assert_allclose(iou, g["ka_iou"], atol=2e-6)
    np.testing.assert_allclose(term, g["ka_term"], atol=2e-6)
    # SURVEY section 4 / 8c values
    np.testing.assert_allclose(iou[0], 0.366509, atol=2e-6)
    np.testing.assert_allclose(term[0], 0.902670, atol=2e-6)
    np.testing.assert_allclose(iou[1], 0.2, atol=1e-6)
    np.testing.assert_allclose(term[1], 1.0307692, atol=2e-6)
    np.testing.assert_allclose(iou[2], 1.0, atol=1e-6)
    np.testing.assert_allclose(term[2], 0.0, atol=1e-6)
    np.testing.assert_allclose(iou[3], 1.0 / 3.0, atol=1e-6)
    np.testing.assert_allclose(term[3], 2.0 / 3.0, atol=1e-6)
    # F5: disjoint boxes, the reference clipper returns the whole pred box
    np.testing.assert_allclose(iou[4], 1.0000002, atol=1e-6)
    np.testing.assert_allclose(term[4], 0.8850825, atol=2e-6)


def test_pairs_values_and_grads(golden):
    g = golden("rgiou_pairs.npz")
    iou, term, grad = og.rgiou_pairs(g["pred"], g["tgt"], giou=True, with_grad=True)
    assert np.abs(iou - g["iou"]).max() <= TOL_IOU
    assert np.abs(term - g["term"]).max() <= TOL_TERM
    assert (iou == g["iou"]).mean() > 0.9          # bit-identical for the vast majority
    d = np.abs(grad - g["grad"])
    assert (d / (np.abs(g["grad"]) + 1e-2)).max() < 5e-3
    # batched call == sum over pairs, shape [1]
    np.testing.assert_allclose(term[:64].astype(np.float64).sum(), g["batch64_loss"][0], rtol=2e-6)
    np.testing.assert_allclose(grad[:64], g["batch64_grad"], atol=2e-5, rtol=5e-3)


def test_shapely_path(golden):
    g = golden("rgiou_pairs.npz")
    n = len(g["shapely_iou"])
    iou, term, grad = og.rgiou_pairs(g["pred"][:n], g["tgt"][:n], giou=False, with_grad=True)
    np.testing.assert_allclose(iou, g["shapely_iou"], atol=1e-6)
    np.testing.assert_allclose(term.astype(np.float64).sum(), g["shapely_loss"][0], rtol=1e-5)
    np.testing.assert_allclose(grad, g["shapely_grad"], atol=1e-6, rtol=1e-3)


def test_anchor_iou(golden):
    g = golden("anchor_iou.npz")
    for G in (76, 38, 19):
        got = og.anchor_iou(g[f"anchors_{G}"], g[f"tboxes_{G}"])
        ref = g[f"ious_{G}"]
        np.testing.assert_allclose(got, ref, atol=2e-6)
        assert (np.argmax(got, 0) == np.argmax(ref, 0)).all()      # best_n bit-exact
        assert ((got > np.float32(0.7)) == (ref > np.float32(0.7))).all()
    got = og.anchor_iou(g["ka_anchors"], g["ka_tboxes"])
    np.testing.assert_allclose(got, g["ka_ious"], atol=2e-6)
    np.testing.assert_allclose(got, [[.1407, .5521], [.2046, .3607], [.2344, .3679]], atol=6e-5)


def test_exact64_vs_cv2():
    """Third, independent implementation: cv2.intersectConvexConvex (fp32 inside)."""
    cv2 = pytest.importorskip("cv2")
    from cy4 import synth
    pred, tgt = synth.make_pairs(200, seed=11)
    for k in range(200):
        pc = og.corners(*pred[k, :4], np.arctan2(pred[k, 4], pred[k, 5]))
        tc = og.corners(*tgt[k, :4], np.arctan2(tgt[k, 4], tgt[k, 5]))
        a, _ = cv2.intersectConvexConvex(pc, tc)
        assert abs(a - og.convex_inter64(pc, tc)) < 2e-3 * max(1.0, a)


def test_f7_noise_floor():
    """Reference-compatible fp32 values stay within the reference's own fp32 noise of the fp64
    truth on overlapping pairs."""
    from cy4 import synth
    pred, tgt = synth.make_pairs(5000, seed=3)
    iou, term = og.rgiou_pairs(pred, tgt, giou=True)
    ei, et = og.rgiou_pairs_exact64(pred, tgt)
    assert np.abs(iou - ei).max() < 5e-4
    assert np.abs(term - et).max() < 5e-4
