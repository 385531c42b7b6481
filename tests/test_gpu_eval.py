"""GPU: the evaluation post-processing kernels (csrc/nms.cu through the C-ABI and the drop-in utils.evaluation_utils)
against the reference fixture and against the CPU oracle at the full output size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BOX_TOL = 2e-4       # merged boxes, pixels up to 608 (1 ulp = 6.1e-5); sums are sequential fp32 in kernel and oracle
IOU_TOL = 1e-5       # fp64 intersection of fp32 corners; only the trig can differ in the last ulp


def _check_against(outs, ref_outs, tol=BOX_TOL):
    assert [o is None for o in outs] == [r is None for r in ref_outs]
    for o, r in zip(outs, ref_outs):
        if r is None:
            continue
        o = o.cpu().numpy()
        assert o.shape == r.shape
        assert np.array_equal(o[:, 6:], r[:, 6:])
        assert np.abs(o[:, :6] - r[:, :6]).max() <= tol


def test_iou_matrix_vs_reference_and_oracle(golden):
    from cy4 import evalops
    from oracle import geometry as og
    import utils.evaluation_utils as ev
    g = golden("eval_nms_b4.npz")
    o0 = g["out_0"]
    t0 = g["targets_px"][g["targets_px"][:, 0] == 0][:, 2:]
    m = evalops.iou_matrix(o0[:, :6], t0).cpu().numpy()
    assert np.abs(m - g["iou_det_vs_tgt"]).max() <= IOU_TOL
    one = ev.iou_rotated_single_vs_multi_boxes_cpu(torch.tensor(o0[3, :6]), torch.tensor(o0[:, :6]))
    assert one.dtype == torch.float32 and not one.is_cuda
    assert np.abs(one.numpy() - g["iou_det_vs_det"][3]).max() <= IOU_TOL
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.uniform(0, 608, (400, 2)), rng.uniform(5, 80, (400, 2)), rng.normal(0, 1, (400, 2))], 1).astype(np.float32)
    b = a[rng.permutation(400)[:300]] + rng.normal(0, 3, (300, 6)).astype(np.float32)
    b[:, 2:4] = np.abs(b[:, 2:4]) + 1
    got = evalops.iou_matrix(a, b).cpu().numpy()
    ref = og.iou_matrix(a, b)
    assert np.abs(got - ref).max() <= IOU_TOL
    assert (got == ref).mean() > 0.9
    assert evalops.iou_matrix(a[:0], b).shape == (0, 300)


def test_post_processing_v2_vs_reference(golden):
    import utils.evaluation_utils as ev
    g = golden("eval_nms_b4.npz")
    outs = ev.post_processing_v2(torch.tensor(g["pred"]), conf_thresh=float(g["conf_thresh"]), nms_thresh=float(g["nms_thresh"]))
    assert all(o is None or (not o.is_cuda and o.dtype == torch.float32) for o in outs)
    _check_against(outs, [g["out_%d" % i] for i in range(4)])
    empty = g["pred"][:1].copy(); empty[:, :, 6] *= 0.1
    assert ev.post_processing_v2(torch.tensor(empty), 0.5, 0.4) == [None]


def test_default_thresholds_vs_reference(golden):
    import utils.evaluation_utils as ev
    g = golden("eval_nms_b3_default_thresh.npz")
    outs = ev.post_processing_v2(torch.tensor(g["pred"]))                     # conf_thresh=0.95, nms_thresh=0.4
    _check_against(outs, [g["out_%d" % i] for i in range(3)])
    st = ev.get_batch_statistics_rotated_bbox(outs, torch.tensor(g["targets_px"]), iou_threshold=0.5)
    for i, (tp, sc, lb) in enumerate(st):
        assert np.array_equal(tp, g["tp_%d" % i]) and np.array_equal(np.asarray(sc), g["score_%d" % i])


def test_batch_statistics_vs_reference(golden):
    import utils.evaluation_utils as ev
    g = golden("eval_nms_b4.npz")
    outs = [torch.tensor(g["out_%d" % i]) for i in range(4)]
    st = ev.get_batch_statistics_rotated_bbox(outs, torch.tensor(g["targets_px"]), iou_threshold=float(g["iou_thresh"]))
    assert len(st) == 4
    for i, (tp, sc, lb) in enumerate(st):
        assert np.array_equal(tp, g["tp_%d" % i])
        assert np.array_equal(np.asarray(sc), g["score_%d" % i]) and np.array_equal(np.asarray(lb), g["label_%d" % i])
    # an image without detections contributes no entry; an image without annotations has no true positives
    st2 = ev.get_batch_statistics_rotated_bbox([None, outs[1]], torch.tensor(g["targets_px"][g["targets_px"][:, 0] == 0]), 0.5)
    assert len(st2) == 1 and st2[0][0].sum() == 0


def test_full_size_vs_oracle():
    """BASELINE-sized output (22,743 rows / image, bs=8): kept sets, class columns and true positives identical to the
    oracle, merged boxes within BOX_TOL; decisions are kept 1e-4 away from the thresholds by construction check."""
    from cy4 import evalops, synth
    from oracle import eval_oracle as eo
    B = 8
    tg = synth.make_targets(B, per_image=8, seed=11)
    pred = synth.make_detections(B, tg, n_rows=22743, dup=8, clutter=150, seed=3)
    ref = eo.post_processing_v2(pred, 0.5, 0.4)
    dets = evalops.nms_v2(torch.tensor(pred), 0.5, 0.4)
    outs = dets.as_list("cpu")
    _check_against(outs, ref)
    tpx = tg.copy(); tpx[:, 2:6] *= 608
    tp = evalops.match(dets, torch.tensor(tpx), 0.5).cpu().numpy()
    st = eo.get_batch_statistics(ref, tpx, 0.5)
    for i, (rtp, _, _) in enumerate(st):
        assert np.array_equal(tp[i, :len(rtp)], rtp.astype(np.uint8))
    assert sum(int(s[0].sum()) for s in st) > 30


def test_edge_cases():
    from cy4 import evalops, synth
    from oracle import eval_oracle as eo
    rng = np.random.default_rng(0)
    # (a) equal scores: ordered by row; (b) identical boxes of one class merge into one; (c) a single candidate
    pred = np.zeros((3, 64, 10), np.float32)
    pred[:, :, 2:4] = 10; pred[:, :, 5] = 1
    pred[0, :8, :2] = rng.uniform(100, 110, (8, 2)); pred[0, :8, 6] = 0.9; pred[0, :8, 7] = 0.8
    pred[1, :5, :2] = 300; pred[1, :5, 6] = np.linspace(0.6, 0.9, 5); pred[1, :5, 8] = 0.7
    pred[2, 17, :2] = 50; pred[2, 17, 6] = 0.99; pred[2, 17, 9] = 0.5
    ref = eo.post_processing_v2(pred, 0.5, 0.4)
    outs = evalops.nms_v2(torch.tensor(pred), 0.5, 0.4).as_list()
    _check_against(outs, ref)
    assert outs[1].shape[0] == 1 and outs[2].shape[0] == 1
    # (d) more candidates than the kernel keeps: loud error, no truncated result
    many = np.zeros((1, 5000, 10), np.float32); many[:, :, 2:4] = 4; many[:, :, 5] = 1; many[:, :, 6] = 0.9; many[:, :, 7] = 0.9
    many[0, :, 0] = np.arange(5000) * 10
    with pytest.raises(RuntimeError):
        evalops.nms_v2(torch.tensor(many), 0.5, 0.4)
    # (e) empty target list
    d = evalops.nms_v2(torch.tensor(pred), 0.5, 0.4)
    assert evalops.match(d, torch.zeros(0, 8), 0.5).sum().item() == 0


def test_v1_is_not_silently_replaced():
    import utils.evaluation_utils as ev
    with pytest.raises(NotImplementedError):
        ev.post_processing(np.zeros((1, 4, 10), np.float32))
