"""CPU: the YOLO-head oracle (oracle/yolo_oracle.py) against golden outputs of the reference's
YoloLayer (values, loss, 18 metrics, gradient, the 13 build_targets tensors incl. duplicate cells)."""
import numpy as np
import pytest
import torch

from oracle import yolo_oracle as yo

CASES = ["g19_giou", "g38_dup_giou", "g19_dup_mse", "g76_giou"]


@pytest.mark.parametrize("case", CASES)
def test_yolo_oracle_vs_reference(golden, case):
    g = golden(f"yolo_{case}.npz")
    x = torch.tensor(g["x"], requires_grad=True)
    tg = torch.tensor(g["targets"])
    anchors = [tuple(a) for a in g["anchors"].tolist()]
    out, loss, met, ex = yo.forward(x, tg, anchors, 3, 608, 0.7, bool(g["use_giou"]))
    loss.backward()
    assert tuple(loss.shape) == tuple(g["loss"].shape)           # [1] with GIoU, 0-dim without (F13)
    np.testing.assert_allclose(out.detach().numpy(), g["output"], atol=1e-5, rtol=1e-6)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=2e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["grad"], atol=2e-6, rtol=1e-4)
    for k, v in zip(g["metric_keys"], g["metric_vals"]):
        np.testing.assert_allclose(met[str(k)], v, rtol=1e-5, atol=1e-7, err_msg=str(k))
    bt = ex["build_targets"]
    for i in range(13):
        key = f"bt{i}"
        if key not in g.files:
            continue
        mine = bt[i].detach().numpy()
        if mine.dtype == np.bool_:
            assert (mine.astype(np.uint8) == g[key]).all(), key      # masks bit-exact
        else:
            np.testing.assert_allclose(mine, g[key], atol=1e-4 if i in (0, 1) else 1e-5, err_msg=key)  # iou on the F7 noise floor


def test_empty_targets_nan():
    """an empty target batch gives a NaN loss in the reference."""
    x = torch.randn(1, 30, 19, 19)
    _, loss, _, _ = yo.forward(x, torch.zeros(0, 8), [(23, 49, 0., 1.), (23, 55, 0., 1.), (24, 53, 0., 1.)])
    assert torch.isnan(loss).all()
