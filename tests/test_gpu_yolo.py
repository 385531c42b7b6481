"""GPU: the fused YOLO head (decode, target assignment, loss, metrics, gradient) through the C-ABI,
against golden outputs of the reference's YoloLayer and against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["g19_giou", "g38_dup_giou", "g19_dup_mse", "g76_giou"]


def _layer(anchors, G):
    from cy4.yolo import YoloLayer
    return YoloLayer(3, [tuple(a) for a in anchors.tolist()], 608 // G, 1.1, 0.7)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("channels_last", [False, True])
def test_layer_vs_reference_golden(golden, case, channels_last):
    g = golden(f"yolo_{case}.npz")
    G = int(g["G"])
    x = torch.tensor(g["x"], device="cuda")
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    tg = torch.tensor(g["targets"], device="cuda")
    layer = _layer(g["anchors"], G)
    out, loss = layer(x, tg, 608, bool(g["use_giou"]))
    loss.backward()
    assert tuple(loss.shape) == tuple(g["loss"].shape)
    np.testing.assert_allclose(out.cpu().numpy(), g["output"], atol=2e-4, rtol=2e-6)   # pixels up to 608
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss"], rtol=1e-5)
    # the head gradient: BASELINE states no tolerance for gradients; fp32 sums in another order
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["grad"], atol=3e-6, rtol=2e-3)
    for k, v in zip(g["metric_keys"], g["metric_vals"]):
        np.testing.assert_allclose(layer.metrics[str(k)], v, rtol=2e-5, atol=1e-4 if str(k) in ("iou_score", "giou_loss") else 1e-6, err_msg=str(k))
    assert list(layer.metrics.keys()) == [str(k) for k in g["metric_keys"]]


@pytest.mark.parametrize("case", CASES)
def test_build_targets_vs_reference_golden(golden, case):
    from oracle import yolo_oracle as yo
    g = golden(f"yolo_{case}.npz")
    G = int(g["G"])
    anchors = [tuple(a) for a in g["anchors"].tolist()]
    d = yo.decode(torch.tensor(g["x"]), anchors, 3, 608)
    layer = _layer(g["anchors"], G)
    layer.img_size = 608; layer.use_giou_loss = bool(g["use_giou"]); layer.device = torch.device("cuda")
    layer.compute_grid_offsets(G)
    bt = layer.build_targets(d["boxes"].cuda(), d["cls"].cuda(), torch.tensor(g["targets"]).cuda(), layer.scaled_anchors)
    assert len(bt) == 13 and bt[3].dtype == torch.bool and bt[4].dtype == torch.bool
    for i in range(13):
        key = f"bt{i}"
        if key not in g.files:
            continue
        mine = bt[i].cpu().numpy()
        if mine.dtype == np.bool_:
            assert (mine.astype(np.uint8) == g[key]).all(), key                 # masks bit-exact
        else:
            np.testing.assert_allclose(mine, g[key], atol=1e-4 if i in (0, 1) else 2e-6, err_msg=key)


def test_decode_only(golden):
    g = golden("yolo_g19_giou.npz")
    layer = _layer(g["anchors"], 19)
    out, zero = layer(torch.tensor(g["x"], device="cuda"), None, 608, True)
    assert zero == 0
    np.testing.assert_allclose(out.cpu().numpy(), g["output"], atol=2e-4, rtol=2e-6)


def test_empty_targets_nan():
    layer = _layer(np.array([(23, 49, 0., 1.), (23, 55, 0., 1.), (24, 53, 0., 1.)], np.float32), 19)
    x = torch.randn(1, 30, 19, 19, device="cuda", requires_grad=True)
    _, loss = layer(x, torch.zeros(0, 8, device="cuda"), 608, True)
    assert torch.isnan(loss).all()


def test_target_out_of_range_raises():
    layer = _layer(np.array([(23, 49, 0., 1.), (23, 55, 0., 1.), (24, 53, 0., 1.)], np.float32), 19)
    layer.check_targets = True
    x = torch.randn(1, 30, 19, 19, device="cuda")
    tg = torch.tensor([[0, 0, 1.0, 0.5, 0.05, 0.08, 0.0, 1.0]], device="cuda")     # x == 1.0 -> gi == G
    with pytest.raises(IndexError):
        layer(x, tg, 608, True)


@pytest.mark.parametrize("B,G,per_image,seed", [(4, 76, 7, 1), (8, 38, 5, 2), (16, 19, 3, 3)])
def test_layer_vs_oracle_larger(B, G, per_image, seed):
    """Seeded larger cases: integer assignment bit-exact, values within the stated tolerances."""
    from cy4 import synth
    from oracle import yolo_oracle as yo
    anchors = {76: [(11, 15, 0., 1.), (10, 24, 0., 1.), (11, 25, 0., 1.)],
               38: [(23, 49, 0., 1.), (23, 55, 0., 1.), (24, 53, 0., 1.)],
               19: [(24, 60, 0., 1.), (27, 63, 0., 1.), (29, 74, 0., 1.)]}[G]
    torch.manual_seed(seed)
    x_cpu = torch.randn(B, 30, G, G) * 0.8
    tg = torch.tensor(synth.make_targets(B, per_image=per_image, seed=seed))
    xo = x_cpu.clone().requires_grad_(True)
    oo, ol, om, ex = yo.forward(xo, tg, anchors, 3, 608, 0.7, True)
    ol.backward()
    layer = _layer(np.array(anchors, np.float32), G)
    xg = x_cpu.cuda().requires_grad_(True)
    out, loss = layer(xg, tg.cuda(), 608, True)
    loss.backward()
    np.testing.assert_allclose(out.cpu().numpy(), oo.detach().numpy(), atol=2e-4, rtol=2e-6)
    np.testing.assert_allclose(loss.item(), ol.item(), rtol=2e-5)
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xo.grad.numpy(), atol=3e-6, rtol=2e-3)
    for k in om:
        np.testing.assert_allclose(layer.metrics[k], om[k], rtol=5e-5, atol=1e-4 if k in ("iou_score", "giou_loss") else 1e-6, err_msg=k)
    # integer indices bit-exact via build_targets
    d = yo.decode(x_cpu, anchors, 3, 608)
    layer.use_giou_loss = True
    bt = layer.build_targets(d["boxes"].cuda(), d["cls"].cuda(), tg.cuda(), layer.scaled_anchors)
    ref = ex["build_targets"]
    assert torch.equal(bt[3].cpu(), ref[3]) and torch.equal(bt[4].cpu(), ref[4])
    assert torch.equal(bt[11].cpu(), ref[11])           # tcls one-hot
    assert torch.equal(bt[2].cpu(), ref[2])             # class_mask
