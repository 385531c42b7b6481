"""GPU: parity of the WHOLE step at the bench configuration, and of a short training run (VERDICT r1 "weak" items 1-2).

* complex_yolov4.cfg, bs=32, 608x608, 160 targets -- exactly bench.py's workload -- against the oracle restated at the engine's
  storage precision (oracle/darknet_oracle.py storage="fp16": same rounding points, plain PyTorch fp32 ops).  The oracle's conv
  stack is evaluated with torch on the same GPU (TF32 off) because 4 TFLOP and ~20 GB of fp32 activations are not a CPU-sized
  job; its loss head is the host restatement (numpy + C geometry) as everywhere else.
* complex_yolov4_tiny.cfg, bs=2: 30 Adam steps, loss trajectory of the engine against the fp32 oracle's, step by step."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_bs32_step_vs_fp16_storage_oracle():
    from cy4 import netdefs, synth
    from cy4.darknet import Darknet
    from oracle import darknet_oracle as do
    B = 32
    cfg = netdefs.cfg_path("complex_yolov4")
    torch.manual_seed(0)
    model = Darknet(cfg, True)
    sd = {k: v.clone().cuda() for k, v in model.state_dict().items()}
    x = synth.make_bev(B)
    tg = torch.tensor(synth.make_targets(B, per_image=5, seed=4321))
    collect = {}
    with torch.no_grad():
        ol, oo, _ = do.forward(do.parse_cfg(cfg), sd, x.cuda(), tg, True, True, collect=collect, storage="fp16")
    keep = sorted(collect)[:12]
    ref_acts = {i: collect[i] for i in keep}
    del collect, sd
    torch.cuda.empty_cache()
    model = model.cuda().train()
    loss, out = model(x.cuda(), tg.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print("bs=32 complex_yolov4: engine loss %.5f, fp16-storage oracle %.5f" % (loss.item(), ol.item()))
    assert abs(loss.item() - ol.item()) <= 1e-2 * abs(ol.item())
    recs = {r["ind"]: r for r in model._engine.plan.convs}
    worst = 0.0
    for ind in keep:                          # the first dozen layers (3x3 s1 / s2, 1x1, Mish, route, shortcut) are in the linear regime
        r = recs[ind]
        if r.get("res") is not None:
            continue                          # fused conv+BN+act+residual: only the shortcut sum is materialised
        v = r["A"]
        a = v.st.buf[..., v.off:v.off + v.C].float().permute(0, 3, 1, 2)
        ref = ref_acts[ind]
        err = (a - ref).abs().max().item() / ref.std().item()
        worst = max(worst, err)
        assert err <= 0.012, (ind, err)
    print("worst first-dozen-layers activation error / sigma: %.4f" % worst)
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert out.shape == (B, 22743, 10) and torch.isfinite(out).all()


def test_30_adam_steps_track_the_fp32_oracle():
    from cy4 import netdefs, synth
    from cy4.darknet import Darknet
    from oracle import darknet_oracle as do
    cfg = netdefs.cfg_path("complex_yolov4_tiny")
    size, B, steps = 256, 2, 30
    torch.manual_seed(2)
    model = Darknet(cfg, True)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    blocks = do.parse_cfg(cfg)
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    batches = [(synth.make_bev(B, img_size=size, seed=100 + i),
                torch.tensor(synth.make_targets(B, per_image=3, seed=200 + i, img_size=size, strides=(16, 32)))) for i in range(4)]
    oopt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3)
    model = model.cuda().train()
    eopt = torch.optim.Adam(model.parameters(), lr=1e-3)
    e_losses, o_losses = [], []
    for t in range(steps):
        x, tg = batches[t % len(batches)]
        ol, _, _ = do.forward(blocks, params, x, tg, True, True, update_running=True)
        oopt.zero_grad(); ol.backward(); oopt.step()
        el, _ = model(x.cuda(), tg.cuda())
        eopt.zero_grad(); el.backward(); eopt.step()
        e_losses.append(el.item()); o_losses.append(ol.item())
    e, o = np.array(e_losses), np.array(o_losses)
    dev = np.abs(e - o) / np.abs(o)
    print("loss trajectory (oracle):", np.round(o[::5], 3), "(engine):", np.round(e[::5], 3), "max relative deviation %.4f at step %d" % (dev.max(), dev.argmax()))
    assert o[-1] < o[0]                            # the run really trains
    assert dev.max() <= 0.02, dev
