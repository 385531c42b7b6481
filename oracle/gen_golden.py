"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference/src) on seeded synthetic inputs.  TEST INFRASTRUCTURE ONLY.

Run in the build container only (the reference tree does not exist on the GPU box):

    python -m oracle.gen_golden            # from the repo root

shapely is not installed, so the reference's GEOS calls go through oracle/shapely_standin.py
(fp64 convex clipping); every fixture records `shapely_kind`.  The GIoU=True path never calls
shapely for pred<->target boxes (only anchor<->target IoU does).
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
from oracle import reference_loader as rl  # noqa: E402
from cy4 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CFG_DIR = os.path.join(ROOT, "complex-yolov4-pytorch_b200", "cy4", "cfg")
ANCHORS_PX = [(11, 15), (10, 24), (11, 25), (23, 49), (23, 55), (24, 53), (24, 60), (27, 63), (29, 74)]


def save(name, **kw):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def gen_pairs(m):
    iou = m["iou"]
    pred, tgt = synth.make_pairs(2000, seed=7, disjoint_frac=0.02)
    T = torch.tensor(tgt)
    ious, terms, grads = [], [], []
    for k in range(len(pred)):
        p = torch.tensor(pred[k:k + 1], requires_grad=True)
        i, l = iou.iou_pred_vs_target_boxes(p, T[k:k + 1], GIoU=True)
        l.backward()
        ious.append(i.item()); terms.append(l.item()); grads.append(p.grad[0].numpy().copy())
    # batched call on the first 64 (sum semantics, shape [1])
    pb = torch.tensor(pred[:64], requires_grad=True)
    ib, lb = iou.iou_pred_vs_target_boxes(pb, T[:64], GIoU=True)
    lb.backward()
    # shapely path (GIoU=False)
    ns = 300
    ps = torch.tensor(pred[:ns], requires_grad=True)
    i_s, l_s = iou.iou_pred_vs_target_boxes(ps, T[:ns], GIoU=False)
    l_s.backward()
    # known answers quoted in SURVEY.md section 4 / 8c (demo blocks of the reference)
    ka_pred = np.array([[100, 100, 60, 10, np.sin(0.5), np.cos(0.5)],
                        [100, 100, 40, 10, 1.0, np.cos(np.pi / 2)],
                        [10, 10, 4, 8, 0, 1], [10, 10, 4, 8, 0, 1],
                        [10, 10, 4, 8, np.sin(.3), np.cos(.3)]], np.float32)
    ka_tgt = np.array([[100, 100, 40, 20, 0, 1], [100, 100, 40, 20, 0, 1], [10, 10, 4, 8, 0, 1],
                       [12, 10, 4, 8, 0, 1], [30, 30, 4, 8, 0, 1]], np.float32)
    ka_i, ka_t = [], []
    for k in range(len(ka_pred)):
        i, l = iou.iou_pred_vs_target_boxes(torch.tensor(ka_pred[k:k + 1]), torch.tensor(ka_tgt[k:k + 1]), GIoU=True)
        ka_i.append(i.item()); ka_t.append(l.item())
    save("rgiou_pairs.npz", pred=pred, tgt=tgt, iou=np.array(ious, np.float32), term=np.array(terms, np.float32),
         grad=np.array(grads, np.float32), batch64_loss=lb.detach().numpy(), batch64_grad=pb.grad.numpy(),
         shapely_iou=i_s.numpy(), shapely_loss=l_s.detach().numpy(), shapely_grad=ps.grad.numpy(),
         ka_pred=ka_pred, ka_tgt=ka_tgt, ka_iou=np.array(ka_i, np.float32), ka_term=np.array(ka_t, np.float32),
         shapely_kind=m["shapely_kind"])


def gen_anchor_iou(m):
    iou = m["iou"]
    tg = synth.make_targets(16, per_image=5, seed=4321)
    res = {}
    for li, (G, mask) in enumerate([(76, (0, 1, 2)), (38, (3, 4, 5)), (19, (6, 7, 8))]):
        stride = 608 / G
        sa = torch.tensor([(ANCHORS_PX[i][0] / stride, ANCHORS_PX[i][1] / stride, 0., 1.) for i in mask], dtype=torch.float)
        tb = torch.cat((torch.tensor(tg[:, 2:6]) * G, torch.tensor(tg[:, 6:8])), dim=-1)
        ap, aa = iou.get_polygons_areas_fix_xy(sa)
        tp, ta = iou.get_polygons_areas_fix_xy(tb[:, 2:6])
        res[f"ious_{G}"] = iou.iou_rotated_boxes_targets_vs_anchors(ap, aa, tp, ta).numpy()
        res[f"anchors_{G}"] = sa.numpy()
        res[f"tboxes_{G}"] = tb[:, 2:6].numpy()
    # known answer quoted in SURVEY.md 8c
    sa = torch.tensor([(11 / 8, 15 / 8, 0., 1.), (10 / 8, 24 / 8, 0., 1.), (11 / 8, 25 / 8, 0., 1.)], dtype=torch.float)
    tb = torch.tensor([[23 / 8, 51 / 8, np.sin(.4), np.cos(.4)], [11 / 8, 16 / 8, np.sin(1.5), np.cos(1.5)]], dtype=torch.float)
    ap, aa = iou.get_polygons_areas_fix_xy(sa); tp, ta = iou.get_polygons_areas_fix_xy(tb)
    res["ka_ious"] = iou.iou_rotated_boxes_targets_vs_anchors(ap, aa, tp, ta).numpy()
    res["ka_anchors"] = sa.numpy(); res["ka_tboxes"] = tb.numpy()
    save("anchor_iou.npz", targets=tg, shapely_kind=m["shapely_kind"], **res)


def gen_yolo_layer(m):
    Y = m["yolo"]
    cases = [("g19_giou", 19, (6, 7, 8), False, True), ("g38_dup_giou", 38, (3, 4, 5), True, True),
             ("g19_dup_mse", 19, (6, 7, 8), True, False), ("g76_giou", 76, (0, 1, 2), False, True)]
    for name, G, mask, dup, giou in cases:
        B = 2
        torch.manual_seed(G + int(dup))
        x = (torch.randn(B, 30, G, G) * 0.7).requires_grad_(True)
        tg = torch.tensor(synth.make_targets(B, per_image=4, seed=5))
        if dup:   # two extra targets landing on already-used cells with another label
            extra = tg[:2].clone(); extra[:, 1] = (extra[:, 1] + 1) % 3; extra[:, 4:6] *= 1.1
            tg = torch.cat([tg, extra])
        anchors = [(ANCHORS_PX[i][0], ANCHORS_PX[i][1], 0., 1.) for i in mask]
        layer = Y.YoloLayer(3, anchors, 608 // G, 1.1, 0.7)
        out, loss = layer(x, tg, 608, giou)
        loss.backward()
        bt = layer.build_targets(*_decode_for_bt(layer, x.detach()), tg, layer.scaled_anchors)
        kw = {f"bt{i}": (t.numpy().astype(np.uint8) if t.dtype == torch.bool else t.detach().numpy()) for i, t in enumerate(bt)}
        if G == 76:   # keep the fixture small: drop the dense float tensors that are all-but-zero
            kw = {k: v for k, v in kw.items() if k in ("bt1", "bt3", "bt4")}
        save(f"yolo_{name}.npz", x=x.detach().numpy(), targets=tg.numpy(), anchors=np.array(anchors, np.float32),
             output=out.detach().numpy().astype(np.float32), loss=loss.detach().numpy(), grad=x.grad.numpy(),
             metric_keys=np.array(list(layer.metrics.keys())), metric_vals=np.array(list(layer.metrics.values()), np.float64),
             use_giou=giou, G=G, shapely_kind=m["shapely_kind"], **kw)


def _decode_for_bt(layer, x):
    """pred_boxes / pred_cls exactly as YoloLayer.forward builds them (yolo_layer.py:156-182)."""
    B, _, G, _ = x.shape
    p = x.view(B, 3, 10, G, G).permute(0, 1, 3, 4, 2).contiguous()
    pb = torch.empty(p[..., :6].shape)
    pb[..., 0] = torch.sigmoid(p[..., 0]) + layer.grid_x
    pb[..., 1] = torch.sigmoid(p[..., 1]) + layer.grid_y
    pb[..., 2] = torch.exp(p[..., 2]).clamp(max=1E3) * layer.anchor_w
    pb[..., 3] = torch.exp(p[..., 3]).clamp(max=1E3) * layer.anchor_h
    pb[..., 4] = p[..., 4]; pb[..., 5] = p[..., 5]
    return pb, torch.sigmoid(p[..., 7:])


def _sample_idx(n, k, seed):
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(n, size=min(k, n), replace=False))


def gen_darknet(m, cfg, tag, batch, n_targets, k_samples=2048):
    D = m["darknet"]
    torch.manual_seed(0)
    model = D.Darknet(os.path.join(CFG_DIR, cfg), use_giou_loss=True)
    model.train()
    x = synth.make_bev(batch)
    tg = torch.tensor(synth.make_targets(batch, seed=4321, total=n_targets))
    sd0 = model.state_dict()
    digest = hashlib.sha256()           # of the freshly initialised weights (torch.manual_seed(0))
    for k in sd0:
        digest.update(k.encode()); digest.update(sd0[k].numpy().tobytes())
    acts = {}
    hooks = []
    for i, mod in enumerate(model.models):
        if isinstance(mod, torch.nn.Sequential):
            def hk(_m, _inp, out, i=i):
                flat = out.detach().reshape(-1)
                sel = _sample_idx(flat.numel(), k_samples, 1000 + i)
                acts[f"act{i}_idx"] = sel
                acts[f"act{i}_val"] = flat[torch.from_numpy(sel)].numpy().copy()
                acts[f"act{i}_stats"] = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()], np.float64)
            hooks.append(mod.register_forward_hook(hk))
    loss, outputs = model(x, tg)
    loss.backward()
    for h in hooks:
        h.remove()
    sd = model.state_dict()
    grads = {}
    for name, p in model.named_parameters():
        g = p.grad.reshape(-1)
        grads["gnorm/" + name] = np.array([g.norm().item(), g.abs().max().item()], np.float64)
        sel = _sample_idx(g.numel(), 64, 7)
        grads["gidx/" + name] = sel
        grads["gval/" + name] = g[torch.from_numpy(sel)].numpy().copy()
    mets = {}
    for li, yl in enumerate(model.yolo_layers):
        mets[f"metrics{li}"] = np.array(list(yl.metrics.values()), np.float64)
    # BN running stats after the step (momentum update)
    rs = {}
    for k, v in model.state_dict().items():
        if "running_mean" in k or "running_var" in k:
            rs["rs/" + k] = np.array([v.mean().item(), v.abs().max().item()], np.float64)
    save(f"darknet_{tag}.npz", cfg=cfg, batch=batch, targets=tg.numpy(), loss=loss.detach().numpy(),
         outputs=outputs.numpy(), weights_sha256=digest.hexdigest(), n_params=sum(p.numel() for p in model.parameters()),
         n_state=len(sd), state_keys=np.array(list(sd.keys())), shapely_kind=m["shapely_kind"],
         metric_keys=np.array(list(model.yolo_layers[0].metrics.keys())), **acts, **grads, **mets, **rs)


def gen_darknet_emu16(cfg, tag, batch, n_targets, k_samples=2048):
    """Same step through oracle/darknet_oracle.py at fp16 storage precision (no reference code involved):
    separates implementation errors from fp16 drift.  Same seeds / sample indices as gen_darknet."""
    from oracle import darknet_oracle as do
    from cy4 import netdefs
    from cy4.darknet import Darknet
    path = netdefs.cfg_path(cfg.replace(".cfg", ""))
    torch.manual_seed(0)
    sd = Darknet(path, True).state_dict()
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    x = synth.make_bev(batch)
    tg = torch.tensor(synth.make_targets(batch, seed=4321, total=n_targets))
    collect = {}
    loss, outputs, mets = do.forward(do.parse_cfg(path), params, x, tg, True, True, collect=collect, storage="fp16")
    loss.backward()
    acts, grads = {}, {}
    for i, a in collect.items():
        flat = a.detach().reshape(-1)
        sel = _sample_idx(flat.numel(), k_samples, 1000 + i)
        acts[f"act{i}_idx"] = sel
        acts[f"act{i}_val"] = flat[torch.from_numpy(sel)].numpy().copy()
        acts[f"act{i}_stats"] = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()], np.float64)
    for name, p in params.items():
        if p.grad is None:
            continue
        g = p.grad.reshape(-1)
        grads["gnorm/" + name] = np.array([g.norm().item(), g.abs().max().item()], np.float64)
        sel = _sample_idx(g.numel(), 4096, 7)
        grads["gidx/" + name] = sel
        grads["gval/" + name] = g[torch.from_numpy(sel)].numpy().copy()
    mk = list(mets[0].keys())
    save(f"darknet_{tag}_emu16.npz", cfg=cfg, batch=batch, targets=tg.numpy(), loss=loss.detach().numpy(),
         outputs=outputs.detach().numpy(), metric_keys=np.array(mk),
         **{f"metrics{i}": np.array([m[k] for k in mk], np.float64) for i, m in enumerate(mets)}, **acts, **grads)


def gen_eval(m):
    """Section 8 row f1: the reference's post_processing_v2, one-vs-many IoU, true-positive matching and AP
    (utils/evaluation_utils.py) on seeded synthetic detections."""
    ev = m["eval"]
    B, N = 4, 3000
    tg = synth.make_targets(B, per_image=6, seed=77)
    pred = synth.make_detections(B, tg, n_rows=N, dup=5, clutter=25, seed=99)
    conf_thresh, nms_thresh, iou_thresh = 0.5, 0.4, 0.5
    outs = ev.post_processing_v2(torch.tensor(pred), conf_thresh=conf_thresh, nms_thresh=nms_thresh)
    tgt_px = torch.tensor(tg).clone()
    tgt_px[:, 2:6] *= 608                                            # evaluate.py:41
    stats = ev.get_batch_statistics_rotated_bbox(outs, tgt_px, iou_threshold=iou_thresh)
    res = {"pred": pred, "targets_px": tgt_px.numpy(), "conf_thresh": conf_thresh, "nms_thresh": nms_thresh, "iou_thresh": iou_thresh,
           "none_mask": np.array([o is None for o in outs])}
    for i, o in enumerate(outs):
        if o is not None:
            res["out_%d" % i] = o.numpy()
    for i, (tp, sc, lb) in enumerate(stats):
        res["tp_%d" % i] = np.asarray(tp); res["score_%d" % i] = np.asarray(sc); res["label_%d" % i] = np.asarray(lb)
    tp = np.concatenate([s_[0] for s_ in stats]); sc = np.concatenate([np.asarray(s_[1]) for s_ in stats])
    lb = np.concatenate([np.asarray(s_[2]) for s_ in stats])
    import io, contextlib
    with contextlib.redirect_stderr(io.StringIO()):
        p_, r_, ap_, f1_, cls_ = ev.ap_per_class(tp, sc, lb, tg[:, 1].tolist())
    res.update(ap_precision=p_, ap_recall=r_, ap_AP=ap_, ap_f1=f1_, ap_classes=cls_)
    # one-vs-many IoU on its own: detections of image 0 against that image's targets and against themselves
    o0 = outs[0]
    t0 = tgt_px[tgt_px[:, 0] == 0][:, 2:]
    res["iou_det_vs_tgt"] = np.stack([ev.iou_rotated_single_vs_multi_boxes_cpu(o0[i, :6], t0).numpy() for i in range(o0.shape[0])])
    res["iou_det_vs_det"] = np.stack([ev.iou_rotated_single_vs_multi_boxes_cpu(o0[i, :6], o0[:, :6]).numpy() for i in range(min(16, o0.shape[0]))])
    # an image without any confident row -> None
    empty = pred[:1].copy(); empty[:, :, 6] *= 0.1
    res["empty_is_none"] = np.array([ev.post_processing_v2(torch.tensor(empty), conf_thresh, nms_thresh)[0] is None])
    res["shapely_kind"] = np.array(m["shapely_kind"])
    save("eval_nms_b4.npz", **res)
    # second case at post_processing_v2's default thresholds (0.95 / 0.4): fewer, more confident candidates
    tg2 = synth.make_targets(3, per_image=9, seed=5)
    pred2 = synth.make_detections(3, tg2, n_rows=1500, dup=7, clutter=30, seed=17, conf_lo=0.9)
    outs2 = ev.post_processing_v2(torch.tensor(pred2))
    t2 = torch.tensor(tg2).clone(); t2[:, 2:6] *= 608
    stats2 = ev.get_batch_statistics_rotated_bbox(outs2, t2, iou_threshold=0.5)
    res2 = {"pred": pred2, "targets_px": t2.numpy(), "none_mask": np.array([o is None for o in outs2])}
    for i, o in enumerate(outs2):
        if o is not None:
            res2["out_%d" % i] = o.numpy()
    for i, (tp_, sc_, lb_) in enumerate(stats2):
        res2["tp_%d" % i] = np.asarray(tp_); res2["score_%d" % i] = np.asarray(sc_); res2["label_%d" % i] = np.asarray(lb_)
    save("eval_nms_b3_default_thresh.npz", **res2)
    print("default thresholds: kept", [None if o is None else o.shape[0] for o in outs2], "TP", [int(s_[0].sum()) for s_ in stats2])
    print("kept per image:", [None if o is None else o.shape[0] for o in outs], " TP:", int(tp.sum()), "of", len(tp), " AP:", ap_)


def gen_bev(m):
    """Section 8 row f3: the reference's removePoints + makeBVFeature + build_yolo_target
    (data_process/kitti_bev_utils.py) on a seeded synthetic LiDAR frame."""
    kb = sys.modules["_ref_data_process.kitti_bev_utils"]
    cnf = sys.modules["_ref_config.kitti_config"]
    pts = synth.make_point_cloud(30000, seed=5)
    b = kb.removePoints(pts.copy(), cnf.boundary)
    rgb = kb.makeBVFeature(b, cnf.DISCRETIZATION, cnf.boundary)            # float64 [3,608,608]
    nz = np.flatnonzero((rgb != 0).any(axis=0))
    rng = np.random.default_rng(8)
    labels = np.stack([rng.integers(0, 3, 40).astype(np.float64), rng.uniform(-5, 60, 40), rng.uniform(-30, 30, 40), rng.uniform(-2, 0, 40),
                       rng.uniform(1.4, 1.9, 40), rng.uniform(0.5, 2.0, 40), rng.uniform(0.8, 4.5, 40), rng.uniform(-np.pi, np.pi, 40)], 1)
    tgt = kb.build_yolo_target(labels.astype(np.float32))
    save("bev_raster.npz", points=pts, filtered_rows=np.int64(b.shape[0]), filtered_head=b[:64], nz_cells=nz.astype(np.int32),
         nz_values=rgb.reshape(3, -1)[:, nz], labels=labels.astype(np.float32), yolo_target=tgt,
         discretization=np.float64(cnf.DISCRETIZATION))
    print("points", pts.shape, "inside", b.shape[0], "occupied cells", len(nz), "targets", tgt.shape)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    m = rl.load()
    which = sys.argv[1:] or ["pairs", "anchor", "yolo", "tiny", "v4", "emu", "eval", "bev"]
    if "eval" in which: gen_eval(m)
    if "bev" in which: gen_bev(m)
    if "pairs" in which: gen_pairs(m)
    if "anchor" in which: gen_anchor_iou(m)
    if "yolo" in which: gen_yolo_layer(m)
    if "tiny" in which: gen_darknet(m, "complex_yolov4_tiny.cfg", "tiny_bs2", 2, 8)       # BASELINE config 1
    if "v4" in which: gen_darknet(m, "complex_yolov4.cfg", "v4_bs2", 2, 8)
    if "emu" in which:
        gen_darknet_emu16("complex_yolov4_tiny.cfg", "tiny_bs2", 2, 8)
        gen_darknet_emu16("complex_yolov4.cfg", "v4_bs2", 2, 8)


if __name__ == "__main__":
    main()
