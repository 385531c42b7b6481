"""CPU restatement of the reference YOLO head: decode, target assignment, the 9 loss terms and
the 18 metrics.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/src/models/yolo_layer.py:53-67 (grid offsets), :69-142 (build_targets),
:144-253 (forward / loss / metrics).  Geometry comes from oracle/rbox_oracle.c.  Written
target-by-target with explicit loops (the reference uses advanced-indexing scatters); the
duplicate-cell semantics of those scatters on CPU (last writer wins, multi-hot tcls, GIoU over
all nT pairs: SURVEY F12) are restated explicitly.  fp32 throughout via torch CPU ops.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import geometry as og

METRIC_KEYS = ["loss", "iou_score", "giou_loss", "loss_x", "loss_y", "loss_w", "loss_h", "loss_eular",
               "loss_im", "loss_re", "loss_obj", "loss_cls", "cls_acc", "recall50", "recall75",
               "precision", "conf_obj", "conf_noobj"]


class _RGIoUSum(torch.autograd.Function):
    """sum_k term_k with the reference's autograd semantics (oracle C backward)."""

    @staticmethod
    def forward(ctx, pred6, tgt6, giou):
        iou, term, grad = og.rgiou_pairs(pred6.detach().numpy(), tgt6.detach().numpy(), giou, True)
        ctx.grad = torch.from_numpy(grad)
        acc = np.float32(0)
        for v in term:               # giou_loss += term, sequential fp32 (iou_rotated_boxes_utils.py:133)
            acc = np.float32(acc + v)
        ious = torch.from_numpy(iou)
        ctx.mark_non_differentiable(ious)
        return torch.tensor([acc], dtype=torch.float), ious

    @staticmethod
    def backward(ctx, g, _gi):
        return ctx.grad * g, None, None


def scaled_anchors(anchors, img_size, grid):
    stride = img_size / grid
    return torch.tensor([(aw / stride, ah / stride, im, re) for aw, ah, im, re in anchors], dtype=torch.float)


def decode(x, anchors, num_classes, img_size):
    """yolo_layer.py:156-190. Returns dict of decoded pieces (all [B,nA,G,G,...])."""
    B, _, G, _ = x.shape
    nA = len(anchors)
    p = x.view(B, nA, num_classes + 7, G, G).permute(0, 1, 3, 4, 2).contiguous()
    sa = scaled_anchors(anchors, img_size, G)
    gx = torch.arange(G, dtype=torch.float).view(1, 1, 1, G)
    gy = torch.arange(G, dtype=torch.float).view(1, 1, G, 1)
    d = {
        "px": torch.sigmoid(p[..., 0]), "py": torch.sigmoid(p[..., 1]), "pw": p[..., 2], "ph": p[..., 3],
        "pim": p[..., 4], "pre": p[..., 5], "conf": torch.sigmoid(p[..., 6]), "cls": torch.sigmoid(p[..., 7:]),
        "sa": sa, "stride": img_size / G,
    }
    boxes = torch.stack([
        d["px"] + gx, d["py"] + gy,
        torch.exp(d["pw"]).clamp(max=1e3) * sa[:, 0].view(1, nA, 1, 1),
        torch.exp(d["ph"]).clamp(max=1e3) * sa[:, 1].view(1, nA, 1, 1),
        d["pim"], d["pre"]], dim=-1)
    d["boxes"] = boxes
    d["output"] = torch.cat((boxes[..., :4].reshape(B, -1, 4) * d["stride"], boxes[..., 4:6].reshape(B, -1, 2),
                             d["conf"].reshape(B, -1, 1), d["cls"].reshape(B, -1, num_classes)), dim=-1)
    return d


def build_targets(boxes, pred_cls, target, sa, ignore_thresh, use_giou):
    """yolo_layer.py:69-142.  Returns the reference's 13-tuple plus the integer index arrays."""
    B, nA, G, _, nC = pred_cls.shape
    nT = target.shape[0]
    z = lambda *s: torch.zeros(*s, dtype=torch.float)
    obj = torch.zeros(B, nA, G, G, dtype=torch.bool)
    noobj = torch.ones(B, nA, G, G, dtype=torch.bool)
    class_mask, iou_scores = z(B, nA, G, G), z(B, nA, G, G)
    tx, ty, tw, th, tim, tre = (z(B, nA, G, G) for _ in range(6))
    tcls = z(B, nA, G, G, nC)
    giou_loss = torch.tensor([0.], dtype=torch.float)
    idx = {k: np.zeros(nT, np.int64) for k in ("b", "label", "best_n", "gi", "gj")}
    ious_at = np.zeros((nA, nT), np.float32)
    if nT > 0:
        tb = torch.cat((target[:, 2:6] * G, target[:, 6:8]), dim=-1)          # :97 (fp32 multiply)
        ious_at = og.anchor_iou(sa.numpy(), tb[:, 2:6].numpy())                 # :103-106
        best_n = np.argmax(ious_at, axis=0)                                    # first max (:107)
        b = target[:, 0].long().numpy(); lab = target[:, 1].long().numpy()     # :96
        gi = tb[:, 0].long().numpy(); gj = tb[:, 1].long().numpy()             # :112 truncation
        idx.update(b=b, label=lab, best_n=best_n, gi=gi, gj=gj)
        thr = np.float32(ignore_thresh)
        for t in range(nT):                                                    # CPU index_put order
            c = (b[t], best_n[t], gj[t], gi[t])
            obj[c] = True; noobj[c] = False                                    # :114-115
        for t in range(nT):
            for a in range(nA):
                if ious_at[a, t] > thr:                                        # :118-119 strict >
                    noobj[b[t], a, gj[t], gi[t]] = False
        gsel = boxes[torch.from_numpy(b), torch.from_numpy(best_n), torch.from_numpy(gj), torch.from_numpy(gi)]
        for t in range(nT):
            c = (b[t], best_n[t], gj[t], gi[t])
            tx[c] = tb[t, 0] - tb[t, 0].floor(); ty[c] = tb[t, 1] - tb[t, 1].floor()       # :122-123
            tw[c] = torch.log(tb[t, 2] / sa[best_n[t], 0] + 1e-16)                          # :125
            th[c] = torch.log(tb[t, 3] / sa[best_n[t], 1] + 1e-16)
            tim[c] = tb[t, 4]; tre[c] = tb[t, 5]
            tcls[c + (lab[t],)] = 1                                                         # multi-hot on duplicates
            class_mask[c] = float(int(torch.argmax(pred_cls[c])) == lab[t])                 # :133
        total, ious = _RGIoUSum.apply(gsel, tb, bool(use_giou))                             # :134
        for t in range(nT):
            iou_scores[b[t], best_n[t], gj[t], gi[t]] = ious[t]
        giou_loss = total / nT                                                              # :137-138
    tconf = obj.float()
    return (iou_scores, giou_loss, class_mask, obj, noobj, tx, ty, tw, th, tim, tre, tcls, tconf), idx, ious_at


def forward(x, targets, anchors, num_classes=3, img_size=608, ignore_thresh=0.7, use_giou_loss=True):
    """yolo_layer.py:144-253.  x [B, nA*(7+nC), G, G] fp32 (may require grad); targets [nT,8].
    Returns (output, total_loss, metrics dict, extras)."""
    d = decode(x, anchors, num_classes, img_size)
    if targets is None:
        return d["output"], 0, {}, d
    bt, idx, ious_at = build_targets(d["boxes"], d["cls"], targets, d["sa"], ignore_thresh, use_giou_loss)
    iou_scores, giou_loss, class_mask, obj, noobj, tx, ty, tw, th, tim, tre, tcls, tconf = bt
    mse = lambda a, b: F.mse_loss(a[obj], b[obj], reduction="mean")
    loss_x, loss_y = mse(d["px"], tx), mse(d["py"], ty)
    loss_w, loss_h = mse(d["pw"], tw), mse(d["ph"], th)
    loss_im, loss_re = mse(d["pim"], tim), mse(d["pre"], tre)
    loss_im_re = ((1. - torch.sqrt(d["pim"][obj] ** 2 + d["pre"][obj] ** 2)) ** 2).mean()   # :205-206
    loss_eular = loss_im + loss_re + loss_im_re
    bce = lambda p, t: F.binary_cross_entropy(p, t, reduction="mean")
    loss_conf_obj = bce(d["conf"][obj], tconf[obj])
    loss_conf_noobj = bce(d["conf"][noobj], tconf[noobj])
    loss_cls = bce(d["cls"][obj], tcls[obj])
    if use_giou_loss:                                                                         # :213-215
        loss_obj = loss_conf_obj + loss_conf_noobj
        total = giou_loss * 3.54 + loss_eular * 3.54 + loss_obj * 64.3 + loss_cls * 37.4
    else:                                                                                     # :216-218
        loss_obj = 1 * loss_conf_obj + 100 * loss_conf_noobj
        total = loss_x + loss_y + loss_w + loss_h + loss_eular + loss_obj + loss_cls
    conf50 = (d["conf"] > 0.5).float()
    iou50 = (iou_scores > 0.5).float(); iou75 = (iou_scores > 0.75).float()
    det = conf50 * class_mask * tconf
    vals = [total, iou_scores[obj].mean(), giou_loss, loss_x, loss_y, loss_w, loss_h, loss_eular, loss_im, loss_re,
            loss_obj, loss_cls, 100 * class_mask[obj].mean(),
            torch.sum(iou50 * det) / (obj.sum() + 1e-16), torch.sum(iou75 * det) / (obj.sum() + 1e-16),
            torch.sum(iou50 * det) / (conf50.sum() + 1e-16), d["conf"][obj].mean(), d["conf"][noobj].mean()]
    metrics = {k: float(v.detach().reshape(-1)[0]) for k, v in zip(METRIC_KEYS, vals)}
    extras = dict(d, build_targets=bt, idx=idx, anchor_ious=ious_at)
    return d["output"], total, metrics, extras
