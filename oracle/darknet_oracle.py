"""Plain PyTorch fp32 restatement of the Darknet graph -- the floating-point reference for the
conv / BatchNorm / activation / route / shortcut / pooling / upsample kernels and for the whole
training step.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/src/models/darknet2pytorch.py:162-230 (forward interpreter) and :235-401
(layer semantics) with torch.nn.functional ops on the CPU; parameters are taken from a state_dict
with the reference's key names, the YOLO heads go through oracle/yolo_oracle.py.  Validated
against the reference's own Darknet via tests/golden/darknet_*.npz.
"""
import math

import torch
import torch.nn.functional as F

from . import yolo_oracle as yo


def parse_cfg(path):
    blocks, blk = [], None
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        if line.startswith("["):
            if blk:
                blocks.append(blk)
            blk = {"type": line[1:-1]}
            if blk["type"] == "convolutional":
                blk["batch_normalize"] = 0
        else:
            k, v = line.split("=")
            blk["_type" if k.strip() == "type" else k.strip()] = v.strip()
    if blk:
        blocks.append(blk)
    return blocks


def _r16(t):
    """Round to fp16 storage precision with a straight-through gradient."""
    return t + (t.detach().half().float() - t.detach())


def forward(blocks, sd, x, targets=None, use_giou=True, training=True, collect=None, update_running=False, storage="fp32"):
    """sd: state_dict-like mapping (tensors may require grad).  Returns (loss|None, outputs [B,N,10],
    per-yolo metrics list).  `collect` (dict) receives every conv block's activated output.

    storage="fp16" restates the SAME math at the engine's documented storage precision:
    conv inputs and weights rounded to fp16, fp32 accumulation, BatchNorm statistics from the fp32
    conv result, the raw conv output and the activated output each rounded once to fp16 (fp32 for the
    linear head convs), straight-through gradients.  It isolates implementation errors from the
    ~1e-2..1 relative drift that fp16 rounding alone produces in a randomly initialised 160-layer
    network (chaotic amplification, measured in tests/golden/darknet_*_emu16.npz)."""
    emu = storage == "fp16"
    outs, yolo_out, metrics = {}, [], []
    loss = 0.
    img = x.shape[2]
    ind, conv_id = -2, 0
    for b in blocks:
        ind += 1
        t = b["type"]
        if t == "net":
            continue
        if t == "convolutional":
            conv_id += 1
            k, s = int(b["size"]), int(b["stride"])
            pad = (k - 1) // 2 if int(b["pad"]) else 0
            pre = "models.%d." % ind
            bn = int(b["batch_normalize"])
            w = sd[pre + "conv%d.weight" % conv_id]
            if emu:
                w = _r16(w)
                x = _r16(x)
            x = F.conv2d(x, w, None if bn else sd[pre + "conv%d.bias" % conv_id], s, pad)
            if bn:
                rm, rv = sd[pre + "bn%d.running_mean" % conv_id], sd[pre + "bn%d.running_var" % conv_id]
                if not update_running:
                    rm, rv = rm.clone(), rv.clone()
                gamma, beta = sd[pre + "bn%d.weight" % conv_id], sd[pre + "bn%d.bias" % conv_id]
                if emu and training:
                    # statistics from the fp32 conv result, normalisation applied to the fp16-stored one
                    mean = x.mean((0, 2, 3)); var = x.var((0, 2, 3), unbiased=False)
                    n = x.numel() / x.shape[1]
                    with torch.no_grad():
                        rm.mul_(0.9).add_(0.1 * mean); rv.mul_(0.9).add_(0.1 * var * n / max(n - 1, 1))
                    sc = gamma / torch.sqrt(var + 1e-5)
                    x = _r16(x) * sc.view(1, -1, 1, 1) + (beta - mean * sc).view(1, -1, 1, 1)
                else:
                    if emu:
                        x = _r16(x)
                    x = F.batch_norm(x, rm, rv, gamma, beta, training, 0.1, 1e-5)
            a = b["activation"]
            if a == "leaky":
                x = F.leaky_relu(x, 0.1)
            elif a == "mish":
                x = x * torch.tanh(F.softplus(x))
            if emu and bn:
                x = _r16(x)
            if collect is not None:
                collect[ind] = x
        elif t == "route":
            ls = [int(i) if int(i) > 0 else int(i) + ind for i in b["layers"].split(",")]
            if len(ls) == 1:
                x = outs[ls[0]]
                g = int(b.get("groups", 1))
                if g > 1:
                    c = x.shape[1] // g
                    x = x[:, c * int(b["group_id"]):c * (int(b["group_id"]) + 1)]
            else:
                x = torch.cat([outs[l] for l in ls], 1)
        elif t == "shortcut":
            f = int(b["from"])
            x = outs[f if f > 0 else f + ind] + outs[ind - 1]
        elif t == "maxpool":
            k, s = int(b["size"]), int(b["stride"])
            x = F.max_pool2d(x, k, s, k // 2 if (s == 1 and k % 2) else 0)
        elif t == "upsample":
            x = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
        elif t == "yolo":
            mask = [int(i) for i in b["mask"].split(",")]
            a = [float(i) for i in b["anchors"].split(",")]
            anchors = [(a[i], a[i + 1], math.sin(a[i + 2]), math.cos(a[i + 2])) for i in range(0, len(a), 3)]
            anchors = [anchors[i] for i in mask]
            # (the loss head restatement runs on the host: numpy + C geometry; the conv stack may be evaluated on any device)
            o, l, m, _ = yo.forward(x.cpu() if x.is_cuda else x, targets.cpu() if (targets is not None and targets.is_cuda) else targets,
                                    anchors, int(b["classes"]), img, float(b["ignore_thresh"]), use_giou)
            yolo_out.append(o)
            metrics.append(m)
            if targets is not None:
                loss = loss + l
        else:
            raise NotImplementedError(t)
        outs[ind] = x
    return (loss if targets is not None else None), torch.cat(yolo_out, 1), metrics
