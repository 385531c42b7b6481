"""Darknet(cfgfile, use_giou_loss): the reference's model class (src/models/darknet2pytorch.py:146-451)
on the B200 step engine.

Same constructor, attributes (`blocks, models, yolo_layers, width, height, num_classes, header, seen,
loss`), `state_dict()` keys/shapes, `forward(x, targets=None)` return convention, `print_network()`
and `load_weights()` as the reference, so train.py / evaluate.py / test.py and reference checkpoints
work unchanged.  The nn.Conv2d / nn.BatchNorm2d children only *hold* the fp32 parameters; the
computation runs through cy4.engine (hand-written sm_100a kernels, NHWC fp16 activations, fp32
accumulation) inside a single autograd node.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine
from .yolo import YoloLayer

__all__ = ["Darknet", "Mish", "MaxPoolDark", "Upsample_expand", "Upsample_interpolate", "Reorg", "GlobalAvgPool2d",
           "EmptyModule", "parse_cfg", "print_cfg"]


def parse_cfg(cfgfile):
    """darknet cfg grammar (reference src/models/darknet_utils.py:17-47): `[type]` sections,
    `key=value` lines, `#` comments; `type` inside a section is stored as `_type`;
    [convolutional] defaults batch_normalize=0."""
    blocks, block = [], None
    with open(cfgfile, "r") as fp:
        for raw in fp:
            line = raw.rstrip()
            if not line or line[0] == "#":
                continue
            if line[0] == "[":
                if block:
                    blocks.append(block)
                block = {"type": line.lstrip("[").rstrip("]")}
                if block["type"] == "convolutional":
                    block["batch_normalize"] = 0
            else:
                key, value = line.split("=")
                key = key.strip()
                block["_type" if key == "type" else key] = value.strip()
    if block:
        blocks.append(block)
    return blocks


def print_cfg(blocks):
    """One line per layer: index, type, filters, kernel/stride, input -> output size
    (same information as the reference's print_cfg, src/models/darknet_utils.py:50-196)."""
    print("layer     filters    size              input                output")
    w = h = c = 0
    shapes = []
    ind = -2
    for block in blocks:
        ind += 1
        t = block["type"]
        if t == "net":
            w, h, c = int(block["width"]), int(block["height"]), int(block["channels"])
            continue
        pw, ph, pc = w, h, c
        desc = ""
        if t == "convolutional":
            k, s = int(block["size"]), int(block["stride"])
            pad = (k - 1) // 2 if int(block["pad"]) else 0
            w, h, c = (pw + 2 * pad - k) // s + 1, (ph + 2 * pad - k) // s + 1, int(block["filters"])
            desc = "%4d  %d x %d / %d" % (c, k, k, s)
        elif t == "maxpool":
            k, s = int(block["size"]), int(block["stride"])
            w, h = (pw // s, ph // s) if s > 1 else (pw, ph)
            desc = "      %d x %d / %d" % (k, k, s)
        elif t == "upsample":
            s = int(block["stride"])
            w, h = pw * s, ph * s
            desc = "           * %d" % s
        elif t == "route":
            layers = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
            w, h = shapes[layers[0]][0], shapes[layers[0]][1]
            c = sum(shapes[l][2] for l in layers)
            if len(layers) == 1 and int(block.get("groups", 1)) > 1:
                c //= int(block["groups"])
            desc = "route " + " ".join(str(l) for l in layers)
        elif t == "shortcut":
            frm = int(block["from"])
            desc = "shortcut %d" % (frm if frm > 0 else frm + ind)
        elif t == "yolo":
            desc = "detection"
        shapes.append((w, h, c))
        print("%5d %-10s %-18s %4d x%4d x%4d   ->  %4d x%4d x%4d" % (ind, t[:10], desc, pw, ph, pc, w, h, c))


class Mish(nn.Module):
    """x * tanh(softplus(x)) (reference darknet2pytorch.py:22-28); the engine fuses it into the
    BatchNorm apply pass, this module is the plain definition for stand-alone use."""

    def forward(self, x):
        return x * torch.tanh(F.softplus(x))


class MaxPoolDark(nn.Module):
    """darknet-style max pool with asymmetric replicate padding (reference :31-61)."""

    def __init__(self, size=2, stride=1):
        super().__init__()
        self.size, self.stride = size, stride

    def forward(self, x):
        p = self.size // 2
        pads = []
        for dim in (3, 2):
            lo = (self.size - 1) // 2
            hi = lo + 1 if ((x.shape[dim] - 1) // self.stride) != ((x.shape[dim] + 2 * p - self.size) // self.stride) else lo
            pads += [lo, hi]
        return F.max_pool2d(F.pad(x, tuple(pads), mode="replicate"), self.size, stride=self.stride)


class Upsample_expand(nn.Module):
    """nearest-neighbour upsampling by an integer factor (reference :64-79)."""

    def __init__(self, stride=2):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        return x.repeat_interleave(self.stride, dim=2).repeat_interleave(self.stride, dim=3)


class Upsample_interpolate(nn.Module):
    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        return F.interpolate(x, size=(x.shape[2] * self.stride, x.shape[3] * self.stride), mode="nearest")


class Reorg(nn.Module):
    """space-to-depth (reference :98-117); not used by the complex-yolo cfgs."""

    def __init__(self, stride=2):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        return F.pixel_unshuffle(x, self.stride)


class GlobalAvgPool2d(nn.Module):
    def forward(self, x):
        return x.mean(dim=(2, 3))


class EmptyModule(nn.Module):
    """placeholder for route / shortcut (reference :136-141)."""

    def forward(self, x):
        return x


class Darknet(nn.Module):
    def __init__(self, cfgfile, use_giou_loss):
        super().__init__()
        self.use_giou_loss = use_giou_loss
        self.blocks = parse_cfg(cfgfile)
        self.width = int(self.blocks[0]["width"])
        self.height = int(self.blocks[0]["height"])
        self.models = self.create_network(self.blocks)
        self.yolo_layers = [layer for layer in self.models if layer.__class__.__name__ == "YoloLayer"]
        self.loss = self.models[len(self.models) - 1]
        self.header = torch.IntTensor([0, 0, 0, 0])
        self.seen = 0
        # engine knobs
        self.grad_scale_target = 256.0  # the fp16 gradient tensors are scaled so that max |d loss / d head| ~ this
        self.sync_outputs = False       # True: the CPU detections are complete when forward returns (training)
        self.use_cuda_graph = False     # True: after 2 eager steps the fwd / bwd launch sequences are replayed as CUDA graphs
        self.wgrad_overlap = 2          # backward: weight gradients on a second stream, next to the HBM-bound BN passes (0: one stream; 1: fork before dgrad)
        self.wgrad_priority = 0         # CUDA stream priority of that second stream (-1: its kernels are placed before the compute stream's)
        self.dy_ring = 4                # number of dY buffers the overlapped backward rotates through
        self.bn_shifted_stats = True    # training: BN statistics summed about the previous step's batch mean (no E[y^2]-E[y]^2 cancellation)
        self.engine_allreduce = False   # True (set by models.model_utils.make_data_parallel): the engine averages gradients over the ranks
                                        # itself, overlapped with backward; DDP then carries a no-op communication hook
        self.allreduce_groups = 6       # number of gradient groups of that exchange
        self.fuse_bn_backward = 0       # 0 (default): BN/activation backward as two bandwidth-bound passes.  1 / 2: the dgrad epilogue of a
                                        # tensor's last gradient writer does the first pass (cy4_conv_dgrad_fused) on long-K layers / wherever
                                        # possible -- measured SLOWER on B200 (profiles/r2_bn_backward_fusion.md), kept as a tested option
        self.fuse_eval = True           # eval() forward without grad: BatchNorm folded into the weights, activation in the conv epilogue
        self.outputs_on_device = False  # True: eval() forward returns the detections as a CUDA tensor (default: CPU, like the reference)
        self._engine = None

    # ------------------------------------------------------------------ forward
    def forward(self, x, targets=None):
        """x [B,3,H,W] fp32.  Returns the detections [B, sum(nA*G*G), 7+nC] on the CPU (reference
        :228) when targets is None, else (loss [1], detections)."""
        if self._engine is None:
            self._engine = engine.StepEngine(self)
        return self._engine.run(x, targets)

    def print_network(self):
        print_cfg(self.blocks)

    # ------------------------------------------------------------------ construction
    def create_network(self, blocks):
        """cfg blocks -> nn.ModuleList with the reference's child names (conv{k}/bn{k}/leaky{k}|mish{k})
        so that state_dict keys match reference checkpoints (reference :235-401)."""
        models = nn.ModuleList()
        prev_filters, conv_id, prev_stride = 3, 0, 1
        out_filters, out_strides = [], []
        for block in blocks:
            t = block["type"]
            if t == "net":
                prev_filters = int(block["channels"])
                continue
            if t == "convolutional":
                conv_id += 1
                bn = int(block["batch_normalize"])
                filters, k, stride = int(block["filters"]), int(block["size"]), int(block["stride"])
                pad = (k - 1) // 2 if int(block["pad"]) else 0
                act = block["activation"]
                seq = nn.Sequential()
                seq.add_module("conv%d" % conv_id, nn.Conv2d(prev_filters, filters, k, stride, pad, bias=not bn))
                if bn:
                    seq.add_module("bn%d" % conv_id, nn.BatchNorm2d(filters))
                if act == "leaky":
                    seq.add_module("leaky%d" % conv_id, nn.LeakyReLU(0.1, inplace=True))
                elif act == "relu":
                    seq.add_module("relu%d" % conv_id, nn.ReLU(inplace=True))
                elif act == "mish":
                    seq.add_module("mish%d" % conv_id, Mish())
                prev_filters = filters
                prev_stride *= stride
                models.append(seq)
            elif t == "maxpool":
                k, stride = int(block["size"]), int(block["stride"])
                if stride == 1 and k % 2:
                    models.append(nn.MaxPool2d(kernel_size=k, stride=stride, padding=k // 2))
                elif stride == k:
                    models.append(nn.MaxPool2d(kernel_size=k, stride=stride, padding=0))
                else:
                    models.append(MaxPoolDark(k, stride))
                prev_stride *= stride
            elif t == "upsample":
                stride = int(block["stride"])
                prev_stride //= stride
                models.append(Upsample_expand(stride))
            elif t == "route":
                ind = len(models)
                layers = [int(i) if int(i) > 0 else int(i) + ind for i in block["layers"].split(",")]
                if len(layers) == 1:
                    g = int(block.get("groups", 1))
                    prev_filters = out_filters[layers[0]] // g
                    prev_stride = out_strides[layers[0]]
                else:
                    prev_filters = sum(out_filters[l] for l in layers)
                    prev_stride = out_strides[layers[0]]
                models.append(EmptyModule())
            elif t == "shortcut":
                ind = len(models)
                prev_filters = out_filters[ind - 1]
                prev_stride = out_strides[ind - 1]
                models.append(EmptyModule())
            elif t == "yolo":
                mask = [int(i) for i in block["mask"].split(",")]
                a = [float(i) for i in block["anchors"].split(",")]
                anchors = [(a[i], a[i + 1], math.sin(a[i + 2]), math.cos(a[i + 2])) for i in range(0, len(a), 3)]
                anchors = [anchors[i] for i in mask]
                self.num_classes = int(block["classes"])
                models.append(YoloLayer(num_classes=self.num_classes, anchors=anchors, stride=prev_stride,
                                        scale_x_y=float(block["scale_x_y"]), ignore_thresh=float(block["ignore_thresh"])))
            elif t == "reorg":
                stride = int(block["stride"])
                prev_filters *= stride * stride
                prev_stride *= stride
                models.append(Reorg(stride))
            elif t == "avgpool":
                models.append(GlobalAvgPool2d())
            else:
                print("unknown type %s" % t)
                models.append(EmptyModule())
            out_filters.append(prev_filters)
            out_strides.append(prev_stride)
        return models

    # ------------------------------------------------------------------ darknet .weights
    def load_weights(self, weightfile):
        """darknet binary weights (reference :403-451 with darknet_utils.py:199-231): 5 int32 header,
        then per conv: [bn bias, bn weight, running mean, running var | conv bias], conv weight."""
        with open(weightfile, "rb") as fp:
            header = np.fromfile(fp, count=5, dtype=np.int32)
            buf = np.fromfile(fp, dtype=np.float32)
        self.header = torch.from_numpy(header)
        self.seen = self.header[3]
        start = 0

        def take(t):
            nonlocal start
            n = t.numel()
            with torch.no_grad():         # in-place copy that bumps t._version: the engine re-packs its fp16 weights
                t.copy_(torch.from_numpy(buf[start:start + n]).reshape(t.shape))
            start += n

        ind = -2
        for block in self.blocks:
            if start >= buf.size:
                break
            ind += 1
            if block["type"] != "convolutional":
                continue
            seq = self.models[ind]
            conv = seq[0]
            if int(block["batch_normalize"]):
                bn = seq[1]
                take(bn.bias); take(bn.weight); take(bn.running_mean); take(bn.running_var)
            else:
                take(conv.bias)
            take(conv.weight)
