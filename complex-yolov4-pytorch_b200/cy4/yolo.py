"""YoloLayer: the reference's loss head (src/models/yolo_layer.py) on hand-written sm_100a kernels.

Same constructor, attributes, `forward(x, targets, img_size, use_giou_loss)` and
`build_targets(pred_boxes, pred_cls, target, anchors)` as the reference.  Differences a caller can
observe are limited to performance: `metrics` is filled lazily (one device->host copy on first
read instead of 18 `.item()` syncs per layer per step), and nothing in forward
synchronises the host.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import geometry

METRIC_KEYS = ("loss", "iou_score", "giou_loss", "loss_x", "loss_y", "loss_w", "loss_h", "loss_eular", "loss_im",
               "loss_re", "loss_obj", "loss_cls", "cls_acc", "recall50", "recall75", "precision", "conf_obj",
               "conf_noobj")


class LazyMetrics(dict):
    """dict of the 18 reference metrics (yolo_layer.py:232-251) materialised on first access."""

    def __init__(self, tensor=None):
        super().__init__()
        self._t = tensor

    def _fill(self):
        if self._t is not None:
            vals = self._t.detach().to("cpu").tolist()
            self._t = None
            super().update(zip(METRIC_KEYS, vals))

    def __getitem__(self, k): self._fill(); return super().__getitem__(k)
    def __iter__(self): self._fill(); return super().__iter__()
    def __len__(self): self._fill(); return super().__len__()
    def __contains__(self, k): self._fill(); return super().__contains__(k)
    def __repr__(self): self._fill(); return super().__repr__()
    def keys(self): self._fill(); return super().keys()
    def values(self): self._fill(); return super().values()
    def items(self): self._fill(); return super().items()
    def get(self, k, d=None): self._fill(); return super().get(k, d)
    def copy(self): self._fill(); return dict(self)


def make_desc(B, G, nA, nC, strides, img_size, ignore_thresh, use_giou):
    d = _lib.YoloDesc()
    d.B, d.G, d.nA, d.nC = int(B), int(G), int(nA), int(nC)
    d.sB, d.sC, d.sH, d.sW = (int(s) for s in strides)
    d.img_size = float(img_size)
    d.ignore_thresh = float(ignore_thresh)
    d.use_giou = 1 if use_giou else 0
    return d


def check_status(status, where="yolo"):
    """Raises the reference's IndexError if a target fell outside the grid / batch (needs a sync)."""
    if int(status.item()) & 1:
        raise IndexError("%s: target index out of range (image id, class or x/y == 1.0); "
                         "reference yolo_layer.py:114 raises here too" % where)


class _YoloLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, targets, anchors4, nC, img_size, ignore_thresh, use_giou):
        L = _lib.lib()
        B, C, G, _ = x.shape
        nA = anchors4.shape[0]
        xs = x.detach()
        if xs.dtype != torch.float32:
            xs = xs.float()
        d = make_desc(B, G, nA, nC, xs.stride(), img_size, ignore_thresh, use_giou)
        nT = int(targets.shape[0])
        tg = targets.detach().to(x.device, torch.float32).contiguous()
        out = torch.empty(B, nA * G * G, 7 + nC, device=x.device, dtype=torch.float32)
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        metrics = torch.empty(18, device=x.device, dtype=torch.float32)
        status = torch.empty(1, device=x.device, dtype=torch.int32)
        ws = torch.empty(L.cy4_yolo_workspace_bytes(ctypes.byref(d), nT), device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device):
            _lib.check(L.cy4_yolo_loss_fwd(ctypes.byref(d), xs.data_ptr(), anchors4.data_ptr(),
                                           tg.data_ptr() if nT else None, nT, out.data_ptr(), loss.data_ptr(),
                                           metrics.data_ptr(), status.data_ptr(), ws.data_ptr(), _lib.stream()),
                       "yolo_loss_fwd")
        ctx.save_for_backward(xs, tg, anchors4, ws)
        ctx.desc = d
        ctx.nT = nT
        ctx.mark_non_differentiable(out, metrics, status)
        return out, loss, metrics, status

    @staticmethod
    def backward(ctx, _go, gloss, _gm, _gs):
        L = _lib.lib()
        xs, tg, anchors4, ws = ctx.saved_tensors
        d = ctx.desc
        dx = torch.empty(xs.shape, device=xs.device, dtype=torch.float32)   # contiguous NCHW
        g = gloss.reshape(-1)[:1].to(torch.float32).contiguous()
        sb, sc, sh, sw = dx.stride()
        with torch.cuda.device(xs.device):
            _lib.check(L.cy4_yolo_loss_bwd(ctypes.byref(d), xs.data_ptr(), anchors4.data_ptr(),
                                           tg.data_ptr() if ctx.nT else None, ctx.nT, g.data_ptr(), ws.data_ptr(),
                                           dx.data_ptr(), sb, sc, sh, sw, _lib.stream()), "yolo_loss_bwd")
        return dx, None, None, None, None, None, None


class YoloLayer(nn.Module):
    """Yolo layer (reference src/models/yolo_layer.py:27-253)."""

    def __init__(self, num_classes, anchors, stride, scale_x_y, ignore_thresh):
        super(YoloLayer, self).__init__()
        self.num_classes = num_classes
        self.anchors = anchors
        self.num_anchors = len(anchors)
        self.stride = stride
        self.scale_x_y = scale_x_y          # stored, unused -- as in the reference
        self.ignore_thresh = ignore_thresh
        self.noobj_scale = 100
        self.obj_scale = 1
        self.lgiou_scale = 3.54
        self.leular_scale = 3.54
        self.lobj_scale = 64.3
        self.lcls_scale = 37.4
        self.seen = 0
        self.grid_size = 0
        self.img_size = 0
        self.metrics = {}
        self.check_targets = False          # True: sync and raise IndexError like the reference
        self._status = None

    def compute_grid_offsets(self, grid_size):
        """Reference :53-67.  Keeps the same attributes (grid_x, grid_y, scaled_anchors, ...)."""
        self.grid_size = grid_size
        g = self.grid_size
        self.stride = self.img_size / self.grid_size
        self.grid_x = torch.arange(g, device=self.device, dtype=torch.float).repeat(g, 1).view([1, 1, g, g])
        self.grid_y = torch.arange(g, device=self.device, dtype=torch.float).repeat(g, 1).t().view([1, 1, g, g])
        self.scaled_anchors = torch.tensor(
            [(a_w / self.stride, a_h / self.stride, im, re) for a_w, a_h, im, re in self.anchors], device=self.device,
            dtype=torch.float)
        self.anchor_w = self.scaled_anchors[:, 0:1].view((1, self.num_anchors, 1, 1))
        self.anchor_h = self.scaled_anchors[:, 1:2].view((1, self.num_anchors, 1, 1))
        self.scaled_anchors_polygons, self.scaled_anchors_areas = geometry.get_polygons_areas_fix_xy(self.scaled_anchors)

    def build_targets(self, pred_boxes, pred_cls, target, anchors):
        """Reference :69-142.  Returns the same 13-tuple (masks as torch.bool)."""
        L = _lib.lib()
        _lib.require_device()
        dev_in = pred_boxes.device
        pb = pred_boxes.detach().float().cuda().contiguous() if not pred_boxes.is_cuda else pred_boxes.detach().float().contiguous()
        pc = pred_cls.detach().float().to(pb.device).contiguous()
        tg = target.detach().float().to(pb.device).contiguous()
        an = anchors.detach().float().to(pb.device).contiguous()
        nB, nA, nG, _, nC = pc.shape
        nT = tg.shape[0]
        use_giou = bool(getattr(self, "use_giou_loss", False))
        img_size = self.img_size if self.img_size else nG * float(self.stride)
        d = make_desc(nB, nG, nA, nC, (0, 0, 0, 0), img_size, self.ignore_thresh, use_giou)
        f = lambda *s: torch.empty(*s, device=pb.device, dtype=torch.float32)
        iou_scores, class_mask = f(nB, nA, nG, nG), f(nB, nA, nG, nG)
        tx, ty, tw, th, tim, tre, tconf = (f(nB, nA, nG, nG) for _ in range(7))
        tcls = f(nB, nA, nG, nG, nC)
        obj = torch.empty(nB, nA, nG, nG, device=pb.device, dtype=torch.uint8)
        noobj = torch.empty(nB, nA, nG, nG, device=pb.device, dtype=torch.uint8)
        giou_loss = f(1)
        status = torch.empty(1, device=pb.device, dtype=torch.int32)
        ws = torch.empty(L.cy4_yolo_workspace_bytes(ctypes.byref(d), nT), device=pb.device, dtype=torch.uint8)
        with torch.cuda.device(pb.device):
            _lib.check(L.cy4_build_targets(ctypes.byref(d), pb.data_ptr(), pc.data_ptr(), tg.data_ptr() if nT else None, nT,
                                           an.data_ptr(), iou_scores.data_ptr(), giou_loss.data_ptr(), class_mask.data_ptr(),
                                           obj.data_ptr(), noobj.data_ptr(), tx.data_ptr(), ty.data_ptr(), tw.data_ptr(),
                                           th.data_ptr(), tim.data_ptr(), tre.data_ptr(), tcls.data_ptr(), tconf.data_ptr(),
                                           None, status.data_ptr(), ws.data_ptr(), _lib.stream()), "build_targets")
        check_status(status, "build_targets")
        outs = (iou_scores, giou_loss, class_mask, obj.type(torch.bool), noobj.type(torch.bool),
                tx, ty, tw, th, tim, tre, tcls, tconf)
        return tuple(o.to(dev_in) for o in outs)

    def forward(self, x, targets=None, img_size=608, use_giou_loss=False):
        """Reference :144-253.
        x [B, nA*(7+nC), G, G]; targets [nT, 8] (image, class, x, y, w, l, im, re) or None.
        Returns (output [B, nA*G*G, 7+nC], 0) or (output, total_loss)."""
        _lib.require_device()
        L = _lib.lib()
        self.img_size = img_size
        self.use_giou_loss = use_giou_loss
        dev_in = x.device
        if not x.is_cuda:
            x = x.cuda()
        self.device = x.device
        num_samples, _, _, grid_size = x.size()
        if grid_size != self.grid_size or getattr(self, "scaled_anchors", None) is None \
                or self.scaled_anchors.device != x.device:
            self.compute_grid_offsets(grid_size)
        anchors4 = self.scaled_anchors
        if targets is None:
            xs = x.detach().float()
            d = make_desc(num_samples, grid_size, self.num_anchors, self.num_classes, xs.stride(), img_size,
                          self.ignore_thresh, use_giou_loss)
            out = torch.empty(num_samples, self.num_anchors * grid_size * grid_size, 7 + self.num_classes,
                              device=x.device, dtype=torch.float32)
            with torch.cuda.device(x.device):
                _lib.check(L.cy4_yolo_decode(ctypes.byref(d), xs.data_ptr(), anchors4.data_ptr(), out.data_ptr(),
                                             _lib.stream()), "yolo_decode")
            return out.to(dev_in), 0
        self.reduction = 'mean'
        out, loss, metrics, status = _YoloLossFn.apply(x, targets, anchors4, self.num_classes, float(img_size),
                                                       float(self.ignore_thresh), bool(use_giou_loss))
        self._status = status
        if self.check_targets:
            check_status(status, "YoloLayer.forward")
        self.metrics = LazyMetrics(metrics)
        total_loss = loss if use_giou_loss else loss.reshape(())      # shape [1] vs 0-dim
        return out.to(dev_in), total_loss.to(dev_in)
