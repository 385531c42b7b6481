"""Rotated-box geometry on the GPU: the device-side replacements of the reference's
src/utils/iou_rotated_boxes_utils.py and src/utils/cal_intersection_rotated_boxes.py.

Same function names, argument meaning and return conventions as the reference; every function
runs hand-written sm_100a kernels through the C-ABI (include/cy4.h).  CPU tensors are moved to
the current CUDA device, computed there and moved back (there is no CPU implementation).
"""
import torch

from . import _lib

F_GIOU = 1


def _dev(t):
    """-> (cuda fp32 contiguous tensor, original device)"""
    _lib.require_device()
    dev = t.device
    if not t.is_cuda:
        t = t.cuda()
    return t.detach().to(torch.float32).contiguous(), dev


class _RGIoUPairs(torch.autograd.Function):
    """Element-wise pred/target pairs -> (ious [n], term [n]); backward follows the reference's
    autograd semantics (detached intersection points)."""

    @staticmethod
    def forward(ctx, pred, target, giou):
        L = _lib.lib()
        p, dev = _dev(pred)
        t, _ = _dev(target)
        n = p.shape[0]
        iou = torch.empty(n, device=p.device, dtype=torch.float32)
        term = torch.empty(n, device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            _lib.check(L.cy4_rgiou_pairs(p.data_ptr(), t.data_ptr(), n, F_GIOU if giou else 0, iou.data_ptr(),
                                         term.data_ptr(), None, None, _lib.stream()), "rgiou_pairs")
        ctx.save_for_backward(p, t)
        ctx.giou = bool(giou)
        ctx.dev = dev
        ctx.mark_non_differentiable(iou)
        return iou.to(dev), term.to(dev)

    @staticmethod
    def backward(ctx, _giou, gterm):
        L = _lib.lib()
        p, t = ctx.saved_tensors
        n = p.shape[0]
        g = gterm.to(p.device, torch.float32).contiguous()
        iou = torch.empty(n, device=p.device, dtype=torch.float32)
        term = torch.empty(n, device=p.device, dtype=torch.float32)
        gp = torch.empty(n, 6, device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            _lib.check(L.cy4_rgiou_pairs(p.data_ptr(), t.data_ptr(), n, F_GIOU if ctx.giou else 0, iou.data_ptr(),
                                         term.data_ptr(), g.data_ptr(), gp.data_ptr(), _lib.stream()), "rgiou_pairs bwd")
        return gp.to(ctx.dev), None, None


def rgiou_pairs(pred_boxes, target_boxes, GIoU=True):
    """(ious [n] detached, terms [n] differentiable w.r.t. pred_boxes)."""
    return _RGIoUPairs.apply(pred_boxes, target_boxes, GIoU)


def iou_pred_vs_target_boxes(pred_boxes, target_boxes, GIoU=False, DIoU=False, CIoU=False):
    """Reference src/utils/iou_rotated_boxes_utils.py:98-142: element-wise pairs [n,6]
    (x, y, w, l, im, re) -> (ious [n] (no grad), giou_loss [1] = sum of the per-pair terms)."""
    assert pred_boxes.size() == target_boxes.size(), "Unmatch size of pred_boxes and target_boxes"
    if DIoU or CIoU:
        raise NotImplementedError
    n = pred_boxes.size(0)
    if n == 0:
        return (torch.tensor([], device=pred_boxes.device, dtype=torch.float),
                torch.tensor([0.], device=pred_boxes.device, dtype=torch.float))
    ious, terms = _RGIoUPairs.apply(pred_boxes, target_boxes, GIoU)
    return ious, _SeqSum.apply(terms)


class _SeqSum(torch.autograd.Function):
    """sum in the reference's accumulation order (`giou_loss += term`), shape [1]."""

    @staticmethod
    def forward(ctx, terms):
        L = _lib.lib()
        t, dev = _dev(terms)
        out = torch.empty(1, device=t.device, dtype=torch.float32)
        with torch.cuda.device(t.device):
            _lib.check(L.cy4_sum_f32_seq(t.data_ptr(), t.numel(), out.data_ptr(), _lib.stream()), "sum")
        ctx.n = t.numel()
        return out.to(dev)

    @staticmethod
    def backward(ctx, g):
        return g.expand(ctx.n)


def get_corners_vectorize(x, y, w, l, yaw):
    """Reference :34-61 -> [n,4,2] corners (front-left, rear-left, rear-right, front-right)."""
    L = _lib.lib()
    xs, dev = _dev(x)
    ys, ws, ls, yaws = (_dev(v)[0] for v in (y, w, l, yaw))
    n = xs.numel()
    out = torch.empty(n, 4, 2, device=xs.device, dtype=torch.float32)
    with torch.cuda.device(xs.device):
        _lib.check(L.cy4_corners(xs.data_ptr(), ys.data_ptr(), ws.data_ptr(), ls.data_ptr(), yaws.data_ptr(), n,
                                 out.data_ptr(), _lib.stream()), "corners")
    return out.to(dev)


class BoxSet:
    """What get_polygons_areas_fix_xy returns in place of the reference's list of shapely polygons:
    the (w, l, im, re) rows on the device.  Opaque to callers, consumed by
    iou_rotated_boxes_targets_vs_anchors."""

    def __init__(self, wlimre):
        self.wlimre = wlimre

    def __len__(self):
        return self.wlimre.shape[0]


def get_polygons_areas_fix_xy(boxes, fix_xy=100.):
    """Reference :64-79.  boxes [n,4] (w, l, im, re) -> (polygons, areas [n])."""
    assert float(fix_xy) == 100., "the device kernel places boxes at the reference's fix_xy=100"
    b, dev = _dev(boxes)
    return BoxSet(b), (boxes[:, 0] * boxes[:, 1])


def iou_rotated_boxes_targets_vs_anchors(anchors_polygons, anchors_areas, targets_polygons, targets_areas):
    """Reference :82-95 -> ious [nA, nT] fp32."""
    L = _lib.lib()
    a, t = anchors_polygons.wlimre, targets_polygons.wlimre
    nA, nT = a.shape[0], t.shape[0]
    out = torch.zeros(nA, nT, device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _lib.check(L.cy4_anchor_iou(a.data_ptr(), nA, t.data_ptr(), nT, out.data_ptr(), _lib.stream()), "anchor_iou")
    return out.to(anchors_areas.device)


def intersection_area(rect1, rect2):
    """Reference src/utils/cal_intersection_rotated_boxes.py:42-90 for one pair of [4,2] quads
    (or a batch [n,4,2]); returns a 0-dim tensor (or [n])."""
    L = _lib.lib()
    r1, dev = _dev(rect1)
    r2, _ = _dev(rect2)
    single = r1.dim() == 2
    r1 = r1.reshape(-1, 4, 2); r2 = r2.reshape(-1, 4, 2)
    n = r1.shape[0]
    out = torch.empty(n, device=r1.device, dtype=torch.float32)
    with torch.cuda.device(r1.device):
        _lib.check(L.cy4_quad_intersection_area(r1.data_ptr(), r2.data_ptr(), n, out.data_ptr(), _lib.stream()), "inter")
    out = out.to(dev)
    return out[0] if single else out


def PolyArea2D(pts):
    """Reference src/utils/cal_intersection_rotated_boxes.py:93-96 (k <= 16 vertices)."""
    L = _lib.lib()
    p_, dev = _dev(pts)
    out = torch.empty(1, device=p_.device, dtype=torch.float32)
    with torch.cuda.device(p_.device):
        _lib.check(L.cy4_poly_area(p_.data_ptr(), p_.shape[0], out.data_ptr(), _lib.stream()), "poly_area")
    return out.to(dev)[0]
