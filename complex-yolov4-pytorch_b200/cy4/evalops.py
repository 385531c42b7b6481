"""Device-side evaluation post-processing: rotated IoU one-vs-many, post_processing_v2
(confidence filter, score sort, rotated NMS with confidence-weighted merging) and the true-positive matching of
evaluate.py, as hand-written sm_100a kernels behind include/cy4.h.  Replaces the shapely loops of the reference's
src/utils/evaluation_utils.py:152-210,322-357.  No CPU implementation: CPU tensors are moved to the GPU."""
import torch

from . import _lib


def _dev32(t):
    _lib.require_device()
    t = torch.as_tensor(t)
    if not t.is_cuda:
        t = t.cuda()
    return t.detach().to(torch.float32).contiguous()


def iou_matrix(a6, b6):
    """[n,6] x [m,6] -> [n,m] fp32 on the device (evaluation_utils.py:186-210 for every pair)."""
    L = _lib.lib()
    a, b = _dev32(a6).reshape(-1, 6), _dev32(b6).reshape(-1, 6)
    out = torch.empty(a.shape[0], b.shape[0], device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        _lib.check(L.cy4_rbox_iou_matrix(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], out.data_ptr(), _lib.stream()), "rbox_iou_matrix")
    return out


class Detections:
    """Device-resident result of nms_v2: boxes [B, max, 9], counts [B]."""

    def __init__(self, out9, counts):
        self.out9, self.counts = out9, counts
        self._host_counts = None

    def host_counts(self):
        if self._host_counts is None:
            self._host_counts = self.counts.cpu().tolist()
        return self._host_counts

    def as_list(self, device="cpu"):
        """The reference's return value: per image a [k,9] tensor, or None when nothing passed the filter."""
        cnt = self.host_counts()
        kmax = max(cnt) if cnt else 0
        rows = self.out9[:, :kmax].to(device) if kmax else None
        return [rows[i, :k].clone() if k else None for i, k in enumerate(cnt)]


def nms_v2(prediction, conf_thresh=0.95, nms_thresh=0.4):
    """prediction [B, N, 7+nC] -> Detections (post_processing_v2, evaluation_utils.py:322-357)."""
    L = _lib.lib()
    pred = _dev32(prediction)
    assert pred.dim() == 3 and pred.shape[2] >= 8, "prediction must be [B, N, 7 + num_classes]"
    B, N, row = pred.shape
    cap = L.cy4_nms_max_candidates()
    out9 = torch.empty(B, cap, 9, device=pred.device, dtype=torch.float32)
    counts = torch.zeros(B, device=pred.device, dtype=torch.int32)
    found = torch.zeros(B, device=pred.device, dtype=torch.int32)
    ws = torch.empty(max(L.cy4_nms_workspace_bytes(B), 1), device=pred.device, dtype=torch.uint8)
    with torch.cuda.device(pred.device):
        _lib.check(L.cy4_nms_rotated_v2(pred.data_ptr(), B, N, row - 7, float(conf_thresh), float(nms_thresh), out9.data_ptr(),
                                        counts.data_ptr(), found.data_ptr(), ws.data_ptr(), _lib.stream()), "nms_rotated_v2")
    fmax = int(found.max().item()) if B else 0
    if fmax > cap:
        raise RuntimeError("cy4 nms_v2: %d rows pass conf_thresh=%g in one image, more than the %d the kernel keeps "
                           "(the reference would need %d^2 shapely calls here); raise conf_thresh" % (fmax, conf_thresh, cap, fmax))
    return Detections(out9, counts)


def match(dets, targets, iou_threshold):
    """True-positive flags [B, max] (uint8, device) of Detections against targets [nT,8] (img, cls, x, y, w, l, im, re
    with x..l in pixels) -- get_batch_statistics_rotated_bbox, evaluation_utils.py:152-183."""
    L = _lib.lib()
    tg = _dev32(targets).reshape(-1, 8)
    B, cap = dets.out9.shape[0], dets.out9.shape[1]
    tp = torch.empty(B, cap, device=tg.device, dtype=torch.uint8)
    n_ann = torch.zeros(B, device=tg.device, dtype=torch.int32)
    with torch.cuda.device(tg.device):
        _lib.check(L.cy4_eval_match(dets.out9.data_ptr(), dets.counts.data_ptr(), B, cap, tg.data_ptr() if tg.numel() else None, tg.shape[0],
                                    float(iou_threshold), tp.data_ptr(), n_ann.data_ptr(), _lib.stream()), "eval_match")
    amax = int(n_ann.max().item()) if B else 0
    if amax > L.cy4_eval_max_annotations():
        raise RuntimeError("cy4 eval_match: %d annotations in one image (kernel limit %d)" % (amax, L.cy4_eval_max_annotations()))
    return tp
