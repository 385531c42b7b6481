"""Drop-in for the reference's src/utils/evaluation_utils.py: the names evaluate.py and
test.py import, with the rotated-IoU / NMS / matching work on the GPU (cy4/evalops.py -> csrc/nms.cu) instead of
shapely loops on the CPU.  Return conventions follow the reference (CPU tensors / numpy arrays, None for images
without detections)."""
import numpy as np
import torch

from cy4 import evalops

__all__ = ["post_processing_v2", "post_processing", "iou_rotated_single_vs_multi_boxes_cpu", "get_batch_statistics_rotated_bbox",
           "ap_per_class", "compute_ap", "load_classes", "rescale_boxes", "get_corners_vectorize"]


def load_classes(path):
    """class names, one per line (reference :43-49)."""
    with open(path, "r") as fp:
        return fp.read().split("\n")[:-1]


def rescale_boxes(boxes, current_dim, original_shape):
    """Undo the letterbox padding of axis-aligned boxes (reference :52-67)."""
    orig_h, orig_w = original_shape
    scale = current_dim / max(original_shape)
    pad_x, pad_y = max(orig_h - orig_w, 0) * scale, max(orig_w - orig_h, 0) * scale
    unpad_h, unpad_w = current_dim - pad_y, current_dim - pad_x
    for c, (pad, unpad, orig) in enumerate([(pad_x, unpad_w, orig_w), (pad_y, unpad_h, orig_h)] * 2):
        boxes[:, c] = ((boxes[:, c] - pad // 2) / unpad) * orig
    return boxes


def get_corners_vectorize(x, y, w, l, yaw):
    """[n] arrays -> [n,4,2] fp32 corners, front-left / rear-left / rear-right / front-right (reference :213-239)."""
    x, y, w, l, yaw = (np.asarray(v, np.float32) for v in (x, y, w, l, yaw))
    c, s = np.cos(yaw), np.sin(yaw)
    out = np.zeros((x.shape[0], 4, 2), np.float32)
    for k, (sw, sl) in enumerate([(-1, -1), (-1, 1), (1, 1), (1, -1)]):
        out[:, k, 0] = x + sw * (w / 2 * c) + sl * (l / 2 * s)
        out[:, k, 1] = y + sw * (w / 2 * s) - sl * (l / 2 * c)
    return out


def iou_rotated_single_vs_multi_boxes_cpu(single_box, multi_boxes):
    """IoU of one (x, y, w, l, im, re) box with m boxes -> float tensor [m] on the CPU (reference :186-210)."""
    return evalops.iou_matrix(torch.as_tensor(single_box).reshape(1, 6), multi_boxes)[0].cpu()


def post_processing_v2(prediction, conf_thresh=0.95, nms_thresh=0.4):
    """Confidence filter + rotated NMS with confidence-weighted merging (reference :322-357).  Returns per image a
    [k, 9] tensor (x, y, w, l, im, re, conf, cls_conf, cls_pred) on the input's device, or None."""
    prediction = torch.as_tensor(prediction)
    return evalops.nms_v2(prediction, conf_thresh, nms_thresh).as_list(prediction.device)


def post_processing(outputs, conf_thresh=0.95, nms_thresh=0.4):
    """The older class-agnostic NMS of the reference (:277-319) is not on the evaluation path (evaluate.py and
    test.py call post_processing_v2); it is not rebuilt."""
    raise NotImplementedError("post_processing (v1) is not part of the rebuilt path; use post_processing_v2")


def get_batch_statistics_rotated_bbox(outputs, targets, iou_threshold):
    """[true_positives, scores, labels] per image that has detections (reference :152-183).  `outputs` is the list
    post_processing_v2 returned; the IoUs and the greedy matching run on the GPU."""
    B = len(outputs)
    cap = max([o.shape[0] for o in outputs if o is not None] + [1])
    out9 = torch.zeros(B, cap, 9, dtype=torch.float32)
    counts = torch.zeros(B, dtype=torch.int32)
    for i, o in enumerate(outputs):
        if o is not None:
            out9[i, :o.shape[0]] = torch.as_tensor(o, dtype=torch.float32).cpu()
            counts[i] = o.shape[0]
    dets = evalops.Detections(out9.cuda(), counts.cuda())
    tp = evalops.match(dets, targets, iou_threshold).cpu().numpy()
    metrics = []
    for i, o in enumerate(outputs):
        if o is None:
            continue
        o = torch.as_tensor(o).cpu()
        metrics.append([tp[i, :o.shape[0]].astype(np.float64), o[:, 6], o[:, -1]])
    return metrics


def compute_ap(recall, precision):
    """Area under the monotone envelope of the precision/recall curve (reference :128-149)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    steps = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[steps + 1] - mrec[steps]) * mpre[steps + 1])


def ap_per_class(tp, conf, pred_cls, target_cls):
    """precision, recall, AP, f1 per ground-truth class (reference :70-125)."""
    tp, conf, pred_cls, target_cls = (np.asarray(v) for v in (tp, conf, pred_cls, target_cls))
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes = np.unique(target_cls)
    ap, p, r = [], [], []
    for c in classes:
        sel = pred_cls == c
        n_gt, n_p = (target_cls == c).sum(), sel.sum()
        if n_p == 0 and n_gt == 0:
            continue
        if n_p == 0 or n_gt == 0:
            ap.append(0); r.append(0); p.append(0)
            continue
        tpc = tp[sel].cumsum()
        fpc = (1 - tp[sel]).cumsum()
        recall, precision = tpc / (n_gt + 1e-16), tpc / (tpc + fpc)
        r.append(recall[-1]); p.append(precision[-1]); ap.append(compute_ap(recall, precision))
    p, r, ap = np.array(p), np.array(r), np.array(ap)
    return p, r, ap, 2 * p * r / (p + r + 1e-16), classes.astype("int32")
