"""CPU restatement of the reference's evaluation post-processing.
TEST INFRASTRUCTURE ONLY -- imported by tests/ and bench tooling as the checker, never by the product.

Follows /root/reference/src/utils/evaluation_utils.py:
  post_processing_v2                  :322-357
  iou_rotated_single_vs_multi_boxes_cpu :186-210 (oracle/rbox_oracle.c orc_iou_matrix)
  get_batch_statistics_rotated_bbox   :152-183
  ap_per_class / compute_ap           :70-149
Pinned against the unmodified reference by oracle/gen_golden.py (tests/golden/eval_*.npz).
"""
import numpy as np

from . import geometry as og

F32 = np.float32


def iou_one_vs_many(single6, multi6):
    return og.iou_matrix(np.asarray(single6, F32).reshape(1, 6), multi6)[0]


def post_processing_v2(prediction, conf_thresh=0.95, nms_thresh=0.4):
    """prediction [B, N, 7+nC] fp32 -> list of [k, 9] fp32 arrays (merged box 6, conf, cls_conf, cls_pred) or None.
    Score ties are ordered by the lower row index (the reference's argsort leaves them undefined)."""
    prediction = np.asarray(prediction, F32)
    output = [None] * len(prediction)
    for image_i, image_pred in enumerate(prediction):
        image_pred = image_pred[image_pred[:, 6] >= F32(conf_thresh)]                       # :331
        if not image_pred.shape[0]:
            continue
        score = image_pred[:, 6] * image_pred[:, 7:].max(axis=1)                            # :336 (fp32 product)
        image_pred = image_pred[np.argsort(-score, kind="stable")]                          # :338
        class_confs = image_pred[:, 7:].max(axis=1, keepdims=True)
        class_preds = image_pred[:, 7:].argmax(axis=1)[:, None].astype(F32)                 # first maximum
        det = np.concatenate([image_pred[:, :7], class_confs, class_preds], axis=1).astype(F32)
        keep = []
        while det.shape[0]:
            large = iou_one_vs_many(det[0, :6], det[:, :6]) > F32(nms_thresh)               # :346
            invalid = large & (det[0, -1] == det[:, -1])                                    # :347-349
            w = det[invalid, 6:7]
            merged = det[0].copy()
            acc = np.zeros(6, F32)
            wsum = F32(0)
            for k in range(w.shape[0]):                                                     # :352, summed in list order, fp32
                acc = (acc + (w[k] * det[invalid][k, :6]).astype(F32)).astype(F32)
                wsum = F32(wsum + w[k, 0])
            with np.errstate(invalid="ignore", divide="ignore"):
                merged[:6] = acc / wsum
            keep.append(merged)
            invalid[0] = True                  # the reference never terminates if the head does not suppress itself
            det = det[~invalid]
        output[image_i] = np.stack(keep).astype(F32)
    return output


def get_batch_statistics(outputs, targets, iou_threshold):
    """outputs: list of [k,9] arrays / None; targets [nT,8] (img, cls, x, y, w, l, im, re; pixels).
    Returns [[true_positives, scores, labels], ...] for the images that have detections (:152-183)."""
    targets = np.asarray(targets, F32)
    batch_metrics = []
    for sample_i, output in enumerate(outputs):
        if output is None:
            continue
        pred_boxes, pred_scores, pred_labels = output[:, :6], output[:, 6], output[:, -1]
        tp = np.zeros(pred_boxes.shape[0])
        ann = targets[targets[:, 0] == sample_i][:, 1:]
        if len(ann) > 0:
            target_labels, target_boxes = ann[:, 0], ann[:, 1:]
            detected = []
            for pred_i in range(pred_boxes.shape[0]):
                if len(detected) == len(ann):
                    break
                if pred_labels[pred_i] not in target_labels:
                    continue
                ious = iou_one_vs_many(pred_boxes[pred_i], target_boxes)
                box_index = int(np.argmax(ious))                                            # first maximum
                if ious[box_index] >= F32(iou_threshold) and box_index not in detected:
                    tp[pred_i] = 1
                    detected.append(box_index)
        batch_metrics.append([tp, pred_scores, pred_labels])
    return batch_metrics


def compute_ap(recall, precision):
    """Area under the monotone precision envelope (:128-149)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def ap_per_class(tp, conf, pred_cls, target_cls):
    """(:70-125) -> precision, recall, AP, f1, classes."""
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
        fpc, tpc = (1 - tp[sel]).cumsum(), tp[sel].cumsum()
        recall = tpc / (n_gt + 1e-16)
        precision = tpc / (tpc + fpc)
        r.append(recall[-1]); p.append(precision[-1]); ap.append(compute_ap(recall, precision))
    p, r, ap = np.array(p), np.array(r), np.array(ap)
    return p, r, ap, 2 * p * r / (p + r + 1e-16), classes.astype("int32")
