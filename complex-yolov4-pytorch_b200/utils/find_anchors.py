"""Drop-in for the reference's src/utils/find_anchors.py: class Find_Anchors with the same
constructor, attributes and methods (compute_iou, avg_iou, kmeans, cluster / loop_cnt ...).  KITTI label loading is the
reference's own code (inherited when its module is importable); the num_boxes x num_anchors shapely loop of compute_iou /
kmeans / avg_iou runs as one kernel launch per iteration (cy4.anchors -> cy4_kmeans_iou).
`Find_Anchors.from_boxes(boxes_wh)` builds a solver from an [n,3] (w, l, yaw) array without a dataset."""
import importlib.util
import os
import sys

import numpy as np

from cy4 import anchors as _an


def _load_reference_class():
    here = os.path.dirname(os.path.abspath(__file__))
    for p in list(sys.path):
        cand = os.path.join(p, "utils", "find_anchors.py")
        if os.path.exists(cand) and os.path.dirname(os.path.abspath(cand)) != here:
            spec = importlib.util.spec_from_file_location("_cy4_ref_find_anchors", cand)
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
                return mod.Find_Anchors
            except Exception:            # shapely / dataset helpers missing: only from_boxes() is usable then
                return None
    return None


_Ref = _load_reference_class()


class Find_Anchors(_Ref if _Ref is not None else object):
    def __init__(self, dataset_dir, img_size, use_yaw_label=False):
        if _Ref is None:
            raise RuntimeError("the reference's utils/find_anchors.py (KITTI label loading) is not importable; "
                               "use Find_Anchors.from_boxes(boxes_wh) with an [n,3] (w, l, yaw) array")
        self.dataset_dir, self.img_size, self.use_yaw_label = dataset_dir, img_size, use_yaw_label
        self.lidar_dir = os.path.join(dataset_dir, 'training', "velodyne")
        self.image_dir = os.path.join(dataset_dir, 'training', "image_2")
        self.calib_dir = os.path.join(dataset_dir, 'training', "calib")
        self.label_dir = os.path.join(dataset_dir, 'training', "label_2")
        split_txt_path = os.path.join(dataset_dir, 'ImageSets', 'trainval.txt')
        self.image_idx_list = [x.strip() for x in open(split_txt_path).readlines()]
        self.sample_id_list = self.remove_invalid_idx(self.image_idx_list)
        self.boxes_wh = self.load_full_boxes_wh()
        self.num_boxes = self.boxes_wh.shape[0]
        print("number of sample_id_list: {}, num_boxes: {}".format(len(self.sample_id_list), self.num_boxes))

    @classmethod
    def from_boxes(cls, boxes_wh, img_size=608, use_yaw_label=True):
        self = cls.__new__(cls)
        self.dataset_dir, self.img_size, self.use_yaw_label = None, img_size, use_yaw_label
        self.boxes_wh = np.array(boxes_wh, dtype=np.float64)
        self.num_boxes = self.boxes_wh.shape[0]
        return self

    def compute_iou(self, i):
        """IoU of box i with every current cluster (reference :53-59), float32 [num_anchors]."""
        return _an.iou_matrix(self.boxes_wh[i:i + 1], self.cluster).cpu().numpy()[0]

    def avg_iou(self):
        return _an.avg_iou(self.boxes_wh, self.cluster)

    def kmeans(self, num_anchors):
        self.cluster, self.loop_cnt = _an.kmeans(self.boxes_wh, num_anchors, seed=0, verbose=True)
