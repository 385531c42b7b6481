"""Replaces the reference's src/utils/iou_rotated_boxes_utils.py with the device kernels
(same function names and return conventions; `polygons` objects are opaque device handles)."""
from cy4.geometry import (BoxSet, get_corners_vectorize, get_polygons_areas_fix_xy, iou_pred_vs_target_boxes,  # noqa: F401
                          iou_rotated_boxes_targets_vs_anchors)
from utils.cal_intersection_rotated_boxes import intersection_area, PolyArea2D  # noqa: F401


def cvt_box_2_polygon(box):
    """Reference :24-31 returns a shapely Polygon; here the [4,2] corner tensor itself is the polygon
    handle accepted by this module's functions."""
    return box
