"""Replaces the reference's src/utils/cal_intersection_rotated_boxes.py (device kernels)."""
import torch

from cy4.geometry import PolyArea2D, intersection_area  # noqa: F401


class Line:
    """ax + by + c = 0 through p1, p2 (reference :16-39); kept for API completeness."""

    def __init__(self, p1, p2):
        self.a = p2[1] - p1[1]
        self.b = p1[0] - p2[0]
        self.c = p2[0] * p1[1] - p2[1] * p1[0]
        self.device = p1.device

    def cal_values(self, pts):
        return self.a * pts[:, 0] + self.b * pts[:, 1] + self.c

    def find_intersection(self, other):
        if not isinstance(other, Line):
            return NotImplemented
        w = self.a * other.b - self.b * other.a
        return torch.stack([(self.b * other.c - self.c * other.b) / w, (self.c * other.a - self.a * other.c) / w]).detach()
