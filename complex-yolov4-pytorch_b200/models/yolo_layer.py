"""Replaces the reference's src/models/yolo_layer.py (same public names)."""
from cy4.yolo import YoloLayer  # noqa: F401
