"""Replaces the reference's src/models/darknet2pytorch.py (same public names)."""
from cy4.darknet import (Darknet, EmptyModule, GlobalAvgPool2d, MaxPoolDark, Mish, Reorg,  # noqa: F401
                         Upsample_expand, Upsample_interpolate)
