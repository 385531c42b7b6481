"""Drop-in `utils` package (see models/__init__.py): replaced modules live here, the rest falls
through to the reference's src/utils/ when it is on sys.path."""
import os
import sys

for _p in list(sys.path):
    _cand = os.path.join(_p, "utils")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)) and \
            os.path.exists(os.path.join(_cand, "iou_rotated_boxes_utils.py")):
        __path__.append(_cand)
        break
