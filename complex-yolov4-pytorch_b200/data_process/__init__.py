"""Drop-in `data_process` package (see models/__init__.py): kitti_bev_utils is replaced here, the other modules
(kitti_dataset, transformation, ...) fall through to the reference's src/data_process/ when it is on sys.path."""
import os
import sys

for _p in list(sys.path):
    _cand = os.path.join(_p, "data_process")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)) and \
            os.path.exists(os.path.join(_cand, "kitti_dataset.py")):
        __path__.append(_cand)
        break
