"""Import stubs that let the UNMODIFIED reference scripts run in this image.  TEST INFRASTRUCTURE ONLY.

The reference imports three packages the image does not have:
  * shapely      (src/utils/iou_rotated_boxes_utils.py:16, evaluation_utils.py:7) -> oracle/shapely_standin.py
  * easydict     (src/config/train_config.py:15, kitti_config.py)                 -> a dict with attribute access
  * matplotlib   (src/utils/train_utils.py:18, only used by its plotting helper)  -> an empty module tree
Real packages win when they are importable.
"""
import importlib.util
import sys
import types


class EasyDict(dict):
    """What the reference uses of easydict.EasyDict: attribute <-> item access, nested dicts converted."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def install():
    """Registers the stand-ins in sys.modules (idempotent).  Returns {package: 'real' | 'stand-in'}."""
    kinds = {}
    from . import shapely_standin
    kinds["shapely"] = shapely_standin.install()
    if importlib.util.find_spec("easydict") is None:
        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
        kinds["easydict"] = "stand-in"
    else:
        kinds["easydict"] = "real"
    if importlib.util.find_spec("matplotlib") is None:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        kinds["matplotlib"] = "stand-in"
    else:
        kinds["matplotlib"] = "real"
    return kinds
