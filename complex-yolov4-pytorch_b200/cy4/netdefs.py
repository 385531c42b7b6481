"""Network definitions for the two BASELINE.json architectures, built programmatically.

The reference describes its networks as darknet `.cfg` files (src/config/cfg/complex_yolov4.cfg,
complex_yolov4_tiny.cfg) and `Darknet(cfgfile)` takes a path.  The reference tree does not exist
on the GPU box, so the same block lists are constructed here from the architecture itself
(CSPDarknet53 + SPP + PANet, and the CSP-tiny variant) and written out in darknet cfg grammar on
demand (`cfg_path`).  Only the keys the reference's create_network reads are emitted
(src/models/darknet2pytorch.py:235-401); tests/test_netdefs.py checks block-for-block equality
with the reference files when /root/reference is present.
"""
import os

ANCHORS_V4 = "11, 15, 0, 10, 24, 0, 11, 25, 0, 23, 49, 0, 23, 55, 0, 24, 53, 0, 24, 60, 0, 27, 63, 0, 29, 74, 0"
ANCHORS_TINY = "11, 15, 0, 11, 25, 0, 23, 49, 0, 23, 55, 0, 24, 53, 0, 25, 61, 0"


def _conv(filters, size, stride=1, act="leaky", bn=1):
    return {"type": "convolutional", "batch_normalize": str(bn), "filters": str(filters), "size": str(size),
            "stride": str(stride), "pad": "1", "activation": act}


def _route(*layers, groups=None, group_id=None):
    b = {"type": "route", "layers": ",".join(str(l) for l in layers)}
    if groups is not None:
        b["groups"] = str(groups); b["group_id"] = str(group_id)
    return b


def _shortcut(frm):
    return {"type": "shortcut", "from": str(frm), "activation": "linear"}


def _yolo(mask, anchors, num, scale_x_y):
    return {"type": "yolo", "mask": ",".join(str(m) for m in mask), "anchors": anchors, "classes": "3",
            "num": str(num), "ignore_thresh": ".7", "scale_x_y": str(scale_x_y)}


def _csp_stage(width, n_res, first):
    """One CSPDarknet53 stage: stride-2 conv, two 1x1 splits, n_res residual units, merge."""
    half = width if first else width // 2
    blocks = [_conv(width, 3, 2, "mish"), _conv(half, 1, 1, "mish"), _route(-2), _conv(half, 1, 1, "mish")]
    for _ in range(n_res):
        blocks += [_conv(width // 2 if first else half, 1, 1, "mish"), _conv(half, 3, 1, "mish"), _shortcut(-3)]
    blocks += [_conv(half, 1, 1, "mish"), _route(-1, -(3 * n_res + 4)), _conv(width, 1, 1, "mish")]
    return blocks


def _five(a, b):
    """The PANet 1x1/3x3/1x1/3x3/1x1 bottleneck run."""
    return [_conv(a, 1), _conv(b, 3), _conv(a, 1), _conv(b, 3), _conv(a, 1)]


def complex_yolov4():
    net = {"type": "net", "width": "608", "height": "608", "channels": "3"}
    b = [_conv(32, 3, 1, "mish")]
    for width, n_res in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        b += _csp_stage(width, n_res, first=(width == 64))
    # SPP
    b += [_conv(512, 1), _conv(1024, 3), _conv(512, 1)]
    b += [{"type": "maxpool", "stride": "1", "size": "5"}, _route(-2),
          {"type": "maxpool", "stride": "1", "size": "9"}, _route(-4),
          {"type": "maxpool", "stride": "1", "size": "13"}, _route(-1, -3, -5, -6)]
    b += [_conv(512, 1), _conv(1024, 3), _conv(512, 1)]
    # PANet top-down
    b += [_conv(256, 1), {"type": "upsample", "stride": "2"}, _route(85), _conv(256, 1), _route(-1, -3)]
    b += _five(256, 512)
    b += [_conv(128, 1), {"type": "upsample", "stride": "2"}, _route(54), _conv(128, 1), _route(-1, -3)]
    b += _five(128, 256)
    # heads + bottom-up
    b += [_conv(256, 3), _conv(30, 1, 1, "linear", bn=0), _yolo((0, 1, 2), ANCHORS_V4, 9, 1.2)]
    b += [_route(-4), _conv(256, 3, 2), _route(-1, -16)] + _five(256, 512)
    b += [_conv(512, 3), _conv(30, 1, 1, "linear", bn=0), _yolo((3, 4, 5), ANCHORS_V4, 9, 1.1)]
    b += [_route(-4), _conv(512, 3, 2), _route(-1, -37)] + _five(512, 1024)
    b += [_conv(1024, 3), _conv(30, 1, 1, "linear", bn=0), _yolo((6, 7, 8), ANCHORS_V4, 9, 1.05)]
    return [net] + b


def complex_yolov4_tiny():
    net = {"type": "net", "width": "416", "height": "416", "channels": "3"}
    b = [_conv(32, 3, 2), _conv(64, 3, 2)]
    for w in (64, 128, 256):
        b += [_conv(w, 3), _route(-1, groups=2, group_id=1), _conv(w // 2, 3), _conv(w // 2, 3), _route(-1, -2),
              _conv(w, 1), _route(-6, -1), {"type": "maxpool", "size": "2", "stride": "2"}]
    b += [_conv(512, 3), _conv(256, 1), _conv(512, 3), _conv(30, 1, 1, "linear", bn=0),
          _yolo((3, 4, 5), ANCHORS_TINY, 6, 1.05)]
    b += [_route(-4), _conv(128, 1), {"type": "upsample", "stride": "2"}, _route(-1, 23), _conv(256, 3),
          _conv(30, 1, 1, "linear", bn=0), _yolo((0, 1, 2), ANCHORS_TINY, 6, 1.05)]
    return [net] + b


NETS = {"complex_yolov4": complex_yolov4, "complex_yolov4_tiny": complex_yolov4_tiny}


def to_cfg_text(blocks):
    out = []
    for blk in blocks:
        out.append("[%s]" % blk["type"])
        for k, v in blk.items():
            if k != "type":
                out.append("%s=%s" % ("type" if k == "_type" else k, v))
        out.append("")
    return "\n".join(out)


def cfg_path(name):
    """Writes (once) and returns the path of a darknet cfg for `name` (with or without .cfg)."""
    name = name[:-4] if name.endswith(".cfg") else name
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, name + ".cfg")
    text = to_cfg_text(NETS[name]())
    if not os.path.exists(path) or open(path).read() != text:
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "w") as f:
            f.write(text)
        os.replace(tmp, path)
    return path
