"""CPU: the static analysis of the step plan (cy4/engine.py Plan._analyse -- pure host logic, no device): layer shapes,
shortcut fusion and in-place concat placement for both cfgs.  DESIGN.md section 3 quotes these numbers."""
import types


def _analyse(cfg, size):
    from cy4 import engine, netdefs
    from cy4.darknet import Darknet
    model = Darknet(netdefs.cfg_path(cfg), True)
    plan = engine.Plan.__new__(engine.Plan)          # no buffers, no library: only the analysis pass
    plan.model, plan.H, plan.W = model, size, size
    return model, plan._analyse()


def test_complex_yolov4_plan():
    model, (info, consumers, fused, placement) = _analyse("complex_yolov4", 608)
    convs = [i for i, v in info.items() if v["type"] == "convolutional"]
    assert len(convs) == 110 and len(info) == 162 == len(model.models)
    # every one of the 23 residual units is fused: conv+BN writes the shortcut output, adding the residual
    shortcuts = [i for i, v in info.items() if v["type"] == "shortcut"]
    assert len(shortcuts) == 23 and sorted(fused) == shortcuts
    assert all(info[c]["type"] == "convolutional" and consumers[c] == [s] for s, c in fused.items())
    # multi-input routes: every source is produced in place inside the concat buffer (no copy kernels)
    routes = [i for i, v in info.items() if v["type"] == "route" and len(v["srcs"]) > 1]
    assert len(routes) == 10
    for r in routes:
        off = 0
        for s in info[r]["srcs"]:
            o = info[s]["origin"]
            assert placement.get(o) == (r, off), (r, s, o, placement.get(o))
            off += info[s]["C"]
        assert off == info[r]["C"]
    assert len(placement) == sum(len(info[r]["srcs"]) for r in routes)
    # head geometry of SURVEY appendix B: three linear 30-channel convs at 76 / 38 / 19
    yolos = [i for i, v in info.items() if v["type"] == "yolo"]
    assert [(info[y]["H"], info[y]["C"]) for y in yolos] == [(76, 30), (38, 30), (19, 30)]
    # SPP: three stride-1 max pools on the 19 x 19 x 512 tensor, concatenated to 2048 channels
    pools = [i for i, v in info.items() if v["type"] == "maxpool"]
    assert len(pools) == 3 and all(info[p]["H"] == 19 and info[p]["C"] == 512 for p in pools)
    assert any(info[r]["C"] == 2048 for r in routes)
    # conv MAC census (63,612,520,448 MAC per image)
    macs = 0
    for i in convs:
        b = info[i]["block"]
        cin = info[info[i]["srcs"][0]]["C"] if info[i]["srcs"] else 3
        macs += info[i]["H"] * info[i]["W"] * info[i]["C"] * cin * int(b["size"]) ** 2
    assert macs == 63612520448


def test_tiny_plan_with_group_routes():
    model, (info, consumers, fused, placement) = _analyse("complex_yolov4_tiny", 416)
    assert sum(v["type"] == "convolutional" for v in info.values()) == 21
    assert not fused                                           # the tiny cfg has no shortcut blocks
    # `groups=2, group_id=1` routes are channel slices: half the channels, never placed in a concat buffer
    slices = [i for i, v in info.items() if v["type"] == "route" and len(v["srcs"]) == 1 and int(v["block"].get("groups", 1)) > 1]
    assert slices and all(info[i]["C"] * 2 == info[info[i]["srcs"][0]]["C"] and info[i]["origin"] is None for i in slices)
    assert [(info[y]["H"], info[y]["C"]) for y, v in info.items() if v["type"] == "yolo"] == [(13, 30), (26, 30)]
