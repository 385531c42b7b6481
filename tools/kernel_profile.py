"""Per-kernel GPU time of one training step (CUPTI via torch.profiler; low overhead, no replays).
    python tools/kernel_profile.py [cfg] [batch] [out.json]
Timing summaries for the judge come from ncu (profiles/); this is the fast inner-loop view."""
import collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from cy4 import netdefs, synth
from cy4.darknet import Darknet
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "complex_yolov4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
net = Darknet(netdefs.cfg_path(cfg), True).cuda().train()
for kv in sys.argv[4:]:                  # engine knobs, e.g. wgrad_overlap=0 (per-kernel times without concurrent kernels); opt:NAME=INT -> cy4_set_option
    k_, v_ = kv.split("=")
    if k_.startswith("opt:"):
        from cy4 import _lib
        _lib.check(_lib.lib().cy4_set_option(k_[4:].encode(), int(v_)), kv)
    else:
        setattr(net, k_, int(v_))
opt = bench.make_optimizer(net)
x = synth.make_bev(B).cuda(); tg = torch.tensor(synth.make_targets(B, per_image=5)).cuda()


def step():
    loss, _ = net(x, tg); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower():
        name = e.name.split("(")[0].replace("void ", "").replace("cy4::", "")
        if not name.startswith("bn_act"):                  # (keep the <activation, ...> template arguments of the BN passes)
            name = re.sub(r"<.*", "", name)
        name = name[-56:]
        agg[name][0] += 1; agg[name][1] += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
        tot += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
seq = [(re.sub(r"<.*", "", e.name.split("(")[0]).replace("void ", "").replace("cy4::", ""), (e.device_time if hasattr(e, "device_time") else e.cuda_time))
       for e in sorted(prof.events(), key=lambda e: e.time_range.start) if e.device_type is not None and "cuda" in str(e.device_type).lower()]
print("total kernel ms %.2f" % tot)
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for k, v in rows[:28]:
    print("%-58s n=%4d %8.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
if len(sys.argv) > 3:
    json.dump({"agg": {k: v for k, v in rows}, "conv_tc_us": [t for n, t in seq if n.endswith("conv_tc_kernel")],
               "wgrad_us": [t for n, t in seq if n.endswith("conv_wgrad_kernel")]}, open(sys.argv[3], "w"), indent=0)
