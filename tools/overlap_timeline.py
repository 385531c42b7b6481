"""Kernel timeline of ONE training step with the weight gradients on their second stream (CUPTI through torch.profiler's
chrome trace): how much of the conv_wgrad time actually runs next to a BN / activation pass, and what the step's kernels do
meanwhile.      python tools/overlap_timeline.py [out.json] [model knob NAME=INT ...] [--graph]
Prints: union busy time, sum of kernel durations, time with >= 2 kernels in flight, per-family overlap of the wgrad launches."""
import collections, json, os, re, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from cy4 import netdefs, synth
from cy4.darknet import Darknet
import bench

args = [a for a in sys.argv[1:] if not a.startswith("--")]
out = args[0] if args else None
torch.manual_seed(0)
net = Darknet(netdefs.cfg_path("complex_yolov4"), True).cuda().train()
net.use_cuda_graph = "--graph" in sys.argv
for kv in args[1:]:
    k_, v_ = kv.split("="); setattr(net, k_, int(v_))
opt = bench.make_optimizer(net)
x = synth.make_bev(32).cuda(); tg = torch.tensor(synth.make_targets(32, per_image=5)).cuda()


def step():
    loss, _ = net(x, tg); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)


for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = []
for e in json.load(open(path))["traceEvents"]:
    if e.get("cat") == "kernel":
        name = re.sub(r"<.*", "", e["name"].split("(")[0]).replace("void ", "").replace("cy4::", "")
        ev.append((float(e["ts"]), float(e["ts"]) + float(e["dur"]), name, int(e.get("args", {}).get("stream", 0))))
ev.sort()
t0 = ev[0][0]
streams = collections.Counter(s for _, _, _, s in ev)
main = streams.most_common(1)[0][0]
pts = sorted([(a, 1) for a, _, _, _ in ev] + [(b, -1) for _, b, _, _ in ev])
busy = multi = 0.0; depth = 0; last = pts[0][0]
for t, d in pts:
    if depth >= 1: busy += t - last
    if depth >= 2: multi += t - last
    depth += d; last = t
tot = sum(b - a for a, b, _, _ in ev)
print("kernels %d  streams %s  span %.2f ms  busy(union) %.2f ms  sum of durations %.2f ms  >=2 kernels in flight %.2f ms"
      % (len(ev), dict(streams), (ev[-1][1] - t0) / 1e3, busy / 1e3, tot / 1e3, multi / 1e3))
# overlap of each side-stream kernel with main-stream kernels, by main-stream family
side = [e for e in ev if e[3] != main and "wgrad" in e[2]]
mains = [e for e in ev if e[3] == main]
ov = collections.defaultdict(float); sdur = 0.0
j0 = 0
for a, b, n, s in side:
    sdur += b - a
    while j0 < len(mains) and mains[j0][1] < a: j0 += 1
    j = j0
    while j < len(mains) and mains[j][0] < b:
        o = min(b, mains[j][1]) - max(a, mains[j][0])
        if o > 0: ov[mains[j][2]] += o
        j += 1
print("side-stream wgrad launches: %d, %.2f ms in total; of that, next to main-stream kernels:" % (len(side), sdur / 1e3))
for k, v in sorted(ov.items(), key=lambda kv: -kv[1])[:8]:
    print("   %-40s %.2f ms" % (k, v / 1e3))
print("   (alone: %.2f ms)" % ((sdur - sum(ov.values())) / 1e3))
fam = collections.defaultdict(lambda: [0, 0.0])
for a, b, n, s in ev:
    fam[n][0] += 1; fam[n][1] += b - a
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[:10]:
    print("%-44s n=%4d %8.3f ms" % (k, v[0], v[1] / 1e3))
# a window of the backward pass for eyeballing: 60 kernels starting at the 40th wgrad launch
if side:
    k0 = ev.index(side[min(40, len(side) - 1)])
    for a, b, n, s in ev[max(0, k0 - 10):k0 + 50]:
        print("%9.1f us  +%7.1f  %s %s" % ((a - t0), b - a, "side" if s != main else "main", n))
if out:
    json.dump([[round(a - t0, 2), round(b - a, 2), n, 0 if s == main else 1] for a, b, n, s in ev], open(out, "w"))
