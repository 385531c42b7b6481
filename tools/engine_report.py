"""Runs the B200 step engine on a golden case and prints per-layer activation errors, loss,
metrics and parameter-gradient errors against the reference's golden outputs."""
import sys, os, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import numpy as np, torch
from cy4 import netdefs, synth
from cy4.darknet import Darknet

tag = sys.argv[1] if len(sys.argv) > 1 else "tiny_bs2"
g = np.load(os.path.join(ROOT, "tests", "golden", "darknet_%s.npz" % tag))
if len(sys.argv) > 2:
    model_scale = float(sys.argv[2])
else:
    model_scale = None
cfg = str(g["cfg"])
torch.manual_seed(0)
model = Darknet(netdefs.cfg_path(cfg), True).cuda()
model.train()
if model_scale is not None:
    model.grad_scale_target = model_scale
x = synth.make_bev(int(g["batch"])).cuda()
tg = torch.tensor(g["targets"]).cuda()
loss, out = model(x, tg)
loss.backward()
torch.cuda.synchronize()
print("loss", loss.item(), "ref", float(g["loss"][0]), "rel", abs(loss.item() - float(g["loss"][0])) / float(g["loss"][0]))
o = out.numpy(); ro = g["outputs"]
print("outputs maxabs diff", np.abs(o - ro).max(), "max rel-to-scale", (np.abs(o - ro) / (np.abs(ro) + 1.0)).max())
plan = model._engine.plan
worst = 0
for rec in plan.convs:
    i = rec["ind"]
    key = "act%d_idx" % i
    if key not in g.files:
        continue
    if rec.get("res") is not None:
        continue
    if "A" in rec:
        v = rec["A"]
        t = v.st.buf[..., v.off:v.off + v.C].float().permute(0, 3, 1, 2).reshape(-1)
    else:
        t = rec["P"].buf[..., :rec["Cout"]].float().permute(0, 3, 1, 2).reshape(-1)
    idx = torch.from_numpy(g[key]).cuda()
    got = t[idx].cpu().numpy(); ref = g["act%d_val" % i]
    st = g["act%d_stats" % i]
    err = np.abs(got - ref).max()
    worst = max(worst, err / max(st[1], 1e-6))
    print("layer %3d  C=%4d  maxerr %.3e  std %.3e  err/std %.3e  max|ref| %.2f" % (i, rec["Cout"], err, st[1], err / max(st[1], 1e-6), st[2]))
print("worst err/std", worst)
gw = 0
for name, p in model.named_parameters():
    if p.grad is None:
        print("NO GRAD", name); continue
    gn = g["gnorm/" + name]
    idx = torch.from_numpy(g["gidx/" + name]).cuda()
    got = p.grad.reshape(-1)[idx].cpu().numpy(); ref = g["gval/" + name]
    rel = np.abs(got - ref).max() / (gn[1] + 1e-12)
    nrel = abs(p.grad.norm().item() - gn[0]) / (gn[0] + 1e-12)
    gw = max(gw, rel)
    if not torch.isfinite(p.grad).all():
        print("NONFINITE grad", name, int((~torch.isfinite(p.grad)).sum()))
    if rel > 0.05 or nrel > 0.05:
        print("grad %-40s maxerr/maxabs %.3e  norm rel %.3e" % (name, rel, nrel))
print("worst grad err/max", gw)
for li, yl in enumerate(model.yolo_layers):
    ref = g["metrics%d" % li]
    mine = np.array(list(yl.metrics.values()))
    print("yolo", li, "metrics max rel", (np.abs(mine - ref) / (np.abs(ref) + 1e-3)).max())
