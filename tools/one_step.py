"""Two training steps of complex_yolov4 at bs=32 (target for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import netdefs, synth
from cy4.darknet import Darknet
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
net = Darknet(netdefs.cfg_path("complex_yolov4"), True).cuda().train()
opt = bench.make_optimizer(net)
x = synth.make_bev(B).cuda(); tg = torch.tensor(synth.make_targets(B, per_image=5)).cuda()
for _ in range(2):
    loss, _ = net(x, tg); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("ok", loss.item())
