"""A/B timing of kernel options on the full training step (one process, same network and inputs).
    python tools/ab_options.py [batch]
Prints ms/step for: defaults, the cy4_set_option tunables, the engine knobs (fused BN backward on/off), CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import _lib, netdefs, synth
from cy4.darknet import Darknet
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()
torch.manual_seed(0)
net = Darknet(netdefs.cfg_path("complex_yolov4"), True).cuda().train()
opt = bench.make_optimizer(net)
x = synth.make_bev(B).cuda(); tg = torch.tensor(synth.make_targets(B, per_image=5)).cuda()


def step():
    loss, _ = net(x, tg); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)


def timed(n=6, warm=3):
    for _ in range(warm):
        step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


base = {b"conv_cluster": 1, b"wgrad_cluster": 1, b"kblocks_per_slot": 4, b"conv1x1_matrix": 0, b"conv_pair": 1, b"wgrad_pair": 1}
for label, setting in [("defaults", {}), ("conv_pair=0", {b"conv_pair": 0}), ("conv1x1_matrix=1", {b"conv1x1_matrix": 1}),
                       ("wgrad_pair=0", {b"wgrad_pair": 0}), ("defaults again", {})]:
    for k, v in {**base, **setting}.items():
        L.cy4_set_option(k, v)
    print("%-34s %.3f ms/step" % (label, timed()), flush=True)
for k, v in base.items():
    L.cy4_set_option(k, v)
net.use_cuda_graph = True
try:
    print("%-34s %.3f ms/step" % ("defaults + CUDA graph", timed(warm=5)), flush=True)
except Exception as e:      # noqa: BLE001
    print("CUDA graph replay failed:", repr(e)[:300], flush=True)
net.use_cuda_graph = False
for mode in (1, 2, 0):
    net.fuse_bn_backward = mode
    print("%-34s %.3f ms/step" % ("fuse_bn_backward=%d" % mode, timed()), flush=True)
