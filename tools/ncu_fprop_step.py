"""One training-mode forward of complex_yolov4 (bs=32, 608x608) between cudaProfilerStart/Stop, eager launches, for an ncu
metric pass over exactly the conv fprop launches of a step (conv_pair_kernel + conv_tc_kernel, the stem's GEMM included):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \\
        --clock-control none --profile-from-start off -k regex:"conv_(tc|pair)_kernel" --csv --log-file gpurun_out/r2_fprop_metrics.csv \\
        python tools/ncu_fprop_step.py
    python tools/ncu_summarise.py fprop gpurun_out/r2_fprop_metrics.csv profiles/r2_ncu_fprop_launches.md

Three full training steps warm everything up outside the capture window.  Numbers under ncu are never bench values."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import netdefs, synth
from cy4.darknet import Darknet
import bench

torch.manual_seed(0)
net = Darknet(netdefs.cfg_path("complex_yolov4"), True).cuda().train()
net.use_cuda_graph = False
net.wgrad_overlap = 0
opt = bench.make_optimizer(net)
x = synth.make_bev(32).cuda(); tg = torch.tensor(synth.make_targets(32, per_image=5)).cuda()
for _ in range(3):
    loss, _ = net(x, tg); loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
with torch.no_grad():
    net(x, tg)                      # training-mode forward: batch statistics, BN passes, loss head -- no backward launches
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("captured one forward")
