"""Launch the SHIPPED hot kernels once each on bench-sized operands, for `ncu --set full` (one wide and two narrow
fprop layers, dgrad, wgrad, the three BatchNorm/activation passes, the rotated-GIoU pair kernel at 10^7 pairs).
Only the launches between cudaProfilerStart/Stop are captured:

    ncu --set full --clock-control none --import-source on --profile-from-start off \\
        -o gpurun_out/r2_hot_kernels python tools/ncu_targets.py

Each target is warmed up twice outside the capture window.  Shapes are layers of complex_yolov4.cfg @608, B=32
(SURVEY.md Appendix A).  Not a benchmark: numbers under ncu are never bench values."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import _lib, convops as co, synth
from cy4 import geometry as cg

B = 32
L = _lib.lib()
_lib.require_device()
st = _lib.stream()
only = set(sys.argv[1:])


def capture(name, fn):
    if only and name.split(":")[0] not in only:
        return
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("captured", name, flush=True)


def conv_targets(Cin, Cout, k, stride, H, which):
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // stride + 1
    torch.manual_seed(0)
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5)
    wp, wd = co.pack_fprop(w), co.pack_dgrad(w)
    dy = torch.randn(B, Ho, Ho, Cout, device="cuda").half()
    y = torch.empty(B, Ho, Ho, Cout, device="cuda", dtype=torch.float16)
    dx = torch.empty(B, H, H, Cin, device="cuda", dtype=torch.float16)
    s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
    acc = torch.zeros((Cout + 31) // 32 * 32, k * k, Cin, device="cuda")
    tag = "%dx%d_%dto%d_s%d_%d" % (k, k, Cin, Cout, stride, H)
    if "f" in which:
        capture("fprop:" + tag, lambda: co.conv_fwd(x, wp, Cout, k, stride, pad, out=y, stats=(s1, s2)))
    if "d" in which:
        capture("dgrad:" + tag, lambda: co.conv_dgrad(dy, wd, H, H, Cin, k, stride, pad, out=dx))
    if "w" in which:
        capture("wgrad:" + tag, lambda: co.conv_wgrad(x, dy, Cin, Cout, k, stride, pad, acc=acc))


conv_targets(256, 512, 3, 1, 38, "fdw")
conv_targets(128, 128, 3, 1, 76, "fdw")
conv_targets(64, 64, 3, 1, 152, "fdw")
conv_targets(512, 256, 1, 1, 38, "fd")
conv_targets(64, 128, 3, 2, 304, "fd")

# BatchNorm + activation passes on a 76x76x128 Mish layer (M = 184,832 rows) and a 304x304x64 one
for (H, C, act) in ((76, 128, 2), (304, 64, 2), (38, 512, 1)):
    M = B * H * H
    y16 = (torch.randn(M, C, device="cuda") * 1.5).half()
    g16 = torch.randn(M, C, device="cuda").half()
    out = torch.empty_like(y16); dy = torch.empty_like(y16)
    q = torch.rand(4, C, device="cuda") + 0.5
    sums = torch.zeros(2, C, device="cuda")
    tag = "%dx%dx%d_act%d" % (H, H, C, act)
    capture("bnfwd:" + tag, lambda: _lib.check(L.cy4_bn_act_fwd(y16.data_ptr(), C, q[0].data_ptr(), q[1].data_ptr(), act, None, 0, out.data_ptr(), C, M, C, st)))
    capture("bnred:" + tag, lambda: _lib.check(L.cy4_bn_act_bwd_reduce(y16.data_ptr(), C, g16.data_ptr(), C, q[0].data_ptr(), q[1].data_ptr(), q[2].data_ptr(),
                                                                      q[3].data_ptr(), act, M, C, sums[0].data_ptr(), sums[1].data_ptr(), st)))
    capture("bnapp:" + tag, lambda: _lib.check(L.cy4_bn_act_bwd_apply(y16.data_ptr(), C, g16.data_ptr(), C, q[0].data_ptr(), q[1].data_ptr(), q[2].data_ptr(),
                                                                     q[3].data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), 1.0 / M, 1, act, 1,
                                                                     dy.data_ptr(), C, M, C, st)))

p_, t_ = synth.make_pairs(10_000_000, seed=7)
pd, td = torch.tensor(p_, device="cuda"), torch.tensor(t_, device="cuda")
capture("rgiou:1e7", lambda: cg.rgiou_pairs(pd, td, True))
