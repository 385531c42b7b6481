"""Which resource bounds the conv kernels?  Times fprop / wgrad of a few layer shapes with the MMAs
or the TMA loads switched off (cy4_set_option("debug", 1|2)); results of those runs are garbage."""
import os, sys
# The "debug" switches only exist in the probe side build:
#   CY4_LIB_NAME=libcy4_probe.so CY4_EXTRA_NVCC_FLAGS=-DCY4_PROBE python complex-yolov4-pytorch_b200/csrc/build.py
# With the product library every column of the table is the full kernel.
if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "complex-yolov4-pytorch_b200", "csrc", "libcy4_probe.so")):
    os.environ.setdefault("CY4_LIB_NAME", "libcy4_probe.so")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import _lib, convops as co
L = _lib.lib()
B = 32
shapes = [(256, 512, 3, 38), (128, 128, 3, 76), (256, 256, 3, 38), (512, 1024, 3, 19), (64, 64, 3, 152), (32, 64, 3, 304), (64, 64, 1, 304)]


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (Cin, Cout, k, H) in shapes:
    pad = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k, device="cuda") / 30
    dy = torch.randn(B, H, H, rup := ((Cout + 63) // 64 * 64), device="cuda").half()
    wp = co.pack_fprop(w)
    y = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
    acc = torch.zeros((Cout + 31) // 32 * 32, k * k, Cin, device="cuda")
    fl = 2.0 * B * H * H * Cout * k * k * Cin
    row = []
    for dbg in (0, 1, 2):
        L.cy4_set_option(b"debug", dbg)
        t_f = timeit(lambda: co.conv_fwd(x, wp, Cout, k, 1, pad, out=y))
        t_w = timeit(lambda: co.conv_wgrad(x, dy[..., :Cout], Cin, Cout, k, 1, pad, acc=acc))
        row.append((t_f, t_w))
    L.cy4_set_option(b"debug", 0)
    print("%4d->%4d k%d @%3d | fprop full %6.0f us (%4.0f TF)  no-MMA %6.0f  no-load %6.0f | wgrad full %6.0f us (%4.0f TF)  no-MMA %6.0f  no-load %6.0f"
          % (Cin, Cout, k, H, row[0][0], fl / row[0][0] / 1e6, row[1][0], row[2][0], row[0][1], fl / row[0][1] / 1e6, row[1][1], row[2][1]))
