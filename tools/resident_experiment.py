import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/complex-yolov4-pytorch_b200")
import torch
from cy4 import _lib, convops as co
L = _lib.lib()
B = 32
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (Cin, Cout, k, H, stride) in [(32, 64, 3, 304, 1), (64, 64, 3, 152, 1), (64, 64, 1, 304, 1), (128, 64, 1, 304, 1), (64, 32, 1, 304, 1), (128, 128, 1, 152, 1), (256, 128, 1, 76, 1)]:
    pad = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k, device="cuda") / 30
    dy = torch.randn(B, H, H, Cout, device="cuda").half()
    wp, wd = co.pack_fprop(w), co.pack_dgrad(w)
    y = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
    dx = torch.empty(B, H, H, Cin, device="cuda", dtype=torch.float16)
    r = []
    for res in (0, 1):
        L.cy4_set_option(b"resident_weights", res)
        r.append((timeit(lambda: co.conv_fwd(x, wp, Cout, k, 1, pad, out=y)), timeit(lambda: co.conv_dgrad(dy, wd, H, H, Cin, k, 1, pad, out=dx))))
    L.cy4_set_option(b"resident_weights", 0)
    print("%4d->%4d k%d @%3d fprop stream %6.0f us resident %6.0f us | dgrad stream %6.0f resident %6.0f" % (Cin, Cout, k, H, r[0][0], r[1][0], r[0][1], r[1][1]))
