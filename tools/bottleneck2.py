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
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (Cin, Cout, k, H) in [(32, 64, 3, 304), (64, 64, 3, 152), (64, 64, 1, 304), (128, 128, 3, 76), (128, 128, 1, 152), (256, 256, 3, 38), (256, 512, 3, 38)]:
    pad = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k, device="cuda") / 30
    wp = co.pack_fprop(w)
    y = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
    res = {}
    for name, dbg, tma in [("full", 0, 1), ("noMMA", 1, 1), ("noLoad", 2, 1), ("noMMA+noLoad", 3, 1), ("noEpilogueStore", 4, 1), ("noMMA+noLoad+arrive", 6, 1)]:
        L.cy4_set_option(b"debug", dbg); L.cy4_set_option(b"tma_store", tma)
        res[name] = timeit(lambda: co.conv_fwd(x, wp, Cout, k, 1, pad, out=y))
    L.cy4_set_option(b"debug", 0); L.cy4_set_option(b"tma_store", 1)
    for kps in (1, 2, 3, 4):
        L.cy4_set_option(b"kblocks_per_slot", kps)
        res["kps%d" % kps] = timeit(lambda: co.conv_fwd(x, wp, Cout, k, 1, pad, out=y))
    L.cy4_set_option(b"kblocks_per_slot", 4)
    print("%d->%d k%d @%d: " % (Cin, Cout, k, H) + "  ".join("%s %.0f" % kv for kv in res.items()))
