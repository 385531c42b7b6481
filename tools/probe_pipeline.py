"""Cycle probes of the conv kernel's TMA-producer and MMA-issuer loops (block 0 only).
Build first with  CY4_LIB_NAME=libcy4_probe.so CY4_EXTRA_NVCC_FLAGS=-DCY4_PROBE python complex-yolov4-pytorch_b200/csrc/build.py
(the product build has no probes and ignores the "debug" option)."""
import ctypes, os, sys
os.environ["CY4_LIB_NAME"] = "libcy4_probe.so"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import _lib, convops as co
L = _lib.lib()
raw = ctypes.CDLL(os.path.join(ROOT, "complex-yolov4-pytorch_b200", "csrc", "libcy4_probe.so"))
B = 32
buf = (ctypes.c_ulonglong * 16)()
names_m = ["wait_tmem_empty", "wait_full", "fence", "mma_issue", "commit", "syncwarp"]
names_p = ["wait_empty", "issue_loads"]
for (Cin, Cout, k, H) in [(32, 64, 3, 304), (64, 64, 3, 152), (128, 128, 3, 76), (256, 512, 3, 38), (512, 1024, 3, 19)]:
    pad = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k, device="cuda") / 30
    wp = co.pack_fprop(w)
    y = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
    for name, dbg in [("full", 0), ("noMMA", 1), ("noLoad", 2), ("noMMA+noLoad", 3), ("noMMA+noLoad+arrive", 6)]:
        L.cy4_set_option(b"debug", dbg)
        co.conv_fwd(x, wp, Cout, k, 1, pad, out=y); torch.cuda.synchronize()
        raw.cy4_probe_read(buf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); co.conv_fwd(x, wp, Cout, k, 1, pad, out=y); b.record(); torch.cuda.synchronize()
        raw.cy4_probe_read(buf)
        v = list(buf)
        nm, npd = max(v[6], 1), max(v[10], 1)
        print("%d->%d k%d @%d %-20s %6.0f us | MMA thread per slot (n=%d): %s | producer per slot (n=%d): %s" % (
            Cin, Cout, k, H, name, a.elapsed_time(b) * 1e3, v[6],
            " ".join("%s %.0f" % (names_m[i], v[i] / nm) for i in range(6)), v[10],
            " ".join("%s %.0f" % (names_p[i], v[8 + i] / npd) for i in range(2))))
    L.cy4_set_option(b"debug", 0)
