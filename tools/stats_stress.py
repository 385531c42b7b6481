"""Stress test of the fused BatchNorm statistics of the conv epilogues: repeat the same launch many times and compare the
per-channel sums between repetitions (they may differ only by fp32 atomic summation order, ~1e-6 relative).
    python tools/stats_stress.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import _lib, convops as co

L = _lib.lib()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = 32
for (Cin, Cout, k, s, H) in [(512, 1024, 3, 2, 38), (256, 512, 3, 1, 38), (128, 128, 3, 1, 76), (512, 256, 1, 1, 38), (64, 64, 3, 1, 152)]:
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // s + 1
    torch.manual_seed(1)
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
    wp = co.pack_fprop(w)
    y = torch.empty(B, Ho, Ho, Cout, device="cuda", dtype=torch.float16)
    for pair in (1, 0):
        L.cy4_set_option(b"conv_pair", pair)
        res = []
        for _ in range(reps):
            s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
            co.conv_fwd(x, wp, Cout, k, s, pad, out=y, stats=(s1, s2))
            res.append(torch.stack([s1, s2]))
        torch.cuda.synchronize()
        R = torch.stack(res)                       # [reps, 2, Cout]
        med = R.median(0).values
        rel = ((R - med).abs() / (med.abs() + 1e-3 * med.abs().max()))
        bad = (rel > 1e-4)
        nbad = int(bad.any(2).any(1).sum())
        worst = rel.max().item()
        msg = ""
        if nbad:
            r, t, c = [int(v) for v in bad.nonzero()[0]]
            msg = " first bad: rep %d %s channel %d value %.6f median %.6f (diff/median %.5f; one warp of 32 rows = %.5f of the %d rows)" % (
                r, "sum" if t == 0 else "sumsq", c, R[r, t, c].item(), med[t, c].item(), (R[r, t, c] - med[t, c]).item() / med[t, c].item(), 32.0 / (B * Ho * Ho), B * Ho * Ho)
        print("%s pair=%d: %d of %d repetitions deviate > 1e-4 from the median, worst rel %.2e%s" % ((Cin, Cout, k, s, H), pair, nbad, reps, worst, msg), flush=True)
    L.cy4_set_option(b"conv_pair", 1)
