"""Per-layer-shape A/B of the tensor-core kernels at the bench batch (complex_yolov4.cfg @608, B=32):
fprop (+BN statistics), dgrad and wgrad of every distinct conv shape, timed alone with CUDA events (5 repetitions after 2
warm-ups), for the default kernels and for the optional ones (cy4_set_option conv_pair=1, wgrad_variant=2).  Prints a table
and writes JSON; the per-shape winners drive the host-side kernel selection (csrc/conv_api.cu).
    python tools/conv_shape_bench.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import torch
from cy4 import _lib, convops as co

sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_conv_shapes import SHAPES          # (Cin, Cout, k, stride, Hin)
COUNT = {(512, 1024, 3, 1, 19): 5, (256, 512, 3, 1, 38): 5, (128, 128, 3, 1, 76): 8, (256, 256, 3, 1, 38): 8, (128, 256, 3, 1, 76): 3,
         (512, 512, 3, 1, 19): 4, (64, 64, 3, 1, 152): 2, (512, 256, 1, 1, 38): 9, (1024, 512, 1, 1, 19): 8, (64, 64, 1, 1, 304): 3,
         (256, 128, 1, 1, 76): 6, (128, 128, 1, 1, 76): 9, (256, 256, 1, 1, 38): 9, (512, 512, 1, 1, 19): 5, (128, 64, 1, 1, 152): 2,
         (64, 64, 1, 1, 152): 3}
B = 32
L = _lib.lib()
_lib.require_device()


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3          # us


rows = []
for (Cin, Cout, k, s, H) in SHAPES:
    if Cout % 32:
        continue
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // s + 1
    torch.manual_seed(0)
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
    wp, wd = co.pack_fprop(w), co.pack_dgrad(w)
    cpad = (Cout + 63) // 64 * 64
    dy = torch.randn(B, Ho, Ho, cpad, device="cuda").half()
    y = torch.empty(B, Ho, Ho, Cout, device="cuda", dtype=torch.float16)
    dx = torch.empty(B, H, H, Cin, device="cuda", dtype=torch.float16)
    s1 = torch.zeros(Cout, device="cuda"); s2 = torch.zeros(Cout, device="cuda")
    acc = torch.zeros((Cout + 31) // 32 * 32, k * k, Cin, device="cuda")
    fl = 2.0 * B * Ho * Ho * Cout * k * k * Cin
    rec = {"shape": [Cin, Cout, k, s, H], "count": COUNT.get((Cin, Cout, k, s, H), 1), "gflop": fl / 1e9}
    for tag, opts in (("base", {b"conv_pair": 0}), ("pair", {b"conv_pair": 1})):
        for o, v in opts.items():
            L.cy4_set_option(o, v)
        try:
            rec["fprop_" + tag] = timed(lambda: co.conv_fwd(x, wp, Cout, k, s, pad, out=y, stats=(s1, s2)))
            rec["dgrad_" + tag] = timed(lambda: co.conv_dgrad(dy[..., :Cout], wd, H, H, Cin, k, s, pad, out=dx))
        finally:
            L.cy4_set_option(b"conv_pair", 1)
    for tag, var in (("base", 0), ("persistent", 1)):      # ("persistent" column = wgrad_pair: the cta_group::2 weight-gradient kernel)
        L.cy4_set_option(b"wgrad_pair", var)
        try:
            rec["wgrad_" + tag] = timed(lambda: co.conv_wgrad(x, dy, Cin, Cout, k, s, pad, acc=acc))
        finally:
            L.cy4_set_option(b"wgrad_pair", 1)
    rows.append(rec)
    print("%-26s x%d  fprop %7.1f / pair %7.1f us (%6.0f TF/s) | dgrad %7.1f / pair %7.1f | wgrad %7.1f / pair %7.1f" % (
        rec["shape"], rec["count"], rec["fprop_base"], rec["fprop_pair"], fl / min(rec["fprop_base"], rec["fprop_pair"]) / 1e6,
        rec["dgrad_base"], rec["dgrad_pair"], rec["wgrad_base"], rec["wgrad_persistent"]), flush=True)
tot = lambda key: sum(r[key] * r["count"] for r in rows) / 1e3
print("step totals (ms): fprop base %.3f pair %.3f best %.3f | dgrad base %.3f pair %.3f best %.3f | wgrad base %.3f pair %.3f best %.3f" % (
    tot("fprop_base"), tot("fprop_pair"), sum(min(r["fprop_base"], r["fprop_pair"]) * r["count"] for r in rows) / 1e3,
    tot("dgrad_base"), tot("dgrad_pair"), sum(min(r["dgrad_base"], r["dgrad_pair"]) * r["count"] for r in rows) / 1e3,
    tot("wgrad_base"), tot("wgrad_persistent"), sum(min(r["wgrad_base"], r["wgrad_persistent"]) * r["count"] for r in rows) / 1e3))
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
