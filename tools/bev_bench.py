"""LiDAR -> BEV rasteriser on a training-sized batch: 32 frames x 120 k points.
    python tools/bev_bench.py [batch] [points]
Algorithmic bytes per frame: 16 B per point read + 3 x 608 x 608 x 4 B written (HBM bound)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import numpy as np
import torch
from cy4 import bevops, synth
from oracle import bev_oracle as bo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 120000
clouds = [synth.make_point_cloud(N, seed=100 + i, ties=False) for i in range(B)]
dev = [torch.tensor(c).cuda() for c in clouds]
pinned = [torch.tensor(c).pin_memory() for c in clouds]


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ms_dev = timeit(lambda: bevops.rasterize(dev))
t0 = time.perf_counter()
for _ in range(5):
    bevops.rasterize(pinned); torch.cuda.synchronize()
ms_host = (time.perf_counter() - t0) / 5 * 1e3
t0 = time.perf_counter()
for c in clouds[:8]:
    bo.make_bv_feature(bo.remove_points(c))
cpu_ms_frame = (time.perf_counter() - t0) / 8 * 1e3
bytes_alg = B * (N * 16 + 3 * 608 * 608 * 4)
peak = 6582.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
print(json.dumps({"batch": B, "points_per_frame": N, "gpu_resident_ms": round(ms_dev, 3), "gpu_from_pinned_host_ms": round(ms_host, 3),
                  "frames_per_s_resident": round(B / (ms_dev / 1e3), 1), "frames_per_s_from_host": round(B / (ms_host / 1e3), 1),
                  "algorithmic_GBps": round(bytes_alg / (ms_dev / 1e3) / 1e9, 1), "hbm_frac": round(bytes_alg / (ms_dev / 1e3) / 1e9 / peak, 3),
                  "cpu_numpy_port_ms_per_frame": round(cpu_ms_frame, 2), "cpu_frames_per_s": round(1e3 / cpu_ms_frame, 1),
                  "note": "gpu_resident includes the torch.cat of the frames and the workspace memset; the CPU port is the numpy lexsort/unique restatement (one core)"}))
