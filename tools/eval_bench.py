"""Evaluation post-processing on a BASELINE-sized batch: post_processing_v2 + true-positive
matching on the device vs the CPU oracle port, same synthetic detections.
    python tools/eval_bench.py [batch] [dup] [clutter]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "complex-yolov4-pytorch_b200"))
import numpy as np
import torch
from cy4 import evalops, synth
from oracle import eval_oracle as eo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dup = int(sys.argv[2]) if len(sys.argv) > 2 else 8
clutter = int(sys.argv[3]) if len(sys.argv) > 3 else 150
tg = synth.make_targets(B, per_image=8, seed=11)
pred = synth.make_detections(B, tg, n_rows=22743, dup=dup, clutter=clutter, seed=3)
tpx = tg.copy(); tpx[:, 2:6] *= 608
pd, td = torch.tensor(pred).cuda(), torch.tensor(tpx).cuda()
host = torch.tensor(pred).pin_memory()


def gpu_resident():
    d = evalops.nms_v2(pd, 0.5, 0.4)
    return d, evalops.match(d, td, 0.5)


def gpu_from_host():
    d = evalops.nms_v2(host.cuda(non_blocking=True), 0.5, 0.4)
    tp = evalops.match(d, td, 0.5)
    return d.as_list("cpu"), tp.cpu()


def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


res = {"batch": B, "rows_per_image": 22743, "candidates_per_image": int((pred[0, :, 6] >= 0.5).sum())}
res["gpu_resident_ms"] = round(timeit(gpu_resident), 3)
res["gpu_from_pinned_host_ms"] = round(timeit(gpu_from_host), 3)
t0 = time.perf_counter()
ref = eo.post_processing_v2(pred, 0.5, 0.4)
st = eo.get_batch_statistics(ref, tpx, 0.5)
res["cpu_oracle_port_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
res["kept_per_image"] = int(np.mean([r.shape[0] for r in ref]))
res["images_per_s_gpu_from_host"] = round(B / (res["gpu_from_pinned_host_ms"] / 1e3), 1)
res["note"] = "python reference (shapely loops): ~1.3 ms per IoU pair => seconds per image at these candidate counts (BASELINE.md section 4)"
print(json.dumps(res))
