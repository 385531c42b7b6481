#!/usr/bin/env python
"""bench.py -- BEV-images/s of one Complex-YOLOv4 training step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

A step = forward + rotated-GIoU YOLO loss + backward + Adam update of complex_yolov4.cfg on a
synthetic 608x608x3 BEV batch (32 images per GPU, 5 rotated targets per image), i.e. what
train.py's loop body does (reference src/train.py:203-221).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "complex-yolov4-pytorch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

GFLOP_FWD_PER_IMG = {"complex_yolov4": 127.225, "complex_yolov4_tiny": 14.506}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_evt = index, [], threading.Event()

    def run(self):
        while not self.stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_evt.wait(0.2)

    def summary(self):
        self.stop_evt.set()
        self.join(timeout=3)
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max([float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(self.rows)}


def make_optimizer(model, fused=False):
    """Reference create_optimizer (src/utils/train_utils.py:21-50), optimizer_type='adam', lr 1e-3, wd 5e-4.
    fused=True selects PyTorch's single-launch multi-tensor Adam (same update rule, one kernel per parameter group
    instead of ~60 foreach launches); the CPU arms keep the default."""
    pg0, pg1, pg2 = [], [], []
    for k, v in model.named_parameters():
        if ".bias" in k:
            pg2.append(v)
        elif "conv" in k and ".weight" in k:
            pg1.append(v)
        else:
            pg0.append(v)
    opt = torch.optim.Adam(pg0, lr=1e-3, **({"fused": True} if fused else {}))
    opt.add_param_group({"params": pg1, "weight_decay": 5e-4})
    opt.add_param_group({"params": pg2})
    return opt


def run_ours(args):
    import torch.distributed as dist
    from cy4 import _lib, netdefs, synth
    from cy4.darknet import Darknet
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 or args.force_ddp:
        if world == 1:      # diagnostic: the DDP wrapper's host overhead without a second GPU
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
            os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    L = _lib.lib()
    for o in args.opt:                       # kernel experiments: cy4_set_option(NAME, INT)
        name, val = o.split("=")
        _lib.check(L.cy4_set_option(name.encode(), int(val)), "cy4_set_option(%s)" % o)
    B = args.batch
    torch.manual_seed(0)
    net = Darknet(netdefs.cfg_path(args.cfg), use_giou_loss=True).to(dev).train()
    net.use_cuda_graph = bool(args.cuda_graph)
    for o in args.model_opt:
        name, val = o.split("=")
        if not hasattr(net, name):
            raise SystemExit("--model-opt: the model has no attribute %r" % name)
        setattr(net, name, int(val))
    model = net
    if world > 1 or args.force_ddp:
        # unchanged PyTorch DDP (parameter broadcast, bucket views); the gradient all-reduce itself is issued by the engine in
        # groups while backward is still running (models/model_utils.py overlap_gradient_exchange), unless --ddp-stock
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=128)
        if not args.ddp_stock:
            from models.model_utils import overlap_gradient_exchange
            overlap_gradient_exchange(model)
    opt = make_optimizer(net, fused=not args.adam_foreach)
    x_host = synth.make_bev(B, seed=1234 + rank).pin_memory()
    tg_host = torch.tensor(synth.make_targets(B, per_image=5, seed=4321 + rank)).pin_memory()
    x = x_host.to(dev)
    tg = tg_host.to(dev)

    def step(xd, td):
        loss, _ = model(xd, td)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph_note = "off"
    try:
        for _ in range(args.warmup):
            loss = step(x, tg)
        sync()
        if net.use_cuda_graph:
            gs = net._engine.plan.graph_state
            graph_note = "fwd+bwd launch sequences replayed as CUDA graphs" if (gs and gs.get("bwd") is not None) else "requested, not captured"
    except Exception as e:      # noqa: BLE001 -- a failed capture must not cost the bench line: fall back to eager launches
        if not net.use_cuda_graph:
            raise
        sys.stderr.write("bench: CUDA-graph capture failed (%r); falling back to eager launches\n" % (e,))
        graph_note = "capture failed, eager launches"
        net.use_cuda_graph = False
        net._engine = None
        torch.cuda.synchronize()
        opt.zero_grad(set_to_none=True)
        for _ in range(args.warmup):
            loss = step(x, tg)
        sync()
    assert torch.isfinite(loss).all(), "loss is not finite after warm-up: %r" % loss
    # host time to ENQUEUE one step into an empty launch queue (no back-pressure): if this approaches ms_per_step the step is
    # bound by the Python / ctypes launch path, not by the GPU
    t_h = time.perf_counter()
    loss = step(x, tg)
    host_enqueue_ms = (time.perf_counter() - t_h) * 1e3
    sync()
    # ---- timed region: K steps, inputs resident in HBM (the batch, weights and activations are far
    # larger than the 126 MB L2, so no explicit flush is needed between iterations)
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    L.cy4_kernel_launches(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        loss = step(x, tg)
    e1.record()
    sync()
    launches = int(L.cy4_kernel_launches(1))
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clk = clocks.summary() if clocks else None
    # ---- end to end through the public API with HOST buffers: pinned H2D of the batch every step,
    # D2H of the loss (and of the detections the reference API returns) inside the timed region
    net.sync_outputs = False     # default training behaviour: detections go to pinned host memory asynchronously,
                                 # complete at the loss.item() synchronisation below
    copy_stream = torch.cuda.Stream(device=dev)

    def h2d():
        """pinned host -> device copy of one batch on the copy stream (overlaps the previous step's kernels)"""
        with torch.cuda.stream(copy_stream):
            xd_ = x_host.to(dev, non_blocking=True)
            td_ = tg_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xd_, td_, ev

    # Every step: H2D of its own batch from pinned memory (issued one step ahead on a copy stream), D2H of the step's loss
    # and detections into pinned memory.  The host reads step i's loss while step i+1 is already enqueued (a training loop
    # that logs the previous step's loss), so the ~850 kernel launches of a step are issued ahead of the GPU instead of
    # after a per-step host synchronisation; the last step's results are read inside the timed region.
    loss_host = torch.zeros(2, 1, pin_memory=True)
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        nxt = h2d()
        lv, out_ = None, None
        for i in range(n):
            xd, td, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            xd.record_stream(torch.cuda.current_stream()); td.record_stream(torch.cuda.current_stream())
            loss_, out_ = model(xd, td)
            loss_.backward()
            if i + 1 < n:
                nxt = h2d()                  # every step copies its own batch from the host, one step ahead
            opt.step()
            opt.zero_grad(set_to_none=True)
            loss_host[i & 1].copy_(loss_.detach().reshape(1), non_blocking=True)     # device -> host read of the step's result
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                lv = float(loss_host[(i - 1) & 1])
        loss_ev[(n - 1) & 1].synchronize()
        lv = float(loss_host[(n - 1) & 1])
        return lv, out_

    e2e_loop(2)                              # untimed: first-use allocations of the per-step input tensors / pinned staging
    sync()
    t0 = time.perf_counter()
    lval, out = e2e_loop(args.steps)
    sync()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d_bytes = x_host.numel() * 4 + tg_host.numel() * 4
    d2h_bytes = out.numel() * 4 + 4

    result = None
    if rank == 0:
        pk = peaks()
        ms_per_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total / 1e3)
        roof, giou, nxt, fwd_ms = None, None, None, None

        def extras():
            """Rank 0 only, after the timed regions: per-launch roofline pass, forward-only timing, rotated-GIoU microbench,
            the "next" rows.  Guarded as a whole below: a failure here must never cost the headline line (nor leave the other
            ranks waiting at the final barrier for longer than it takes to report it)."""
            roof, giou, nxt, fwd_ms = None, None, None, None
            # ---- roofline pass: CUDA-event timing of every tensor-core conv launch for 2 steps
            plan = net._engine.plan
            net.use_cuda_graph = False          # events around individual launches need eager launches
            net.engine_allreduce = False        # rank 0 runs this alone: the engine must not issue its gradient all-reduces here
            plan.prof = []
            for _ in range(2):                  # on the bare module: no collective may run on rank 0 alone
                l_, _o = net(x, tg)
                l_.backward()
                opt.step()
                opt.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            prof, plan.prof = plan.prof, None
            agg = {}
            for name, a, b, fl, shape in prof:
                t = a.elapsed_time(b)
                k = agg.setdefault(name, [0.0, 0.0, 0])
                k[0] += t; k[1] += fl; k[2] += 1
            # per layer shape (Cin, Cout, k, stride, Ho): where the tensor-core time goes, for the next optimisation round
            try:
                by_shape = {}
                for name, a, b, fl, shape in prof:
                    k2 = by_shape.setdefault((name, tuple(shape)), [0.0, 0.0, 0])
                    k2[0] += a.elapsed_time(b); k2[1] += fl; k2[2] += 1
                top_shapes = [{"kernel": n_, "cin_cout_k_s_ho": list(sh), "launches_per_step": v[2] // 2, "ms_per_step": round(v[0] / 2, 3),
                               "tflops": round(v[1] / (v[0] * 1e9), 1) if v[0] else 0.0}
                              for (n_, sh), v in sorted(by_shape.items(), key=lambda kv: -kv[1][0])]
            except Exception as e:      # noqa: BLE001 -- diagnostics must never cost the headline line
                top_shapes = [{"error": repr(e)[:200]}]
            conv_ms = sum(v[0] for v in agg.values()) / 2
            tc = {n: {"ms_per_step": v[0] / 2, "tflops": v[1] / (v[0] * 1e9) if v[0] else 0.0, "launches_per_step": v[2] // 2} for n, v in agg.items()}
            # fprop family: cy4_conv_fwd (heads, eval) + cy4_conv_fwd_stats (training-mode BN layers, shifted statistics)
            fwd = [1e-9, 0.0, 0]
            for n_ in ("cy4_conv_fwd", "cy4_conv_fwd_stats"):
                if n_ in agg:
                    fwd = [fwd[0] + agg[n_][0], fwd[1] + agg[n_][1], fwd[2] + agg[n_][2]]
            achieved = fwd[1] / (fwd[0] * 1e9)
            # dram read+write bytes per fprop launch: from this round's `ncu --set full` capture of the shipped kernel over the
            # same workload (profiles/r2_conv_fprop_traffic.json, written by tools/ncu_summarise.py); null when no capture of
            # the current kernel exists -- never a stale constant
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "r2_conv_fprop_traffic.json")
            if os.path.exists(tpath) and args.cfg == "complex_yolov4" and B == 32:
                tj = json.load(open(tpath))
                traffic, traffic_src = tj.get("bytes_per_launch"), tj.get("source")
            # ---- the whole forward (110 convs + BN/activation passes + routes + loss head), CUDA events: the north_star's
            # ">= 40 % of tensor peak on the forward" is about THIS, not about the fprop launches alone
            fwd_ms = None
            try:
                with torch.no_grad():
                    for _ in range(2):
                        net(x, tg)
                    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize(); fa.record()
                    for _ in range(5):
                        net(x, tg)
                    fb.record(); torch.cuda.synchronize()
                    fwd_ms = fa.elapsed_time(fb) / 5
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("forward-only timing failed: %r\n" % (e,))
            roof = {"bound": "tensor", "kernel": "conv fprop launches (conv_pair_kernel, conv_tc_kernel for the shapes the pair kernel does not take)", "achieved": round(achieved, 1), "peak": pk["tf_sustained"],
                    "unit": "TFLOP/s", "frac": round(achieved / pk["tf_sustained"], 4), "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": pk["src"] + ", sustained bf16 cuBLAS figure (kernel timed inside a long step)",
                    "flops_per_step": fwd[1] / 2, "avg_launch_ms": fwd[0] / max(fwd[2], 1), "by_kernel": tc,
                    "top_layer_shapes": top_shapes, "conv_share_of_step": round(conv_ms / ms_per_step, 3)}
            # ---- rotated-GIoU microbench (BASELINE config 4): 100k pairs (latency) and 10^7 pairs (bandwidth)
            from cy4 import geometry as cg
            giou = {}
            for n in (100_000, 10_000_000):
                p_, t_ = synth.make_pairs(n, seed=7)
                pd, td_ = torch.tensor(p_, device=dev), torch.tensor(t_, device=dev)
                for _ in range(3):
                    cg.rgiou_pairs(pd, td_, True)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                reps = 20 if n <= 100_000 else 5
                for _ in range(reps):
                    cg.rgiou_pairs(pd, td_, True)
                b.record(); torch.cuda.synchronize()
                t = a.elapsed_time(b) / reps
                if n == 100_000:        # the CPU legs on the same pairs, for the north_star's ">= 100x CPU" bar
                    from oracle import geometry as og
                    t0 = time.perf_counter(); og.rgiou_pairs(p_, t_, True); cpu_s = time.perf_counter() - t0
                    giou["cpu_port_pairs_per_s_1thread"] = round(n / cpu_s, 0)
                    giou["reference_python"] = reference_giou_rate(p_, t_)
                giou[str(n)] = {"us": round(t * 1e3, 2), "pairs_per_s": round(n / (t / 1e3), 0), "GBps": round(n * 56 / (t / 1e3) / 1e9, 1),
                                "hbm_frac": round(n * 56 / (t / 1e3) / 1e9 / pk["hbm_gbs"], 4)}
            # ---- the "next" rows built so far, outside the timed region; a failure here must
            # not cost the headline line, so each is guarded and reports its error instead
            nxt = {}
            try:
                from cy4 import evalops
                tgs = synth.make_targets(B, per_image=8, seed=11)
                dets_h = synth.make_detections(B, tgs, n_rows=22743, dup=8, clutter=150, seed=3)
                pd_ = torch.tensor(dets_h, device=dev)
                tpx = tgs.copy(); tpx[:, 2:6] *= 608
                td2 = torch.tensor(tpx, device=dev)
                for _ in range(2):
                    dd = evalops.nms_v2(pd_, 0.5, 0.4); evalops.match(dd, td2, 0.5)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(5):
                    dd = evalops.nms_v2(pd_, 0.5, 0.4); evalops.match(dd, td2, 0.5)
                b.record(); torch.cuda.synchronize()
                nxt["f1_rotated_nms_and_matching"] = {"ms_per_batch": round(a.elapsed_time(b) / 5, 3), "batch": B, "rows_per_image": 22743,
                                                      "candidates_per_image": int((dets_h[0, :, 6] >= 0.5).sum())}
            except Exception as e:      # noqa: BLE001
                nxt["f1_rotated_nms_and_matching"] = {"error": repr(e)[:300]}
            try:
                from cy4 import bevops
                clouds = [torch.tensor(synth.make_point_cloud(120000, seed=100 + i, ties=False), device=dev) for i in range(B)]
                for _ in range(2):
                    bevops.rasterize(clouds)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(5):
                    bevops.rasterize(clouds)
                b.record(); torch.cuda.synchronize()
                t_ = a.elapsed_time(b) / 5
                nxt["f3_lidar_to_bev"] = {"ms_per_batch": round(t_, 3), "batch": B, "points_per_frame": 120000,
                                          "algorithmic_GBps": round(B * (120000 * 16 + 3 * 608 * 608 * 4) / (t_ / 1e3) / 1e9, 1)}
            except Exception as e:      # noqa: BLE001
                nxt["f3_lidar_to_bev"] = {"error": repr(e)[:300]}
            return roof, giou, nxt, fwd_ms

        if not args.no_roofline:
            try:
                roof, giou, nxt, fwd_ms = extras()
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("bench: the roofline / microbench pass failed (%r); the headline line is unaffected\n" % (e,))
        result = {
            "metric": "BEV-images/sec training step (bs=32, 608x608)", "value": round(value, 2), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16 (fp32 accumulate, fp32 master weights / loss head)",
            "data": "synthetic",
            "config": {"workload": workload_name(args.cfg, B),
                       "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "no flush needed: per-step working set (>20 GB) exceeds the 126 MB L2",
                       "optimizer": "torch.optim.Adam(%s), reference parameter groups (train_utils.py:21-50)" % ("foreach" if args.adam_foreach else "fused=True"),
                       "cuda_graph": graph_note, **({"options": args.opt} if args.opt else {}),
                       **({"model_options": args.model_opt} if args.model_opt else {})},
            "e2e": {"value": round(world * B * args.steps / float(e2e_s.item()), 2), "unit": "img/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "last_loss": lval},
            "gpu_launches": launches, "gpu_launches_per_step": launches // args.steps,
            "host_enqueue_ms_per_step": round(host_enqueue_ms, 2),
            "clocks": clk, "roofline": roof,
            "forward_only": None if not fwd_ms else {
                "ms": round(fwd_ms, 3), "tflops": round(GFLOP_FWD_PER_IMG.get(args.cfg, 0) * B * 1e-3 / (fwd_ms / 1e3), 1),
                "frac_of_sustained_peak": round(GFLOP_FWD_PER_IMG.get(args.cfg, 0) * B * 1e-3 / (fwd_ms / 1e3) / pk["tf_sustained"], 4),
                "what": "whole training-mode forward incl. BN/activation passes and the loss head, conv FLOPs only in the numerator"},
            "library_baseline": library_baseline(),
            "step_tflops": round(3 * GFLOP_FWD_PER_IMG.get(args.cfg, 0) * B * 1e-3 / (ms_per_step / 1e3), 1),
            "rgiou_microbench": giou,
            "next_rows": nxt,
        }
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_step_baseline(args.cfg, budget_s=25.0)
    if world > 1 or args.force_ddp:
        if world > 1:
            dist.barrier()
        dist.destroy_process_group()
    return result


def library_baseline():
    """Same-box PyTorch/cuDNN step of the conv stack (tools/cudnn_comparator.py), measured this round on a
    B200 of this pool and committed under profiles/: context for the headline, not a bench leg."""
    out = {}
    for mode in ("fp16", "tf32"):
        path = os.path.join(ROOT, "profiles", "r2_cudnn_comparator_%s.json" % mode)
        if os.path.exists(path):
            try:
                out[mode] = json.load(open(path))
            except Exception:       # noqa: BLE001
                pass
    return out or None


def reference_giou_rate(pred, tgt, n=1000):
    """iou_pred_vs_target_boxes(GIoU=True) of the UNMODIFIED reference on the first n pairs, host CPU."""
    mods = reference_modules()
    if mods is None:
        return {"unavailable": "no oracle/_ref on this box"}
    try:
        import contextlib
        torch.set_num_threads(1)       # the loop is Python-bound; one thread is the faster setting
        p = torch.tensor(pred[:n]); t = torch.tensor(tgt[:n])
        with contextlib.redirect_stdout(sys.stderr):
            mods["iou"].iou_pred_vs_target_boxes(p[:8], t[:8], GIoU=True)
            t0 = time.perf_counter()
            mods["iou"].iou_pred_vs_target_boxes(p, t, GIoU=True)
            dt = time.perf_counter() - t0
        return {"pairs_per_s": round(n / dt, 1), "pairs_timed": n, "threads": 1,
                "what": "utils.iou_rotated_boxes_utils.iou_pred_vs_target_boxes(GIoU=True), unmodified, torch CPU"}
    except Exception as e:          # noqa: BLE001
        return {"error": repr(e)[:200]}


def usable_cores():
    """Host threads this process may really use: CPU affinity capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0]); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, min(n, 64))          # beyond ~64 threads the oneDNN convs of this size stop scaling


def workload_name(cfg, batch):
    """config.workload, shared by both arms (the reference arm times a bounded sample of the same workload)."""
    return ("%s.cfg training step (fwd + rotated-GIoU loss + bwd + Adam), bs=%d/GPU, 608x608x3 synthetic BEV, "
            "5 targets/img, GIoU on" % (cfg, batch))


def reference_modules():
    """The UNMODIFIED reference modules (oracle/_ref copy on the GPU box, /root/reference/src in the build container)
    with the import stand-ins of oracle/ref_stubs.py, or None when no reference tree travelled with the snapshot."""
    try:
        from oracle import reference_loader as rl, ref_stubs
        if not rl.available():
            return None
        kinds = ref_stubs.install()
        mods = rl.load()
        mods["kinds"] = kinds
        mods["src"] = rl.REF_SRC
        return mods
    except Exception as e:      # noqa: BLE001 -- fall back to the port, and say so
        sys.stderr.write("bench: reference modules unavailable (%r); timing the oracle port instead\n" % (e,))
        return None


class CpuStep:
    """One training step of the hot path on the host cores: the reference's own Darknet(cfg).forward + backward + Adam
    (kind "reference") when oracle/_ref is present, else the oracle port (kind "port")."""

    def __init__(self, cfg, batch):
        from cy4 import netdefs, synth
        self.batch = batch
        self.cores = usable_cores()
        torch.set_num_threads(self.cores)
        self.x = synth.make_bev(batch)
        self.tg = torch.tensor(synth.make_targets(batch, per_image=5, seed=4321))
        mods = reference_modules()
        if mods is not None:
            self.kind = "reference"
            cfgfile = os.path.join(mods["src"], "config", "cfg", cfg + ".cfg")
            torch.manual_seed(0)
            self.model = mods["darknet"].Darknet(cfgfile=cfgfile, use_giou_loss=True).train()
            self.opt = make_optimizer(self.model)
            self.what = ("UNMODIFIED reference Darknet(%s.cfg).forward + loss.backward + Adam on the host CPU (shapely: %s)"
                         % (cfg, mods["kinds"].get("shapely")))
        else:
            from cy4.darknet import Darknet
            from oracle import darknet_oracle as do
            self.kind = "port"
            path = netdefs.cfg_path(cfg)
            torch.manual_seed(0)
            sd = Darknet(path, True).state_dict()
            self.params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
            self.blocks = do.parse_cfg(path)
            self.do = do
            self.opt = torch.optim.Adam([p for p in self.params.values() if p.requires_grad], lr=1e-3)
            self.what = "fp32 oracle port of the same step (torch CPU ops + C rotated-box geometry)"

    def step(self):
        t0 = time.perf_counter()
        if self.kind == "reference":
            loss, _ = self.model(self.x, self.tg)
        else:
            loss, _, _ = self.do.forward(self.blocks, self.params, self.x, self.tg, True, True, update_running=True)
        loss.backward()
        self.opt.step(); self.opt.zero_grad()
        return time.perf_counter() - t0


def cpu_step_baseline(cfg, budget_s=25.0, batch=2):
    """The same training step on the host cores, on a bounded sample: `batch` images per step, ~budget_s seconds."""
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):        # the reference prints while it builds the network
        runner = CpuStep(cfg, batch)
        times = []
        t_start = time.perf_counter()
        while True:
            times.append(runner.step())
            if (len(times) >= 2 and (time.perf_counter() - t_start > budget_s or len(times) >= 8)) or time.perf_counter() - t_start > 4 * budget_s:
                break
    best = min(times[1:]) if len(times) > 1 else times[0]
    return {"value": round(batch / best, 3), "unit": "img/s", "cores": runner.cores, "kind": runner.kind,
            "sample": "%d steps of bs=%d of the bench workload (%s); best step after 1 warm-up" % (len(times), batch, runner.what)}


def run_reference(args):
    """`--impl reference`: W untimed + K timed steps of the reference's own CPU implementation of the step, each step a
    bounded sample (bs=2) of the bench workload; `steps`, `warmup` and `ms_per_step` describe exactly what ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if os.environ.get("CY4_BENCH_TEST_TINY"):       # CPU unit test of the harness: tiny net
        args.cfg = "complex_yolov4_tiny"
    sample_bs = 2
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):        # the reference prints while it builds the network
        runner = CpuStep(args.cfg, sample_bs)
        for _ in range(args.warmup):
            runner.step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            runner.step()
        total = time.perf_counter() - t0
    ms_per_step = total / max(args.steps, 1) * 1e3
    value = round(sample_bs * args.steps / total, 3)
    base = {"value": value, "unit": "img/s", "cores": runner.cores, "kind": runner.kind,
            "sample": "%d timed steps (+%d warm-up) of bs=%d of the bench workload: %s" % (args.steps, args.warmup, sample_bs, runner.what)}
    return {"impl": "reference", "metric": "BEV-images/sec training step (bs=32, 608x608)", "value": value, "unit": "img/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": workload_name(args.cfg, args.batch), "global_batch": args.batch * world, "parallelism": "host cpu",
                       "sample": "each step is a bounded sample of that workload: bs=%d on the host cores (img/s is what is compared)" % sample_bs},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--cfg", default="complex_yolov4")
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--cuda-graph", dest="cuda_graph", type=int, default=1,
                    help="1 (default): after two eager steps the fwd / bwd launch sequences (~850 kernels) are replayed as CUDA graphs; 0: eager launches")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=INT",
                    help="cy4_set_option(NAME, INT) before the run (kernel experiments, e.g. conv_cluster=2); recorded in config")
    ap.add_argument("--model-opt", dest="model_opt", action="append", default=[], metavar="NAME=INT",
                    help="engine experiments: setattr(model, NAME, INT) (e.g. wgrad_overlap=0, dy_ring=2, bn_shifted_stats=0); recorded in config")
    ap.add_argument("--adam-foreach", dest="adam_foreach", action="store_true", help="torch.optim.Adam's default foreach path instead of fused=True")
    ap.add_argument("--force-ddp", dest="force_ddp", action="store_true", help="wrap the model in DDP even with one rank (host-overhead diagnostic)")
    ap.add_argument("--ddp-stock", dest="ddp_stock", action="store_true",
                    help="N>1: let stock DDP do the (un-overlapped) bucketed all-reduce instead of the engine's overlapped exchange")
    ap.add_argument("--no-roofline", dest="no_roofline", action="store_true", help="skip the per-launch roofline pass (quick A/B runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    res = run_reference(args) if args.impl == "reference" else run_ours(args)
    if res is not None:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
