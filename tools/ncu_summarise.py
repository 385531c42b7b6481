"""Turn ncu output into the small tracked summaries under profiles/.

  python tools/ncu_summarise.py hot   gpurun_out/r2_hot_kernels.ncu-rep  profiles/r2_ncu_hot_kernels.md
        per-launch table of an `ncu --set full` report (tools/ncu_targets.py): duration, tensor-pipe % of peak,
        DRAM bytes / GB/s, L2 throughput, achieved occupancy, top stall reasons

  python tools/ncu_summarise.py step  gpurun_out/r2_step_metrics.csv     profiles/r2_step_metrics_summary.md
        `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor...`
        over one training step of bench.py: per-kernel-family totals of the last complete step, and
        profiles/r2_conv_fprop_traffic.json = mean DRAM bytes per fprop launch (bench.py's roofline.traffic)
"""
import collections, csv, io, json, os, re, subprocess, sys

NCU = "/usr/local/cuda/bin/ncu"


def short(name):
    s = re.sub(r"^void\s+", "", name)
    s = re.split(r"[<(]", s)[0].replace("cy4::", "")
    return "at::" if s.startswith("at::") else s


def fnum(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def hot(rep, out):
    raw = subprocess.run([NCU, "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, key):
        i = col.get(key)
        return fnum(r[i]) if i is not None and r[i] != "" else None

    want = [
        ("dur_us", "gpu__time_duration.sum", 1e-3),
        ("tensor_pipe_pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("tensor_inst_pct", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", 1),
        ("dram_rd_MB", "dram__bytes_read.sum", None), ("dram_wr_MB", "dram__bytes_write.sum", None),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("issue_active_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
        ("occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("regs", "launch__registers_per_thread", 1),
        ("smem_KB", "launch__shared_mem_per_block_dynamic", None),
    ]
    tensor_keys = [h for h in hdr if "pipe_tensor_cycles_active" in h and "pct_of_peak" in h] or \
                  [h for h in hdr if "pipe_tensor" in h and "pct_of_peak" in h]
    stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")] or \
                 [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled")]
    lines = ["# ncu --set full, shipped kernels on bench-sized operands (tools/ncu_targets.py), B200, --clock-control none",
             "# source report: %s (kept out of git: gpurun_out/ is scratch); per-launch values, cold-ish cache, serialised" % os.path.basename(rep),
             "", "| # | kernel | grid x block | dur us | tensor pipe % | DRAM rd+wr MB | DRAM GB/s | DRAM % | L2 % | SM % | occ % | regs | top stalls |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    js = []
    for n, r in enumerate(data):
        name = short(r[col["Kernel Name"]])
        dur = g(r, "gpu__time_duration.sum")
        du = units[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else "ns"
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1e-3) if dur is not None else None

        def bytes_of(key):
            v = g(r, key)
            if v is None:
                return None
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[key]], 1)

        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        tens = max([g(r, k) or 0.0 for k in tensor_keys] or [0.0])
        stalls = sorted(((g(r, k) or 0.0, re.sub(r".*issue_stalled_|_per_issue_active.ratio|\.ratio|\.pct", "", k)) for k in stall_keys), reverse=True)[:3]
        gbs = (rd + wr) / (dur_us * 1e-6) / 1e9 if rd is not None and wr is not None and dur_us else None
        grid = "%s x %s" % (r[col["Grid Size"]], r[col["Block Size"]]) if "Grid Size" in col else ""
        rec = dict(i=n, kernel=name, grid=grid, dur_us=dur_us, tensor_pipe_pct=tens, dram_read=rd, dram_write=wr, dram_GBps=gbs,
                   dram_pct=g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), l2_pct=g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                   sm_pct=g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), occ_pct=g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                   regs=g(r, "launch__registers_per_thread"), stalls=[(s, round(v, 2)) for v, s in stalls])
        js.append(rec)
        f = lambda v, p=1: "" if v is None else ("%%.%df" % p) % v
        lines.append("| %d | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            n, name, grid, f(dur_us), f(tens), f((rd + wr) / 1e6 if rd is not None and wr is not None else None), f(gbs, 0), f(rec["dram_pct"]),
            f(rec["l2_pct"]), f(rec["sm_pct"]), f(rec["occ_pct"]), f(rec["regs"], 0), ", ".join("%s %.2f" % (s, v) for v, s in stalls)))
    open(out, "w").write("\n".join(lines) + "\n")
    json.dump(js, open(os.path.splitext(out)[0] + ".json", "w"), indent=1)
    print("\n".join(lines))


def step(path, out):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = fnum(r["Metric Value"])
        u = r["Metric Unit"]
        if v is not None:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        k[r["Metric Name"]] = v
    rows = list(per.values())
    starts = [i for i, r in enumerate(rows) if "stem_im2col" in r["name"]]
    if len(starts) < 2:
        sys.exit("need at least two steps in the list")
    stp = rows[starts[-2]:starts[-1]]
    bwd0 = next((i for i, r in enumerate(stp) if "yolo_dense_bwd" in r["name"] or "yolo_targets_bwd" in r["name"]), len(stp))
    agg = collections.OrderedDict()
    for i, r in enumerate(stp):
        nm = short(r["name"])
        if nm == "conv_tc_kernel":
            nm += " (fprop)" if i < bwd0 else " (dgrad)"
        a = agg.setdefault(nm, collections.defaultdict(float))
        a["n"] += 1
        for m, v in r.items():
            if m != "name" and v is not None:
                a[m] += v
    tot = sum(a.get("gpu__time_duration.sum", 0) for a in agg.values())
    txt = ["# ncu metric pass over ONE training step (complex_yolov4, bs=32): last complete step of %s" % os.path.basename(path),
           "# serialised / cold-cache durations: compare SHARES, not absolutes. DRAM bytes are per-step totals.",
           "", "| kernel | launches | sum dur ms | share | DRAM read GB | DRAM write GB | DRAM GB/s over its own time | tensor-pipe % active (mean over launches) |", "|---|---|---|---|---|---|---|---|"]
    tkey = next((m for a in agg.values() for m in a if "pipe_tensor_cycles_active" in m and "pct" in m), None)
    for nm, a in sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
        d = a.get("gpu__time_duration.sum", 0.0)
        rd, wr = a.get("dram__bytes_read.sum", 0.0), a.get("dram__bytes_write.sum", 0.0)
        txt.append("| %s | %d | %.3f | %.1f%% | %.3f | %.3f | %.0f | %s |" % (nm, a["n"], d / 1e3, 100 * d / tot if tot else 0, rd / 1e9, wr / 1e9,
                                                                          (rd + wr) / (d * 1e-6) / 1e9 if d else 0, "" if not tkey else "%.1f" % (a.get(tkey, 0) / a["n"])))
    open(out, "w").write("\n".join(txt) + "\n")
    print("\n".join(txt))
    fp = agg.get("conv_tc_kernel (fprop)")
    if fp:
        json.dump({"source": "%s (ncu metric pass of this round's shipped conv_tc_kernel, %d fprop launches of one bs=32 step)" % (os.path.basename(out), int(fp["n"])),
                   "dram_bytes_read": fp.get("dram__bytes_read.sum"), "dram_bytes_write": fp.get("dram__bytes_write.sum"), "launches": int(fp["n"]),
                   "bytes_per_launch": (fp.get("dram__bytes_read.sum", 0) + fp.get("dram__bytes_write.sum", 0)) / fp["n"],
                   "ncu_time_ms_cold": fp.get("gpu__time_duration.sum", 0) / 1e3},
                  open(os.path.join(os.path.dirname(out), "r2_conv_fprop_traffic.json"), "w"), indent=1)


def fprop(path, out):
    """`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active... -k regex:conv_tc_kernel
    --launch-skip 720 --launch-count 110` over bench.py: exactly the 110 fprop launches of the timed step (a step launches 240
    conv_tc kernels, the forward ones first).  Writes the per-launch table and r2_conv_fprop_traffic.json (roofline.traffic)."""
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = per.setdefault(int(r["ID"]), {"grid": r["Grid Size"], "block": r["Block Size"]})
        v = fnum(r["Metric Value"])
        if v is not None:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
        k[r["Metric Name"]] = v
    rows = list(per.values())
    if len(rows) > 110:
        print("note: %d launches in the file, keeping the first 110 (the forward ones)" % len(rows))
        rows = rows[:110]
    rd = sum(r.get("dram__bytes_read.sum", 0) for r in rows); wr = sum(r.get("dram__bytes_write.sum", 0) for r in rows)
    dur = sum(r.get("gpu__time_duration.sum", 0) for r in rows)
    tk = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tw = sum(r.get(tk, 0) * r.get("gpu__time_duration.sum", 0) for r in rows) / dur if dur else 0
    txt = ["# ncu metric pass over the %d conv fprop launches (conv_pair_kernel / conv_tc_kernel) of one bs=32 training-mode forward (%s)" % (len(rows), os.path.basename(path)),
           "# serialised, cold-cache durations; DRAM bytes are what roofline.traffic reports (mean per launch)",
           "", "launches %d | sum dur %.3f ms | DRAM read %.3f GB | DRAM write %.3f GB | mean bytes/launch %.1f MB | algorithmic 14.4 GB/fwd (SURVEY 8d) -> ratio %.2f | tensor-pipe active (time-weighted) %.1f %%"
           % (len(rows), dur / 1e3, rd / 1e9, wr / 1e9, (rd + wr) / max(len(rows), 1) / 1e6, (rd + wr) / 14.4e9, tw),
           "", "| # | dur us | DRAM rd MB | DRAM wr MB | tensor pipe % |", "|---|---|---|---|---|"]
    for i, r in enumerate(rows):
        txt.append("| %d | %.1f | %.1f | %.1f | %.1f |" % (i, r.get("gpu__time_duration.sum", 0), r.get("dram__bytes_read.sum", 0) / 1e6,
                                                        r.get("dram__bytes_write.sum", 0) / 1e6, r.get(tk, 0)))
    open(out, "w").write("\n".join(txt) + "\n")
    print("\n".join(txt[:5]))
    json.dump({"source": "%s (ncu metric pass of the shipped conv_pair_kernel / conv_tc_kernel, the %d fprop launches of one bs=32 step, tools/ncu_fprop_step.py)" % (os.path.basename(out), len(rows)),
               "dram_bytes_read": rd, "dram_bytes_write": wr, "launches": len(rows), "bytes_per_launch": (rd + wr) / max(len(rows), 1),
               "ncu_time_ms_cold": dur / 1e3, "tensor_pipe_active_pct_time_weighted": tw},
              open(os.path.join(os.path.dirname(out), "r2_conv_fprop_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    {"hot": hot, "step": step, "fprop": fprop}[sys.argv[1]](sys.argv[2], sys.argv[3])
