"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: per-kernel totals of the
LAST complete training step in the list (a step starts at stem_im2col_kernel).
    python tools/launch_summary.py profiles/r1_launches.csv > profiles/r1_launches_summary.txt"""
import collections, csv, re, sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")) * (1e-3 if r["Metric Unit"] == "ns" else 1.0)))
starts = [i for i, r in enumerate(rows) if "stem_im2col" in r[1]]
if len(starts) < 2:
    sys.exit("need at least two steps in the list")
a, b = starts[-2], starts[-1]
step = rows[a:b]
agg = collections.OrderedDict()
for _id, name, us in step:
    short = re.sub(r"^void\s+", "", name)
    short = re.split(r"[<(]", short)[0].replace("cy4::", "")
    if short.startswith("at::"):
        short = "at::"
    k = agg.setdefault(short, [0, 0.0])
    k[0] += 1; k[1] += us
tot = sum(v[1] for v in agg.values())
print("# ncu launch list, one training step (complex_yolov4, bs=32), `--metrics gpu__time_duration.sum --clock-control none`")
print("# source: %s, launches %d..%d (the last complete step of the list); cold-cache serialised times: compare SHARES" % (path, step[0][0], step[-1][0]))
print("kernels in step: %d, sum of durations: %.2f ms\n" % (len(step), tot / 1e3))
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-54s n=%4d  %9.3f ms  %4.1f%%" % (name, n, us / 1e3, 100 * us / tot))
