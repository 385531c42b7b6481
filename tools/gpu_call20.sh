#!/bin/bash
# call 20: separable argmax max-pool, interleaved parity classes of the large stride-2 input gradients
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/c20_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/c20_pytest.log
for v in 1 0; do
  timeout 300 python tools/kernel_profile.py complex_yolov4 32 $O/c20_kp_il$v.json wgrad_overlap=0 opt:dgrad_interleave=$v > $O/c20_kernel_profile_il$v.txt 2>&1
  echo "== dgrad_interleave=$v"; grep -E "total kernel|conv_pair|conv_tc|maxpool" $O/c20_kernel_profile_il$v.txt
done
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c20_bench_$tag.json 2> $O/c20_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c20_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c20_bench_$tag.json)"; }
run il1
run il0 --opt dgrad_interleave=0
run il1b
python - <<'PY'
import json
for v in (1, 0):
    d = json.load(open("gpurun_out/c20_kp_il%d.json" % v))
    print("interleave", v, "longest conv_tc launches (us):", sorted(d["conv_tc_us"], reverse=True)[:8])
PY
