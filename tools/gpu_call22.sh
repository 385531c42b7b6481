#!/bin/bash
# call 22: grouped TMA store + early accumulator release + resident weights in the 1-CTA conv kernel: parity, per-kernel and whole-step A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/c22_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/c22_pytest.log
kp() { tag=$1; shift; timeout 300 python tools/kernel_profile.py complex_yolov4 32 $O/c22_kp_$tag.json wgrad_overlap=0 "$@" > $O/c22_kernel_profile_$tag.txt 2>&1; echo "== $tag"; grep -E "total kernel|conv_pair|conv_tc" $O/c22_kernel_profile_$tag.txt; }
kp default
kp nogroup opt:group_store=0
kp nobres opt:b_resident=0
kp late opt:early_acc_release=0
kp none opt:group_store=0 opt:early_acc_release=0 opt:b_resident=0
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c22_bench_$tag.json 2> $O/c22_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c22_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c22_bench_$tag.json)"; }
run default
run none --opt group_store=0 --opt early_acc_release=0 --opt b_resident=0
run default2
python - <<'PY'
import json
for tag in ("default", "nogroup", "nobres", "late", "none"):
    d = json.load(open("gpurun_out/c22_kp_%s.json" % tag))
    print(tag, "longest conv_tc launches (us):", [round(x) for x in sorted(d["conv_tc_us"], reverse=True)[:14]])
PY
