#!/bin/bash
# call 13: slab statistics + packed fp32x2 BN passes + carve-out preference: parity, per-kernel times, whole-step A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c13_pytest.log 2>&1; echo "pytest rc $?"; tail -15 $O/c13_pytest.log
timeout 300 python tools/kernel_profile.py complex_yolov4 32 $O/c13_kernel_profile.json wgrad_overlap=0 > $O/c13_kernel_profile_cupti.txt 2>&1; head -14 $O/c13_kernel_profile_cupti.txt
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c13_bench_$tag.json 2> $O/c13_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c13_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c13_bench_$tag.json)"; }
run default
run accstats --opt slab_stats=0
run nocarve --opt ew_carveout=0
run ov0 --model-opt wgrad_overlap=0
run ov0_nocarve --model-opt wgrad_overlap=0 --opt ew_carveout=0
run ov1 --model-opt wgrad_overlap=1
run default2
