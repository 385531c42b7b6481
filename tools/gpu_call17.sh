#!/bin/bash
# call 17: grid cap of the BN passes (one resident wave vs two), new colsum kernel; engine tests
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_engine_bench_config.py -m gpu -q -x > $O/c17_pytest_engine.log 2>&1; echo "pytest engine rc $?"; tail -3 $O/c17_pytest_engine.log
for v in 6 3 2 4; do
  timeout 300 python tools/kernel_profile.py complex_yolov4 32 $O/c17_kp_$v.json wgrad_overlap=0 opt:ew_blocks_per_sm=$v > $O/c17_kernel_profile_bpsm$v.txt 2>&1
  echo "== ew_blocks_per_sm=$v"; grep -E "total kernel|bn_act|colsum" $O/c17_kernel_profile_bpsm$v.txt
done
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c17_bench_$tag.json 2> $O/c17_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c17_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c17_bench_$tag.json)"; }
run b6
run b3 --opt ew_blocks_per_sm=3
run b2 --opt ew_blocks_per_sm=2
run b4 --opt ew_blocks_per_sm=4
run b6_again
