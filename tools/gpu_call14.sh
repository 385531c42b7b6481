#!/bin/bash
# call 14: (a) packed fp32x2 Mish vs the scalar forms (side build libcy4_scalar.so), per kernel and whole step;
#          (b) priority of the weight-gradient stream
mkdir -p gpurun_out
O=gpurun_out
for lib in libcy4.so libcy4_scalar.so; do
  CY4_LIB_NAME=$lib timeout 300 python tools/kernel_profile.py complex_yolov4 32 $O/c14_kp_$lib.json wgrad_overlap=0 > $O/c14_kernel_profile_$lib.txt 2>&1
  echo "== $lib"; grep -E "total kernel|bn_act|conv_pair|conv_tc" $O/c14_kernel_profile_$lib.txt
done
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c14_bench_$tag.json 2> $O/c14_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c14_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c14_bench_$tag.json)"; }
run default
CY4_LIB_NAME=libcy4_scalar.so run scalar_act
run ov2_prio --model-opt wgrad_priority=-1
run ov1_prio --model-opt wgrad_overlap=1 --model-opt wgrad_priority=-1
run ov0 --model-opt wgrad_overlap=0
run ov2_prio_eager --model-opt wgrad_priority=-1 --cuda-graph 0
run ov2_eager --cuda-graph 0
run default2
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x > $O/c14_pytest_engine.log 2>&1; echo "pytest engine rc $?"; tail -3 $O/c14_pytest_engine.log
