#!/bin/bash
# call 16: programmatic dependent launch -- parity (whole GPU suite with CY4_PDL=1), then whole-step A/B
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "programmatic" -s > $O/c16_pytest_pdl.log 2>&1; echo "pdl test rc $?"; tail -6 $O/c16_pytest_pdl.log
CY4_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x > $O/c16_pytest_all_pdl.log 2>&1; echo "pytest (CY4_PDL=1) rc $?"; tail -4 $O/c16_pytest_all_pdl.log
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c16_bench_$tag.json 2> $O/c16_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c16_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c16_bench_$tag.json) $(grep -o '"cuda_graph": "[^"]*"' $O/c16_bench_$tag.json)"; }
run plain
run pdl --opt pdl=1
run pdl_ov0 --opt pdl=1 --model-opt wgrad_overlap=0
run plain_ov0 --model-opt wgrad_overlap=0
run pdl_eager --opt pdl=1 --cuda-graph 0
run plain_eager --cuda-graph 0
run pdl2 --opt pdl=1
run plain2
tail -3 $O/c16_bench_pdl.err
