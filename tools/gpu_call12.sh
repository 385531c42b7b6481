#!/bin/bash
# call 12: overlapped weight-gradient stream -- parity tests, then whole-step A/B of the stream modes / ring sizes
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "overlapped" -s > $O/c12_pytest_overlap.log 2>&1; echo "overlap tests rc $?"; tail -12 $O/c12_pytest_overlap.log
timeout 900 python -m pytest tests -m gpu -q > $O/c12_pytest.log 2>&1; echo "pytest rc $?"; tail -4 $O/c12_pytest.log
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
for mo in "wgrad_overlap=2" "wgrad_overlap=0" "wgrad_overlap=1" "wgrad_overlap=2 --model-opt dy_ring=2" "wgrad_overlap=2 --model-opt dy_ring=8" "wgrad_overlap=2 --model-opt bn_shifted_stats=0" "wgrad_overlap=0 --model-opt bn_shifted_stats=0"; do
  tag=$(echo "$mo" | tr -c 'a-z0-9=_\n' '_')
  timeout 300 python bench.py $Q --model-opt $mo > $O/c12_bench_$tag.json 2> $O/c12_bench_$tag.err
  echo "$mo: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c12_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c12_bench_$tag.json)"
done
timeout 300 python bench.py $Q --cuda-graph 0 --model-opt wgrad_overlap=2 > $O/c12_bench_eager_ov2.json 2>/dev/null; echo "eager ov2: $(grep -o '"ms_per_step": [0-9.]*' $O/c12_bench_eager_ov2.json | head -1)"
timeout 300 python bench.py $Q --cuda-graph 0 --model-opt wgrad_overlap=0 > $O/c12_bench_eager_ov0.json 2>/dev/null; echo "eager ov0: $(grep -o '"ms_per_step": [0-9.]*' $O/c12_bench_eager_ov0.json | head -1)"
