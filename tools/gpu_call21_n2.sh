#!/bin/bash
# call 21 (2 GPUs): the gradient exchange captured into the backward graph.
# HISTORICAL: ran at commit 79e7e6c (model.graph_allreduce existed there); the in-graph exchange hung and was reverted,
# so `--model-opt graph_allreduce=0` and the graph-replay DDP test no longer exist at HEAD.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddp.py -q -m gpu -s > $O/c21_pytest_ddp.log 2>&1; echo "ddp tests rc $?"; tail -6 $O/c21_pytest_ddp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
Q="--gpus 2 --steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
timeout 400 $TR bench.py $Q > $O/c21_bench_n2_in_graph.json 2> $O/c21_bench_n2_in_graph.err; echo "n2 in-graph rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"cuda_graph": "[^"]*"' $O/c21_bench_n2_in_graph.json | head -6
timeout 400 $TR bench.py $Q --model-opt graph_allreduce=0 > $O/c21_bench_n2_after.json 2> $O/c21_bench_n2_after.err; echo "n2 after-replay rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c21_bench_n2_after.json | head -4
timeout 400 $TR bench.py $Q > $O/c21_bench_n2_in_graph2.json 2> /dev/null; grep -o '"ms_per_step": [0-9.]*' $O/c21_bench_n2_in_graph2.json | head -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline > $O/c21_bench_n1.json 2> $O/c21_bench_n1.err; echo "n1 rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c21_bench_n1.json | head -4
tail -3 $O/c21_bench_n2_in_graph.err
