#!/bin/bash
# call 19 (2 GPUs): the multi-GPU row with the overlapped backward / graph replay at HEAD
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -m gpu > $O/c19_pytest_ddp.log 2>&1; echo "ddp tests rc $?"; tail -4 $O/c19_pytest_ddp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline > $O/c19_bench_n2.json 2> $O/c19_bench_n2.err; echo "n2 rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"cuda_graph": "[^"]*"' $O/c19_bench_n2.json | head -6
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline --cuda-graph 0 > $O/c19_bench_n2_eager.json 2> $O/c19_bench_n2_eager.err; echo "n2 eager rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c19_bench_n2_eager.json | head -4
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline --cuda-graph 0 --model-opt wgrad_overlap=0 > $O/c19_bench_n2_eager_ov0.json 2> $O/c19_bench_n2_eager_ov0.err; echo "n2 eager ov0 rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c19_bench_n2_eager_ov0.json | head -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline > $O/c19_bench_n1.json 2> $O/c19_bench_n1.err; echo "n1 rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c19_bench_n1.json | head -4
timeout 400 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/c19_bench_ref_n2.json 2> $O/c19_bench_ref_n2.err; echo "ref arm n2 rc $?"; cut -c1-300 $O/c19_bench_ref_n2.json
tail -3 $O/c19_bench_n2.err
