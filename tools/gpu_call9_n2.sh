#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_ddp.py tests/test_gpu_engine.py -q -k "ddp or two_rank or elementwise or step_vs" > $O/c9_pytest.log 2>&1; echo "tests rc $?"; tail -4 $O/c9_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline > $O/c9_bench_n2.json 2> $O/c9_bench_n2.err; echo "n2 rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"cuda_graph": "[^"]*"' $O/c9_bench_n2.json | head -6
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline --cuda-graph 0 > $O/c9_bench_n2_eager.json 2> $O/c9_bench_n2_eager.err; echo "n2 eager rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c9_bench_n2_eager.json | head -4
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-roofline --cuda-graph 0 --ddp-stock > $O/c9_bench_n2_stock.json 2> $O/c9_bench_n2_stock.err; echo "n2 stock rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c9_bench_n2_stock.json | head -4
python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline > $O/c9_bench_n1.json 2> $O/c9_bench_n1.err; echo "n1 rc $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' $O/c9_bench_n1.json | head -4
python tools/kernel_profile.py complex_yolov4 32 > $O/c9_cupti.txt 2>&1; head -10 $O/c9_cupti.txt | tail -7
