#!/bin/bash
# Round-2 2-GPU call: NCCL gradient-equality test, weak-scaling bench with the engine's overlapped exchange and with stock DDP.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/c4_gpus.txt
python -m pytest tests/test_gpu_ddp.py -q -s > $O/r2_ddp_2gpu_test.log 2>&1; echo "ddp test rc $?"; tail -6 $O/r2_ddp_2gpu_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline > $O/c4_bench_n2.json 2> $O/c4_bench_n2.err; echo "n2 rc $?"; cut -c1-700 $O/c4_bench_n2.json
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-roofline --ddp-stock > $O/c4_bench_n2_stock.json 2> $O/c4_bench_n2_stock.err; echo "n2 stock rc $?"; cut -c1-700 $O/c4_bench_n2_stock.json
python bench.py --steps 10 --warmup 3 --no-roofline --no-cpu-baseline > $O/c4_bench_n1.json 2> $O/c4_bench_n1.err; echo "n1 rc $?"; cut -c1-500 $O/c4_bench_n1.json
