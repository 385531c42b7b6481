#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q -k "not bench_shape" > $O/c5_pytest.log 2>&1; echo "pytest rc $?"; tail -8 $O/c5_pytest.log
python bench.py --steps 20 --warmup 5 > $O/c5_bench.json 2> $O/c5_bench.err; echo "bench rc $?"; cut -c1-350 $O/c5_bench.json
python bench.py --steps 20 --warmup 5 --cuda-graph 0 --no-roofline --no-cpu-baseline > $O/c5_bench_eager.json 2> $O/c5_bench_eager.err; echo "eager rc $?"; cut -c1-350 $O/c5_bench_eager.json
python bench.py --steps 20 --warmup 5 --cuda-graph 0 --no-roofline --no-cpu-baseline --force-ddp > $O/c5_bench_eager_ddp1.json 2> $O/c5_bench_eager_ddp1.err; echo "eager ddp1 rc $?"; grep -o '"ms_per_step": [0-9.]*\|"host_enqueue_ms_per_step": [0-9.]*' $O/c5_bench_eager_ddp1.json
python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --force-ddp > $O/c5_bench_graph_ddp1.json 2> $O/c5_bench_graph_ddp1.err; echo "graph ddp1 rc $?"; grep -o '"ms_per_step": [0-9.]*\|"host_enqueue_ms_per_step": [0-9.]*' $O/c5_bench_graph_ddp1.json
python tools/kernel_profile.py complex_yolov4 32 > $O/c5_cupti.txt 2>&1; echo "cupti rc $?"
python tools/infer_bench.py > $O/c5_infer.json 2> $O/c5_infer.err; echo "infer rc $?"; cat $O/c5_infer.json
grep -o '"ms_per_step": [0-9.]*\|"host_enqueue_ms_per_step": [0-9.]*\|"cuda_graph": "[^"]*"' $O/c5_bench.json $O/c5_bench_eager.json
