#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_shapes.py -q -x -k "wgrad or bench_shape" > $O/c10_pytest_conv.log 2>&1; echo "conv pytest rc $?"; tail -5 $O/c10_pytest_conv.log
python -m pytest tests -m gpu -q -k "not conv" > $O/c10_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/c10_pytest.log
timeout 300 python tools/conv_shape_bench.py $O/r2_conv_shape_bench_d.json > $O/c10_shape_bench.txt 2>&1; echo "shape bench rc $?"; grep "^\[32, " $O/c10_shape_bench.txt; tail -1 $O/c10_shape_bench.txt
python bench.py --steps 20 --warmup 5 > $O/c10_bench.json 2> $O/c10_bench.err; echo "bench rc $?"; cut -c1-300 $O/c10_bench.json
python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --opt wgrad_wide32=0 > $O/c10_bench_nowide.json 2> $O/c10_bench_nowide.err; echo "rc $?"; grep -o '"ms_per_step": [0-9.]*' $O/c10_bench_nowide.json | head -1
