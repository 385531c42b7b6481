#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_fused.py tests/test_gpu_conv_shapes.py -q -x > $O/c8_pytest_conv.log 2>&1; echo "conv pytest rc $?"; tail -5 $O/c8_pytest_conv.log
python -m pytest tests -m gpu -q -k "not conv" > $O/c8_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/c8_pytest.log
timeout 300 python tools/conv_shape_bench.py $O/r2_conv_shape_bench_c.json > $O/c8_shape_bench.txt 2>&1; echo "shape bench rc $?"; tail -2 $O/c8_shape_bench.txt
python bench.py --steps 20 --warmup 5 > $O/c8_bench.json 2> $O/c8_bench.err; echo "bench rc $?"; cut -c1-300 $O/c8_bench.json
python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --opt accum_tma=0 > $O/c8_bench_noaccumtma.json 2> $O/c8_bench_noaccumtma.err; echo "rc $?"; grep -o '"ms_per_step": [0-9.]*' $O/c8_bench_noaccumtma.json | head -1
python tools/kernel_profile.py complex_yolov4 32 > $O/c8_cupti.txt 2>&1; echo "cupti rc $?"; head -10 $O/c8_cupti.txt | tail -7
