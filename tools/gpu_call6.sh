#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 180 python -m pytest tests/test_gpu_conv.py -q -x -k "wgrad_pair" > $O/c6_wgrad_pair.log 2>&1; echo "wgrad pair test rc $?"; tail -4 $O/c6_wgrad_pair.log
python -m pytest tests -m gpu -q -k "not bench_shape and not wgrad_pair" > $O/c6_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/c6_pytest.log
timeout 300 python tools/conv_shape_bench.py $O/r2_conv_shape_bench_b.json > $O/c6_shape_bench.txt 2>&1; echo "shape bench rc $?"; tail -2 $O/c6_shape_bench.txt
timeout 600 python tools/ab_options.py > $O/c6_ab.txt 2>&1; echo "ab rc $?"; grep "ms/step" $O/c6_ab.txt
python bench.py --steps 20 --warmup 5 > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc $?"; cut -c1-300 $O/c6_bench.json
python tools/kernel_profile.py complex_yolov4 32 > $O/c6_cupti.txt 2>&1; echo "cupti rc $?"
