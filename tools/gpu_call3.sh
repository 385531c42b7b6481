#!/bin/bash
# Round-2 third GPU call: re-validate the fixed paths, per-shape kernel A/B, step A/B of the BN-backward fusion modes, bench.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q -k "not bench_shape" > $O/c3_pytest.log 2>&1; echo "pytest rc $?"; tail -8 $O/c3_pytest.log
python tools/conv_shape_bench.py $O/r2_conv_shape_bench.json > $O/c3_shape_bench.txt 2>&1; echo "shape bench rc $?"; tail -3 $O/c3_shape_bench.txt
timeout 600 python tools/ab_options.py > $O/c3_ab.txt 2>&1; echo "ab rc $?"; grep "ms/step" $O/c3_ab.txt
python bench.py --steps 10 --warmup 3 > $O/c3_bench.json 2> $O/c3_bench.err; echo "bench rc $?"; cut -c1-300 $O/c3_bench.json
python tools/infer_bench.py > $O/c3_infer.json 2> $O/c3_infer.err; echo "infer rc $?"; cat $O/c3_infer.json
python tools/kernel_profile.py complex_yolov4 32 > $O/c3_cupti.txt 2>&1; echo "cupti rc $?"
nvidia-smi --query-gpu=name,clocks.sm --format=csv
