#!/bin/bash
# Round-2 second GPU call: parity of the new fused paths, A/B of the kernel options, bench, inference bench, bounded ncu passes.
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q -k "not bench_shape" > $O/c2_pytest.log 2>&1; echo "pytest rc $?"; tail -8 $O/c2_pytest.log
timeout 600 python tools/ab_options.py > $O/c2_ab.txt 2>&1; echo "ab rc $?"; cat $O/c2_ab.txt | grep "ms/step"
python bench.py --steps 10 --warmup 3 > $O/c2_bench.json 2> $O/c2_bench.err; echo "bench rc $?"; cut -c1-400 $O/c2_bench.json
python tools/infer_bench.py > $O/c2_infer.json 2> $O/c2_infer.err; echo "infer rc $?"; cat $O/c2_infer.json
python tools/kernel_profile.py complex_yolov4 32 > $O/c2_cupti.txt 2>&1; echo "cupti rc $?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --metrics $M --clock-control none -k regex:conv_tc_kernel --launch-skip 720 --launch-count 110 --csv --log-file $O/r2_fprop_metrics.csv python bench.py --steps 1 --warmup 3 --no-roofline --no-cpu-baseline > $O/c2_ncu_fprop.log 2>&1; echo "ncu fprop rc $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2600 --launch-count 1400 --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-roofline --no-cpu-baseline > $O/c2_ncu_list.log 2>&1; echo "ncu list rc $?"
nvidia-smi --query-gpu=name,clocks.sm --format=csv
