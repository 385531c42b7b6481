#!/bin/bash
# Round-2 first GPU call: parity at the bench configuration, same-box cuDNN comparator, bench, ncu of the shipped kernels.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/c1_smi.txt
nproc >> $O/c1_smi.txt
python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/c1_pytest.log
python tools/cudnn_comparator.py 32 fp16 > $O/r2_cudnn_comparator_fp16.json 2> $O/c1_cudnn_fp16.err; echo "cudnn fp16 rc $?"; cat $O/r2_cudnn_comparator_fp16.json
python tools/cudnn_comparator.py 32 tf32 > $O/r2_cudnn_comparator_tf32.json 2> $O/c1_cudnn_tf32.err; echo "cudnn tf32 rc $?"; cat $O/r2_cudnn_comparator_tf32.json
python bench.py --steps 10 --warmup 3 > $O/c1_bench.json 2> $O/c1_bench.err; echo "bench rc $?"; cut -c1-600 $O/c1_bench.json
python bench.py --impl reference --steps 4 --warmup 1 > $O/c1_bench_ref.json 2> $O/c1_bench_ref.err; echo "ref rc $?"; cut -c1-300 $O/c1_bench_ref.json
python tools/kernel_profile.py complex_yolov4 32 > $O/c1_cupti.txt 2>&1; echo "cupti rc $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/r2_hot_kernels -f python tools/ncu_targets.py > $O/c1_ncu_hot.log 2>&1; echo "ncu hot rc $?"; tail -3 $O/c1_ncu_hot.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/r2_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-roofline --no-cpu-baseline > $O/c1_ncu_step.log 2>&1; echo "ncu step rc $?"
# the two never-run drafts, last (a protocol error in a draft kernel shows up as a hang: bounded by timeout)
CY4_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_conv.py -q -x -k "persistent" > $O/c1_exp_wgrad2.log 2>&1; echo "exp wgrad2 rc $?"; tail -5 $O/c1_exp_wgrad2.log
CY4_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_conv.py -q -x -k "pair" > $O/c1_exp_pair.log 2>&1; echo "exp pair rc $?"; tail -5 $O/c1_exp_pair.log
nvidia-smi --query-gpu=name,clocks.sm --format=csv
