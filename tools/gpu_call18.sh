#!/bin/bash
# call 18: whole GPU suite at HEAD, BN grid caps, full bench line, ncu of the shipped kernels (hot-kernel set + fprop metric pass)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/c18_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/c18_pytest.log
for v in 2 1; do
  timeout 300 python tools/kernel_profile.py complex_yolov4 32 $O/c18_kp_bwd$v.json wgrad_overlap=0 opt:ew_bwd_blocks_per_sm=$v > $O/c18_kernel_profile_bwd$v.txt 2>&1
  echo "== ew_bwd_blocks_per_sm=$v"; grep -E "total kernel|bn_act|colsum|maxpool" $O/c18_kernel_profile_bwd$v.txt
done
Q="--steps 20 --warmup 5 --no-roofline --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 python bench.py $Q "$@" > $O/c18_bench_$tag.json 2> $O/c18_bench_$tag.err; echo "$tag: rc $? $(grep -o '"ms_per_step": [0-9.]*' $O/c18_bench_$tag.json | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $O/c18_bench_$tag.json)"; }
run default
run bwd1 --opt ew_bwd_blocks_per_sm=1
run fwd6bwd6 --opt ew_fwd_blocks_per_sm=6 --opt ew_bwd_blocks_per_sm=6
timeout 600 python bench.py --steps 20 --warmup 5 > $O/c18_bench_full.json 2> $O/c18_bench_full.err; echo "full bench rc $?"; cut -c1-400 $O/c18_bench_full.json
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/r2b_hot_kernels -f python tools/ncu_targets.py > $O/c18_ncu_hot.log 2>&1; echo "ncu hot rc $?"; tail -2 $O/c18_ncu_hot.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 500 ncu --metrics $M --clock-control none --profile-from-start off -k regex:"conv_(tc|pair)_kernel" --csv --log-file $O/r2b_fprop_metrics.csv python tools/ncu_fprop_step.py > $O/c18_ncu_fprop.log 2>&1; echo "ncu fprop rc $?"; tail -2 $O/c18_ncu_fprop.log
ls -la $O/r2b_hot_kernels.ncu-rep $O/r2b_fprop_metrics.csv
