#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python tools/stats_stress.py 300 > $O/c11_stats_stress.txt 2>&1; echo "stress rc $?"; cat $O/c11_stats_stress.txt
python -m pytest tests -m gpu -q > $O/c11_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/c11_pytest.log
python bench.py --steps 20 --warmup 5 > $O/c11_bench.json 2> $O/c11_bench.err; echo "bench rc $?"; cut -c1-300 $O/c11_bench.json
python bench.py --steps 20 --warmup 5 --no-roofline --no-cpu-baseline --opt conv_pair=0 > $O/c11_bench_nopair.json 2>/dev/null; grep -o '"ms_per_step": [0-9.]*' $O/c11_bench_nopair.json | head -1
