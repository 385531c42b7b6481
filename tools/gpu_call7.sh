#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q > $O/c7_pytest.log 2>&1; echo "pytest rc $?"; tail -6 $O/c7_pytest.log
python __graft_entry__.py smoke > $O/c7_smoke.log 2>&1; echo "smoke rc $?"; tail -4 $O/c7_smoke.log
python bench.py --steps 20 --warmup 5 > $O/c7_bench.json 2> $O/c7_bench.err; echo "bench rc $?"; cut -c1-300 $O/c7_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/c7_bench_ref.json 2> $O/c7_bench_ref.err; echo "ref rc $?"; cut -c1-200 $O/c7_bench_ref.json
python tools/kernel_profile.py complex_yolov4 32 > $O/c7_cupti.txt 2>&1; echo "cupti rc $?"
grep -n "trajectory\|max relative deviation\|first dozen\|bs=32 complex" $O/c7_pytest.log | head
