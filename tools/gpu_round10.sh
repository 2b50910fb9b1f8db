#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_adjoint.py -q -k forward_solve --timeout 60 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_tc.log 2>&1
echo "ncu launches exit $?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_tc_kernel -s 3 -c 1 -o gpurun_out/prof_solve_tc_bench \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full_tc.log | cut -c1-200
