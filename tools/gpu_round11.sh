#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_tc.log; tail -2 gpurun_out/bench_tc.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_tc.log 2>&1
echo "ncu launches exit $?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_tc_kernel -s 3 -c 1 -o gpurun_out/prof_solve_tc_bench \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1
echo "ncu full exit $?"
