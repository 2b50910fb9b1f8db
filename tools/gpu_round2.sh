#!/bin/bash
# tests + bench + ncu launch list + ncu full capture of the solve kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -o gpurun_out/prof_solve \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
