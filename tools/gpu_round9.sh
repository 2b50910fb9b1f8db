#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_tc.log; tail -2 gpurun_out/bench_tc.log | cut -c1-1800
NFC_SCHED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
NFC_TC=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
