#!/bin/bash
# One GPU visit: sanity -> pytest -m gpu -> smoke -> bench (+ optional ncu passes given as $1 = "ncu")
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/gpu_sanity.py > gpurun_out/sanity.log 2>&1; echo "sanity exit $?" >> gpurun_out/sanity.log
tail -12 gpurun_out/sanity.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
