#!/bin/bash
# Re-entry verification: pytest -m gpu -> smoke -> bench -> RKC2 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
timeout 200 python tools/gpu_rkc_bench.py > gpurun_out/rkc_bench.log 2>&1; echo "rkc exit $?" >> gpurun_out/rkc_bench.log
tail -9 gpurun_out/rkc_bench.log
