#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_adjoint_debug.py > gpurun_out/adjoint_debug.log 2>&1; echo "adjoint debug exit $?" >> gpurun_out/adjoint_debug.log
cat gpurun_out/adjoint_debug.log | tail -70
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
