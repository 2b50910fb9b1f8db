#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
NFC_TC2_SCALES=1e-5 NFC_TC2_B=65536,262144 timeout 160 python tools/gpu_tc2_check.py 2>&1 | tail -2 | cut -c1-330
