#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_forward.py -x -q --timeout 120 -k "two_tile" > gpurun_out/pytest_tc2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc2.log; tail -15 gpurun_out/pytest_tc2.log
