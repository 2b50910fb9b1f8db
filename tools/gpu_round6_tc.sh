#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export NFC_TC=1
timeout 120 python tools/gpu_tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc check exit $?" >> gpurun_out/tc_check.log; tail -4 gpurun_out/tc_check.log
timeout 60 python tools/gpu_sanity.py > gpurun_out/sanity_tc.log 2>&1; echo "sanity exit $?" >> gpurun_out/sanity_tc.log; tail -9 gpurun_out/sanity_tc.log
timeout 240 python -m pytest tests/test_gpu_forward.py -q -x --timeout 60 > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log; tail -15 gpurun_out/pytest_tc.log
