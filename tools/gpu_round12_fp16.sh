#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export NFC_TC=1 NFC_TC_SPLIT=fp16
timeout 60 python tools/gpu_tc_check.py 2>&1 | tail -8
timeout 60 python tools/gpu_sanity.py > gpurun_out/sanity_fp16.log 2>&1; echo "sanity exit $?" >> gpurun_out/sanity_fp16.log; tail -7 gpurun_out/sanity_fp16.log
NFC_TC_DEBUG=1 timeout 60 python tools/prof_tc_uniform.py 2>&1 | tail -5
timeout 240 python -m pytest tests/test_gpu_forward.py -q -x --timeout 60 > gpurun_out/pytest_fp16.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fp16.log; tail -5 gpurun_out/pytest_fp16.log
