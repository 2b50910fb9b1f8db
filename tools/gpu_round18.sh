#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_TC2_SCALES=1e-5 NFC_TC2_B=98304,131072,196608,524288,1048576 timeout 200 python tools/gpu_tc2_check.py > gpurun_out/tc2_sizes.log 2>&1; echo "exit $?" >> gpurun_out/tc2_sizes.log; tail -7 gpurun_out/tc2_sizes.log
