#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_TC2_B=${NFC_TC2_B:-1,129,300,4096,65536} timeout 120 python tools/gpu_tc2_check.py > gpurun_out/tc2_check.log 2>&1; echo "tc2 exit $?" >> gpurun_out/tc2_check.log; tail -14 gpurun_out/tc2_check.log
