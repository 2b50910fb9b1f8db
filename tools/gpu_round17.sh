#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_TC_DEBUG=1 NFC_TC_TILES=2 timeout 60 python tools/prof_tc_uniform.py > gpurun_out/tc2_dbg.log 2>&1; tail -3 gpurun_out/tc2_dbg.log
NFC_TC2_B=1,127,300,4096,65536,262144 timeout 120 python tools/gpu_tc2_check.py > gpurun_out/tc2_check.log 2>&1; echo "tc2 exit $?" >> gpurun_out/tc2_check.log; tail -13 gpurun_out/tc2_check.log
