#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export NFC_TC=1
timeout 120 python tools/prof_tc.py > gpurun_out/prof_tc_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_tc_kernel -s 1 -c 1 -o gpurun_out/prof_solve_tc python tools/prof_tc.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_tc.log
