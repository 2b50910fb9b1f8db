#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/prof_tc2.py > gpurun_out/prof_tc2_plain.log 2>&1; echo "plain exit $?" >> gpurun_out/prof_tc2_plain.log; tail -2 gpurun_out/prof_tc2_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_tc2_kernel -s 2 -c 1 -o gpurun_out/prof_solve_tc2 python tools/prof_tc2.py > gpurun_out/ncu_tc2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_tc2.log
