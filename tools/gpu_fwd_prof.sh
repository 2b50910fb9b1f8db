#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:solve_tc_kernel -s 1 -c 1 -o gpurun_out/fwd_tc1 python tools/prof_tc.py > gpurun_out/fwd_ncu.log 2>&1; echo "exit $?" >> gpurun_out/fwd_ncu.log
tail -3 gpurun_out/fwd_ncu.log
timeout 300 python -m pytest tests/test_gpu_forward.py -q -m gpu -k set_option 2>&1 | tail -3
