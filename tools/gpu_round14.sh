#!/bin/bash
# adjoint evidence: plain timings (uniform + adaptive at two batch sizes), then ONE ncu --set full capture of adjoint_kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/prof_adj.py > gpurun_out/prof_adj_plain.log 2>&1; echo "plain exit $?" >> gpurun_out/prof_adj_plain.log; tail -2 gpurun_out/prof_adj_plain.log
NFC_ADJ_B=16384 timeout 200 python tools/gpu_adj_time.py > gpurun_out/adj_time_16k.log 2>&1; tail -4 gpurun_out/adj_time_16k.log
NFC_ADJ_B=65536 timeout 300 python tools/gpu_adj_time.py > gpurun_out/adj_time_64k.log 2>&1; tail -4 gpurun_out/adj_time_64k.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adjoint_kernel -s 1 -c 1 -o gpurun_out/prof_adjoint python tools/prof_adj.py > gpurun_out/ncu_adj.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_adj.log
