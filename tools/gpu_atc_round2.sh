#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_ADJ_B=65536 NFC_ADJ_IT=2 timeout 300 python tools/gpu_adj_time.py > gpurun_out/atc_time65k.log 2>&1; echo "exit $?" >> gpurun_out/atc_time65k.log
tail -8 gpurun_out/atc_time65k.log
NFC_ADJ_B=16384 NFC_ADJ_IT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:adjoint_tc_kernel -c 1 -o gpurun_out/atc_16k python tools/gpu_adj_time.py > gpurun_out/atc_ncu.log 2>&1; echo "exit $?" >> gpurun_out/atc_ncu.log
tail -3 gpurun_out/atc_ncu.log
