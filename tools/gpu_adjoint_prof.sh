#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_atc_check.py 200 > gpurun_out/atc_check.log 2>&1; echo "exit $?" >> gpurun_out/atc_check.log
tail -5 gpurun_out/atc_check.log
NFC_ADJ_B=16384 NFC_ADJ_IT=2 timeout 300 python tools/gpu_adj_time.py > gpurun_out/atc_time16k.log 2>&1; echo "exit $?" >> gpurun_out/atc_time16k.log
grep "fwd(record)" gpurun_out/atc_time16k.log
NFC_ADJ_B=65536 NFC_ADJ_IT=2 timeout 300 python tools/gpu_adj_time.py > gpurun_out/atc_time65k.log 2>&1; echo "exit $?" >> gpurun_out/atc_time65k.log
grep "fwd(record)" gpurun_out/atc_time65k.log
if [ "$1" = "ncu" ]; then
NFC_ADJ_B=16384 NFC_ADJ_IT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:adjoint_tc_kernel -c 1 -o gpurun_out/atc_16k python tools/gpu_adj_time.py > gpurun_out/atc_ncu.log 2>&1; echo "exit $?" >> gpurun_out/atc_ncu.log
tail -2 gpurun_out/atc_ncu.log
fi
