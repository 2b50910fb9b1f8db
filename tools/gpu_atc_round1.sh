#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/gpu_atc_check.py 9 200 > gpurun_out/atc_check.log 2>&1; echo "exit $?" >> gpurun_out/atc_check.log
tail -12 gpurun_out/atc_check.log
NFC_ADJ_B=16384 NFC_ADJ_IT=2 timeout 300 python tools/gpu_adj_time.py > gpurun_out/atc_time16k.log 2>&1; echo "exit $?" >> gpurun_out/atc_time16k.log
tail -8 gpurun_out/atc_time16k.log
