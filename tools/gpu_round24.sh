#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint.py tests/test_gpu_forward.py -q --timeout 120 > gpurun_out/pytest_rec.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_rec.log; tail -12 gpurun_out/pytest_rec.log
NFC_ADJ_B=65536 NFC_ADJ_IT=3 timeout 300 python tools/gpu_adj_time.py > gpurun_out/adj_time_64k.log 2>&1; tail -8 gpurun_out/adj_time_64k.log
NFC_TC_RECORD=0 NFC_ADJ_B=65536 NFC_ADJ_IT=3 timeout 300 python tools/gpu_adj_time.py > gpurun_out/adj_time_64k_simt.log 2>&1; tail -4 gpurun_out/adj_time_64k_simt.log
NFC_ADJS_B=9 timeout 120 python tools/gpu_adj_small.py 2>&1 | tail -1
