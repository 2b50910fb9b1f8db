#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_ADJ_B=65536 NFC_ADJ_IT=2 timeout 300 python tools/gpu_adj_time.py > gpurun_out/atc_time65k.log 2>&1; echo "exit $?" >> gpurun_out/atc_time65k.log
grep "fwd(record)" gpurun_out/atc_time65k.log
bash tools/gpu_tests.sh tests/test_gpu_adjoint_tc.py tests/test_gpu_adjoint.py
