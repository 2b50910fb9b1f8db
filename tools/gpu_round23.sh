#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_adjoint.py -q --timeout 120 > gpurun_out/pytest_adj.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_adj.log; tail -5 gpurun_out/pytest_adj.log
NFC_ADJS_B=9,296,2960,4440,5920 timeout 120 python tools/gpu_adj_small.py > gpurun_out/adj_small.log 2>&1; echo "exit $?" >> gpurun_out/adj_small.log; tail -6 gpurun_out/adj_small.log
