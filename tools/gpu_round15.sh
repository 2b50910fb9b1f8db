#!/bin/bash
# adjoint: batched tape prefetch + history-keyed longest-first backward queue
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint.py -x -q --timeout 120 > gpurun_out/pytest_adj.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_adj.log; tail -4 gpurun_out/pytest_adj.log
timeout 120 python tools/prof_adj.py > gpurun_out/prof_adj_plain.log 2>&1; echo "plain exit $?" >> gpurun_out/prof_adj_plain.log; tail -2 gpurun_out/prof_adj_plain.log
NFC_ADJ_B=16384 NFC_ADJ_IT=3 timeout 200 python tools/gpu_adj_time.py > gpurun_out/adj_time_16k.log 2>&1; tail -6 gpurun_out/adj_time_16k.log
NFC_ADJ_B=65536 NFC_ADJ_IT=3 timeout 300 python tools/gpu_adj_time.py > gpurun_out/adj_time_64k.log 2>&1; tail -6 gpurun_out/adj_time_64k.log
