#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/comm2.log
timeout 600 python -m pytest tests/test_gpu_comm.py -q -m gpu --timeout 300 >> gpurun_out/comm2.log 2>&1; echo "pytest exit $?" >> gpurun_out/comm2.log
tail -15 gpurun_out/comm2.log
