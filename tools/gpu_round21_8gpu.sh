#!/bin/bash
# BASELINE config 4 on 8 GPUs (1,048,576 columns, forward + loss/gradient + one all-reduce) with the two-tile forward kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/gpu_config4.py > gpurun_out/config4_8gpu.log 2>&1; echo "config4 exit $?" >> gpurun_out/config4_8gpu.log
grep -v "^W\|^\*\*\*" gpurun_out/config4_8gpu.log | tail -10
