#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale_8.log 2>&1; echo "exit $?" >> gpurun_out/scale_8.log
grep "^{" gpurun_out/scale_8.log | cut -c1-400; tail -1 gpurun_out/scale_8.log
