#!/bin/bash
# run with: gpurun --gpus 2 -- bash tools/gpu_round5_multi.sh 2
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist exit $?" >> gpurun_out/dist_check_$N.log
tail -5 gpurun_out/dist_check_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_$N.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_$N.log
tail -3 gpurun_out/bench_$N.log
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1.log 2>&1; tail -1 gpurun_out/bench_1.log
