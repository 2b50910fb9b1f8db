#!/bin/bash
# scaling run on an N-GPU box: dist check at N, bench at 1,2,4,..,N
NMAX=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_$NMAX.log 2>&1; echo "dist exit $?" >> gpurun_out/dist_check_$NMAX.log
tail -3 gpurun_out/dist_check_$NMAX.log
for N in 1 2 4 8; do
  if [ $N -le $NMAX ]; then
    if [ $N -eq 1 ]; then timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.log 2>&1
    else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_$N.log 2>&1; fi
    tail -1 gpurun_out/scale_$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=%d value %.4e e2e %.4e ms/step %.2f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']))"
  fi
done
