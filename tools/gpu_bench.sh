#!/bin/bash
# bench.py as the driver runs it.  Usage: tools/gpu_bench.sh N [extra args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}; shift
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 "$@" > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "exit $?" >> gpurun_out/bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?" >> gpurun_out/bench_n$N.err
fi
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_n$N.json") if x.startswith("{")]
d=json.loads(l[-1])
print("value %.4g e2e %.4g ms/step %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]))
print("cold", d.get("cold_call"))
print("adjoint", json.dumps(d.get("adjoint"))[:1500])
print("config4", json.dumps(d.get("config4"))[:2500])
print("other chains", json.dumps(d.get("other_chains")))
print("cpu", d.get("cpu_baseline"), d.get("cpu_affinity"))
PY
