#!/bin/bash
# Evidence for profiles/: bench line (1 GPU), ncu launch list of the same command, full captures of the forward and the backward kernels.
# The .ncu-rep files (~28 MB each) stay on the box: their raw-metric and per-instruction pages are exported as CSV (gpurun_out/ holds <= 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
bash tools/gpu_bench.sh 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "launch list exit $?"
cap() {   # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name capture exit $?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
}
cap solve_tc 'solve_tc_kernel' 1 python tools/prof_tc.py
NFC_TC_TILES=2 NFC_PROF_B=262144 cap solve_tc2 'solve_tc2_kernel' 1 python tools/prof_tc.py
NFC_ADJ_B=16384 NFC_ADJ_IT=1 cap adjoint_tc 'adjoint_tc_kernel' 0 python tools/gpu_adj_time.py
ls -la gpurun_out/*.csv
