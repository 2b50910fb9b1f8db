#!/bin/bash
# Evidence for profiles/: bench line (1 GPU), ncu launch list of the same command, full captures of the forward and the backward kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_bench.sh 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "launch list exit $?"
bash tools/gpu_fwd_prof.sh
NFC_ADJ_B=16384 NFC_ADJ_IT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:adjoint_tc_kernel -c 1 -o gpurun_out/atc_16k python tools/gpu_adj_time.py > gpurun_out/atc_ncu.log 2>&1; echo "adjoint capture exit $?"
NFC_TC_TILES=2 NFC_PROF_B=262144 timeout 600 ncu --set full --import-source on --clock-control none -k regex:solve_tc2_kernel -s 1 -c 1 -o gpurun_out/prof_solve_tc2 python tools/prof_tc.py > gpurun_out/ncu_tc2.log 2>&1; echo "two-tile capture exit $?"
