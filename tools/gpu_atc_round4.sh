#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== default tolerances" > gpurun_out/atc_tol.log
timeout 300 python tools/gpu_atc_check.py 200 >> gpurun_out/atc_tol.log 2>&1
echo "== adjoint reltol 1e-6" >> gpurun_out/atc_tol.log
ATC_ADJ_RTOL=1e-6 timeout 300 python tools/gpu_atc_check.py 200 >> gpurun_out/atc_tol.log 2>&1
echo "== fixed dt 1/512" >> gpurun_out/atc_tol.log
ATC_FIXED_DT=0.001953125 timeout 300 python tools/gpu_atc_check.py 200 >> gpurun_out/atc_tol.log 2>&1
cat gpurun_out/atc_tol.log
