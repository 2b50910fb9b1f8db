cd "$(dirname "$0")/.."
NFC_ADJ_B=16384 NFC_ADJ_IT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:adjoint_tc_kernel -c 1 -o /tmp/adjoint_tc python tools/gpu_adj_time.py > gpurun_out/ncu_adjoint_tc.log 2>&1; echo "capture exit $?"
ncu -i /tmp/adjoint_tc.ncu-rep --page raw --csv > gpurun_out/adjoint_tc_raw.csv 2>/dev/null
ncu -i /tmp/adjoint_tc.ncu-rep --page source --csv > gpurun_out/adjoint_tc_source.csv 2>/dev/null
ls -la gpurun_out/adjoint_tc_*.csv
