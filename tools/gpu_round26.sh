#!/bin/bash
# final round-1 evidence: bench line, ncu launch list of the same command, full capture of the headline kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-large-batch > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_tc_kernel -s 1 -c 1 -o gpurun_out/prof_solve_tc python tools/prof_tc.py > gpurun_out/ncu_tc.log 2>&1; echo "ncu full exit $?"; tail -2 gpurun_out/ncu_tc.log
