#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_TC_DEBUG=1 NFC_TC_TILES=2 timeout 60 python tools/prof_tc_uniform.py > gpurun_out/tc2_dbg.log 2>&1; tail -3 gpurun_out/tc2_dbg.log
