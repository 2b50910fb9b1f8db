#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NFC_TC_DEBUG=1 NFC_TC_TILES=1 timeout 60 python tools/prof_tc_uniform.py 2>&1 | tail -5 | cut -c1-230
NFC_TC_DEBUG=1 NFC_TC_TILES=2 timeout 60 python tools/prof_tc_uniform.py 2>&1 | tail -3 | cut -c1-330
NFC_TC2_SCALES=1e-5 NFC_TC2_B=300,65536,262144 timeout 120 python tools/gpu_tc2_check.py 2>&1 | tail -3 | cut -c1-330
