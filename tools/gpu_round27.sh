#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_embedded.py -q --timeout 120 > gpurun_out/pytest_emb.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_emb.log; tail -12 gpurun_out/pytest_emb.log
timeout 200 python tools/gpu_embedded_bench.py > gpurun_out/embedded_bench.log 2>&1; tail -4 gpurun_out/embedded_bench.log
