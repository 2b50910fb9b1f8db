#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_adjoint.py -q --timeout 120 -k "one_column" > gpurun_out/pytest_adj.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_adj.log; tail -25 gpurun_out/pytest_adj.log

