#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adjoint.py -q --timeout 120 -k record_pass > gpurun_out/pytest_rec.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_rec.log; tail -12 gpurun_out/pytest_rec.log
