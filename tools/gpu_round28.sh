#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out

timeout 600 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/pytest_fw.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_fw.log; tail -6 gpurun_out/pytest_fw.log
