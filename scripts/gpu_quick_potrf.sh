#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_check_potrf.py > gpurun_out/potrf_check.log 2>&1
echo "exit $?" >> gpurun_out/potrf_check.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
head -14 gpurun_out/potrf_check.log; grep -E "^time|exit" gpurun_out/potrf_check.log; tail -3 gpurun_out/pytest_gpu.log
