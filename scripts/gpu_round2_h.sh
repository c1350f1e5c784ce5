#!/bin/bash
# GPU box: sanitizers over every kernel family incl. the INT8 Cholesky
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python scripts/gpu_sanitize_small.py > gpurun_out/r2h_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2h_memcheck.log
