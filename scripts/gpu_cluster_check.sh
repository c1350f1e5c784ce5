#!/bin/bash
# cluster Cholesky (small batches): LAPACK check, latency A/B against the single-CTA path, tests
mkdir -p gpurun_out
timeout 600 python scripts/gpu_check_potrf.py 2>&1 | grep -v "^time B=\(296\|1184\|2368\|148\)\|DMMA\|cuBLAS" > gpurun_out/cluster_check.log; cat gpurun_out/cluster_check.log
SPB_NO_CLUSTER=1 timeout 600 python scripts/gpu_check_potrf.py 2>&1 | grep "time B=1 " | sed 's/^/single-CTA path: /'
timeout 600 python scripts/gpu_configs.py 2>&1 | tail -5
SPB_NO_CLUSTER=1 timeout 600 python scripts/gpu_configs.py 2>&1 | head -2 | sed 's/^/single-CTA path: /'
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -3
