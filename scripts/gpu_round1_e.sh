#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_check_potrf.py > gpurun_out/potrf_check.log 2>&1
echo "potrf_check exit $?" >> gpurun_out/potrf_check.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err
tail -14 gpurun_out/potrf_check.log | head -8;  tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_sweep.json; tail -5 gpurun_out/bench_sweep.err
