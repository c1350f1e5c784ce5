#!/bin/bash
# re-entry sanity pass: tests, smoke, default bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/bench_sweep.json | head -c 1500
