#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/gpu_time_design.py > gpurun_out/time_design.log 2>&1
timeout 900 python scripts/gpu_check_potrf.py > gpurun_out/potrf_check.log 2>&1
timeout 900 python scripts/gpu_parity_probe.py 64 > gpurun_out/parity_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'design_rows' -c 1 \
   -o gpurun_out/prof_design2 -f python scripts/prof_driver.py design 16 1 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/time_design.log; tail -12 gpurun_out/potrf_check.log; cat gpurun_out/parity_probe.log
