#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3; do
  echo "== SPB_DESIGN_VARIANT=$v" >> gpurun_out/time_design_variants.log
  SPB_DESIGN_VARIANT=$v timeout 300 python scripts/gpu_time_design.py 2>&1 | grep "design matrix" >> gpurun_out/time_design_variants.log
done
timeout 300 python scripts/gpu_check_potrf.py > gpurun_out/potrf_check.log 2>&1
echo "potrf_check exit $?" >> gpurun_out/potrf_check.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'design_rows' -c 1 \
   -o gpurun_out/prof_design3 -f python scripts/prof_driver.py design 16 1 > gpurun_out/ncu_full3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'potrf_lnlike' -c 1 \
   -o gpurun_out/prof_potrf3 -f python scripts/prof_driver.py sweep 592 1 >> gpurun_out/ncu_full3.log 2>&1
cat gpurun_out/time_design_variants.log; tail -16 gpurun_out/potrf_check.log;  tail -5 gpurun_out/pytest_gpu.log
