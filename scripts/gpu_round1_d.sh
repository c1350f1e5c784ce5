#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/time_design_variants.log
( timeout 600 python -m pytest tests -m gpu -x -q -k "design or wigner or tensordot or flux_operator" ) > gpurun_out/pytest_gpu_design.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_design.log
for v in 0 1; do
  echo "== SPB_DESIGN_VARIANT=$v" >> gpurun_out/time_design_variants.log
  SPB_DESIGN_VARIANT=$v timeout 300 python scripts/gpu_time_design.py 2>&1 | grep -E "design matrix|Error|error" >> gpurun_out/time_design_variants.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'design_rows' -c 1 \
   -o gpurun_out/prof_design4 -f python scripts/prof_driver.py design 16 1 > gpurun_out/ncu_full4.log 2>&1
tail -5 gpurun_out/pytest_gpu_design.log; cat gpurun_out/time_design_variants.log
