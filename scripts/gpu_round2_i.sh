#!/bin/bash
# GPU box: ncu --set full of the round-2 INT8 / tabulated kernels of one sweep step and of one conditional call
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'syrk_i8_kernel|slice_rows_kernel|rowsum_sym_uniform' -c 3 \
   -o gpurun_out/r2i_sweep -f python scripts/prof_driver.py sweep 592 1 > gpurun_out/r2i_ncu_a.log 2>&1
tail -2 gpurun_out/r2i_ncu_a.log
timeout 900 ncu --set full --clock-control none -k regex:'gemm_i8_lower_kernel|slice256_kernel' -c 3 -s 3 \
   -o gpurun_out/r2i_cond -f python scripts/gpu_cond_one.py > gpurun_out/r2i_ncu_b.log 2>&1
tail -2 gpurun_out/r2i_ncu_b.log
for f in r2i_sweep r2i_cond; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out/r2i_*
