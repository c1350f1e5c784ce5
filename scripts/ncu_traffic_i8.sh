#!/bin/bash
# GPU box: ncu --set full capture of potrf_i8_kernel (592 matrices, nt = 1000, 7 planes of 8-bit digits)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:potrf_i8 -c 1 -s 1 -o gpurun_out/i8_full_592 python scripts/gpu_i8_one.py 592 78 1000 > gpurun_out/i8_ncu592.log 2>&1
echo rc=$?; tail -2 gpurun_out/i8_ncu592.log
ncu -i gpurun_out/i8_full_592.ncu-rep --page raw --csv > gpurun_out/i8_full_592_raw.csv 2>/dev/null
