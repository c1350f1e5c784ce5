#!/bin/bash
# round 2, evidence pass: sanitizers, launch list of one step, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/gpu_sanitize_small.py > gpurun_out/r2f_memcheck.log 2>&1
tail -4 gpurun_out/r2f_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python scripts/gpu_sanitize_small.py > gpurun_out/r2f_racecheck.log 2>&1
grep -c "Race reported\|hazard" gpurun_out/r2f_racecheck.log; tail -3 gpurun_out/r2f_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r2f_launches_sweep_B1184.csv python scripts/prof_driver.py sweep 1184 2 > gpurun_out/r2f_ncu_list.log 2>&1
python scripts/launch_agg.py gpurun_out/r2f_launches_sweep_B1184.csv > gpurun_out/r2f_launches_sweep_B1184_summary.txt 2>&1
tail -25 gpurun_out/r2f_launches_sweep_B1184_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'potrf_lnlike|rowsum_sym|moments_k|gemm_nt|marginal' -c 14 \
   -o gpurun_out/r2f_prof_sweep -f python scripts/prof_driver.py sweep 592 1 > gpurun_out/r2f_ncu_full.log 2>&1
tail -2 gpurun_out/r2f_ncu_full.log
ls -la gpurun_out/r2f_prof_sweep.ncu-rep
