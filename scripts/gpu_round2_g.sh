#!/bin/bash
# GPU box: launch list of one sweep step with the INT8-tensor-core Cholesky as the automatic choice
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/r2g_launches_sweep_B1184.csv python scripts/prof_driver.py sweep 1184 2 > gpurun_out/r2g_ncu_list.log 2>&1
python scripts/launch_agg.py gpurun_out/r2g_launches_sweep_B1184.csv > gpurun_out/r2g_launches_sweep_B1184_summary.txt 2>&1
tail -22 gpurun_out/r2g_launches_sweep_B1184_summary.txt
