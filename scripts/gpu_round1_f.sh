#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_check_potrf.py 2>&1 | grep -E "^time|exit" > gpurun_out/potrf_time.log
timeout 900 python bench.py --cpu-evals 8 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches_sweep.csv python scripts/prof_driver.py sweep 1184 2 > gpurun_out/ncu_list.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/potrf_time.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_sweep.json'))
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['stage_ms_total'], d['parity_vs_reference_golden'])
PY
tail -3 gpurun_out/bench_sweep.err
