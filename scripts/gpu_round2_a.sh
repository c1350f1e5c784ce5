#!/bin/bash
# round 2, pass A: GPU tests (incl. the new round-2 file), smoke, bench N=1 (strong == weak at N=1)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 ) > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1
timeout 1200 python bench.py > gpurun_out/r2a_bench_sweep.json 2> gpurun_out/r2a_bench_sweep.err
tail -25 gpurun_out/r2a_pytest_gpu.log; tail -2 gpurun_out/r2a_smoke.log; tail -3 gpurun_out/r2a_bench_sweep.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2a_bench_sweep.json'))
    print("value", d['value'], "ms", d['ms_per_step'], "frac", d['roofline']['frac'], d['roofline']['stage_ms_total'])
    print("e2e", d['e2e']); print("parity", d['parity_vs_reference_golden']); print("cpu", d['cpu_baseline'])
    for k,v in d['phases'].items():
        print(k, {a:b for a,b in v.items() if a in ('value','unit','ms_per_call','ms_per_launch','frac','achieved','error','parity_vs_reference_golden','finite_fraction')}, (v.get('roofline') or {}).get('frac'))
except Exception as e: print("ERR", e)
PY
