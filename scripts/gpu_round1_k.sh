#!/bin/bash
# moments iteration pass: full GPU tests + default bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_sweep.json'))
st=d['roofline']['stage_ms_total']
print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], {k: round(v/d['steps'],2) for k,v in st.items()}, d['parity_vs_reference_golden'], round(d['e2e']['value']), d['gpu_launches'])
PY
