#!/bin/bash
# same-box A/B: library of commit 6142274 (round-1 Cholesky + fence) against the current one
mkdir -p gpurun_out
for rep in 1 2; do
for lib in libspb200_r2base.so libspb200.so; do
  SPB200_LIB=$PWD/starry_process_b200/$lib timeout 600 python bench.py --batch 4096 --no-phases --steps 10 > gpurun_out/r2d.json 2> gpurun_out/r2d.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2d.json'))
print("$lib: %.1f evals/s  %.3f ms/step  frac %.3f  stages" % (d['value'], d['ms_per_step'], d['roofline']['frac']), {k: round(v/d['steps'],3) for k,v in d['roofline']['stage_ms_total'].items()})
PY
done; done
