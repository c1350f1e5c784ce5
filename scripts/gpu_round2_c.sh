#!/bin/bash
# bench A/B of the Cholesky geometries inside the real step (affine-fused loads, marginal sweep)
mkdir -p gpurun_out
run() { # name env B
  env $2 timeout 600 python bench.py --batch $3 --no-phases --steps 10 > gpurun_out/r2c_$1_B$3.json 2> gpurun_out/r2c_$1_B$3.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2c_$1_B$3.json'))
print("$1 B $3: %.1f evals/s  %.3f ms/step  frac %.3f  stages" % (d['value'], d['ms_per_step'], d['roofline']['frac']), {k: round(v/d['steps'],3) for k,v in d['roofline']['stage_ms_total'].items()})
PY
}
run t128 SPB_CHOL_TILE=128 4096
run t128nopf "SPB_CHOL_TILE=128 SPB_NO_PREFETCH=1" 4096
run t64 SPB_CHOL_TILE=64 4096
run t64nopf "SPB_CHOL_TILE=64 SPB_NO_PREFETCH=1" 4096
run auto SPB_X=1 4096
run auto SPB_X=1 512
