#!/bin/bash
# GPU box: INT8-tensor-core Cholesky -- parity tests, check script, gate on/off timing
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_i8.py -x -q -m gpu 2>&1 | tail -6
for g in 1 0; do
  echo "== SPB_I8_GATE=$g"
  SPB_I8_GATE=$g timeout 400 python scripts/gpu_potrf_i8.py > gpurun_out/potrf_i8_gate$g.log 2>&1
  grep -A8 "timing of the" gpurun_out/potrf_i8_gate$g.log
done
grep -B2 -A22 "C ABI" gpurun_out/potrf_i8_gate1.log | grep -v "planes=7" | head -20
