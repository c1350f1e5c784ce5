#!/bin/bash
# multi-GPU pass: NCCL test of the sharded calls + the bench at N ranks (strong = default; weak under phases)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
tail -3 gpurun_out/r2e_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2e_bench_n$N.json'))
print("N=$N", d['scaling'], "value %.1f evals/s  %.3f ms/step  e2e %.1f" % (d['value'], d['ms_per_step'], d['e2e']['value']), "frac", d['roofline']['frac'], {k: round(v/d['steps'],3) for k,v in d['roofline']['stage_ms_total'].items()})
print("   phases", {k:(v.get('value'), v.get('ms_per_step')) for k,v in d['phases'].items()}, "parity", d['parity_vs_reference_golden'])
PY
