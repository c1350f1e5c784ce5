#!/bin/bash
# round 2, pass B: what one GPU of an 8-way strong split sees (512 matrices per step): bench line,
# launch list, and a step-time sweep over batch sizes
mkdir -p gpurun_out
timeout 600 python bench.py --batch 512 --no-phases --steps 20 > gpurun_out/r2b_bench_B512.json 2> gpurun_out/r2b_bench_B512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/r2b_launches_B512.csv python scripts/prof_driver.py sweep 512 2 > gpurun_out/r2b_ncu_list.log 2>&1
python scripts/launch_agg.py gpurun_out/r2b_launches_B512.csv > gpurun_out/r2b_launches_B512_summary.txt 2>&1
cat gpurun_out/r2b_launches_B512_summary.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2b_batch_sweep.log
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench, starry_process_b200 as spb
dev = torch.device("cuda:0")
for B in (256, 296, 444, 512, 592, 1024, 2048, 4096):
    hp, t, flux, _ = bench.synthetic_inputs(B, seed=1234)
    hd = {k: torch.tensor(v, device=dev) for k, v in hp.items()}
    td, fd = torch.tensor(t, device=dev), torch.tensor(flux, device=dev)
    for _ in range(3):
        spb.StarryProcess(**hd).log_likelihood(td, fd, 1e-6, p=1.0, u=[0.4, 0.26])
    stage = {}
    spb.StarryProcess._stage_ms = stage
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    K = 10
    for _ in range(K):
        spb.StarryProcess(**hd).log_likelihood(td, fd, 1e-6, p=1.0, u=[0.4, 0.26])
    e1.record(); torch.cuda.synchronize()
    spb.StarryProcess._stage_ms = None
    ms = e0.elapsed_time(e1) / K
    st = {k: sum(a.elapsed_time(b) for a, b in v) / K for k, v in stage.items()}
    print("B=%5d  %.3f ms/step  %.1f evals/s  us/eval %.2f  stages %s" % (B, ms, B / ms * 1e3, ms / B * 1e3, {k: round(v, 3) for k, v in st.items()}))
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_B512.json'))
print("B512 value", d['value'], "ms", d['ms_per_step'], "frac", d['roofline']['frac'], d['roofline']['stage_ms_total'], "e2e", d['e2e']['value'])
PY
