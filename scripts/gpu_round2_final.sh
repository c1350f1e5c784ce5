#!/bin/bash
# GPU box: final pass of the round -- full GPU suite, smoke, bench N = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_final.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
ph = d.get("phases", {})
for k, v in ph.items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    if k == "long_baseline_nt4096":
        print(k, v["value"], v["ms_per_call"], v["cholesky_kernel"], v["roofline"]["frac"], v["parity_vs_reference_golden"], "| fp64:", v["fp64_kernel"]["value"], v["fp64_kernel"]["roofline"]["frac"])
    elif k == "int8_cholesky_nt1000":
        for kk, vv in v.items(): print("  ", kk, vv)
    else:
        print(k, {a: b for a, b in v.items() if a in ("value", "ms_per_call", "ms", "scaling")})
PY
