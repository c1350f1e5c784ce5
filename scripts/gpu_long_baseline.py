import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, bench
import starry_process_b200 as spb
dev = torch.device("cuda")
ctx = spb.get_context()
out = bench.long_baseline_phase(spb, dev, torch, 37.16)
print(out["value"], out["ms_per_call"], out["roofline"]["ms_per_launch"], out["parity_vs_reference_golden"], "| fp64", out["fp64_kernel"]["value"], out["fp64_kernel"]["ms_per_call"], out["fp64_kernel"]["parity_vs_reference_golden"])
