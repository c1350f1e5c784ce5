"""World-size-2 gloo tests (CPU) of the sharded product calls of starry_process_b200.distributed
(SURVEY.md section 8(e)): the split / collective logic runs for real, the per-rank evaluation is
the CPU oracle injected through the ``evaluate`` hook (the product itself has no CPU path).  The GPU
counterparts (NCCL, real kernels) are tests/test_gpu_distributed.py."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from oracle import sp_oracle as so
from starry_process_b200 import distributed as D

rank = int(sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=2)
rng = np.random.default_rng(3)
B, nt = 5, 40
hyper = dict(r=torch.tensor(rng.uniform(10, 30, B)), mu=torch.tensor(rng.uniform(0, 80, B)),
             sigma=torch.tensor(rng.uniform(5, 30, B)), c=0.1, n=torch.tensor(rng.uniform(1, 10, B)))
t = np.linspace(0, 2, nt)
flux = 1e-3 * rng.standard_normal(nt)

calls = []
def evaluate(shard, t_, f_, dc, **kw):
    # the CPU oracle, one element at a time: stands in for StarryProcess(**shard).log_likelihood
    n_local = max(int(torch.as_tensor(v).numel()) for v in shard.values())
    calls.append(n_local)
    out = []
    for k in range(n_local):
        hp = {key: (float(torch.as_tensor(v).reshape(-1)[k]) if torch.as_tensor(v).numel() > 1 else float(v))
              for key, v in shard.items()}
        f_k = f_[k] if getattr(f_, "ndim", 1) == 3 else f_
        inc = kw.get("i", 60.0)
        inc = float(inc[k]) if hasattr(inc, "ndim") and inc.ndim == 1 else float(inc)
        o = so.OracleProcess(marginalize_over_inclination=False, **hp)
        out.append(o.log_likelihood(t_, np.asarray(f_k), dc, i=inc, p=1.0))
    return torch.tensor(out, dtype=torch.float64)

# (1) configs[2]/[3]: one sweep sharded, ragged shards (3 + 2), per-sample inclinations sliced too
inc = torch.tensor(rng.uniform(20, 80, B))
full = D.log_likelihood_sharded(hyper, t, flux, 1e-6, evaluate=evaluate, i=inc)
assert calls == [3 if rank == 0 else 2], calls
ref = evaluate(hyper, t, flux, 1e-6, i=inc)
assert full.shape == (B,) and torch.equal(full, ref), (full, ref)
loc, (b0, b1) = D.log_likelihood_sharded(hyper, t, flux, 1e-6, evaluate=evaluate, i=inc, gather=False)
assert (b0, b1) == D.shard_range(B, rank, 2) and torch.equal(loc, ref[b0:b1])
# per-sample light curves (B, M, nt) are sliced with the batch
f3 = 1e-3 * rng.standard_normal((B, 1, nt))
full3 = D.log_likelihood_sharded(hyper, t, f3, 1e-6, evaluate=evaluate, i=inc)
assert torch.equal(full3, evaluate(hyper, t, f3, 1e-6, i=inc))

# (2) configs[1]: ensemble sharing one K, light curves split, one all-reduce of a scalar
M = 7
fens = 1e-3 * rng.standard_normal((M, nt))
fid = dict(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
def eval_ens(hp, t_, f_, dc, **kw):
    o = so.OracleProcess(marginalize_over_inclination=False, **hp)
    return torch.tensor([o.log_likelihood(t_, np.asarray(f_), dc, i=60.0, p=1.0)], dtype=torch.float64)
joint = D.ensemble_log_likelihood_sharded(fid, t, fens, 1e-6, evaluate=eval_ens)
ref_joint = eval_ens(fid, t, fens, 1e-6)[0]
assert abs(float(joint) - float(ref_joint)) <= 1e-12 * abs(float(ref_joint)), (joint, ref_joint)
# -inf on one rank (its columns only) must survive the reduction
def eval_inf(hp, t_, f_, dc, **kw):
    return torch.tensor([-float("inf") if rank == 1 else 1.0], dtype=torch.float64)
assert float(D.ensemble_log_likelihood_sharded(fid, t, fens, 1e-6, evaluate=eval_inf)) == -float("inf")

# (3) configs[4]: design matrix split along the inclination axis (3 inclinations: 2 + 1) and,
#     for a single inclination, along time; no collective
incs = np.array([30.0, 60.0, 85.0])
o = so.OracleProcess(**fid)
def eval_design(t_, i_, p_, u_):
    return torch.tensor(np.stack([o.design_matrix(np.asarray(t_), float(x), p_, [0.0, 0.0] if u_ is None else u_)
                                  for x in np.atleast_1d(np.asarray(i_))]))
A, (i0, i1), axis = D.design_matrix_sharded(None, t, incs, evaluate=eval_design)
assert axis == "i" and (i0, i1) == D.shard_range(3, rank, 2) and A.shape == (i1 - i0, nt, 256)
assert torch.equal(A, eval_design(t, incs[i0:i1], 1.0, None))
A, (t0, t1), axis = D.design_matrix_sharded(None, t, incs[:1], evaluate=eval_design)
assert axis == "t" and (t0, t1) == D.shard_range(nt, rank, 2) and A.shape == (1, t1 - t0, 256)
assert torch.equal(A, eval_design(t[t0:t1], incs[:1], 1.0, None))
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_calls_two_ranks_gloo(tmp_path, oracle):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker_sharded.py"
    script.write_text(_WORKER % {"root": ROOT})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("ok" in o for o in outs)
