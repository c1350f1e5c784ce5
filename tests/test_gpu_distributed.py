"""GPU (-m gpu): the sharded product calls of starry_process_b200.distributed with the real kernels.
On one GPU the calls degenerate to world_size 1 and must reproduce the plain StarryProcess results
bit for bit; with >= 2 GPUs the same script runs as two NCCL ranks (one process per GPU)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import FID, U_LD

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import bench
import starry_process_b200 as spb

rank, world = int(sys.argv[1]), int(sys.argv[2])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
if world > 1:
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + sys.argv[3], rank=rank,
                            world_size=world, device_id=dev)
U = [0.4, 0.26]
B = 37                                   # ragged shards
hp, t, flux, _ = bench.synthetic_inputs(B, seed=1234)
hd = {k: torch.tensor(v, device=dev) for k, v in hp.items()}
td, fd = torch.tensor(t[:200], device=dev), torch.tensor(flux[:200], device=dev)
full = spb.log_likelihood_sharded(hd, td, fd, 1e-6, p=1.0, u=U)
ref = spb.StarryProcess(**hd).log_likelihood(td, fd, 1e-6, p=1.0, u=U)
assert full.shape == (B,) and torch.equal(full, ref), float((full - ref).abs().max())
# conditional branch with one inclination per sample (sliced with the batch)
inc = torch.linspace(10.0, 80.0, B, dtype=torch.float64, device=dev)
fullc = spb.log_likelihood_sharded(hd, td, fd, 1e-6, i=inc, p=1.0, u=U,
                                   process_kwargs=dict(marginalize_over_inclination=False))
refc = spb.StarryProcess(marginalize_over_inclination=False, **hd).log_likelihood(td, fd, 1e-6, i=inc, p=1.0, u=U)
assert torch.equal(fullc, refc)
# ensemble: replicated factorisation, light curves split, one all-reduce
fid = dict(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
te, fe = bench.ensemble_flux(300)
ted, fed = torch.tensor(te, device=dev), torch.tensor(fe, device=dev)
joint = spb.ensemble_log_likelihood_sharded(fid, ted, fed, 1e-6, p=1.0, u=U)
refj = spb.StarryProcess(**fid).log_likelihood(ted, fed, 1e-6, p=1.0, u=U)
assert abs(float(joint) - float(refj)) <= 1e-11 * abs(float(refj)), (float(joint), float(refj))
# design matrix: inclination axis split, no collective
gp = spb.StarryProcess(**fid)
incs = torch.tensor([5.0, 30.0, 60.0, 85.0, 89.0], dtype=torch.float64)
A, (i0, i1), axis = spb.design_matrix_sharded(gp, te[:64], incs, p=1.0, u=U)
Afull = gp.design_matrix(te[:64], incs, 1.0, U)
assert axis == "i" and torch.equal(A, Afull[i0:i1])
A, (t0, t1), axis = spb.design_matrix_sharded(gp, te[:64], incs[:1], p=1.0, u=U, axis="t")
assert axis == "t" and (t0, t1) == spb.shard_range(64, rank, world)
assert torch.equal(A, gp.design_matrix(te[:64], incs[:1], 1.0, U)[:, t0:t1])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
print("ok", rank)
"""


def _run(world, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / ("worker_gpu_sharded_%d.py" % world)
    script.write_text(_WORKER % {"root": ROOT})
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("ok" in o for o in outs)


def test_sharded_calls_single_gpu(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run(1, tmp_path)


def test_sharded_calls_two_gpus_nccl(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, tmp_path)
