#!/usr/bin/env python
"""bench.py -- lnlike evals/sec (ydeg=15, nt=1000, fp64) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload sweep|ensemble] [--batch B]

Workload `sweep` (BASELINE.json configs[2], the configuration the metric is quoted on; it fits
one GPU): an MCMC-style sweep of B hyperparameter samples x 1 synthetic light curve (nt=1000),
marginalize_over_inclination=True, normalized=True (the reference defaults), limb darkened.
One STEP = one pass of the hot path over the batch: Ylm moments -> marginal kernel -> covariance
assembly -> batched Cholesky/solve/logdet -> B log-likelihoods.  Weak scaling: every rank owns B
samples (no data-path collective; one all-gather of the B*N log-likelihoods per step).

Workload `ensemble` (configs[1]): 1024 light curves sharing one hyperparameter set.

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = the same metric
through the public API from pinned HOST buffers (H2D of the hyperparameters / light curve and D2H
of the log-likelihoods inside the timed region).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NT = 1000
U_LD = [0.4, 0.26]
METRIC = "lnlike evals/sec (ydeg=15, nt=1000, fp64)"


def synthetic_inputs(B, seed):
    """Hyperparameter draws from the reference's own stability prior with the contrast range
    narrowed so that the normalised process stays inside its validity range (as in the golden
    sweep, oracle/gen_golden.py), plus the fiducial synthetic light curve."""
    rng = np.random.default_rng(seed)
    hp = dict(r=rng.uniform(10, 30, B), c=rng.uniform(0.01, 0.15, B), n=rng.uniform(1, 12, B),
              mu=rng.uniform(0, 85, B), sigma=rng.uniform(5, 40, B))
    g = np.load(os.path.join(ROOT, "tests", "golden", "fiducial_nt1000.npz"))
    return hp, g["t"].copy(), g["flux_norm"].copy(), g["flux_ens_norm"].copy()


class ClockSampler(object):
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): NVML polled
    every 10 ms from a thread (the same counters `nvidia-smi --query-gpu=clocks.sm,
    clocks_event_reasons.*` prints), falling back to an `nvidia-smi -lms 100` subprocess."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None
        self.nvml = None
        self._stop = threading.Event()
        self.sm, self.smax, self.reasons = [], [], set()

    def _nvml_index(self):
        # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(x) for x in vis.split(",") if x.strip() != ""]
            return ids[self.index] if ids else self.index
        except (ValueError, IndexError):
            return self.index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.smax.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, attr in bits:
                    if r & getattr(nv, attr, 0):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": max(self.smax) if self.smax else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (reference C++ compiled in place when oracle/_ref is
# available, else the plain-C restatement) + NumPy/SciPy LAPACK on the host cores
# ------------------------------------------------------------------------------------------
_POOL = None


def _worker_init():
    # one LAPACK thread per worker process: the host cores are used by evaluating independent
    # hyperparameter samples in parallel (the reference itself is single-threaded per evaluation)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    from oracle import sp_oracle  # noqa: F401  (load the shared objects once per worker)


def _eval_one(job):
    from oracle import sp_oracle as so

    r, mu, sigma, c, n, t, flux, native = job
    o = so.OracleProcess(r=r, mu=mu, sigma=sigma, c=c, n=n, native=native)
    return o.log_likelihood(t, flux, 1e-6, p=1.0, u=U_LD)


def close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.shutdown(wait=True)
        _POOL = None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_evals(hp, t, flux, n_eval, workload, flux_ens=None):
    """The reference's CPU implementation of the path (oracle/_ref: its own compiled C++ + NumPy /
    SciPy LAPACK) on the host cores.  sweep: independent evaluations spread over one worker
    process per core.  ensemble: one joint evaluation (LAPACK multithreaded)."""
    global _POOL
    from oracle import sp_oracle as so

    native = "ref" if so.ref_available(15, 2) else "port"
    kind = "reference" if native == "ref" else "port"
    out = []
    if workload == "sweep":
        cores = host_cores()
        if _POOL is None and cores > 1:
            import multiprocessing as mp
            from concurrent.futures import ProcessPoolExecutor

            _POOL = ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn"),
                                        initializer=_worker_init)
            list(_POOL.map(_eval_one, [(hp["r"][0], hp["mu"][0], hp["sigma"][0], hp["c"][0],
                                        hp["n"][0], t[:50], flux[:50], native)] * cores))  # spin up
        jobs = [(hp["r"][s], hp["mu"][s], hp["sigma"][s], hp["c"][s], hp["n"][s], t, flux, native)
                for s in range(n_eval)]
        t0 = time.perf_counter()
        out = list(_POOL.map(_eval_one, jobs)) if _POOL is not None else [_eval_one(j) for j in jobs]
        dt = time.perf_counter() - t0
        n_done = n_eval
    else:
        t0 = time.perf_counter()
        o = so.OracleProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0, native=native)
        out.append(o.log_likelihood(t, flux_ens, 1e-6, p=1.0, u=U_LD))
        dt = time.perf_counter() - t0
        n_done = flux_ens.shape[0]
    return n_done / dt, kind, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import sp_oracle as so  # noqa: F401  (fails loudly if the oracle is missing)

    B = args.batch
    hp, t, flux, fens = synthetic_inputs(max(B, 64), seed=1234)
    n_eval = args.ref_evals
    if args.workload == "ensemble":
        fens = np.tile(fens, (128, 1))[:1024]
    for _ in range(args.warmup):
        cpu_evals(hp, t, flux, 1, args.workload, fens[:8])
    rates = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_, kind, _ = cpu_evals(hp, t, flux, n_eval, args.workload, fens)
        rates.append(r_)
    total = time.perf_counter() - t0
    value = float(np.mean(rates))
    cores = host_cores()
    sample = ("%d lnlike evaluations per step of the same workload (each: Ylm moments + marginal "
              "kernel + 1000x1000 Cholesky), reference C++ + NumPy/SciPy on the host, one worker "
              "process per core" % n_eval) if args.workload == "sweep" else \
        "one joint lnlike of 1024 light curves per step (1 factorisation + 1024-RHS solve)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "nt": NT, "ydeg": 15, "batch_per_gpu": B},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_name(args):
    if args.workload == "sweep":
        return ("configs[2]: MCMC-style sweep of %d hyperparameter samples per GPU x 1 light curve "
                "(nt=1000), marginalize_over_inclination=True, normalized=True, u=[0.4,0.26]"
                % args.batch)
    return ("configs[1]: ensemble of 1024 light curves (nt=1000) sharing one hyperparameter set: "
            "one K factorisation + 1024-RHS forward solve")


def ncu_traffic(kernel, units):
    """DRAM bytes (read + write) per launch from the committed ncu --set full capture of the same
    kernel (profiles/ncu_traffic.json, scripts/ncu_traffic.py), scaled by the units in this launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return float(json.load(fh)[kernel]["dram_bytes_per_unit"]) * units
    except (OSError, KeyError, ValueError):
        return None


def hbm_peak():
    """Measured HBM copy bandwidth (driver-written MEASURED_PEAKS.json), else the recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def design_phase(spb, _lib, ctx, dev, torch, I=64, nt=100000, reps=5):
    lib, h = ctx.lib, ctx.handle
    P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
    gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
    rta1 = gp._rTA1(U_LD)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    t = torch.linspace(0, 40, nt, dtype=torch.float64, device=dev)
    inc = torch.arccos(torch.rand(I, dtype=torch.float64, device=dev, generator=gen))  # isotropic
    per = torch.ones(I, dtype=torch.float64, device=dev)
    A = torch.empty(I, nt, 256, dtype=torch.float64, device=dev)       # 13.1 GB >> L2
    nb = lib.spb_design_matrix_workspace_bytes(h, I, nt)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    times = []
    for r in range(reps + 3):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_design_matrix(h, I, nt, P(t), P(inc), P(per), P(rta1), 0, P(A), P(ws), nb,
                                         stream))
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    alg_bytes = I * nt * (2048.0 + 8.0)
    peak, src = hbm_peak()
    ach = alg_bytes / (ms * 1e-3) / 1e9
    del A, ws
    return {"bound": "hbm", "kernel": "design_rows_kernel (+ rx_kernel, design_v_kernel prologue)",
            "workload": "configs[4]: design matrix, nt=1e5 timestamps x 64 inclinations, u=[0.4,0.26]",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
            "ms_per_launch": ms, "rows_per_s": I * nt / (ms * 1e-3),
            "algorithmic_bytes_per_launch": alg_bytes,
            "traffic": ncu_traffic("design_rows_kernel", float(I) * nt)}


def sample_ylm_phase(spb, _lib, ctx, dev, torch, dmma_peak, nsamples=1000000, reps=5):
    """configs[4], second half: prior draws y = mean + L u (sp.py:505-509) for 1e6 samples.  One NT
    GEMM on the FP64 tensor pipe; per draw 2 * 256^2 flop against 2048 B read (u) + 2048 B written
    (y), i.e. 32 flop/B: on B200 (37 TFLOP/s FP64 tensor vs 6.5 TB/s) it sits on the tensor side
    of the ridge (5.7 flop/B), so both fractions are reported."""
    lib, h = ctx.lib, ctx.handle
    P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
    gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
    L = gp.cho_cov_ylm.reshape(1, 256, 256).contiguous()
    mean = gp.mean_ylm.reshape(1, 256).contiguous()
    gen = torch.Generator(device=dev)
    gen.manual_seed(6)
    un = torch.randn(1, nsamples, 256, dtype=torch.float64, device=dev, generator=gen)  # 2 GB >> L2
    y = torch.empty(1, nsamples, 256, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    times = []
    for r in range(reps + 3):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.spb_sample_ylm(h, 1, nsamples, P(mean), P(L), P(un), P(y), stream))
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    flops = 2.0 * 256 * 256 * nsamples
    alg_bytes = nsamples * 4096.0
    peak_bw, src = hbm_peak()
    out = {"bound": "tensor", "kernel": "gnt::gemm_nt_kernel<EPI_ADD_ROWVEC> (DMMA)",
           "workload": "configs[4]: sample_ylm, 1e6 prior draws of the 256 Ylm coefficients",
           "achieved": flops / (ms * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
           "frac": flops / (ms * 1e-3) / 1e12 / dmma_peak, "ms_per_launch": ms,
           "draws_per_s": nsamples / (ms * 1e-3), "algorithmic_flops_per_launch": flops,
           "hbm": {"achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak_bw, "unit": "GB/s",
                   "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak_bw, "peak_source": src,
                   "algorithmic_bytes_per_launch": alg_bytes}}
    del un, y
    return out


# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import starry_process_b200 as spb
    from starry_process_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    hp, t, flux, fens = synthetic_inputs(B, seed=1234 + rank)
    ctx = spb.get_context(local)
    lib, h = ctx.lib, ctx.handle

    # FP64 tensor-pipe peak, measured in this run (MEASURED_PEAKS.json carries no fp64 entry)
    tf = ctypes.c_double()
    ms = ctypes.c_double()
    _lib.check(lib.spb_dmma_peak(h, 40000, ctypes.byref(tf), ctypes.byref(ms)))
    dmma_peak = tf.value

    t_d = torch.tensor(t, device=dev)
    f_d = torch.tensor(flux, device=dev)
    hp_d = {k: torch.tensor(v, device=dev) for k, v in hp.items()}
    if args.workload == "ensemble":
        fens = np.tile(fens, (128, 1))[:1024]
        fe_d = torch.tensor(fens, device=dev)
    # pinned host copies for the end-to-end arm
    hp_h = {k: torch.tensor(v).pin_memory() for k, v in hp.items()}
    t_h = torch.tensor(t).pin_memory()
    f_h = torch.tensor(flux).pin_memory()
    out_h = torch.empty(B, dtype=torch.float64).pin_memory()
    if args.workload == "ensemble":
        fe_h = torch.tensor(fens).pin_memory()

    stage_ms = {}

    def step_device():
        if args.workload == "sweep":
            gp = spb.StarryProcess(**hp_d)
            gp._stage_ms = stage_ms
            ll = gp.log_likelihood(t_d, f_d, 1e-6, p=1.0, u=U_LD)
        else:
            gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
            gp._stage_ms = stage_ms
            ll = gp.log_likelihood(t_d, fe_d, 1e-6, p=1.0, u=U_LD).reshape(1)
        if world > 1:
            ll = spb.gather_lnlike(ll, equal_shards=True)
        return ll

    def step_e2e():
        if args.workload == "sweep":
            hd = {k: v.to(dev, non_blocking=True) for k, v in hp_h.items()}
            gp = spb.StarryProcess(**hd)
            ll = gp.log_likelihood(t_h.to(dev, non_blocking=True), f_h.to(dev, non_blocking=True),
                                   1e-6, p=1.0, u=U_LD)
            out_h.copy_(ll, non_blocking=True)
        else:
            gp = spb.StarryProcess(r=10.0, mu=30.0, sigma=5.0, c=0.1, n=10.0)
            ll = gp.log_likelihood(t_h.to(dev, non_blocking=True), fe_h.to(dev, non_blocking=True),
                                   1e-6, p=1.0, u=U_LD).reshape(1)
            out_h[:1].copy_(ll, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_, op=dist.ReduceOp.MAX)
        return ms_.item(), out

    for _ in range(max(args.warmup, 3)):
        step_device()
    stage_ms.clear()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    total_ms, ll = timed(step_device, args.steps)
    launches = ctx.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    stage_snapshot = {k: list(v) for k, v in stage_ms.items()}
    for _ in range(2):
        step_e2e()
    e2e_ms, _ = timed(step_e2e, args.steps)

    units = (B if args.workload == "sweep" else 1024) * world
    value = units * args.steps / (total_ms * 1e-3)
    e2e_value = units * args.steps / (e2e_ms * 1e-3)
    if args.workload == "sweep":
        h2d = (5 * B + 2 * NT) * 8
        d2h = 8 * B
    else:
        h2d = (1024 * NT + NT) * 8
        d2h = 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (batched Cholesky): events recorded on the launching
    # stream around every launch inside the timed region
    torch.cuda.synchronize()
    chol = stage_snapshot.get("cholesky", [])
    chol_ms = float(np.sum([a.elapsed_time(b) for a, b in chol])) if chol else float("nan")
    n_launch = max(len(chol), 1)
    if args.workload == "sweep":
        flops_total = args.steps * B * (NT ** 3 / 3.0 + NT ** 2 * 1.0)
    else:
        flops_total = args.steps * (NT ** 3 / 3.0 + NT ** 2 * 1024.0)
    achieved = flops_total / (chol_ms * 1e-3) / 1e12 if chol else None
    shares = {}
    for k, v in stage_snapshot.items():
        shares[k] = float(np.sum([a.elapsed_time(b) for a, b in v]))
    roofline = {
        "bound": "tensor", "kernel": "potrf_lnlike_kernel (DMMA m8n8k4.f64 left-looking Cholesky + "
                                     "augmented forward solve)",
        "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
        "frac": (achieved / dmma_peak) if achieved else None,
        "traffic": ncu_traffic("potrf_lnlike_kernel", flops_total / n_launch / (NT ** 3 / 3.0 + NT ** 2)
                               if args.workload == "sweep" else 1.0),
        "traffic_note": "DRAM read+write bytes per launch: per-matrix figure of the ncu capture "
                        "profiles/r01_ncu_sweep_B592_final2.txt (profiles/ncu_traffic.json) x matrices in the launch",
        "peak_source": "FP64 mma.sync peak measured in this run by spb_dmma_peak "
                       "(MEASURED_PEAKS.json has no fp64 entry; cuBLAS DGEMM 8192^3 on this pool: "
                       "35.5 TFLOP/s)",
        "launches_timed": n_launch, "ms_per_launch": chol_ms / n_launch if chol else None,
        "algorithmic_flops_per_launch": flops_total / n_launch,
        "stage_ms_total": shares,
    }
    if args.workload != "sweep":
        roofline["note"] = ("configs[1] is ONE 1000 x 1000 factorisation (an 8-SM thread-block "
                            "cluster, latency-bound by its 16 dependent panels) plus 1024 "
                            "right-hand-side rows: a latency workload, not a throughput one -- the "
                            "headline roofline is the sweep workload's")
    # ---- design-matrix phase (BASELINE configs[4]: nt = 1e5 timestamps x 64 inclinations), timed in
    # the same process with CUDA events on the launching stream, against the measured HBM peak
    phases = {"design_matrix": design_phase(spb, _lib, ctx, dev, torch),
              "sample_ylm": sample_ylm_phase(spb, _lib, ctx, dev, torch, dmma_peak)}
    # ---- CPU baseline: bounded sample of the same workload on the host cores
    cb_value, cb_kind, cb_out = cpu_evals(hp, t, flux, args.cpu_evals, args.workload, fens)
    parity = None
    parity_golden = None
    if args.workload == "sweep":
        llall = ll.cpu().numpy()
        llh = llall[: args.cpu_evals]
        ref = np.array(cb_out)
        fin = np.isfinite(ref)
        parity = float(np.max(np.abs(llh[fin] - ref[fin]) / np.abs(ref[fin]))) if fin.any() else None
        # the same first samples evaluated by the UNMODIFIED reference in the build container
        # (tests/golden/bench_sweep_seed1234.npz, oracle/gen_golden_bench.py).  This is the parity
        # figure: the live oracle on this host is only reproducible to ~1e-6 across CPUs
        # (DESIGN.md "numerical fragility"), the fixture is what every golden file was made with.
        gpath = os.path.join(ROOT, "tests", "golden", "bench_sweep_seed1234.npz")
        if os.path.exists(gpath):
            gref = np.load(gpath)["lnlike_m1_n1"]
            ng = min(len(gref), B)
            gref = gref[:ng]
            gfin = np.isfinite(gref)
            same_inf = bool(np.array_equal(np.isneginf(gref), np.isneginf(llall[:ng])))
            gerr = np.abs(llall[:ng][gfin] - gref[gfin]) / np.abs(gref[gfin])
            parity_golden = {
                "max_rel": float(np.max(gerr)), "median_rel": float(np.median(gerr)),
                "n": int(ng), "neg_inf_pattern_equal": same_inf, "tolerance": 1e-8}
    cpu_baseline = {
        "value": cb_value, "unit": "evals/s", "cores": host_cores(), "kind": cb_kind,
        "sample": "%d evaluations of the same workload, reference C++ + NumPy/SciPy on the host, one "
                  "worker process per core" % (args.cpu_evals if args.workload == "sweep" else 1024),
        "max_rel_lnlike_diff_vs_gpu_on_sample": parity,
        "note": "live oracle on THIS host; the reference algorithm moves by up to ~3e-6 between "
                "CPUs (LAPACK kernel selection), see parity_vs_reference_golden for the pinned check",
    }
    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "nt": NT, "ydeg": 15, "batch_per_gpu": B,
                   "l2": "per-step working set (B x 8 MB covariance matrices) is far larger than "
                         "the 126 MB L2; no explicit flush needed"},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "parity_vs_reference_golden": parity_golden, "phases": phases,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Anything a library prints on fd 1 (e.g. NCCL's version banner) goes to stderr; the one JSON
    line is written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sweep", choices=["sweep", "ensemble"])
    ap.add_argument("--batch", type=int, default=4096, help="hyperparameter samples per GPU")
    ap.add_argument("--cpu-evals", type=int, default=64, help="CPU-baseline sample size")
    ap.add_argument("--ref-evals", type=int, default=96, help="reference arm: evaluations per step")
    args = ap.parse_args()
    quiet_stdout()
    try:
        if args.impl == "reference":
            return run_reference(args)
        return run_b200(args)
    finally:
        close_pool()


if __name__ == "__main__":
    sys.exit(main())
