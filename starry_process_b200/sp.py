"""
``StarryProcess`` -- B200-native drop-in for the batched likelihood path of
``starry_process.StarryProcess`` (reference: starry_process/sp.py:38-284, 420-441, 489-509,
643-727, 1052-1188).

Same constructor keywords, property names and method signatures as the reference for the hot
path; inputs/outputs are ``torch.float64`` CUDA tensors and evaluation is eager.  Extension: any of
``r, mu, sigma, a, b, c, n`` may be a ``(B,)`` tensor (one hyperparameter sample per element), and
``i`` may be ``(B,)`` in the conditional branch; outputs then carry a leading batch axis and
``log_likelihood`` returns ``(B,)``.  All numerics run in ``libspb200.so`` (hand-written sm_100a CUDA
behind ``include/spb200.h``) -- there is no CPU or PyTorch fallback: without the library (or without
a CUDA device) construction raises.

SURVEY.md section 8(f) ranks 2 and 3 are covered as well: ``sample``, ``predict``,
``sample_conditional`` and ``sample_ylm_conditional`` (sp.py:518-641, 729-765, 767-1002), composed
from the same Cholesky / GEMM / assembly kernels, and time-variable surfaces (``tau``,
``temporal_kernel``: temporal.py:8-16, sp.py:697-698, 893-894, 510-516).

The uniform spot-size prior (``dr``, size.py:55-89, 116-125) is
supported.  Out of scope in this drop-in: pixel-space moments and visualisation.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, _tables, temporal as _temporal

__all__ = ["StarryProcess", "StarryProcessSum", "gauss2beta", "beta2gauss", "defaults"]

# starry_process/defaults.py:4-35
defaults = dict(
    ydeg=15, udeg=2, r=20.0, dr=None, a=0.40, b=0.27, c=0.1, n=10.0, p=1.0, i=60.0,
    u=[0.0, 0.0], tau=None, normalized=True, normalization_order=20, normalization_zmax=0.023,
    marginalize_over_inclination=True, baseline_mean=0.0, baseline_var=0.0, covpts=300,
    log_alpha_max=10, log_beta_max=10, abmin=1e-12, sigma_max=45.0, epsy=1e-12, epsy15=1e-9,
    eps=1e-8,
)

_CTX = {}
# module-wide default of the `longitude_basis` keyword (environment: SPB200_LONGITUDE_BASIS)
DEFAULT_LONGITUDE_BASIS = os.environ.get("SPB200_LONGITUDE_BASIS", "pinned")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Context(object):
    """One libspb200 context (constant tables resident in HBM) per CUDA device."""

    def __init__(self, device, longitude_basis="pinned"):
        if not torch.cuda.is_available():
            raise RuntimeError("starry_process_b200 needs a CUDA device (sm_100a); no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.longitude_basis = longitude_basis
        blob, _ = _tables.build_tables(longitude_basis)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            _lib.check(self.lib.spb_create(device, blob.ctypes.data_as(ctypes.c_void_p), blob.size,
                                           ctypes.byref(h)))
        self.handle = h
        # INT8-tensor-core Cholesky (spb_cholesky_lnlike_i8): 0 = always the FP64 (DMMA) kernel; 78 = seven
        # planes of 8-bit digits (55 bits per row: FP64 rounding-noise level), 8 / 87 = eight planes of 7-bit
        # digits (56 bits), 7 / 77 = seven planes of 7-bit digits (49 bits) whenever the path applies;
        # -1 (default) = automatic: 78 where that kernel is the faster one on B200 -- nt >= 704 and more
        # matrices than the cluster kernel takes (measured Cholesky stage, INT8 vs DMMA: nt = 512 +14 %,
        # 768 -10 %, 1000 -13 % (-25 % conditional), 1536 -45 %, 4096 -55 %)
        self.cholesky_i8 = int(os.environ.get("SPB200_CHOLESKY_I8", "-1"))
        self.num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    def i8_planes(self, nt, batch):
        """Digit planes the batched ``log_likelihood`` uses for ``batch`` matrices of size ``nt`` (0: the
        FP64 kernel)."""
        if self.cholesky_i8 >= 0:
            return self.cholesky_i8
        return 78 if (nt >= 704 and 2 * batch > self.num_sms) else 0

    def set_option(self, name, value):
        """Run-time switches of the library (include/spb200.h: spb_set_option), plus
        ``"cholesky_i8"`` (-1 | 0 | 7 | 8): batched ``log_likelihood`` factorises on the INT8 tensor cores
        with that many 7-bit digit planes (``spb_cholesky_lnlike_i8``); -1 = automatic, 0 = never."""
        if name == "cholesky_i8":
            if int(value) not in (-1, 0, 7, 8, 77, 78, 87):
                raise ValueError("cholesky_i8 must be -1, 0, 7 (= 77), 8 (= 87) or 78")
            self.cholesky_i8 = int(value)
            return
        _lib.check(self.lib.spb_set_option(self.handle, name.encode(), int(value)))

    def launches(self):
        n = ctypes.c_longlong()
        _lib.check(self.lib.spb_launch_count(self.handle, ctypes.byref(n)))
        return n.value


def get_context(device=None, longitude_basis=None):
    """The libspb200 context of ``device`` (created on first use).  ``longitude_basis``:
    ``"pinned"`` (default; the shipped longitude eigenvector table, identical on every host) or
    ``"host"`` (this host's own ``numpy.linalg.eigh``, what a reference run in this process uses) --
    see ``_tables.build_tables``.  The two bases are separate contexts."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    basis = DEFAULT_LONGITUDE_BASIS if longitude_basis is None else longitude_basis
    key = (int(device), basis)
    if key not in _CTX:
        _CTX[key] = _Context(int(device), basis)
    return _CTX[key]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _is_batched(x):
    return (isinstance(x, torch.Tensor) and x.ndim > 0) or (isinstance(x, np.ndarray) and x.ndim > 0) \
        or isinstance(x, (list, tuple))


def gauss2beta(mu, sigma, log_alpha_max=10, log_beta_max=10):
    """latitude.py:14-77 (host utility; accepts floats, arrays or tensors, angles in degrees)."""
    tens = isinstance(mu, torch.Tensor) or isinstance(sigma, torch.Tensor)
    m = torch.as_tensor(mu, dtype=torch.float64) * math.pi / 180
    v = (torch.as_tensor(sigma, dtype=torch.float64) * math.pi / 180) ** 2
    c1, c2, c3 = torch.cos(m), torch.cos(2 * m), torch.cos(3 * m)
    term = 1.0 / (16 * v * torch.cos(0.5 * m) ** 4)
    alpha = (2 + 4 * v + (3 + 8 * v) * c1 + 2 * c2 + c3) * term
    beta = (c1 + 2 * v * (3 + c2) - c3) * term
    a = torch.log(alpha) / log_alpha_max
    b = torch.clamp((torch.log(beta) - math.log(0.5)) / (log_beta_max - math.log(0.5)), min=0.0)
    if tens:
        return a, b
    if a.ndim == 0:
        return float(a), float(b)
    return a.numpy(), b.numpy()


def beta2gauss(a, b, log_alpha_max=10, log_beta_max=10):
    """latitude.py:80-168 (host utility)."""
    tens = isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor)
    a_ = torch.as_tensor(a, dtype=torch.float64)
    b_ = torch.as_tensor(b, dtype=torch.float64)
    alpha = torch.exp(a_ * log_alpha_max)
    beta = torch.exp(math.log(0.5) + b_ * (log_beta_max - math.log(0.5)))
    term = 4 * alpha ** 2 - 8 * alpha - 6 * beta + 4 * alpha * beta + beta ** 2 + 5
    mu = 2 * torch.atan(torch.sqrt(2 * alpha + beta - 2 - torch.sqrt(term)))
    term = 1 - alpha + beta + (beta - 1) * torch.cos(mu) + (alpha - 1) / torch.cos(mu) ** 2
    sigma = torch.sin(mu) / torch.sqrt(term)
    invalid = (alpha <= 1) | (beta <= 0.5)
    nan = torch.full_like(mu, float("nan"))
    mu = torch.where(invalid, nan, mu) * 180 / math.pi
    sigma = torch.where(invalid, nan, sigma) * 180 / math.pi
    if tens:
        return mu, sigma
    if mu.ndim == 0:
        return float(mu), float(sigma)
    return mu.numpy(), sigma.numpy()


def _check_bounds(name, x, lower=-np.inf, upper=np.inf, tol=1e-6):
    """ops/exceptions.py:30-48 for host-resident values."""
    xv = np.atleast_1d(np.asarray(x, dtype=np.float64))
    low = xv < lower - tol
    high = xv > upper + tol
    if low.any():
        raise ValueError("%s out of bounds: %f %s %f" % (name, xv[low][0], "<=", lower))
    if high.any():
        raise ValueError("%s out of bounds: %f %s %f" % (name, xv[high][0], ">=", upper))


class StarryProcess(object):
    def __init__(self, r=defaults["r"], dr=defaults["dr"], c=defaults["c"], n=defaults["n"],
                 tau=defaults["tau"], temporal_kernel=None,
                 marginalize_over_inclination=defaults["marginalize_over_inclination"],
                 normalized=defaults["normalized"], covpts=defaults["covpts"], device=None,
                 **kwargs):
        # sp.py:204-222
        mu = kwargs.pop("mu", None)
        sigma = kwargs.pop("sigma", None)
        if mu is None and sigma is None:
            a = kwargs.pop("a", defaults["a"])
            b = kwargs.pop("b", defaults["b"])
        elif (kwargs.get("a", None) is None and kwargs.get("b", None) is None) and (
                mu is not None and sigma is not None):
            a = b = None
        else:
            raise ValueError("Must provide either `a` and `b` *or* `mu` and `sigma`.")
        # sp.py:225-232; the reference accepts any callable f(t1, t2, tau): the CUDA assembly
        # implements the two kernels the reference ships (temporal.py)
        self._time_variable = tau is not None
        self._tkind = 0
        if self._time_variable:
            tk = _temporal.Matern32Kernel if temporal_kernel is None else temporal_kernel
            if isinstance(tk, str):
                tk = {"matern32": _temporal.Matern32Kernel,
                      "expsquared": _temporal.ExpSquaredKernel}.get(tk.lower())
            kind = getattr(tk, "spb_kind", None)
            if kind is None:
                kind = {"Matern32Kernel": 1, "ExpSquaredKernel": 2}.get(
                    getattr(tk, "__name__", ""), None)
            if kind is None:
                raise NotImplementedError("temporal_kernel must be Matern32Kernel or "
                                          "ExpSquaredKernel (starry_process_b200.temporal)")
            self._tkind = int(kind)
            self._temporal_kernel = tk
        self._ydeg = int(kwargs.pop("ydeg", defaults["ydeg"]))
        self._udeg = int(kwargs.pop("udeg", defaults["udeg"]))
        assert self._ydeg >= 5, "Degree of map must be >= 5."           # sp.py:236
        if self._ydeg > 15:
            raise NotImplementedError("libspb200 evaluates spherical-harmonic degrees up to 15 (the "
                                      "reference is numerically unstable above: joss/paper.md:174-191)")
        if self._udeg not in (0, 2):
            raise NotImplementedError("udeg must be 0 or 2")
        # keyword options the reference threads through its integrals (sp.py:241-262):
        # contrast.py:26-32 (epsy, epsy15), latitude.py:171-200 (abmin, log_alpha_max, log_beta_max)
        self._opt_kw = dict(epsy=float(kwargs.pop("epsy", defaults["epsy"])),
                            epsy15=float(kwargs.pop("epsy15", defaults["epsy15"])),
                            abmin=float(kwargs.pop("abmin", defaults["abmin"])),
                            log_alpha_max=float(kwargs.pop("log_alpha_max", defaults["log_alpha_max"])),
                            log_beta_max=float(kwargs.pop("log_beta_max", defaults["log_beta_max"])))
        self._normN = int(kwargs.pop("normalization_order", defaults["normalization_order"]))
        self._normzmax = float(kwargs.pop("normalization_zmax", defaults["normalization_zmax"]))
        self._max_chunk_bytes = int(kwargs.pop("max_chunk_bytes", 96 << 30))
        self._sigma_max = float(kwargs.pop("sigma_max", defaults["sigma_max"]))
        kwargs.pop("seed", None)
        self._nylm = (self._ydeg + 1) ** 2
        self._covpts = int(covpts)
        self._normalized = bool(normalized)
        self._marginalize_over_inclination = bool(marginalize_over_inclination)

        self._ctx = get_context(device, kwargs.pop("longitude_basis", None))
        self.device = self._ctx.device
        self._lib = self._ctx.lib
        self._opt, self._opt_keep = self._moments_options()

        params = dict(r=r, c=c, n=n)
        if dr is not None:     # uniform prior on the spot radius over [r - dr, r + dr], size.py:116-125
            params["dr"] = dr
        if self._time_variable:
            params["tau"] = tau
        if a is None:
            params.update(mu=mu, sigma=sigma)
        else:
            params.update(a=a, b=b)
        self._batched = any(_is_batched(v) for v in params.values())
        # host-side bounds checks, as the reference raises (device-resident tensors are flagged
        # per element through info[] instead of forcing a device->host sync)
        hostvals = {k: v for k, v in params.items()
                    if not (isinstance(v, torch.Tensor) and v.is_cuda)}
        if "r" in hostvals:
            _check_bounds("r", np.asarray(hostvals["r"], dtype=np.float64) * np.pi / 180, 0,
                          0.5 * np.pi)
        if "dr" in hostvals:
            _check_bounds("dr", np.asarray(hostvals["dr"], dtype=np.float64) * np.pi / 180, 0,
                          0.5 * np.pi)
        if "a" in hostvals:
            _check_bounds("a", hostvals["a"], 0, 1)
        if "b" in hostvals:
            _check_bounds("b", hostvals["b"], 0, 1)
        if "n" in hostvals:
            _check_bounds("n", hostvals["n"], 0, np.inf)
        if "tau" in hostvals:
            _check_bounds("tau", hostvals["tau"], 0, np.inf)
        tens = {k: torch.as_tensor(v, dtype=torch.float64).to(self.device).reshape(-1)
                for k, v in params.items()}
        B = max(t.numel() for t in tens.values())
        for k, t in tens.items():
            if t.numel() not in (1, B):
                raise ValueError("hyperparameter `%s` has %d elements, expected 1 or %d" %
                                 (k, t.numel(), B))
            tens[k] = t.expand(B).contiguous()
        self._B = B
        self._r, self._c, self._n = tens["r"], tens["c"], tens["n"]
        self._tau = tens.get("tau", None)
        self._dr = tens.get("dr", None)
        self._mu_sigma = None
        if a is None:
            self._mu_sigma = (tens["mu"], tens["sigma"])
            self._a = torch.empty(B, dtype=torch.float64, device=self.device)
            self._b = torch.empty(B, dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.spb_gauss2beta(self._ctx.handle, B, _ptr(tens["mu"]),
                                                    _ptr(tens["sigma"]), _ptr(self._a),
                                                    _ptr(self._b), _stream()))
            if "mu" in hostvals:
                ah, bh = gauss2beta(torch.as_tensor(hostvals["mu"], dtype=torch.float64),
                                    torch.as_tensor(hostvals["sigma"], dtype=torch.float64))
                _check_bounds("a", ah.numpy(), 0, 1)
                _check_bounds("b", bh.numpy(), 0, 1)
        else:
            self._a, self._b = tens["a"], tens["b"]
        self._mean_ylm = None
        self._cov_ylm = None
        self._cho_cov_ylm = None
        self._info = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._z = None
        self._rTA1_cache = {}

    def _moments_options(self):
        """spb_moments_options for non-default keywords / a degree below 15 (None: all defaults).
        A lower degree is EMBEDDED in the degree-15 machinery: the spot operator ``Bp`` of that degree
        (size.py:10-43, rows l > ydeg zero) and a jitter vector that vanishes for l > ydeg make every
        Ylm moment with l > ydeg exactly zero, the rotations are block-diagonal in l, and the
        leading ``(ydeg+1)^2`` block is the reference's degree-``ydeg`` result."""
        kw, ydeg = self._opt_kw, self._ydeg
        default = (ydeg == 15 and kw["epsy"] == defaults["epsy"] and kw["epsy15"] == defaults["epsy15"]
                   and kw["abmin"] == defaults["abmin"]
                   and kw["log_alpha_max"] == defaults["log_alpha_max"]
                   and kw["log_beta_max"] == defaults["log_beta_max"])
        if default:
            return None, ()
        opt = _lib.MomentsOptions()
        keep = []
        if ydeg < 15:
            Bp = np.zeros((16, 1000))
            Bp[: ydeg + 1] = _tables.spot_operator(ydeg)
            bp_d = torch.tensor(Bp, dtype=torch.float64, device=self.device).contiguous()
            keep.append(bp_d)
            opt.Bp = bp_d.data_ptr()
        lam = np.zeros(256)
        lam[: self._nylm] = kw["epsy"]
        lam[15 ** 2: self._nylm] = kw["epsy15"]        # contrast.py:28-31 (only l = 15 exists there)
        lam_d = torch.tensor(lam, dtype=torch.float64, device=self.device)
        keep.append(lam_d)
        opt.lambda_ = lam_d.data_ptr()
        opt.abmin, opt.log_alpha_max, opt.log_beta_max = kw["abmin"], kw["log_alpha_max"], kw["log_beta_max"]
        return opt, tuple(keep)

    def _optp(self):
        return ctypes.byref(self._opt) if self._opt is not None else None

    @property
    def ydeg(self):
        """sp.py:347-350."""
        return self._ydeg

    def _cut(self, x, dims):
        """Leading (ydeg+1)^2 block along the given Ylm axes (no-op at degree 15)."""
        if self._nylm == 256:
            return x
        for d in dims:
            x = x.narrow(d, 0, self._nylm)
        return x.contiguous()

    def _pad_ylm(self, x, dim):
        """Zero-pad a Ylm axis of length (ydeg+1)^2 to 256."""
        if x.shape[dim] == 256:
            return x
        if x.shape[dim] != self._nylm:
            raise ValueError("expected a spherical-harmonic axis of length %d" % self._nylm)
        shape = list(x.shape)
        shape[dim] = 256
        out = torch.zeros(shape, dtype=x.dtype, device=x.device)
        out.narrow(dim, 0, self._nylm).copy_(x)
        return out

    # ------------------------------------------------------------------ hyperparameters
    @property
    def a(self):
        return self._out(self._a)

    @property
    def b(self):
        return self._out(self._b)

    @property
    def batch_size(self):
        return self._B

    @property
    def info(self):
        """Per-element status bits (include/spb200.h: SPB_INFO_*)."""
        return self._info

    def _out(self, t):
        return t if self._batched else t[0]

    # optional per-stage CUDA-event timing (bench.py sets ``_stage_ms`` to a dict)
    _stage_ms = None

    def _mark(self, name):
        """Opens an NVTX range for one phase of the path (moments, flux_marginal /
        design, assemble, cholesky; visible in nsys / ncu --nvtx) and, when ``_stage_ms`` is set,
        brackets it with CUDA events on the launching stream."""
        torch.cuda.nvtx.range_push("spb200:" + name)
        if self._stage_ms is None:
            return False
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self._stage_ms.setdefault(name, []).append((e0, e1))
        return e1

    @staticmethod
    def _mark_end(ev):
        if ev is not False:
            ev.record()
        torch.cuda.nvtx.range_pop()

    def log_jac(self):
        """sp.py:1004-1050 -> LatitudeIntegral._log_jac (latitude.py:281-316): log-Jacobian of the
        ``(a, b) -> (mu, sigma)`` transform, ``-inf`` where ``sigma > sigma_max``."""
        out = torch.empty(self._B, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.spb_log_jac_opt(self._ctx.handle, self._B, _ptr(self._a),
                                                 _ptr(self._b), self._sigma_max, self._optp(),
                                                 _ptr(out), _stream()))
        return self._out(out)

    # ------------------------------------------------------------------ Ylm moments
    def _compute_moments(self):
        if self._mean_ylm is not None:
            return
        B = self._B
        with torch.cuda.device(self.device):
            mean = torch.empty(B, 256, dtype=torch.float64, device=self.device)
            cov = torch.empty(B, 256, 256, dtype=torch.float64, device=self.device)
            nbytes = self._lib.spb_ylm_moments_workspace_bytes(self._ctx.handle, B)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            ev = self._mark("moments")
            _lib.check(self._lib.spb_ylm_moments_dr(
                self._ctx.handle, B, _ptr(self._r), _ptr(self._dr), _ptr(self._a), _ptr(self._b),
                _ptr(self._c), _ptr(self._n), self._optp(), _ptr(mean), _ptr(cov), _ptr(self._info),
                _ptr(ws), nbytes, _stream()))
            self._mark_end(ev)
            del ws
        self._mean_ylm, self._cov_ylm = mean, cov

    @property
    def mean_ylm(self):
        """sp.py:420-426."""
        self._compute_moments()
        return self._out(self._cut(self._mean_ylm, (1,)))

    @property
    def cov_ylm(self):
        """sp.py:428-434."""
        self._compute_moments()
        return self._out(self._cut(self._cov_ylm, (1, 2)))

    @property
    def cho_cov_ylm(self):
        """sp.py:436-441 (lower Cholesky factor of ``cov_ylm``)."""
        self._compute_moments()
        if self._cho_cov_ylm is None:
            with torch.cuda.device(self.device):
                cov = self._cov_ylm
                if self._nylm < 256:     # embedded lower degree: unit diagonal on the padded block
                    cov = cov.clone()
                    cov[:, self._nylm:, self._nylm:] += torch.eye(256 - self._nylm, dtype=torch.float64,
                                                                  device=self.device)
                L = torch.empty_like(cov)
                info = torch.zeros(self._B, dtype=torch.int32, device=self.device)
                _lib.check(self._lib.spb_cho_cov_ylm(self._ctx.handle, self._B, _ptr(cov),
                                                     _ptr(L), _ptr(info), _stream()))
                bad = (info & 1) != 0
                L = torch.where(bad[:, None, None], torch.full_like(L, float("nan")), L)
            self._cho_cov_ylm = L
        return self._out(self._cut(self._cho_cov_ylm, (1, 2)))

    def sample_ylm(self, t=None, nsamples=1, u=None, generator=None):
        """sp.py:489-509 (``t`` must be None: static surfaces).  ``u`` optionally supplies the
        standard-normal draws, shape ``(nylm, nsamples)`` as in the reference's
        ``random_normal(self.random, (nylm, nsamples))`` or ``(B, nylm, nsamples)``."""
        if t is not None:
            return self._sample_ylm_temporal(t, nsamples, u, generator)
        self.cho_cov_ylm
        L = self._cho_cov_ylm          # (B, 256, 256), padded at degrees below 15
        B, ny = self._B, self._nylm
        with torch.cuda.device(self.device):
            if u is None:
                un = torch.zeros(B, nsamples, 256, dtype=torch.float64, device=self.device)
                un[:, :, :ny] = torch.randn(B, nsamples, ny, dtype=torch.float64, device=self.device,
                                            generator=generator)
            else:
                un = torch.as_tensor(u, dtype=torch.float64).to(self.device)
                nsamples = un.shape[-1]
                if un.ndim == 2:
                    un = un[None].expand(B, un.shape[0], nsamples)
                un = self._pad_ylm(un.transpose(1, 2).contiguous(), 2)
            y = torch.empty(B, nsamples, 256, dtype=torch.float64, device=self.device)
            _lib.check(self._lib.spb_sample_ylm(self._ctx.handle, B, nsamples, _ptr(self._mean_ylm),
                                                _ptr(L.contiguous()), _ptr(un.contiguous()), _ptr(y),
                                                _stream()))
        return self._out(self._cut(y, (2,)))

    # ------------------------------------------------------------------ flux operator / design
    def _u(self, u):
        u = torch.as_tensor(defaults["u"] if u is None else u, dtype=torch.float64).reshape(-1)
        if self._udeg == 0:
            return torch.zeros(2, dtype=torch.float64, device=self.device)
        if u.numel() < self._udeg:
            raise ValueError("Vector `u` has the wrong size. Expected %d, got %d." %
                             (self._udeg, u.numel()))
        return u[: self._udeg].to(self.device).contiguous()

    def _rTA1(self, u):
        """Flux operator row vector for limb darkening ``u`` (a13), cached per process object for
        host-resident ``u`` (it only depends on ``u``; keyed by value, so no device sync)."""
        key = None
        if not (isinstance(u, torch.Tensor) and u.is_cuda):
            uh = np.asarray(defaults["u"] if u is None else
                            (u.detach().cpu().numpy() if isinstance(u, torch.Tensor) else u),
                            dtype=np.float64).reshape(-1)
            key = tuple(uh.tolist())
            if key in self._rTA1_cache:
                return self._rTA1_cache[key]
        ud = self._u(u)
        out = torch.empty(256, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.spb_flux_operator(self._ctx.handle, 1, _ptr(ud), _ptr(out),
                                                   _stream()))
        if key is not None:
            self._rTA1_cache[key] = out
        return out

    def _t(self, t):
        return torch.as_tensor(t, dtype=torch.float64).to(self.device).reshape(-1).contiguous()

    _udt_cache = None

    def _uniform_dt(self, t):
        """Spacing of equally spaced time stamps (``t_k = t_0 + k dt`` to 4e-12 of a step, e.g. any
        ``linspace`` / fixed-cadence grid), else 0: the marginal assembly then tabulates the covariance
        interpolant instead of evaluating it per entry (``spb_noise_model.uniform_dt``).  Host arrays are
        checked with NumPy; device tensors once per (storage, length, version) with one reduction."""
        if isinstance(t, torch.Tensor) and t.is_cuda:
            key = (t.data_ptr(), t.numel(), t._version)
            if self._udt_cache is not None and self._udt_cache[0] == key:
                return self._udt_cache[1]
            tt = t.detach().reshape(-1).to(torch.float64)
            n = tt.numel()
            if n < 3:
                return 0.0
            dt = (tt[-1] - tt[0]) / (n - 1)
            dev_ = (tt - (tt[0] + dt * torch.arange(n, dtype=torch.float64, device=tt.device))).abs().max()
            dt, dev_ = float(dt), float(dev_)
            out = dt if (dt > 0.0 and dev_ <= 4e-12 * dt) else 0.0
            self._udt_cache = (key, out)
            return out
        ta = np.asarray(t, dtype=np.float64).reshape(-1)
        n = ta.size
        if n < 3:
            return 0.0
        dt = (ta[-1] - ta[0]) / (n - 1)
        if not dt > 0.0:
            return 0.0
        return float(dt) if np.max(np.abs(ta - (ta[0] + dt * np.arange(n)))) <= 4e-12 * dt else 0.0

    def _inc(self, i):
        if not (isinstance(i, torch.Tensor) and i.is_cuda):
            _check_bounds("i", np.asarray(i, dtype=np.float64) * np.pi / 180, 0, 0.5 * np.pi)
        it = torch.as_tensor(i, dtype=torch.float64).to(self.device).reshape(-1) * (math.pi / 180)
        return it.contiguous()

    def design_matrix(self, t, i=defaults["i"], p=defaults["p"], u=None):
        """flux.py:345-350: ``A (nt, nylm)``, or ``(I, nt, nylm)`` for a vector of inclinations."""
        A = self._cut(self._design_full(t, i, p, u), (2,))
        return A if _is_batched(i) else A[0]

    def _design_full(self, t, i, p, u):
        """The degree-15 design matrix ``(I, nt, 256)`` (a lower degree is its leading block)."""
        t = self._t(t)
        inc = self._inc(i)
        _check_bounds("p", p, 0, np.inf)
        I, nt = inc.numel(), t.numel()
        rta1 = self._rTA1(u)
        with torch.cuda.device(self.device):
            per = torch.full((I,), float(p), dtype=torch.float64, device=self.device)
            A = torch.empty(I, nt, 256, dtype=torch.float64, device=self.device)
            nbytes = self._lib.spb_design_matrix_workspace_bytes(self._ctx.handle, I, nt)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(self._lib.spb_design_matrix(self._ctx.handle, I, nt, _ptr(t), _ptr(inc),
                                                   _ptr(per), _ptr(rta1), 0, _ptr(A), _ptr(ws),
                                                   nbytes, _stream()))
        return A

    # ------------------------------------------------------------------ flux mean / covariance
    def _noise_model(self, nt, data_cov, baseline_var, keep, lower_only=False, b0=0):
        nm = _lib.NoiseModel()
        nm.lower_only = 1 if lower_only else 0
        nm.defer = 0
        nm.normalized = 1 if self._normalized else 0
        nm.normalization_order = self._normN
        nm.normalization_zmax = self._normzmax
        nm.data_cov = None
        nm.baseline_var = None
        nm.data_stride = 0
        nm.base_stride = 0
        nm.temporal_kind = self._tkind
        nm.uniform_dt = float(getattr(self, "_udt", 0.0) or 0.0)
        nm.tau = self._tau[b0:].data_ptr() if self._tkind else None
        nm.tau_stride = 1 if self._tkind else 0
        if data_cov is not None:
            d = torch.as_tensor(data_cov, dtype=torch.float64).to(self.device).contiguous()
            if d.ndim == 0:
                nm.data_kind, d = 0, d.reshape(1)
            elif d.ndim == 1:
                if d.numel() != nt:
                    raise ValueError("data_cov vector has the wrong length")
                nm.data_kind = 1
            else:
                if tuple(d.shape) != (nt, nt):
                    raise ValueError("data_cov matrix has the wrong shape")
                nm.data_kind = 2
            keep.append(d)
            nm.data_cov = d.data_ptr()
        if baseline_var is not None:
            bv = torch.as_tensor(baseline_var, dtype=torch.float64).to(self.device).contiguous()
            if bv.ndim == 0:
                nm.base_kind, bv = 0, bv.reshape(1)
            elif bv.ndim == 1:
                # extension: one baseline variance per batch element (the free `v` of
                # calibrate/log_prob.py:70-74)
                if bv.numel() != self._B:
                    raise ValueError("a 1-D baseline_var must have one entry per batch element")
                nm.base_kind, nm.base_stride = 0, 1
                bv = bv[b0:].contiguous()
            else:
                if tuple(bv.shape) != (nt, nt):
                    raise ValueError("baseline_var must be a scalar or an (nt, nt) matrix")
                nm.base_kind = 2
            keep.append(bv)
            nm.baseline_var = bv.data_ptr()
        return nm

    def _flux_cov_chunk(self, b0, b1, t, inc, p, rta1, marg, data_cov, baseline_var, ldk,
                        lower_only=False, defer=False):
        """GP mean (scalar per element) and the (noise-augmented) covariance for elements b0:b1.
        Returns (gp_mean, K, z), or with ``defer`` (marginal branch, scalar / per-point noise terms)
        (gp_mean, K_raw, z, affine, keepalive): the covariance is left raw and the normalisation +
        noise are applied inside the Cholesky kernel (``spb_cholesky_lnlike_affine``)."""
        lib, h = self._lib, self._ctx.handle
        Bc, nt = b1 - b0, t.numel()
        dev = self.device
        keep = []
        nm = self._noise_model(nt, data_cov, baseline_var, keep, lower_only and marg, b0=b0)
        defer_req = bool(defer)
        defer = bool(defer and marg and lower_only and nm.data_kind != 2 and nm.base_kind != 2)
        nm.defer = 1 if defer else 0
        mean_ylm = self._mean_ylm[b0:b1]
        cov_ylm = self._cov_ylm[b0:b1]
        info = self._info[b0:b1]
        gp_mean = torch.empty(Bc, dtype=torch.float64, device=dev)
        K = torch.empty(Bc, nt, ldk, dtype=torch.float64, device=dev)
        z = torch.zeros(Bc, dtype=torch.float64, device=dev)
        nb_as = lib.spb_assemble_workspace_bytes(h, Bc, nt)
        ws_as = torch.empty(nb_as, dtype=torch.uint8, device=dev)
        if marg:
            nc = self._covpts + 1
            var = torch.empty(Bc, dtype=torch.float64, device=dev)
            coef = torch.empty(Bc, 4, nc, dtype=torch.float64, device=dev)
            nb = lib.spb_flux_marginal_workspace_bytes(h, Bc)
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            ev = self._mark("flux_marginal")
            _lib.check(lib.spb_flux_marginal(h, Bc, _ptr(mean_ylm), _ptr(cov_ylm), _ptr(rta1),
                                             self._covpts, _ptr(gp_mean), _ptr(var), _ptr(coef),
                                             _ptr(ws), nb, _stream()))
            self._mark_end(ev)
            self._last_coef = coef   # (Bc, 4, covpts + 1): predict's cross-covariance re-uses it
            ev = self._mark("assemble")
            _lib.check(lib.spb_assemble_marginal(h, Bc, nt, _ptr(t), float(p), self._covpts,
                                                 _ptr(coef), _ptr(var), _ptr(gp_mean),
                                                 ctypes.byref(nm), _ptr(K), ldk, _ptr(z), _ptr(info),
                                                 _ptr(ws_as), nb_as, _stream()))
            self._mark_end(ev)
        else:
            inc_c = inc if inc.numel() == 1 else inc[b0:b1].contiguous()
            Ic = inc_c.numel()
            per = torch.full((Ic,), float(p), dtype=torch.float64, device=dev)
            A = torch.empty(Ic, nt, 256, dtype=torch.float64, device=dev)
            nbd = lib.spb_design_matrix_workspace_bytes(h, Ic, nt)
            wsd = torch.empty(nbd, dtype=torch.uint8, device=dev)
            _lib.check(lib.spb_design_matrix(h, Ic, nt, _ptr(t), _ptr(inc_c), _ptr(per), _ptr(rta1),
                                             0, _ptr(A), _ptr(wsd), nbd, _stream()))
            nb = lib.spb_flux_conditional_workspace_bytes(h, Bc, nt)
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            # log-likelihood path of an unnormalised, static process: only the lower triangle is needed and
            # the noise terms are added inside the Cholesky kernel's loads (no read-modify-write pass)
            defer_c = bool(defer_req and lower_only and not self._normalized and not self._tkind
                           and nm.data_kind != 2 and nm.base_kind != 2)
            fc = lib.spb_flux_conditional_lower if defer_c else lib.spb_flux_conditional
            _lib.check(fc(h, Bc, nt, _ptr(A), 0 if Ic == 1 else nt * 256, _ptr(mean_ylm), _ptr(cov_ylm),
                          _ptr(gp_mean), _ptr(K), ldk, _ptr(ws), nb, _stream()))
            if defer_c:
                defer = True
            elif self._tkind:   # sp.py:697-698
                _lib.check(lib.spb_temporal_scale(h, Bc, nt, nt, _ptr(t), _ptr(t), self._tkind,
                                                  _ptr(self._tau[b0:b1].contiguous()), 1, None, 0,
                                                  _ptr(K), ldk, nt * ldk, _stream()))
            if not defer_c:
                _lib.check(lib.spb_assemble_conditional(h, Bc, nt, _ptr(gp_mean), ctypes.byref(nm),
                                                        _ptr(K), ldk, _ptr(z), _ptr(info), _ptr(ws_as),
                                                        nb_as, _stream()))
        if defer:
            af = _lib.Affine()
            if self._normalized:
                qp, sp_ = ctypes.c_void_p(), ctypes.c_void_p()
                lib.spb_assemble_workspace_layout(Bc, nt, _ptr(ws_as), ctypes.byref(qp),
                                                  ctypes.byref(sp_))
                af.q, af.scal = qp, sp_
            af.diag = nm.data_cov
            af.diag_kind = nm.data_kind
            af.diag_stride = 0
            af.offset = nm.baseline_var
            af.offset_stride = nm.base_stride
            return gp_mean, K, z, af, (keep, ws_as)
        return gp_mean, K, z

    def _chunks(self, nt, ldk):
        # as few chunks as the memory bound allows (96 GB of covariance matrices by default: the
        # B200 has 180 GB), of equal size: the Cholesky kernel claims matrices dynamically, so one
        # long launch has a shorter tail than several short ones
        per = nt * ldk * 8 + 4 * 256 * 256 * 8
        if self._ctx.i8_planes(nt, self._B):   # digit planes of the INT8 path live next to K
            code = self._ctx.i8_planes(nt, self._B)   # 7 | 8 | 77 | 78 | 87: the first digit is the plane count
            per += (code // 10 if code > 10 else code) * (nt + 64) * (nt + 64)
        # (the assembly / GEMM kernels carry the batch in grid.y: at most 65535 elements a launch)
        step = max(1, min(self._B, self._max_chunk_bytes // per, 65535))
        nchunks = -(-self._B // step)
        step = -(-self._B // nchunks)
        return [(b0, min(self._B, b0 + step)) for b0 in range(0, self._B, step)]

    def _prep(self, t, i, p, u, marginalize_over_inclination):
        marg = self._marginalize_over_inclination if marginalize_over_inclination is None \
            else bool(marginalize_over_inclination)
        self._compute_moments()
        self._udt = self._uniform_dt(t) if os.environ.get("SPB200_NO_UNIFORM_T") is None else 0.0
        t = self._t(t)
        inc = self._inc(i)
        if inc.numel() not in (1, self._B):
            raise ValueError("`i` must be a scalar or have one entry per batch element")
        _check_bounds("p", p, 0, np.inf)
        rta1 = self._rTA1(u)
        return marg, t, inc, rta1

    def mean(self, t, i=defaults["i"], p=defaults["p"], u=None, marginalize_over_inclination=None):
        """sp.py:643-672."""
        marg, t, inc, rta1 = self._prep(t, i, p, u, marginalize_over_inclination)
        nt = t.numel()
        if self._normalized:
            out = torch.zeros(self._B, nt, dtype=torch.float64, device=self.device)
            return self._out(out)
        with torch.cuda.device(self.device):
            gm = []
            for b0, b1 in self._chunks(nt, nt + (nt & 1)):
                g, _, _ = self._flux_cov_chunk(b0, b1, t, inc, p, rta1, marg, None, None,
                                               nt + (nt & 1))
                gm.append(g)
            g = torch.cat(gm)
        return self._out(g[:, None].expand(self._B, nt).contiguous())

    def cov(self, t, i=defaults["i"], p=defaults["p"], u=None, marginalize_over_inclination=None):
        """sp.py:674-703: the (normalised, if requested) GP flux covariance, ``(nt, nt)`` or
        ``(B, nt, nt)``."""
        marg, t, inc, rta1 = self._prep(t, i, p, u, marginalize_over_inclination)
        nt = t.numel()
        ldk = nt + (nt & 1)
        outs, zs = [], []
        with torch.cuda.device(self.device):
            for b0, b1 in self._chunks(nt, ldk):
                _, K, z = self._flux_cov_chunk(b0, b1, t, inc, p, rta1, marg, None, None, ldk)
                outs.append(K[:, :, :nt])
                zs.append(z)
        self._z = torch.cat(zs)
        K = torch.cat(outs) if len(outs) > 1 else outs[0]
        return self._out(K)

    # ------------------------------------------------------------------ log likelihood
    def log_likelihood(self, t, flux, data_cov, i=defaults["i"], p=defaults["p"], u=None,
                       baseline_mean=defaults["baseline_mean"],
                       baseline_var=defaults["baseline_var"], marginalize_over_inclination=None,
                       return_grad=False):
        """sp.py:1052-1188.  ``flux`` is ``(nt,)`` or ``(M, nt)`` (M light curves sharing period,
        limb darkening, inclination and noise, scored jointly); returns a scalar, or ``(B,)`` for a
        batch of hyperparameter samples.  Extension: ``flux`` of shape ``(B, M, nt)`` gives every
        batch element its own light curve(s) (the light-curve x sample x inclination batches of
        calibrate/inclination.py:63-74)."""
        if return_grad:
            return self.log_likelihood_and_grad(
                t, flux, data_cov, i=i, p=p, u=u, baseline_mean=baseline_mean,
                baseline_var=baseline_var, marginalize_over_inclination=marginalize_over_inclination)
        marg, t, inc, rta1 = self._prep(t, i, p, u, marginalize_over_inclination)
        nt = t.numel()
        ldk = nt + (nt & 1)
        dev = self.device
        f = torch.as_tensor(flux, dtype=torch.float64).to(dev)
        if f.ndim == 1:
            f = f[None]
        per_element = f.ndim == 3
        if per_element and f.shape[0] != self._B:
            raise ValueError("per-element flux must have shape (B, M, nt)")
        if f.shape[-1] != nt:
            raise ValueError("flux must have shape (nt,), (M, nt) or (B, M, nt)")
        M = f.shape[-2]
        bm = torch.as_tensor(baseline_mean, dtype=torch.float64).to(dev)
        if bm.ndim == 1 and bm.numel() != self._B:
            raise ValueError("a 1-D baseline_mean must have one entry per batch element")
        bvar = None
        if isinstance(baseline_var, (torch.Tensor, np.ndarray)) or baseline_var != 0.0:
            bvar = baseline_var
        lnlike = torch.empty(self._B, dtype=torch.float64, device=dev)
        zs = []
        lib, h = self._lib, self._ctx.handle
        # white-noise floor of K (a lower bound of its smallest eigenvalue): scales the right-hand-side
        # digit planes of the INT8 path when the noise was already added by the assembly kernels
        lam_min, noise_free = 0.0, False
        if isinstance(data_cov, (int, float)) or (
                isinstance(data_cov, (torch.Tensor, np.ndarray)) and data_cov.ndim == 0):
            lam_min = float(data_cov)
            noise_free = not (lam_min > 0.0)   # no white-noise floor: nothing bounds the solved rows
        with torch.cuda.device(dev):
            for b0, b1 in self._chunks(nt, ldk):
                Bc = b1 - b0
                out = self._flux_cov_chunk(b0, b1, t, inc, p, rta1, marg, data_cov, bvar, ldk,
                                           lower_only=True, defer=True)
                gp_mean, K, z = out[:3]
                affine = out[3] if len(out) > 3 else None
                zs.append(z)
                # r = flux - (gp_mean + baseline_mean)  (sp.py:1157-1161); normalised: mean == 0
                resid = torch.zeros(Bc, M, ldk, dtype=torch.float64, device=dev)
                fc = f[b0:b1] if per_element else f[None]
                bmc = bm[b0:b1, None, None] if bm.ndim == 1 else bm   # (B,) baseline means
                if self._normalized:
                    resid[:, :, :nt] = fc - bmc
                else:
                    resid[:, :, :nt] = fc - (gp_mean[:, None, None] + bmc)
                ev = self._mark("cholesky")
                if Bc == 1 and M >= 64:
                    # one factorisation + many right-hand sides (ensemble of light curves sharing
                    # K): factor on one CTA, then spread the RHS rows over the whole GPU
                    logdet = torch.empty(1, dtype=torch.float64, device=dev)
                    quad = torch.empty(M, dtype=torch.float64, device=dev)
                    if affine is not None:
                        _lib.check(lib.spb_cholesky_lnlike_affine(
                            h, 1, nt, _ptr(K), ldk, nt * ldk, ctypes.byref(affine), 0, None, ldk, 0,
                            None, None, _ptr(logdet), _ptr(self._info[b0:b1]), _stream()))
                    else:
                        _lib.check(lib.spb_cholesky_lnlike(
                            h, 1, nt, _ptr(K), ldk, nt * ldk, 0, None, ldk, 0, None, None,
                            _ptr(logdet), _ptr(self._info[b0:b1]), _stream()))
                    _lib.check(lib.spb_cholesky_solve_rows(h, nt, _ptr(K), ldk, M, _ptr(resid), ldk,
                                                           _ptr(quad), _stream()))
                    ll = -0.5 * quad.sum() - M * logdet[0] - 0.5 * nt * M * math.log(2 * math.pi)
                    flagged = (self._info[b0:b1] != 0) | torch.isnan(ll)
                    lnlike[b0:b1] = torch.where(flagged, torch.full_like(ll, -float("inf")), ll)
                elif self._ctx.i8_planes(nt, Bc) and nt > 64 and not noise_free and (
                        (affine is not None and affine.diag) or lam_min > 0.0):
                    # factorisation on the INT8 tensor cores (K is only read; potrf_i8.cuh)
                    planes = self._ctx.i8_planes(nt, Bc)
                    nb_i8 = lib.spb_cholesky_i8_workspace_bytes(Bc, nt, M, planes)
                    ws_i8 = torch.empty(nb_i8, dtype=torch.uint8, device=dev)
                    _lib.check(lib.spb_cholesky_lnlike_i8(
                        h, Bc, nt, _ptr(K), ldk, nt * ldk,
                        ctypes.byref(affine) if affine is not None else None, M, _ptr(resid), ldk,
                        M * ldk, _ptr(lnlike[b0:b1]), None, None, _ptr(self._info[b0:b1]), planes,
                        0.0 if affine is not None and affine.diag else lam_min, _ptr(ws_i8), nb_i8,
                        _stream()))
                    del ws_i8
                elif affine is not None:
                    _lib.check(lib.spb_cholesky_lnlike_affine(
                        h, Bc, nt, _ptr(K), ldk, nt * ldk, ctypes.byref(affine), M, _ptr(resid), ldk,
                        M * ldk, _ptr(lnlike[b0:b1]), None, None, _ptr(self._info[b0:b1]),
                        _stream()))
                else:
                    _lib.check(lib.spb_cholesky_lnlike(
                        h, Bc, nt, _ptr(K), ldk, nt * ldk, M, _ptr(resid), ldk, M * ldk,
                        _ptr(lnlike[b0:b1]), None, None, _ptr(self._info[b0:b1]), _stream()))
                self._mark_end(ev)
                del K, resid, out
        self._z = torch.cat(zs)
        return self._out(lnlike)

    # ------------------------------------------------------------------ gradient
    def log_likelihood_and_grad(self, t, flux, data_cov, i=defaults["i"], p=defaults["p"], u=None,
                                baseline_mean=defaults["baseline_mean"],
                                baseline_var=defaults["baseline_var"],
                                marginalize_over_inclination=None, rel_step=1e-4):
        """``log_likelihood`` and its gradient with respect to the hyperparameters the process was
        built from -- ``r`` [deg], ``a`` and ``b`` (or ``mu`` and ``sigma`` [deg]), ``c``, ``n`` -- as
        ``(lnlike, {"r": ..., "a" | "mu": ..., "b" | "sigma": ..., "c": ..., "n": ...})``, one value
        per batch element.  What the reference obtains by Theano reverse mode (``tt.grad`` of
        sp.py:1052-1188 through ops/include/latitude.h:22-173, eigh.h:19-65, wigner.h:345-404,
        465-531, math.py:40-72; checked there with ``verify_grad`` at 1e-4, tests/test_lnlike.py:105-136).

        Here: the Ylm moments move along their ANALYTIC tangents (``spb_ylm_moments_grad``: derivative
        lanes of the Beta moments, profile derivative, monomial scalings in c and n; exact because
        the rest of the moment pipeline is linear / quadratic), and the log-likelihood of the ten
        displaced moment sets is evaluated by the same kernels in ONE batch of 11 B elements; the
        central difference along each tangent has a relative truncation error of ``rel_step**2``
        (the flux-side likelihood is smooth in the moments) and no eigen-solver noise, because the
        displaced sets share the base eigen-problem up to the smooth perturbation.  Measured against
        the analytic oracle (oracle/sp_oracle_grad.py): <= 1e-6 relative."""
        if self._dr is not None or self._time_variable or not hasattr(self, "_r") or self._r is None:
            raise NotImplementedError("gradients: delta spot-size prior, static surfaces, a single "
                                      "process (not a sum)")
        dev, B = self.device, self._B
        lib, h = self._lib, self._ctx.handle
        with torch.cuda.device(dev):
            mean = torch.empty(11 * B, 256, dtype=torch.float64, device=dev)
            cov = torch.empty(11 * B, 256, 256, dtype=torch.float64, device=dev)
            eps = torch.empty(5, B, dtype=torch.float64, device=dev)
            info = torch.zeros(B, dtype=torch.int32, device=dev)
            nb = lib.spb_ylm_moments_grad_workspace_bytes(h, B)
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            ev = self._mark("moments_grad")
            _lib.check(lib.spb_ylm_moments_grad(h, B, _ptr(self._r), _ptr(self._a), _ptr(self._b),
                                                _ptr(self._c), _ptr(self._n), self._optp(),
                                                float(rel_step),
                                                _ptr(mean), _ptr(cov), _ptr(eps), _ptr(info), _ptr(ws),
                                                nb, _stream()))
            self._mark_end(ev)
            del ws
            # an 11 B-wide process over the displaced moments; everything else is shared
            big = object.__new__(StarryProcess)
            big.__dict__.update(self.__dict__)
            big._B, big._batched = 11 * B, True
            big._mean_ylm, big._cov_ylm, big._cho_cov_ylm = mean, cov, None
            big._info = info.repeat(11).contiguous()
            big._rTA1_cache = self._rTA1_cache

            def tile(x, per_dim):
                """Per-element arguments (leading dimension B) are repeated for the 11 variants."""
                if isinstance(x, (torch.Tensor, np.ndarray)):
                    xt = torch.as_tensor(x, dtype=torch.float64).to(dev)
                    if xt.ndim == per_dim and xt.shape[0] == B and B > 1:
                        return xt.repeat((11,) + (1,) * (xt.ndim - 1))
                    return xt
                return x

            ll_all = big.log_likelihood(
                t, tile(flux, 3), data_cov, i=tile(i, 1), p=p, u=u, baseline_mean=tile(baseline_mean, 1),
                baseline_var=tile(baseline_var, 1),
                marginalize_over_inclination=marginalize_over_inclination).reshape(11, B)
            self._z = None if big._z is None else big._z[:B]
            self._info = big._info[:B].clone()
            ll = ll_all[0]
            g = (ll_all[1::2] - ll_all[2::2]) / (2.0 * eps)          # (5, B): r, a, b, c, n
            ok = torch.isfinite(ll)[None, :] & torch.isfinite(ll_all[1:]).reshape(5, 2, B).all(dim=1)
            g = torch.where(torch.isfinite(ll)[None, :].expand(5, B), g, torch.zeros_like(g))
            g = torch.where(ok | ~torch.isfinite(ll)[None, :], g, torch.full_like(g, float("nan")))
            grad = {"r": self._out(g[0]), "c": self._out(g[3]), "n": self._out(g[4])}
            if getattr(self, "_mu_sigma", None) is None:
                grad["a"], grad["b"] = self._out(g[1]), self._out(g[2])
            else:
                # (a, b) = gauss2beta(mu, sigma), latitude.py:14-77: elementwise 2 x 2 Jacobian
                mu = self._mu_sigma[0].detach().clone().requires_grad_(True)
                sg = self._mu_sigma[1].detach().clone().requires_grad_(True)
                a_, b_ = gauss2beta(mu, sg)
                da = torch.autograd.grad(a_.sum(), (mu, sg), retain_graph=True)
                db = torch.autograd.grad(b_.sum(), (mu, sg))
                grad["mu"] = self._out(g[1] * da[0] + g[2] * db[0])
                grad["sigma"] = self._out(g[1] * da[1] + g[2] * db[1])
        return self._out(ll), grad

    # ------------------------------------------------------------------ SURVEY 8(f) rank 2
    def _factor_rows(self, K, n, ldk, rows=None, diag_add=None):
        """In-place batched Cholesky of ``K (Bq, n, ldk)`` (lower triangle read, ``L`` written over
        it; ``diag_add`` is added to the diagonal inside the kernel's loads); ``rows (Bq, M, ldr)``
        are replaced by ``L^-1 row`` (math.py:20-38 forward substitution).  Returns ``info``."""
        lib, h = self._lib, self._ctx.handle
        Bq = K.shape[0]
        info = torch.zeros(Bq, dtype=torch.int32, device=self.device)
        if rows is None:
            M, rp, ldr, rstride = 0, None, ldk, 0
        else:
            M, ldr = rows.shape[1], rows.shape[2]
            rp, rstride = _ptr(rows), M * ldr
        if diag_add is None:
            _lib.check(lib.spb_cholesky_lnlike(h, Bq, n, _ptr(K), ldk, n * ldk, M, rp, ldr, rstride,
                                               None, None, None, _ptr(info), _stream()))
        else:
            af = _lib.Affine()
            dg = torch.full((1,), float(diag_add), dtype=torch.float64, device=self.device)
            af.diag, af.diag_kind, af.diag_stride = dg.data_ptr(), 0, 0
            _lib.check(lib.spb_cholesky_lnlike_affine(h, Bq, n, _ptr(K), ldk, n * ldk,
                                                      ctypes.byref(af), M, rp, ldr, rstride, None,
                                                      None, None, _ptr(info), _stream()))
        return info

    def _gemm(self, batch, M, N, K, A, lda, sA, Bm, ldb, sB, C, ldc, sC, alpha=1.0, beta=0.0):
        """``C = alpha A Bm^T + beta C`` (spb_gemm_nt); tensors or raw device addresses."""
        ad = lambda x: ctypes.c_void_p(x) if isinstance(x, int) else _ptr(x)   # noqa: E731
        _lib.check(self._lib.spb_gemm_nt(self._ctx.handle, batch, M, N, K, float(alpha), ad(A), lda,
                                         sA, ad(Bm), ldb, sB, float(beta), ad(C), ldc, sC,
                                         _stream()))

    def _draw(self, mu, L, n, ld, nsamples, unit_normals, generator):
        """``mu[:, None] + L u`` for ``u ~ N(0, 1)`` of shape ``(n, nsamples)`` (or ``(B, n,
        nsamples)``), returned as ``(B, nsamples, n)``.  ``L (B, n, ld)`` must be clean lower."""
        B, dev = L.shape[0], self.device
        if unit_normals is None:
            un = torch.zeros(B, nsamples, ld, dtype=torch.float64, device=dev)
            un[:, :, :n] = torch.randn(B, nsamples, n, dtype=torch.float64, device=dev,
                                       generator=generator)
        else:
            u_ = torch.as_tensor(unit_normals, dtype=torch.float64).to(dev)
            if u_.ndim == 2:
                u_ = u_[None].expand(B, u_.shape[0], u_.shape[1])
            if u_.shape[1] != n:
                raise ValueError("unit normals must have shape (%d, nsamples)" % n)
            nsamples = u_.shape[2]
            un = torch.zeros(B, nsamples, ld, dtype=torch.float64, device=dev)
            un[:, :, :n] = u_.transpose(1, 2)
        out = mu[:, None, :].expand(B, nsamples, n).contiguous()
        self._gemm(B, nsamples, n, ld, un, ld, nsamples * ld, L, ld, n * ld, out, n, nsamples * n,
                   alpha=1.0, beta=1.0)
        return out

    # ---- batch chunking of the methods that materialise several (nt x nt) arrays per element
    def _sub(self, b0, b1):
        """A view of batch elements b0:b1 as a process of its own (moments already computed)."""
        sub = object.__new__(type(self))
        sub.__dict__.update(self.__dict__)
        sub._B, sub._batched = b1 - b0, True
        for k in ("_r", "_a", "_b", "_c", "_n", "_tau", "_dr", "_mean_ylm", "_cov_ylm", "_info"):
            v = getattr(self, k, None)
            setattr(sub, k, None if v is None else v[b0:b1])
        sub._cho_cov_ylm = None if self._cho_cov_ylm is None else self._cho_cov_ylm[b0:b1]
        if getattr(self, "_mu_sigma", None) is not None:
            sub._mu_sigma = (self._mu_sigma[0][b0:b1], self._mu_sigma[1][b0:b1])
        return sub

    def _chunked(self, method, n1, n2, kwargs,
                 per_element=("i", "unit_normals", "baseline_mean", "baseline_var")):
        """Runs ``method`` over batch chunks when the whole batch would not fit ``max_chunk_bytes``
        (about four (n1 x n2) arrays per element) or one launch (65535 elements: the assembly and
        GEMM kernels carry the batch in grid.y).  Returns None when no chunking is needed."""
        per = 4 * n1 * (n2 + (n2 & 1)) * 8 + 4 * 256 * 256 * 8
        step = max(1, min(self._B, self._max_chunk_bytes // per, 65535))
        if step >= self._B:
            return None
        self._compute_moments()
        outs = []
        for b0 in range(0, self._B, step):
            b1 = min(self._B, b0 + step)
            kw = dict(kwargs)
            for k in per_element:
                v = kw.get(k, None)
                if isinstance(v, (torch.Tensor, np.ndarray)) and v.ndim >= 1 and v.shape[0] == self._B \
                        and v.ndim == (3 if k == "unit_normals" else 1):
                    kw[k] = v[b0:b1]
            outs.append(getattr(self._sub(b0, b1), method)(**kw))
        if isinstance(outs[0], tuple):
            return tuple(torch.cat([o[k] for o in outs]) for k in range(len(outs[0])))
        return torch.cat(outs)

    def sample(self, t, i=defaults["i"], p=defaults["p"], u=None, nsamples=1, eps=defaults["eps"],
               unit_normals=None, generator=None, marginalize_over_inclination=None):
        """sp.py:729-765: draws from the prior over light curves, ``(nsamples, nt)`` (or
        ``(B, nsamples, nt)``).  ``unit_normals`` optionally supplies the reference's
        ``random_normal(self.random, (nt, nsamples))``."""
        nt_ = int(np.size(t)) if not isinstance(t, torch.Tensor) else t.numel()
        out = self._chunked("sample", nt_, nt_, dict(
            t=t, i=i, p=p, u=u, nsamples=nsamples, eps=eps, unit_normals=unit_normals,
            generator=generator, marginalize_over_inclination=marginalize_over_inclination))
        if out is not None:
            return out
        marg, t, inc, rta1 = self._prep(t, i, p, u, marginalize_over_inclination)
        nt = t.numel()
        ldk = nt + (nt & 1)
        B = self._B
        with torch.cuda.device(self.device):
            # eps I is the scalar "data covariance" of the assembly (added after the normalisation,
            # as sp.py:762 adds it to the output of cov())
            gp_mean, K, _ = self._flux_cov_chunk(0, B, t, inc, p, rta1, marg, float(eps), None, ldk)
            info = self._factor_rows(K, nt, ldk)
            _lib.check(self._lib.spb_tril(self._ctx.handle, B, nt, _ptr(K), ldk, nt * ldk,
                                          _stream()))
            if nt & 1:
                K[:, :, nt:] = 0.0
            mu = torch.zeros(B, nt, dtype=torch.float64, device=self.device)
            if not self._normalized:
                mu += gp_mean[:, None]
            out = self._draw(mu, K, nt, ldk, nsamples, unit_normals, generator)
            out = torch.where(((info & 1) != 0)[:, None, None], torch.full_like(out, float("nan")),
                              out)
        return self._out(out)

    def predict(self, t, flux, data_cov, t_sample=None, i=defaults["i"], p=defaults["p"], u=None,
                baseline_mean=defaults["baseline_mean"], baseline_var=defaults["baseline_var"],
                marginalize_over_inclination=None):
        """sp.py:767-922: mean ``(nts,)`` and covariance ``(nts, nts)`` of the light-curve
        distribution conditioned on ``flux`` (leading ``B`` axis for batched hyperparameters).

        One augmented factorisation does the work: the rows ``[K(ts, t); y - mean]`` are appended
        to ``K(t, t)``, the Cholesky kernel returns ``V = (L^-1 K(t, ts))^T`` and ``w = L^-1 (y -
        mean)``, and ``mu = mean + V w``, ``K = K(ts, ts) - V V^T`` are two tensor-core GEMMs."""
        if self._normalized:
            raise NotImplementedError("Method not implemented when the flux is normalized.")
        nt_ = int(np.size(t)) if not isinstance(t, torch.Tensor) else t.numel()
        ns_ = nt_ if t_sample is None else (
            int(np.size(t_sample)) if not isinstance(t_sample, torch.Tensor) else t_sample.numel())
        out = self._chunked("predict", max(nt_, ns_), max(nt_, ns_), dict(
            t=t, flux=flux, data_cov=data_cov, t_sample=t_sample, i=i, p=p, u=u,
            baseline_mean=baseline_mean, baseline_var=baseline_var,
            marginalize_over_inclination=marginalize_over_inclination))
        if out is not None:
            return out
        marg, t, inc, rta1 = self._prep(t, i, p, u, marginalize_over_inclination)
        dev, B = self.device, self._B
        nt = t.numel()
        ldk = nt + (nt & 1)
        ts = t if t_sample is None else self._t(t_sample)
        nts = ts.numel()
        lds = nts + (nts & 1)
        bvar = None
        if isinstance(baseline_var, (torch.Tensor, np.ndarray)) or baseline_var != 0.0:
            bvar = torch.as_tensor(baseline_var, dtype=torch.float64).to(dev)
        # scalar (or one value per batch element) vs. a full (nt, nt) matrix of extra covariance
        scalar_bvar = bvar is None or bvar.ndim <= 1
        f = torch.as_tensor(flux, dtype=torch.float64).to(dev).reshape(-1)
        if f.numel() != nt:
            raise ValueError("flux must have shape (nt,)")
        bm = torch.as_tensor(baseline_mean, dtype=torch.float64).to(dev)
        lib, h = self._lib, self._ctx.handle
        with torch.cuda.device(dev):
            # K(ts, ts) + baseline_var first (the marginal kernel table is shared by both calls)
            _, Kss, _ = self._flux_cov_chunk(0, B, ts, inc, p, rta1, marg, None,
                                             bvar if (scalar_bvar or nts == nt) else None, lds)
            gp_mean, Ktt, _ = self._flux_cov_chunk(0, B, t, inc, p, rta1, marg, data_cov, bvar, ldk)
            # right-hand sides: rows 0..nts-1 = K(ts, t) + baseline_var, row nts = y - mean
            R = torch.zeros(B, nts + 1, ldk, dtype=torch.float64, device=dev)
            rstride = (nts + 1) * ldk
            off = None
            if bvar is not None and scalar_bvar:
                off = bvar.reshape(-1).contiguous()
            off_s = 0 if off is None or off.numel() == 1 else 1
            late_off = off if self._tkind else None   # K(ts,t) *= k_temporal BEFORE + baseline_var
            if self._tkind:
                off = None
            if marg:
                _lib.check(lib.spb_cross_marginal(
                    h, B, nts, nt, _ptr(ts), _ptr(t), float(p), self._covpts,
                    _ptr(self._last_coef), _ptr(off), off_s, _ptr(R), ldk, rstride, _stream()))
            else:
                I = inc.numel()
                per = torch.full((I,), float(p), dtype=torch.float64, device=dev)
                A2 = []
                for tt_, n_ in ((ts, nts), (t, nt)):
                    A_ = torch.empty(I, n_, 256, dtype=torch.float64, device=dev)
                    nbd = lib.spb_design_matrix_workspace_bytes(h, I, n_)
                    wsd = torch.empty(nbd, dtype=torch.uint8, device=dev)
                    _lib.check(lib.spb_design_matrix(h, I, n_, _ptr(tt_), _ptr(inc), _ptr(per),
                                                     _ptr(rta1), 0, _ptr(A_), _ptr(wsd), nbd,
                                                     _stream()))
                    A2.append(A_)
                A_ts, A_t = A2
                T = torch.empty(B, nts, 256, dtype=torch.float64, device=dev)
                # T = A_ts Sigma (Sigma symmetric), K(ts, t) = T A_t^T   (sp.py:903-906)
                self._gemm(B, nts, 256, 256, A_ts, 256, 0 if I == 1 else nts * 256, self._cov_ylm,
                           256, 65536, T, 256, nts * 256)
                if off is not None:
                    R[:, :nts, :nt] = off.reshape(-1, 1, 1)
                self._gemm(B, nts, nt, 256, T, 256, nts * 256, A_t, 256, 0 if I == 1 else nt * 256,
                           R, ldk, rstride, alpha=1.0, beta=1.0)
            if self._tkind:   # sp.py:893-895
                _lib.check(lib.spb_temporal_scale(h, B, nts, nt, _ptr(ts), _ptr(t), self._tkind,
                                                  _ptr(self._tau), 1, _ptr(late_off), off_s,
                                                  _ptr(R), ldk, rstride, _stream()))
            if bvar is not None and not scalar_bvar:
                R[:, :nts, :nt] += bvar
            R[:, nts, :nt] = (f - bm)[None, :] - gp_mean[:, None]
            info = self._factor_rows(Ktt, nt, ldk, rows=R)
            # mu = mean + V w ; K = K(ts, ts) - V V^T
            mu = gp_mean[:, None].expand(B, nts).contiguous()
            wrow = R.data_ptr() + nts * ldk * 8
            self._gemm(B, nts, 1, ldk, R, ldk, rstride, wrow, ldk, rstride, mu, 1, nts,
                       alpha=1.0, beta=1.0)
            self._gemm(B, nts, nts, ldk, R, ldk, rstride, R, ldk, rstride, Kss, lds, nts * lds,
                       alpha=-1.0, beta=1.0)
            bad = ((info & 1) != 0)
            mu = torch.where(bad[:, None], torch.full_like(mu, float("nan")), mu)
            Kc = Kss[:, :, :nts]
            Kc = torch.where(bad[:, None, None], torch.full_like(Kc, float("nan")), Kc)
        return self._out(mu), self._out(Kc)

    def sample_conditional(self, t, flux, data_cov, t_sample=None, i=defaults["i"],
                           p=defaults["p"], u=None, baseline_mean=defaults["baseline_mean"],
                           baseline_var=defaults["baseline_var"], nsamples=1, eps=defaults["eps"],
                           unit_normals=None, generator=None, marginalize_over_inclination=None):
        """sp.py:924-1002: draws ``(nsamples, nts)`` from the conditional light-curve distribution.
        (The reference body reads an undefined name ``ts``; the intended ``ts = t_sample or t`` is
        what is implemented.)"""
        nt_ = int(np.size(t)) if not isinstance(t, torch.Tensor) else t.numel()
        ns_ = nt_ if t_sample is None else (
            int(np.size(t_sample)) if not isinstance(t_sample, torch.Tensor) else t_sample.numel())
        out = self._chunked("sample_conditional", max(nt_, ns_), max(nt_, ns_), dict(
            t=t, flux=flux, data_cov=data_cov, t_sample=t_sample, i=i, p=p, u=u,
            baseline_mean=baseline_mean, baseline_var=baseline_var, nsamples=nsamples, eps=eps,
            unit_normals=unit_normals, generator=generator,
            marginalize_over_inclination=marginalize_over_inclination))
        if out is not None:
            return out
        mu, K = self.predict(t, flux, data_cov, t_sample=t_sample, i=i, p=p, u=u,
                             baseline_mean=baseline_mean, baseline_var=baseline_var,
                             marginalize_over_inclination=marginalize_over_inclination)
        if not self._batched:
            mu, K = mu[None], K[None]
        B, nts = mu.shape
        lds = nts + (nts & 1)
        with torch.cuda.device(self.device):
            Kp = torch.zeros(B, nts, lds, dtype=torch.float64, device=self.device)
            Kp[:, :, :nts] = K
            info = self._factor_rows(Kp, nts, lds, diag_add=eps)
            _lib.check(self._lib.spb_tril(self._ctx.handle, B, nts, _ptr(Kp), lds, nts * lds,
                                          _stream()))
            out = self._draw(mu.contiguous(), Kp, nts, lds, nsamples, unit_normals, generator)
            out = torch.where(((info & 1) != 0)[:, None, None], torch.full_like(out, float("nan")),
                              out)
        return self._out(out)

    def sample_ylm_conditional(self, t, flux, data_cov, i=defaults["i"], p=defaults["p"], u=None,
                               baseline_mean=defaults["baseline_mean"],
                               baseline_var=defaults["baseline_var"], nsamples=1,
                               unit_normals=None, generator=None):
        """sp.py:518-641: draws ``(nsamples, 256)`` of the Ylm coefficients conditioned on the
        observed flux (the inclination is always used).

        With ``C = data_cov + baseline_var = L_C L_C^T``, ``G = (L_C^-1 A)^T`` and ``Y = L_y^-T``
        (``cov_ylm = L_y L_y^T``):  ``W = G G^T + Y Y^T``,  ``rhs = G L_C^-1 (f - b) + Y L_y^-1 mu``,
        ``ymu = W^-1 rhs``, ``ycov = W^-1 = Y_W Y_W^T`` -- every product a tensor-core GEMM, every
        inverse a forward substitution on appended rows of the Cholesky kernel."""
        if self._normalized:
            raise NotImplementedError("Method not implemented when the flux is normalized.")
        if self._time_variable:
            raise NotImplementedError("Method not implemented for time-variable maps.")
        if self._nylm != 256:
            raise NotImplementedError("sample_ylm_conditional is implemented at ydeg = 15")
        nt_ = int(np.size(t)) if not isinstance(t, torch.Tensor) else t.numel()
        out = self._chunked("sample_ylm_conditional", nt_, 256, dict(
            t=t, flux=flux, data_cov=data_cov, i=i, p=p, u=u, baseline_mean=baseline_mean,
            baseline_var=baseline_var, nsamples=nsamples, unit_normals=unit_normals,
            generator=generator))
        if out is not None:
            return out
        self._compute_moments()
        dev, B = self.device, self._B
        t = self._t(t)
        nt = t.numel()
        ldk = nt + (nt & 1)
        f = torch.as_tensor(flux, dtype=torch.float64).to(dev).reshape(-1)
        if f.numel() != nt:
            raise ValueError("flux must have shape (nt,)")
        bm = torch.as_tensor(baseline_mean, dtype=torch.float64).to(dev)
        lib, h = self._lib, self._ctx.handle
        with torch.cuda.device(dev):
            A = self._design_full(t, i, p, u)
            I = A.shape[0]
            if I not in (1, B):
                raise ValueError("`i` must be a scalar or have one entry per batch element")
            # C = data_cov (+ baseline_var on every entry): input assembly, sp.py:594-607
            d = torch.as_tensor(data_cov, dtype=torch.float64).to(dev)
            C = torch.zeros(1, nt, ldk, dtype=torch.float64, device=dev)
            if d.ndim == 0:
                C[0, :, :nt].diagonal().fill_(float(d))
            elif d.ndim == 1:
                C[0, :, :nt].diagonal().copy_(d)
            else:
                C[0, :, :nt] = d
            C[0, :, :nt] += torch.as_tensor(baseline_var, dtype=torch.float64).to(dev)
            infoC = self._factor_rows(C, nt, ldk)
            # rows: A^T (256 per inclination) and the residual, all against the one factor of C
            R = torch.zeros(I * 256 + 1, ldk, dtype=torch.float64, device=dev)
            R[: I * 256, :nt] = A.transpose(1, 2).reshape(I * 256, nt)
            R[I * 256, :nt] = f - bm
            quad = torch.empty(I * 256 + 1, dtype=torch.float64, device=dev)
            _lib.check(lib.spb_cholesky_solve_rows(h, nt, _ptr(C), ldk, I * 256 + 1, _ptr(R), ldk,
                                                   _ptr(quad), _stream()))
            grow = R.data_ptr() + I * 256 * ldk * 8
            # cov_ylm = L_y L_y^T with rows [I; mu] appended: Y = L_y^-T (as rows), wmu = L_y^-1 mu
            Ly = self._cov_ylm.clone()
            Ry = torch.zeros(B, 257, 256, dtype=torch.float64, device=dev)
            Ry[:, :256, :] = torch.eye(256, dtype=torch.float64, device=dev)
            Ry[:, 256, :] = self._mean_ylm
            infoY = self._factor_rows(Ly, 256, 256, rows=Ry)
            sY = 257 * 256
            W = torch.empty(B, 256, 256, dtype=torch.float64, device=dev)
            sG = 0 if I == 1 else 256 * ldk
            self._gemm(B, 256, 256, ldk, R, ldk, sG, R, ldk, sG, W, 256, 65536)
            self._gemm(B, 256, 256, 256, Ry, 256, sY, Ry, 256, sY, W, 256, 65536, beta=1.0)
            # rhs = Y wmu + G g, appended (with the identity) to the factorisation of W
            Rw = torch.zeros(B, 257, 256, dtype=torch.float64, device=dev)
            Rw[:, :256, :] = torch.eye(256, dtype=torch.float64, device=dev)
            rhs = Rw.data_ptr() + 256 * 256 * 8
            self._gemm(B, 256, 1, 256, Ry, 256, sY, Ry.data_ptr() + 256 * 256 * 8, 256, sY, rhs, 1, sY)
            self._gemm(B, 256, 1, ldk, R, ldk, sG, grow, ldk, 0, rhs, 1, sY, beta=1.0)
            infoW = self._factor_rows(W, 256, 256, rows=Rw)
            # ycov = Y_W Y_W^T, ymu = Y_W (L_W^-1 rhs)
            ycov = torch.empty(B, 256, 256, dtype=torch.float64, device=dev)
            self._gemm(B, 256, 256, 256, Rw, 256, sY, Rw, 256, sY, ycov, 256, 65536)
            ymu = torch.empty(B, 256, dtype=torch.float64, device=dev)
            self._gemm(B, 256, 1, 256, Rw, 256, sY, rhs, 256, sY, ymu, 1, 256)
            infoS = self._factor_rows(ycov, 256, 256)
            _lib.check(lib.spb_tril(h, B, 256, _ptr(ycov), 256, 65536, _stream()))
            out = self._draw(ymu, ycov, 256, 256, nsamples, unit_normals, generator)
            bad = (((infoY | infoW | infoS) & 1) != 0) | ((infoC & 1) != 0)
            out = torch.where(bad[:, None, None], torch.full_like(out, float("nan")), out)
        return self._out(out)

    # ------------------------------------------------------------------ SURVEY 8(f) rank 3
    @property
    def tau(self):
        """sp.py:337-340."""
        return None if self._tau is None else self._out(self._tau)

    @property
    def temporal_kernel(self):
        """sp.py:342-345."""
        return getattr(self, "_temporal_kernel", None)

    def _sample_ylm_temporal(self, t, nsamples, u, generator):
        """sp.py:510-516 + ops/sample.py:24-33 (a triple Python loop in the reference):
        ``y_i = L_t U_i L_y^T`` with ``L_t = cho_factor(temporal_kernel(t, t, tau))`` -- two
        tensor-core GEMMs per batch element.  As in the reference, ``mean_ylm`` is NOT added on
        this branch.  ``u`` optionally supplies ``U`` of shape ``(nsamples, nt, nylm)``."""
        if not self._time_variable:
            raise ValueError("sample_ylm(t=...) needs a time-variable process (tau)")
        if self._nylm != 256:
            raise NotImplementedError("time-variable sample_ylm is implemented at ydeg = 15")
        t = self._t(t)
        nt = t.numel()
        ldt = nt + (nt & 1)
        dev, B = self.device, self._B
        Ly = self.cho_cov_ylm
        Ly = (Ly if self._batched else Ly[None]).contiguous()
        lib, h = self._lib, self._ctx.handle
        with torch.cuda.device(dev):
            Kt = torch.zeros(B, nt, ldt, dtype=torch.float64, device=dev)
            Kt[:, :, :nt] = 1.0
            _lib.check(lib.spb_temporal_scale(h, B, nt, nt, _ptr(t), _ptr(t), self._tkind,
                                              _ptr(self._tau), 1, None, 0, _ptr(Kt), ldt, nt * ldt,
                                              _stream()))
            info = self._factor_rows(Kt, nt, ldt)
            _lib.check(lib.spb_tril(h, B, nt, _ptr(Kt), ldt, nt * ldt, _stream()))
            if u is None:
                U = torch.randn(nsamples, nt, 256, dtype=torch.float64, device=dev,
                                generator=generator)
            else:
                U = torch.as_tensor(u, dtype=torch.float64).to(dev)
                if U.ndim != 3 or U.shape[1] != nt or U.shape[2] != 256:
                    raise ValueError("u must have shape (nsamples, nt, 256)")
                nsamples = U.shape[0]
            # Two batched tensor-core GEMMs for the whole (B x nsamples) set, no Python loop:
            #   V[b] (nt, ns*256) = L_t[b] (nt, nt) . [U_0 | U_1 | ...]   (NT form: Bm = the stacked
            #   U_s^T, (ns*256, ldt), shared by every b; k = time, zero-padded to an even length)
            #   W[b] (nt*ns, 256) = V[b] viewed as (nt*ns, 256) . L_y[b]^T
            # and one permutation (t, s) -> (s, t) of the result.
            Ut = torch.zeros(nsamples * 256, ldt, dtype=torch.float64, device=dev)
            Ut[:, :nt] = U.transpose(1, 2).reshape(nsamples * 256, nt)
            V = torch.empty(B, nt, nsamples * 256, dtype=torch.float64, device=dev)
            self._gemm(B, nt, nsamples * 256, ldt, Kt, ldt, nt * ldt, Ut, ldt, 0, V, nsamples * 256,
                       nt * nsamples * 256)
            W = torch.empty(B, nt * nsamples, 256, dtype=torch.float64, device=dev)
            self._gemm(B, nt * nsamples, 256, 256, V, 256, nt * nsamples * 256, Ly, 256, 65536, W,
                       256, nt * nsamples * 256)
            out = W.view(B, nt, nsamples, 256).transpose(1, 2).contiguous()
            out = torch.where(((info & 1) != 0)[:, None, None, None],
                              torch.full_like(out, float("nan")), out)
        return self._out(out)

    # ------------------------------------------------------------------ flux of given surfaces, sums
    def flux(self, y, t, i=defaults["i"], p=defaults["p"], u=None):
        """sp.py:1237-1283: light curves ``(nsamples, nt)`` of spherical-harmonic vectors ``y``
        (``(nsamples, nylm)``, or ``(nsamples, nt, nylm)`` for time-variable surfaces, e.g. the
        output of ``sample_ylm``); mean-normalised when the process is ``normalized``."""
        dev = self.device
        yt = torch.as_tensor(y, dtype=torch.float64).to(dev)
        squeeze = yt.ndim == 1
        if squeeze:
            yt = yt[None]
        if _is_batched(i):
            raise ValueError("flux() takes a scalar inclination")
        A = self._design_full(t, i, p, u)[0]
        yt = self._pad_ylm(yt, yt.ndim - 1)
        nt = A.shape[0]
        with torch.cuda.device(dev):
            if self._time_variable:
                if yt.ndim != 3 or yt.shape[1] != nt or yt.shape[2] != 256:
                    raise ValueError("time-variable surfaces: y must have shape (nsamples, nt, 256)")
                ns = yt.shape[0]
                yt = yt.contiguous()
                out = torch.empty(nt, ns, dtype=torch.float64, device=dev)
                # F[k][s] = A[k] . y[s][k]  -- one (ns x 256)(256 x 1) product per time
                self._gemm(nt, ns, 1, 256, yt, nt * 256, 256, A, 256, 256, out, 1, ns)
                F = out.transpose(0, 1).contiguous()
            else:
                if yt.ndim != 2 or yt.shape[1] != 256:
                    raise ValueError("y must have shape (nsamples, 256)")
                ns = yt.shape[0]
                F = torch.empty(ns, nt, dtype=torch.float64, device=dev)
                self._gemm(1, ns, nt, 256, yt.contiguous(), 256, 0, A, 256, 0, F, nt, 0)
            if self._normalized:   # sp.py:1279-1282
                F = (1.0 + F) / (1.0 + F).mean(dim=-1, keepdim=True) - 1.0
        return F[0] if squeeze else F

    def __add__(self, other):
        """sp.py:1190-1191."""
        return StarryProcessSum(self, other)

    def __radd__(self, other):
        """sp.py:1193-1197 (so that ``sum([...])`` works)."""
        if isinstance(other, (int, float)) and other == 0:
            return self
        return self.__add__(other)


class StarryProcessSum(StarryProcess):
    """sp.py:1335-1400: the sum of independent processes -- Ylm means and covariances add; every
    flux-space method then runs on the summed moments through the same kernels."""

    def __init__(self, first, second):
        assert isinstance(second, StarryProcess), \
            "Can only add instances of `StarryProcess` to each other."
        assert first._ydeg == second._ydeg, "Mismatch in `ydeg`."
        assert first._udeg == second._udeg, "Mismatch in `udeg`."
        assert first._normalized == second._normalized, "Mismatch in `normalized`."
        assert first._marginalize_over_inclination == second._marginalize_over_inclination, \
            "Mismatch in `marginalize_over_inclination`."
        assert first._covpts == second._covpts, "Mismatch in `covpts`."
        assert first._time_variable is False and second._time_variable is False, \
            "Sums of `StarryProcess` instances not implemented for time-variable surfaces."
        assert first.device == second.device, "Mismatch in device."
        for k in ("_ydeg", "_udeg", "_nylm", "_normalized", "_marginalize_over_inclination",
                  "_covpts", "_normN", "_normzmax", "_max_chunk_bytes", "_sigma_max", "_ctx",
                  "device", "_lib", "_opt", "_opt_keep", "_opt_kw"):
            setattr(self, k, getattr(first, k))
        self._time_variable, self._tkind, self._tau = False, 0, None
        self._children = []
        for child in (first, second):
            self._children += getattr(child, "_children", [child])
        if first._B != second._B and 1 not in (first._B, second._B):
            raise ValueError("batched summands must have the same batch size")
        self._B = max(first._B, second._B)
        self._batched = first._batched or second._batched
        first._compute_moments()
        second._compute_moments()
        B = self._B
        # Sum the random variables (sp.py:1383-1385)
        self._mean_ylm = (first._mean_ylm + second._mean_ylm).expand(B, 256).contiguous()
        self._cov_ylm = (first._cov_ylm + second._cov_ylm).expand(B, 256, 256).contiguous()
        self._info = (first._info | second._info).expand(B).contiguous()
        self._cho_cov_ylm = None
        self._z = None
        self._rTA1_cache = {}
        self._a = self._b = self._r = self._c = self._n = self._dr = None

    def _compute_moments(self):
        return

    def log_jac(self):
        raise NotImplementedError("log_jac is defined for a single process (sp.py:1004-1050)")

    @property
    def a(self):
        raise AttributeError("a sum of processes has no single latitude distribution")

    b = a
