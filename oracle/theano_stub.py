"""
TEST INFRASTRUCTURE ONLY -- not part of the product path.

An *eager NumPy stand-in* for the handful of ``theano`` / ``aesara_theano_fallback`` entry
points that the reference's Python glue touches, so that the UNMODIFIED reference package under
``/root/reference/starry_process`` can be imported and executed in this container (Theano,
aesara, pymc3, starry and matplotlib are absent and there is no network).  With it

    import oracle.theano_stub as ts; ts.install(); import starry_process

runs the reference's own ``sp.py / flux.py / integrals.py / latitude.py / ...`` line by line; every
``BaseOp`` (the Theano ``ExternalCOp`` wrappers around ``ops/include/*.h``, base_op.py:17-90) is
dispatched to ``oracle/_ref/libspref_y*_u*.so`` -- the reference's own C++ headers compiled by
``oracle/Makefile``.  Nothing here re-implements reference *numerics*: ``tt.dot`` is ``numpy.dot``,
``tt.set_subtensor`` is a copy + ``__setitem__``, ``ifelse`` is a Python conditional, and
``Op.__call__`` runs ``make_node`` + ``perform`` immediately.

Used only by ``oracle/gen_golden.py`` (fixture generation) and by CPU tests that are skipped when
``/root/reference`` is missing.
"""
import ctypes
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SP_REFERENCE_ROOT", "/root/reference")


# ----------------------------------------------------------------------------------------------
# Eager tensor
# ----------------------------------------------------------------------------------------------
class _Placeholder(object):
    """What ``TensorType(...)()`` returns: an output slot filled in by ``perform``."""

    def __init__(self, ndim=None):
        self.ndim = ndim


class _TypeFactory(object):
    def __init__(self, ndim=None):
        self.ndim = ndim

    def __call__(self, *args, **kwargs):
        return _Placeholder(self.ndim)


class T(np.ndarray):
    """ndarray that remembers how it was sliced (for ``set_subtensor``) and compares by identity."""

    __array_priority__ = 100.0

    def __new__(cls, x):
        arr = np.asarray(x)
        if arr.dtype != np.float64 and arr.dtype.kind in "iub":
            pass
        obj = arr.view(cls)
        obj._parent = None
        obj._idx = None
        return obj

    def __array_finalize__(self, obj):
        self._parent = None
        self._idx = None

    def __getitem__(self, idx):
        out = np.ndarray.__getitem__(self, idx)
        if not isinstance(out, np.ndarray):
            out = np.asarray(out).view(T)
        out._parent = self
        out._idx = idx
        return out

    # Theano variables compare by identity
    def __eq__(self, other):
        return self is other

    def __ne__(self, other):
        return self is not other

    __hash__ = object.__hash__

    def eval(self, *args, **kwargs):
        return np.array(self)

    def astype(self, dtype, *a, **k):
        return T(np.asarray(self).astype(dtype))

    @property
    def type(self):
        return _TypeFactory(self.ndim)

    def dimshuffle(self, *pattern):
        arr = np.asarray(self)
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])
        keep = [p for p in pattern if p != "x"]
        arr = arr.transpose(keep)
        idx = tuple(np.newaxis if p == "x" else slice(None) for p in pattern)
        return T(arr[idx])


def _a(x):
    return np.asarray(x)


def _t(x):
    return T(x)


# ----------------------------------------------------------------------------------------------
# Graph objects
# ----------------------------------------------------------------------------------------------
class Node(object):
    pass


class Apply(Node):
    def __init__(self, op, inputs, outputs):
        self.op = op
        self.inputs = list(inputs)
        self.outputs = list(outputs)


class Op(object):
    __props__ = ()

    def make_node(self, *inputs):
        raise NotImplementedError

    def perform(self, node, inputs, output_storage):
        raise NotImplementedError

    def __call__(self, *inputs, **kwargs):
        node = self.make_node(*inputs)
        ins = [np.array(i, dtype=np.float64) if not isinstance(i, _Placeholder) else i
               for i in node.inputs]
        storage = [[None] for _ in node.outputs]
        if isinstance(self, ExternalCOp):
            _c_perform(self, ins, storage)
        else:
            self.perform(node, ins, storage)
        outs = [_t(np.asarray(s[0])) for s in storage]
        return outs[0] if len(outs) == 1 else outs


class ExternalCOp(Op):
    def __init__(self, func_files=None, func_name=None):
        self._stub_func_name = func_name


class Params(object):
    pass


class ParamsType(object):
    pass


# ----------------------------------------------------------------------------------------------
# Dispatch of the reference's C ops to oracle/_ref (the reference's own headers, compiled)
# ----------------------------------------------------------------------------------------------
_LIBS = {}


def ref_lib(ydeg, udeg):
    key = (int(ydeg), int(udeg))
    if key not in _LIBS:
        path = os.path.join(HERE, "_ref", "libspref_y%d_u%d.so" % key)
        if not os.path.exists(path):
            raise RuntimeError(
                "oracle/_ref library for ydeg=%d udeg=%d missing: run `make -C oracle ref`" % key
            )
        lib = ctypes.CDLL(path)
        D = ctypes.c_double
        P = ctypes.c_void_p
        I = ctypes.c_int
        lib.ref_Rx.argtypes = [D, P, P]
        lib.ref_tensordotRz.argtypes = [P, P, I, P]
        lib.ref_special_tensordotRz.argtypes = [P, P, P, I, P]
        lib.ref_rTA1.argtypes = [P]
        lib.ref_rTA1L.argtypes = [P, P]
        lib.ref_latitude.argtypes = [D, D, P, P]
        lib.ref_hyp2f1.argtypes = [D, D, D, D]
        lib.ref_hyp2f1.restype = D
        for name in ("ref_Rx", "ref_tensordotRz", "ref_special_tensordotRz", "ref_rTA1",
                     "ref_rTA1L", "ref_latitude"):
            getattr(lib, name).restype = None
        assert lib.ref_ydeg() == key[0] and lib.ref_udeg() == key[1]
        _LIBS[key] = lib
    return _LIBS[key]


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c_perform(op, ins, storage):
    name = op._stub_func_name
    ydeg = op.ydeg
    udeg = op.udeg
    N = (ydeg + 1) ** 2
    nwig = ((ydeg + 1) * (2 * ydeg + 1) * (2 * ydeg + 3)) // 3
    lib = ref_lib(ydeg, udeg)
    cc = np.ascontiguousarray
    if name == "APPLY_SPECIFIC(Rx)":
        R = np.empty(nwig)
        dR = np.empty(nwig)
        lib.ref_Rx(float(ins[0]), _ptr(R), _ptr(dR))
        storage[0][0] = R
        storage[1][0] = dR
    elif name == "APPLY_SPECIFIC(tensordotRz)":
        M = cc(ins[0])
        th = cc(ins[1]).reshape(-1)
        assert M.ndim == 2 and M.shape[1] == N and M.shape[0] == th.shape[0]
        f = np.empty_like(M)
        lib.ref_tensordotRz(_ptr(M), _ptr(th), th.shape[0], _ptr(f))
        storage[0][0] = f
    elif name == "APPLY_SPECIFIC(special_tensordotRz)":
        Tm = cc(ins[0])
        M = cc(ins[1])
        th = cc(ins[2]).reshape(-1)
        assert Tm.shape == (N, N) and M.shape == (N, N)
        f = np.empty(th.shape[0])
        lib.ref_special_tensordotRz(_ptr(Tm), _ptr(M), _ptr(th), th.shape[0], _ptr(f))
        storage[0][0] = f
    elif name == "APPLY_SPECIFIC(rTA1)":
        f = np.empty(N)
        lib.ref_rTA1(_ptr(f))
        storage[0][0] = f
    elif name == "APPLY_SPECIFIC(rTA1L)":
        u = cc(ins[0]).reshape(-1)
        assert u.shape[0] == udeg
        f = np.empty(N)
        lib.ref_rTA1L(_ptr(u), _ptr(f))
        storage[0][0] = f
    elif name == "APPLY_SPECIFIC(latitude)":
        q = np.empty(N)
        Q = np.empty((N, N))
        lib.ref_latitude(float(ins[0]), float(ins[1]), _ptr(q), _ptr(Q))
        nanv = np.full(N, np.nan)
        nanm = np.full((N, N), np.nan)
        for k, v in enumerate([q, nanv, nanv, Q, nanm, nanm]):  # derivative lanes unused forward
            storage[k][0] = v
    else:
        raise NotImplementedError("C op %s is not on the forward hot path" % name)


# ----------------------------------------------------------------------------------------------
# tensor namespace
# ----------------------------------------------------------------------------------------------
def _build_tt():
    tt = types.ModuleType("aesara_theano_fallback.tensor")

    tt.as_tensor_variable = lambda x, *a, **k: x if isinstance(x, T) else _t(np.asarray(x))
    tt.TensorType = lambda dtype=None, broadcastable=(), **k: _TypeFactory(len(broadcastable))
    tt.dvector = lambda *a, **k: _Placeholder(1)
    tt.dmatrix = lambda *a, **k: _Placeholder(2)
    tt.dscalar = lambda *a, **k: _Placeholder(0)

    def wrap1(f):
        return lambda x, *a, **k: _t(f(_a(x), *a, **k))

    for name, f in dict(exp=np.exp, log=np.log, sqrt=np.sqrt, cos=np.cos, sin=np.sin,
                        arctan=np.arctan, abs_=np.abs, floor=np.floor, isnan=np.isnan,
                        transpose=np.transpose, diag=np.diag, sum=np.sum, mean=np.mean,
                        zeros_like=np.zeros_like, ones_like=np.ones_like, tril=np.tril,
                        triu=np.triu, argmax=np.argmax, prod=np.prod, max=np.max,
                        min=np.min).items():
        setattr(tt, name, wrap1(f))

    tt.dot = lambda a, b: _t(np.dot(_a(a), _a(b)))
    tt.outer = lambda a, b: _t(np.outer(_a(a), _a(b)))
    tt.tensordot = lambda a, b, axes=2: _t(np.tensordot(_a(a), _a(b), axes=axes))
    tt.batched_dot = lambda a, b: _t(np.einsum("ij,ij->i", _a(a), _a(b)))
    tt.zeros = lambda shape, dtype="float64": _t(np.zeros(shape, dtype=dtype))
    tt.ones = lambda shape, dtype="float64": _t(np.ones(shape, dtype=dtype))
    tt.eye = lambda n, *a, **k: _t(np.eye(int(n), *a, **k))
    tt.arange = lambda *a, **k: _t(np.arange(*[_a(x)[()] if isinstance(x, np.ndarray) else x
                                               for x in a], **k))
    tt.reshape = lambda x, shape, ndim=None: _t(np.reshape(_a(x), tuple(int(s) for s in shape)
                                                             if not isinstance(shape, int)
                                                             else shape))
    tt.swapaxes = lambda x, a, b: _t(np.swapaxes(_a(x), a, b))
    tt.tile = lambda x, reps, ndim=None: _t(np.tile(_a(x), reps))
    tt.mod = lambda a, b: _t(np.mod(_a(a), _a(b)))
    tt.cast = lambda x, dtype: _t(_a(x).astype(dtype))
    tt.shape = lambda x: _a(x).shape
    tt.concatenate = lambda xs, axis=0: _t(np.concatenate([_a(x) for x in xs], axis=axis))
    tt.maximum = lambda a, b: _t(np.maximum(_a(a), _a(b)))
    tt.minimum = lambda a, b: _t(np.minimum(_a(a), _a(b)))
    tt.gt = lambda a, b: _t(np.greater(_a(a), _a(b)))
    tt.lt = lambda a, b: _t(np.less(_a(a), _a(b)))
    tt.ge = lambda a, b: _t(np.greater_equal(_a(a), _a(b)))
    tt.le = lambda a, b: _t(np.less_equal(_a(a), _a(b)))
    tt.eq = lambda a, b: _t(np.equal(_a(a), _a(b)))
    tt.or_ = lambda a, b: _t(np.logical_or(_a(a), _a(b)))
    tt.and_ = lambda a, b: _t(np.logical_and(_a(a), _a(b)))
    tt.switch = lambda c, a, b: _t(np.where(_a(c), _a(a), _a(b)))

    def set_subtensor(sub, val, **k):
        assert isinstance(sub, T) and sub._parent is not None, "set_subtensor needs x[idx]"
        out = np.array(sub._parent, copy=True)
        out[sub._idx] = _a(val)
        return _t(out)

    def inc_subtensor(sub, val, **k):
        assert isinstance(sub, T) and sub._parent is not None
        out = np.array(sub._parent, copy=True)
        np.add.at(out, sub._idx, _a(val))
        return _t(out)

    tt.set_subtensor = set_subtensor
    tt.inc_subtensor = inc_subtensor

    extra_ops = types.ModuleType("aesara_theano_fallback.tensor.extra_ops")

    class CpuContiguous(object):
        def __call__(self, x):
            return _t(np.ascontiguousarray(_a(x)))

    extra_ops.CpuContiguous = CpuContiguous
    tt.extra_ops = extra_ops

    nlinalg = types.ModuleType("aesara_theano_fallback.tensor.nlinalg")

    class Eig(Op):
        pass

    nlinalg.Eig = Eig
    tt.nlinalg = nlinalg

    slinalg = types.ModuleType("aesara_theano_fallback.tensor.slinalg")

    class Solve(Op):
        def __init__(self, A_structure="general", lower=False, overwrite_A=False,
                     overwrite_b=False):
            self.A_structure = A_structure
            self.lower = lower

        def make_node(self, A, b):
            return Apply(self, [tt.as_tensor_variable(A), tt.as_tensor_variable(b)],
                         [_Placeholder()])

    class Cholesky(Op):
        def __init__(self, lower=True, on_error="raise"):
            self.lower = lower
            self.destructive = False
            self.on_error = on_error

        def make_node(self, x):
            return Apply(self, [tt.as_tensor_variable(x)], [_Placeholder()])

    slinalg.Solve = Solve
    slinalg.Cholesky = Cholesky
    tt.slinalg = slinalg
    return tt, slinalg, extra_ops, nlinalg


def _ifelse(cond, a, b):
    return a if bool(np.asarray(cond)) else b


class _RandomStream(object):
    def __init__(self, seed=0):
        self._rng = np.random.RandomState(seed)

    def normal(self, size=None, **k):
        return _t(self._rng.normal(size=size))

    def uniform(self, size=None, **k):
        return _t(self._rng.uniform(size=size))


class _Anything(types.ModuleType):
    """Module whose every attribute is another permissive stub (pymc3, matplotlib, ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return self


_INSTALLED = False


def install(reference_root=None):
    """Register the stub modules and put the reference package on ``sys.path``."""
    global _INSTALLED
    if _INSTALLED:
        return
    root = reference_root or REFERENCE_ROOT
    if not os.path.isdir(os.path.join(root, "starry_process")):
        raise RuntimeError("reference tree not found at %s" % root)

    tt, slinalg, extra_ops, nlinalg = _build_tt()

    atf = types.ModuleType("aesara_theano_fallback")
    theano = types.ModuleType("aesara_theano_fallback.aesara")
    theano.config = types.SimpleNamespace(floatX="float64", cast_policy="numpy+floatX",
                                          compute_test_value="ignore")
    theano.gradient = types.SimpleNamespace(DisconnectedType=type("DisconnectedType", (), {}))
    theano.tensor = tt

    def _no_function(*a, **k):
        raise NotImplementedError("eager stub: call the tensors' .eval() instead")

    theano.function = _no_function
    atf.aesara = theano
    atf.tensor = tt
    atf.ifelse = _ifelse
    atf.USE_AESARA = True

    graph = types.ModuleType("aesara_theano_fallback.graph")
    basic = types.SimpleNamespace(Node=Node, Apply=Apply)
    op = types.SimpleNamespace(Op=Op, ExternalCOp=ExternalCOp)
    params_type = types.SimpleNamespace(Params=Params, ParamsType=ParamsType)
    graph.basic, graph.op, graph.params_type, graph.fg = basic, op, params_type, None
    atf.graph = graph

    rutils = types.ModuleType("aesara.tensor.random.utils")
    rutils.RandomStream = _RandomStream

    mods = {
        "aesara_theano_fallback": atf,
        "aesara_theano_fallback.tensor": tt,
        "aesara_theano_fallback.tensor.slinalg": slinalg,
        "aesara_theano_fallback.tensor.extra_ops": extra_ops,
        "aesara_theano_fallback.tensor.nlinalg": nlinalg,
        "aesara_theano_fallback.graph": graph,
        "aesara": _Anything("aesara"),
        "aesara.tensor": _Anything("aesara.tensor"),
        "aesara.tensor.random": _Anything("aesara.tensor.random"),
        "aesara.tensor.random.utils": rutils,
    }
    for name in ("pymc3", "pymc3.blocking", "pymc3.model", "pymc3.aesaraf", "pymc3.theanof",
                 "pymc3.util", "pymc3_ext", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.colors", "matplotlib.animation", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1"):
        mods[name] = _Anything(name)
    # the setuptools_scm-generated version module is absent from a source checkout
    ver = types.ModuleType("starry_process.starry_process_version")
    ver.__version__ = "0.0.0.dev0+reference-checkout"
    mods["starry_process.starry_process_version"] = ver
    for name, mod in mods.items():
        sys.modules.setdefault(name, mod)
    if root not in sys.path:
        sys.path.insert(0, root)
    _INSTALLED = True


def import_reference(reference_root=None):
    install(reference_root)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import starry_process  # noqa: the reference package, unmodified

    return starry_process
