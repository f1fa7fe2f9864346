"""Registered device objectives and the resolution of the reference's ``func`` argument.

The reference takes any Python callable ``f(x, *args)`` (_sda.py:440-444).  The B200 path
evaluates objectives inside the annealing kernel, so ``func`` must name a device functor:

* a :class:`DeviceFunctor` instance (``Rastrigin()``, ``ShiftedRastrigin(3.14159)`` ...),
* the bound method ``Benchmark.fun`` of a catalogue class whose name is registered
  (sdaopt/benchmark/go_benchmark_functions), as the suite passes it (benchmark/bench.py:298-299),
* the function ``sutton_chen`` (sdaopt/benchmark/suttonchen.py:23), matched by name,
* or the registered name as a string.

Anything else raises ``TypeError``: there is no CPU fallback.
"""
import ctypes as C
import hashlib
import os

import numpy as np

from . import _lib

# name -> id, identical to include/sdaopt_b200.h (SDA_FN_*)
FUNCTOR_IDS = dict(Rastrigin=0, ShiftedRastrigin=1, Rosenbrock=2, Ackley01=3, Schwefel01=4,
                   Schwefel26=5, Exponential=6, SuttonChen=7, Constant=8, Sphere=9, Griewank=10)

FUNCTOR_IDS.update({
    'Alpine01': 11,
    'Brown': 12,
    'Cigar': 13,
    'CosineMixture': 14,
    'Deb01': 15,
    'DixonPrice': 16,
    'Easom': 17,
    'Qing': 18,
    'Quintic': 19,
    'Salomon': 20,
    'Schwefel20': 21,
    'Schwefel21': 22,
    'Schwefel22': 23,
    'Step': 24,
    'StyblinskiTang': 25,
    'Trid': 26,
    'Zacharov': 27,
    'SixHumpCamel': 28,
    'Beale': 29,
    'HimmelBlau': 30,
    'Matyas': 31,
    'McCormick': 32,
    'GoldsteinPrice': 33,
    'Levy13': 34,
    'Bohachevsky1': 35,
    'Bukin06': 36,
    'EggCrate': 37,
    'Schaffer01': 38,
    'DropWave': 39,
    'ThreeHumpCamel': 40,
    'Zettl': 41,
    'Leon': 42,
    'Cube': 43,
    'Brent': 44,
    'Price01': 45,
    'Bird': 46,
    'Ackley02': 47,
    'Ackley03': 48,
    'Zirilli': 49,
    'Wavy': 50,
    'Step2': 51,
    'Plateau': 52,
    'Sodp': 53,
    'Trigonometric02': 54,
    'XinSheYang02': 55,
    'SineEnvelope': 56,
    'Pathological': 57,
    'Rana': 58,
    'Adjiman': 59,
})

# name: (default N, n-D capable, box or per-dimension bounds, fglob or fglob(N)) -- the class
# attributes of go_benchmark_functions/go_funcs_*.py (`_bounds`, `fglob`, `change_dimensionality`)
_CATALOGUE2 = {
    'Alpine01': (2, True, (-10.0, 10.0), 0.0),
    'Brown': (2, True, (-1.0, 4.0), 0.0),
    'Cigar': (2, True, (-100.0, 100.0), 0.0),
    'CosineMixture': (2, True, (-1.0, 1.0), lambda n: -0.9 * n),
    'Deb01': (2, True, (-1.0, 1.0), -1.0),
    'DixonPrice': (2, True, (-10.0, 10.0), 0.0),
    'Easom': (2, False, (-100.0, 100.0), -1.0),
    'Qing': (2, True, (-500.0, 500.0), 0.0),
    'Quintic': (2, True, (-10.0, 10.0), 0.0),
    'Salomon': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel20': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel21': (2, True, (-100.0, 100.0), 0.0),
    'Schwefel22': (2, True, (-100.0, 100.0), 0.0),
    'Step': (2, True, (-100.0, 100.0), 0.0),
    'StyblinskiTang': (2, True, (-5.0, 5.0), lambda n: -39.16616570377142 * n),
    'Trid': (6, True, (-20.0, 20.0), -50.0),
    'Zacharov': (2, True, (-5.0, 10.0), 0.0),
    'SixHumpCamel': (2, False, (-5.0, 5.0), -1.031628),
    'Beale': (2, False, (-4.5, 4.5), 0.0),
    'HimmelBlau': (2, False, (-5.0, 5.0), 0.0),
    'Matyas': (2, False, (-10.0, 10.0), 0.0),
    'McCormick': (2, False, [(-1.5, 4.0), (-3.0, 3.0)], -1.913222954981037),
    'GoldsteinPrice': (2, False, (-2.0, 2.0), 3.0),
    'Levy13': (2, False, (-10.0, 10.0), 0.0),
    'Bohachevsky1': (2, False, (-100.0, 100.0), 0.0),
    'Bukin06': (2, False, [(-15.0, -5.0), (-3.0, 3.0)], 0.0),
    'EggCrate': (2, False, (-5.0, 5.0), 0.0),
    'Schaffer01': (2, False, (-100.0, 100.0), 0.0),
    'DropWave': (2, False, (-5.12, 5.12), -1.0),
    'ThreeHumpCamel': (2, False, (-5.0, 5.0), 0.0),
    'Zettl': (2, False, (-5.0, 10.0), -0.003791237220468656),
    'Leon': (2, False, (-1.2, 1.2), 0.0),
    'Cube': (2, False, (-10.0, 10.0), 0.0),
    'Brent': (2, False, (-10.0, 10.0), 0.0),
    'Price01': (2, False, (-500.0, 500.0), 0.0),
    'Bird': (2, False, (-6.283185307179586, 6.283185307179586), -106.7645367198034),
    'Ackley02': (2, False, (-32.0, 32.0), -200.0),
    'Ackley03': (2, False, (-32.0, 32.0), -195.6290282592388),
    'Zirilli': (2, False, (-10.0, 10.0), -0.35238603),
    'Wavy': (2, True, (-3.141592653589793, 3.141592653589793), 0.0),
    'Step2': (2, True, (-100.0, 100.0), 0.5),
    'Plateau': (2, True, (-5.12, 5.12), 30.0),
    'Sodp': (2, True, (-1.0, 1.0), 0.0),
    'Trigonometric02': (2, True, (-500.0, 500.0), 1.0),
    'XinSheYang02': (2, True, (-6.283185307179586, 6.283185307179586), 0.0),
    'SineEnvelope': (2, True, (-100.0, 100.0), 0.0),
    'Pathological': (2, False, (-100.0, 100.0), 0.0),
    'Rana': (2, True, (-500.000001, 500.000001), -500.8021602966615),
    'Adjiman': (2, False, [(-1.0, 2.0), (-1.0, 1.0)], -2.02180678),
}

# default box and known optimum of the catalogue classes (go_funcs_*.py `__init__`)
_CATALOGUE = {
    "Rastrigin": ((-5.12, 5.12), 0.0),      # go_funcs_R.py:80-86
    "Rosenbrock": ((-30.0, 30.0), 0.0),     # go_funcs_R.py:280-286
    "Ackley01": ((-35.0, 35.0), 0.0),       # go_funcs_A.py:38-41
    "Schwefel01": ((-100.0, 100.0), 0.0),   # go_funcs_S.py:324-330
    "Schwefel26": ((-500.0, 500.0), 0.0),   # go_funcs_S.py:606-611
    "Exponential": ((-1.0, 1.0), -1.0),     # go_funcs_E.py:296-301
    "Sphere": ((-5.12, 5.12), 0.0),         # go_funcs_S.py:1149-1154
    "Griewank": ((-100.0, 100.0), 0.0),     # go_funcs_G.py:160-169
}


class DeviceFunctor(object):
    """A registered CUDA objective.  Calling it evaluates on the GPU (one point or a batch)."""

    def __init__(self, name, params=(), dimensions=None):
        if name not in FUNCTOR_IDS:
            raise KeyError("unknown device functor %r" % (name,))
        self.name = name
        self.functor_id = FUNCTOR_IDS[name]
        self.params = np.ascontiguousarray(params, dtype=np.float64).ravel()
        self.N = dimensions
        box, fglob = _CATALOGUE.get(name, (None, float("nan")))
        if name in _CATALOGUE2:
            n_default, _, box, fglob = _CATALOGUE2[name]
            if self.N is None:
                self.N = n_default
            if callable(fglob):
                fglob = fglob(self.N)
        self.fglob = fglob
        self._box = box
        self.library_path = None        # set for user functors: the build that contains them

    @property
    def bounds(self):
        if self._box is None or self.N is None:
            raise AttributeError("%s has no default bounds" % self.name)
        if isinstance(self._box, list):                    # per-dimension bounds (fixed-N classes)
            return list(self._box)
        return [self._box] * int(self.N)

    def fun(self, x, *args):
        return self(x)

    def __call__(self, x, *args):
        x = np.ascontiguousarray(x, dtype=np.float64)
        single = x.ndim == 1
        pts = x.reshape(1, -1) if single else x
        out = evaluate(self, pts)
        return float(out[0]) if single else out

    def __repr__(self):
        return "DeviceFunctor(%s%s)" % (self.name, "" if not self.params.size else
                                        ", params=%s" % self.params.tolist())


def Rastrigin(dimensions=2): return DeviceFunctor("Rastrigin", dimensions=dimensions)
def Rosenbrock(dimensions=2): return DeviceFunctor("Rosenbrock", dimensions=dimensions)
def Ackley01(dimensions=2): return DeviceFunctor("Ackley01", dimensions=dimensions)
def Schwefel01(dimensions=2): return DeviceFunctor("Schwefel01", dimensions=dimensions)
def Schwefel26(dimensions=2): return DeviceFunctor("Schwefel26", dimensions=dimensions)
def Exponential(dimensions=2): return DeviceFunctor("Exponential", dimensions=dimensions)
def Sphere(dimensions=2): return DeviceFunctor("Sphere", dimensions=dimensions)
def Griewank(dimensions=2): return DeviceFunctor("Griewank", dimensions=dimensions)


def ShiftedRastrigin(shift=3.14159, dimensions=None):
    """README.md:61-71: sum((x-s)^2 - 10 cos(2 pi (x-s))) + 10 D"""
    return DeviceFunctor("ShiftedRastrigin", params=[shift], dimensions=dimensions)


def SuttonChen(n_atoms=None):
    """sdaopt/benchmark/suttonchen.py:23-40 (x is the flattened (n_atoms, 3) coordinate array)"""
    return DeviceFunctor("SuttonChen", dimensions=None if n_atoms is None else 3 * n_atoms)


def Constant(value):
    """f(x) = value; the `weirdfunc` of sdaopt/tests/test_sda.py:34"""
    return DeviceFunctor("Constant", params=[value])


def compile_user_functor(source, params=(), name="user", bounds=None, fglob=float("nan"),
                          build_dir=None):
    """User-supplied CUDA device functor (replaces the reference's arbitrary Python ``func``,
    _sda.py:440-444).  ``source`` must define

        template <class X>
        __device__ double sda_user_objective(const X &x, int dim, const double *params);

    where ``x[k]`` is the k-th coordinate.  The annealing / local-search kernels are rebuilt with
    the functor inlined (nvcc, about a minute, cached by source hash under build/user/) and the
    returned DeviceFunctor routes every call to that library.  No CPU evaluation ever happens."""
    from . import build as _build
    digest = hashlib.sha1(source.encode()).hexdigest()[:16]
    root = build_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                     "build", "user", digest)
    os.makedirs(root, exist_ok=True)
    header = os.path.join(root, "user_functor.cuh")
    so = os.path.join(root, "libsdaopt_b200_user.so")
    if not os.path.exists(so):
        with open(header, "w") as fh:
            fh.write("// user functor %s\n#pragma once\n%s\n" % (name, source))
        _build.build_user(header, so)
    f = DeviceFunctor.__new__(DeviceFunctor)
    f.name = name
    f.functor_id = _lib.SDA_FN_USER
    f.params = np.ascontiguousarray(params, dtype=np.float64).ravel()
    f.N = None if bounds is None else len(bounds)
    f.fglob = fglob
    f._box = None if bounds is None else [tuple(b) for b in bounds]
    f.library_path = so
    return f


def resolve(func):
    """Map the reference's ``func`` argument to a DeviceFunctor or raise TypeError."""
    if isinstance(func, DeviceFunctor):
        return func
    if isinstance(func, str):
        return DeviceFunctor(func)
    owner = getattr(func, "__self__", None)
    if owner is not None and getattr(func, "__name__", "") in ("fun", "_funcwrapped"):
        if isinstance(owner, DeviceFunctor):
            return owner
        # Algo._funcwrapped wraps self._k.fun (benchmark/bench.py:298)
        bench = getattr(owner, "_k", owner)
        name = type(bench).__name__
        if name in FUNCTOR_IDS:
            return DeviceFunctor(name, dimensions=getattr(bench, "N", None))
    if getattr(func, "__name__", "") == "sutton_chen":
        return DeviceFunctor("SuttonChen")
    raise TypeError(
        "sdaopt_b200 evaluates objectives on the GPU: `func` must be a registered device functor "
        "(sdaopt_b200.functors), a catalogue Benchmark.fun bound method, `sutton_chen` or a "
        "registered name -- arbitrary Python callables are not supported (no CPU fallback). "
        "Got %r" % (func,))


def make_problem(functor, lower, upper):
    """(SdaProblem with HOST pointers, keep-alive tuple)"""
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    params = functor.params
    p = _lib.SdaProblem(functor.functor_id, int(lower.size), lower.ctypes.data, upper.ctypes.data,
                        params.ctypes.data if params.size else None, int(params.size))
    return p, (lower, upper, params)


def evaluate(functor, points, device=0):
    """Objective values at points[n, D] (replaces Benchmark.fun / sutton_chen per point)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n, dim = pts.shape
    dummy = np.zeros(dim)
    p, keep = make_problem(functor, dummy, dummy + 1)
    out = np.empty(n, dtype=np.float64)
    _lib.require_cuda()
    _lib.check(_lib.lib(functor.library_path).sda_eval_host(C.byref(p), pts.ctypes.data, n,
                                                         out.ctypes.data, device))
    return out
