"""ctypes binding of the CPU oracle (oracle/libsda_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs import this module.  Nothing under ``sdaopt_b200/`` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# functor ids, identical to include/sdaopt_b200.h
FN = dict(Rastrigin=0, ShiftedRastrigin=1, Rosenbrock=2, Ackley01=3, Schwefel01=4, Schwefel26=5,
          Exponential=6, SuttonChen=7, Constant=8, Sphere=9, Griewank=10)
FN.update({'Alpine01': 11, 'Brown': 12, 'Cigar': 13, 'CosineMixture': 14, 'Deb01': 15, 'DixonPrice': 16, 'Easom': 17, 'Qing': 18, 'Quintic': 19, 'Salomon': 20, 'Schwefel20': 21, 'Schwefel21': 22, 'Schwefel22': 23, 'Step': 24, 'StyblinskiTang': 25, 'Trid': 26, 'Zacharov': 27, 'SixHumpCamel': 28, 'Beale': 29, 'HimmelBlau': 30, 'Matyas': 31, 'McCormick': 32, 'GoldsteinPrice': 33, 'Levy13': 34, 'Bohachevsky1': 35, 'Bukin06': 36, 'EggCrate': 37, 'Schaffer01': 38, 'DropWave': 39, 'ThreeHumpCamel': 40, 'Zettl': 41, 'Leon': 42, 'Cube': 43, 'Brent': 44, 'Price01': 45, 'Bird': 46, 'Ackley02': 47, 'Ackley03': 48, 'Zirilli': 49, 'Wavy': 50, 'Step2': 51, 'Plateau': 52, 'Sodp': 53, 'Trigonometric02': 54, 'XinSheYang02': 55, 'SineEnvelope': 56, 'Pathological': 57, 'Rana': 58, 'Adjiman': 59})
RNG_PHILOX, RNG_TAPE = 0, 1
DEC_IMPROVE, DEC_NEW_BEST, DEC_METRO_ACCEPT, DEC_REJECT = 1, 2, 3, 4


class OrcProblem(C.Structure):
    _fields_ = [("functor", C.c_int32), ("dim", C.c_int32), ("lower", C.c_void_p),
                ("upper", C.c_void_p), ("params", C.c_void_p), ("n_params", C.c_int32)]


class OrcConfig(C.Structure):
    _fields_ = [("maxiter", C.c_int64), ("initial_temp", C.c_double), ("visit", C.c_double),
                ("accept", C.c_double), ("maxfun", C.c_double), ("pure_sa", C.c_int32),
                ("temperature_restart", C.c_double), ("rng_mode", C.c_int32),
                ("max_calls", C.c_int64), ("fglob", C.c_double), ("tol", C.c_double),
                ("trace_len", C.c_int64), ("temp_table", C.c_void_p),
                ("sigmax_table", C.c_void_p), ("table_len", C.c_int64)]


class OrcResult(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls", "status", "hit", "ncall_hit",
                 "f_hit", "trace_e", "trace_d", "tape_used")]


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsda_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_temperature.restype = C.c_double
        L.orc_temperature.argtypes = [C.c_double, C.c_double, C.c_int64]
        L.orc_sigmax.restype = C.c_double
        L.orc_sigmax.argtypes = [C.c_double, C.c_double]
        L.orc_eval.restype = C.c_double
        L.orc_eval.argtypes = [C.POINTER(OrcProblem), C.c_void_p]
        L.orc_run.restype = C.c_int
        L.orc_run_mt.restype = C.c_int
        L.orc_visiting.restype = C.c_int64
        L.orc_visit_fn.restype = C.c_int64
        L.orc_local_search.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Problem:
    """Owns the numpy buffers behind an OrcProblem."""

    def __init__(self, functor, bounds, params=None):
        b = np.asarray(bounds, dtype=np.float64)
        self.lower = np.ascontiguousarray(b[:, 0])
        self.upper = np.ascontiguousarray(b[:, 1])
        self.params = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        self.dim = int(b.shape[0])
        fid = FN[functor] if isinstance(functor, str) else int(functor)
        self.c = OrcProblem(fid, self.dim, _ptr(self.lower), _ptr(self.upper), _ptr(self.params),
                            0 if self.params is None else self.params.size)

    def eval(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_eval(C.byref(self.c), _ptr(x))


def philox4x32_10(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return [int(v) for v in out]


def run(problem, n_chains=1, x0=None, seeds=None, tape=None, maxiter=1000, initial_temp=5230.,
        visit=2.62, accept=-5.0, maxfun=1e7, pure_sa=False, max_calls=0, fglob=float("nan"),
        tol=1e-8, trace_len=0, n_threads=1, tables=None):
    """Run ``n_chains`` oracle chains; returns a dict of numpy arrays."""
    L = lib()
    D = problem.dim
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(n_chains, D)
    if seeds is not None:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    tape_len = 0
    if tape is not None:
        tape = np.ascontiguousarray(tape, dtype=np.float64).reshape(n_chains, -1)
        tape_len = tape.shape[1]
    cfg = OrcConfig(int(maxiter), float(initial_temp), float(visit), float(accept), float(maxfun),
                    int(bool(pure_sa)), 0.1, RNG_TAPE if tape is not None else RNG_PHILOX,
                    int(max_calls), float(fglob), float(tol), int(trace_len), None, None, 0)
    if tables is not None:
        tt = np.ascontiguousarray(tables[0], dtype=np.float64)
        st = np.ascontiguousarray(tables[1], dtype=np.float64)
        cfg.temp_table, cfg.sigmax_table, cfg.table_len = _ptr(tt), _ptr(st), tt.size
    out = dict(
        x=np.zeros((n_chains, D)), fun=np.zeros(n_chains), nit=np.zeros(n_chains, np.int64),
        ncall=np.zeros(n_chains, np.int64), ncall_ls=np.zeros(n_chains, np.int64),
        n_ls=np.zeros(n_chains, np.int64), status=np.zeros(n_chains, np.int32),
        hit=np.zeros(n_chains, np.int32), ncall_hit=np.zeros(n_chains, np.int64),
        f_hit=np.zeros(n_chains), trace_e=np.zeros((n_chains, max(trace_len, 1))),
        trace_d=np.zeros((n_chains, max(trace_len, 1)), np.uint8),
        tape_used=np.zeros(n_chains, np.int64))
    res = OrcResult(*[_ptr(out[k]) for k, _ in OrcResult._fields_])
    if n_threads > 1:
        rc = L.orc_run_mt(C.byref(problem.c), C.byref(cfg), C.c_int64(n_chains), _ptr(x0),
                          _ptr(seeds), _ptr(tape), C.c_int64(tape_len), C.byref(res),
                          C.c_int(n_threads))
    else:
        rc = L.orc_run(C.byref(problem.c), C.byref(cfg), C.c_int64(n_chains), _ptr(x0),
                       _ptr(seeds), _ptr(tape), C.c_int64(tape_len), C.byref(res))
    if rc == -2:
        raise ValueError('Bounds are note consistent min < max')
    if rc != 0:
        raise RuntimeError("oracle returned %d" % rc)
    out["trace_e"] = out["trace_e"][:, :trace_len]
    out["trace_d"] = out["trace_d"][:, :trace_len]
    return out


def visiting(problem, qv, x, step, temperature, tape):
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    xv = np.zeros_like(x)
    used = lib().orc_visiting(C.byref(problem.c), C.c_double(qv), _ptr(x), C.c_int64(step),
                              C.c_double(temperature), _ptr(tape), C.c_int64(tape.size), _ptr(xv))
    return xv, int(used)


def visit_fn(qv, temperature, tape, n):
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    out = np.zeros(n)
    used = lib().orc_visit_fn(C.c_double(qv), C.c_double(temperature), _ptr(tape),
                              C.c_int64(tape.size), _ptr(out), C.c_int64(n))
    return out, int(used)


def local_search(problem, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    xo = np.zeros_like(x)
    f = C.c_double()
    nfev = C.c_int64()
    nit = C.c_int64()
    ok = lib().orc_local_search(C.byref(problem.c), _ptr(x), _ptr(xo), C.byref(f), C.byref(nfev),
                                C.byref(nit))
    return bool(ok), xo, f.value, nfev.value
