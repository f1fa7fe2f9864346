"""Minimal driver for ncu: the L-BFGS-B kernel alone (unit hook sda_local_search_host).
    python profiles/prof_ls.py --n 8192 --dim 50 [--functor Rastrigin]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdaopt_b200 as sb  # noqa: E402
from sdaopt_b200 import _lib, functors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--dim", type=int, default=50)
ap.add_argument("--functor", default="Rastrigin")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--near", action="store_true", help="start near local minima (integers +- 0.05), like in-run searches")
a = ap.parse_args()
f = sb.DeviceFunctor(a.functor, dimensions=a.dim)
lo = np.array([b[0] for b in f.bounds]); up = np.array([b[1] for b in f.bounds])
X = lo + np.random.RandomState(1).random_sample((a.n, a.dim)) * (up - lo)
if a.near:
    X = np.clip(np.round(X * 0.6) + np.random.RandomState(2).normal(size=X.shape) * 0.05, lo, up)
p, keep = functors.make_problem(f, lo, up)
xo = np.empty_like(X); fo = np.empty(a.n); ok = np.zeros(a.n, np.int32); nf = np.zeros(a.n, np.int64)
for rep in range(a.reps):
    t0 = time.perf_counter()
    _lib.check(_lib.lib().sda_local_search_host(C.byref(p), X.ctypes.data, a.n, xo.ctypes.data,
                                                 fo.ctypes.data, ok.ctypes.data, nf.ctypes.data, 0))
    dt = time.perf_counter() - t0
    print("rep %d: %.1f ms (incl. copies), %d searches, success %.3f, evals %d -> %.1f Mevals/s, "
          "mean requests %.1f" % (rep, dt * 1e3, a.n, ok.mean(), nf.sum(), nf.sum() / dt / 1e6,
                                  nf.mean() / (1 + 2 * a.dim)))
