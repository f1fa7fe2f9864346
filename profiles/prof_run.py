"""Minimal driver for ncu: launches the annealing kernel a few times on a small batch.
    python profiles/prof_run.py --chains 16384 --maxiter 12 [--pure-sa] [--functor Rastrigin --dim 50]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import sdaopt_b200 as sb  # noqa: E402
from sdaopt_b200.device import DeviceBatch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=16384)
ap.add_argument("--maxiter", type=int, default=12)
ap.add_argument("--dim", type=int, default=50)
ap.add_argument("--functor", default="Rastrigin")
ap.add_argument("--pure-sa", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
f = sb.DeviceFunctor(a.functor, dimensions=a.dim)
b = DeviceBatch(f, f.bounds, a.chains, maxiter=a.maxiter, pure_sa=a.pure_sa)
for rep in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b.launch()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r = b.fetch()
    print("rep %d: %.2f ms, evals %d (ls %d), %.1f Mevals/s" % (
        rep, dt * 1e3, r.ncall.sum(), r.ncall_ls.sum(), r.ncall.sum() / dt / 1e6))
