"""Persistent-kernel cases: small LS batches (D=50, 8192 chains), suite-like 100-chain jobs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdaopt_b200 as sb
def run(label, f, bounds, chains, **kw):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter(); r = sb.sda_batch(f, bounds, chains, **kw); best = min(best, time.perf_counter() - t0)
    print("%-34s %.3f s %.2f Mevals/s" % (label, best, r.ncall.sum() / best / 1e6), flush=True)
f = sb.Rastrigin(50); run("Rastrigin50 8192 LS maxiter150", f, f.bounds, 8192, maxiter=150)
f = sb.Rastrigin(10); run("Rastrigin10 100 suite-like", f, f.bounds, 100, maxiter=int(1e6), max_calls=200000, fglob=0.0, tol=1e-8, seeds=np.arange(100, dtype=np.uint64) + np.uint64(1234))
f = sb.DeviceFunctor("Bukin06"); run("Bukin06 100 suite-like (fails)", f, f.bounds, 100, maxiter=int(1e6), max_calls=200000, fglob=f.fglob, tol=1e-8, seeds=np.arange(100, dtype=np.uint64) + np.uint64(1234))
