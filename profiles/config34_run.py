"""BASELINE configs 3 and 4 shapes, reduced iteration counts (sanity + first throughput numbers).
    python profiles/config34_run.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdaopt_b200 as sb  # noqa: E402

rows = []
CONFIG4_ONLY = os.environ.get("CONFIG4_ONLY") == "1"
for name, dim, chains, maxiter in [] if CONFIG4_ONLY else [("Rastrigin", 10, 65536, 200), ("Ackley01", 10, 65536, 200),
                                   ("Schwefel26", 10, 65536, 200), ("Rastrigin", 100, 65536, 40),
                                   ("Ackley01", 100, 65536, 40), ("Schwefel26", 100, 65536, 40),
                                   ("Rastrigin", 1000, 4096, 4), ("Ackley01", 1000, 4096, 4)]:
    f = sb.DeviceFunctor(name, dimensions=dim)
    dt = 1e30
    for rep in range(2):          # the first call of a process also loads the 8 MB module: report the warm one
        t0 = time.perf_counter()
        r = sb.sda_batch(f, f.bounds, chains, maxiter=maxiter)
        dt = min(dt, time.perf_counter() - t0)
    rows.append((name, dim, chains, maxiter, dt, int(r.ncall.sum()), float(np.median(r.fun)), int((r.status != 0).sum())))
    print("config3 %-10s D=%-4d chains=%-6d maxiter=%-4d %.2f s  %.1f Mevals/s (ls %.0f%%)  median f=%.3g  errors=%d"
          % (name, dim, chains, maxiter, dt, r.ncall.sum() / dt / 1e6, 100.0 * r.ncall_ls.sum() / r.ncall.sum(),
             np.median(r.fun), (r.status != 0).sum()), flush=True)
# chain counts that fill the GPU (one chain per warp from D > 256: 148 SMs x 16 warps = 2,368 resident)
for n_atoms, chains, maxiter in [(6, 8192, 200), (38, 8192, 12), (128, 8192, 3), (256, 4096, 2), (512, 2048, 2)]:
    f = sb.SuttonChen(n_atoms)
    L = 0.7 * (n_atoms / 6.0) ** (1.0 / 3.0)          # constant density box
    bounds = [(-L, L)] * (3 * n_atoms)
    dt = 1e30
    for rep in range(2 if n_atoms < 256 else 1):
        t0 = time.perf_counter()
        r = sb.sda_batch(f, bounds, chains, maxiter=maxiter, pure_sa=(n_atoms > 38))
        dt = min(dt, time.perf_counter() - t0)
    print("config4 Sutton-Chen N=%-4d D=%-5d chains=%-5d maxiter=%-4d %.2f s  %.3f Mevals/s  %.0f G pair terms/s  best f=%.6g  errors=%d"
          % (n_atoms, 3 * n_atoms, chains, maxiter, dt, r.ncall.sum() / dt / 1e6,
             r.ncall.sum() * n_atoms * (n_atoms - 1) / dt / 1e9, r.fun.min(), (r.status != 0).sum()), flush=True)
