"""How fast does ONE long suite chain run as a function of how many share the GPU?  Bukin06 (2-D, never
reaches its fglob: every chain spends its whole budget, ~80 % of it inside L-BFGS-B) with a reduced
budget, n chains in one launch, one chain per warp.  Prints chains/s and the per-chain latency
extrapolated to the 1e6 budget.    python profiles/suite_contention.py [budget]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdaopt_b200 as sb  # noqa: E402
from sdaopt_b200.device import DeviceBatch  # noqa: E402

budget = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
names = sys.argv[2:] or ["Bukin06"]
for name in names:
    f = sb.DeviceFunctor(name)
    for n in [int(v) for v in os.environ.get("SDA_CONT_NS", "37,148,296,592,1184,2368,4736").split(",")]:
        b = DeviceBatch(f, f.bounds, n, maxiter=int(1e6), max_calls=budget, fglob=f.fglob, tol=1e-8)
        b.set_seeds(1234 + np.arange(n))
        b.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b.set_seeds(99991 + np.arange(n))
        e0.record()
        b.launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        r = b.fetch()
        evals = float(r.ncall.sum())
        print("%-10s G=%s chains %5d  %8.1f ms  %7.1f M evals/s  per-chain latency at 1e6 budget %6.2f s  (hit %d, ls evals %.0f%%, %.1f evals per LS)"
              % (name, os.environ.get("SDA_GROUP", "auto"), n, ms, evals / ms / 1e3, ms * 1e-3 * 1e6 / budget,
                 int(r.hit.sum()), 100.0 * r.ncall_ls.sum() / evals, r.ncall_ls.sum() / max(r.n_ls.sum(), 1)), flush=True)
