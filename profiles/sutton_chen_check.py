import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import sdaopt_b200 as sb
for n_atoms, chains, maxiter in [(6, 8192, 200), (128, 1024, 4), (512, 256, 2)]:
    f = sb.SuttonChen(n_atoms)
    L = 0.7 * (n_atoms / 6.0) ** (1.0 / 3.0)
    bounds = [(-L, L)] * (3 * n_atoms)
    best = 1e9
    for rep in range(2):
        t0 = time.perf_counter()
        r = sb.sda_batch(f, bounds, chains, maxiter=maxiter, pure_sa=(n_atoms > 38))
        best = min(best, time.perf_counter() - t0)
    print("G=%s N=%d %.3f s %.3f Mevals/s" % (os.environ.get("SDA_GROUP"), n_atoms, best, r.ncall.sum() / best / 1e6), flush=True)
