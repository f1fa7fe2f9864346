"""The Python reference's own CPU speed on the build host (BASELINE.md section 3, items 1-2 and a
Benchmarker-style subset of item 4), 1 core per run.  /root/reference is imported read-only; build
container only.    python profiles/python_reference_cpu.py > profiles/r2_python_reference_cpu.json"""
import json
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import ref_import  # noqa: E402

ref_import.load()
from sdaopt import sda  # noqa: E402
import sdaopt.benchmark.go_benchmark_functions as gbf  # noqa: E402
from sdaopt.benchmark import bench as refbench  # noqa: E402

out = {"host": {"cpu_count": os.cpu_count(), "numpy": np.__version__, "cores_used_per_run": 1}}


def timed(fn):
    t0 = time.perf_counter()
    r = fn()
    return r, time.perf_counter() - t0


# config 1: README.md:28-39
def rosenbrock(x):
    return np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)


r, dt = timed(lambda: sda(rosenbrock, None, [(-30, 30)] * 2, seed=1234))
out["config1_rosenbrock2_seed1234"] = dict(fun=float(r.fun), nit=int(r.nit), ncall=int(r.ncall), seconds=dt,
                                           evals_per_s=r.ncall / dt)

# config 2: README.md:41-80 (pure_sa trajectory run) and the same with local search on
D = 50
np.random.seed(1234)
lower, upper = np.array([-5.12] * D), np.array([10.24] * D)
x_init = lower + np.random.rand(D) * (upper - lower)


def modified_rastrigin(x):
    return np.sum((x - 3.14159) ** 2 - 10 * np.cos(2 * np.pi * (x - 3.14159))) + 10 * np.size(x)


r, dt = timed(lambda: sda(modified_rastrigin, x_init, list(zip(lower, upper)), seed=np.random.RandomState(1234),
                          pure_sa=True, maxiter=1000))
out["config2_shifted_rastrigin50_pure_sa"] = dict(fun=float(r.fun), ncall=int(r.ncall), seconds=dt,
                                                  evals_per_s=r.ncall / dt)
k = gbf.Rastrigin(dimensions=50)
r, dt = timed(lambda: sda(k.fun, None, k.bounds, seed=1234))
out["rastrigin50_ls_on"] = dict(fun=float(r.fun), ncall=int(r.ncall), seconds=dt, evals_per_s=r.ncall / dt)

# config 5 subset, Benchmarker semantics (bench.py:155-216): 5 runs each of a few jobs incl. a failing one
sub = {}
for name, dim in [("Rastrigin", None), ("Ackley01", 10), ("Rosenbrock", None), ("Bukin06", None)]:
    t0 = time.perf_counter()
    calls = 0
    for i in range(5 if name != "Bukin06" else 2):
        algo = refbench.SDAOptimizer()
        np.random.seed(1234 + i)
        algo.prepare(name, getattr(gbf, name), dim)
        try:
            algo.optimize()
        except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
            pass
        calls += algo.nbcall
    dt = time.perf_counter() - t0
    sub["%s%s" % (name, "" if dim is None else "_%d" % dim)] = dict(evals=int(calls), seconds=dt,
                                                                 evals_per_s=calls / dt)
out["suite_subset"] = sub
print(json.dumps(out, indent=1))
