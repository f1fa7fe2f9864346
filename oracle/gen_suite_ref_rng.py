"""Reference statistics for the two suite jobs whose objective draws random numbers inside fun()
(Stochastic, go_funcs_S.py:1274-1280; XinSheYang01, go_funcs_X.py:47-51), appended to
tests/golden/suite_stats_ref_full.npz.  Same protocol as oracle/gen_suite_ref_full.py (live reference
SDAOptimizer, NB_RUNS = 100, `np.random.seed(1234 + i)` before run i -- which also seeds the draws
inside fun(), as in the reference harness -- budget 200,000 evaluations), but parallel over the
runs: a run of these noisy functions needs up to 45 s on the build host.  Build container only.

    python oracle/gen_suite_ref_rng.py [n_procs]
    python oracle/gen_suite_ref_rng.py --stateless-probe     (design evidence, see below)

Result (numpy 2.3.5, scipy 1.18.1): Stochastic 29 % success, XinSheYang01 64 % (budget 200,000).

Why the device noise carries the call number: `--stateless-probe` runs the live reference on
Stochastic with fun() replaced by noise that is a hash of the point alone -> 0 of 32 runs succeed
(29 % with the reference's own per-call draws).  The successes come from L-BFGS-B line searches that
request many nearly identical points next to the optimum, each with fresh factors; noise that is a
pure function of the point removes exactly that, so the CUDA functors address their Philox noise by
(point, call number).
"""
import multiprocessing
import os
import sys
import time
import warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

BUDGET = 200000
NB_RUNS = 100
JOBS = ("Stochastic", "XinSheYang01")


def run_one(arg):
    name, i = arg
    from oracle import ref_import
    ref_import.load()
    import sdaopt.benchmark.go_benchmark_functions as gbf
    from sdaopt.benchmark import bench as refbench
    algo = refbench.SDAOptimizer()
    np.random.seed(1234 + i)
    algo.prepare(name, getattr(gbf, name), None)
    algo._maxcall = BUDGET
    try:
        with np.errstate(all="ignore"):
            algo.optimize()
    except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
        pass
    ok = bool(algo.success)
    return (name, i, ok, int(algo.fcall_success) if ok else BUDGET,
            float(algo.fsuccess) if algo.fsuccess is not None else np.nan, int(algo._k.N))


def merge(results):
    path = os.path.join(ROOT, "tests", "golden", "suite_stats_ref_full.npz")
    z = np.load(path, allow_pickle=True)
    rows = {str(r["job"]): r for r in z["rows"]}
    for name in JOBS:
        rs = sorted(r for r in results if r[0] == name)
        assert [r[1] for r in rs] == list(range(NB_RUNS)), name
        rows[name] = dict(job=name, functor=name, dim=rs[0][5],
                          success=np.array([r[2] for r in rs], np.uint8),
                          ncall=np.array([r[3] for r in rs], np.int32),
                          fvalue=np.array([r[4] for r in rs]))
    out = np.array([rows[k] for k in sorted(rows)], dtype=object)
    np.savez_compressed(path + ".tmp.npz", rows=out, budget=int(z["budget"]), nb_runs=int(z["nb_runs"]),
                        numpy_version=str(z["numpy_version"]), wall_s=float(z["wall_s"]))
    os.replace(path + ".tmp.npz", path)
    for name in JOBS:
        r = rows[name]
        print("%-14s success %5.1f%%  median ncall %9.1f" % (name, 100.0 * r["success"].mean(),
                                                             np.median(r["ncall"])))


_M = (1 << 64) - 1


def _mix(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def stateless_probe_run(i):
    from oracle import ref_import
    ref_import.load()
    import sdaopt.benchmark.go_benchmark_functions as gbf
    from sdaopt.benchmark import bench as refbench
    algo = refbench.SDAOptimizer()
    np.random.seed(1234 + i)
    algo.prepare("Stochastic", gbf.Stochastic, None)
    algo._maxcall = BUDGET
    k = algo._k

    def fun(x, *a):
        k.nfev += 1
        b = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
        h = 0
        for j, v in enumerate(b):
            h = (h + _mix((int(v) + (j + 1) * 0x9E3779B97F4A7C15) & _M)) & _M
        eps = np.array([(_mix(h ^ ((j * 0xD1B54A32D192ED03) & _M)) >> 11) / 9007199254740992.0
                        for j in range(len(b))])
        return np.sum(eps * np.abs(x - 1.0 / np.arange(1, len(b) + 1)))
    k.fun = fun
    try:
        with np.errstate(all="ignore"):
            algo.optimize()
    except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
        pass
    return bool(algo.success)


if __name__ == "__main__":
    if "--stateless-probe" in sys.argv:
        with multiprocessing.Pool(8) as pool:
            ok = pool.map(stateless_probe_run, range(32), chunksize=1)
        print("point-addressed noise in the live reference: %d of %d runs succeed" % (sum(ok), len(ok)))
        sys.exit(0)
    t0 = time.time()
    args = [(n, i) for i in range(NB_RUNS) for n in JOBS]
    with multiprocessing.Pool(int(sys.argv[1]) if len(sys.argv) > 1 else 4) as pool:
        results = list(pool.imap_unordered(run_one, args, chunksize=1))
    merge(results)
    print("%.0f s" % (time.time() - t0))
