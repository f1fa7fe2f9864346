"""Reference statistics for n-D suite jobs above dimension 10 (bench.py:103-118: Ackley01, Exponential,
Rastrigin, Rosenbrock, Schwefel01 at D = 20..100), appended to tests/golden/suite_stats_ref_full.npz.
Same protocol as oracle/gen_suite_ref_full.py (live reference SDAOptimizer, NB_RUNS = 100,
`np.random.seed(1234 + i)` before run i, budget 200,000 evaluations) but parallel over the runs, so
that a D = 50 job (60,000 evaluations per run at ~3,000 evaluations/s) finishes in minutes on the
build host.  Build container only.

    python oracle/gen_suite_ref_nd.py n_procs Rastrigin_50 Ackley01_20 ...
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


def run_one(arg):
    job, i = arg
    fname, dim = job.rsplit("_", 1)
    from oracle import ref_import
    ref_import.load()
    import sdaopt.benchmark.go_benchmark_functions as gbf
    from sdaopt.benchmark import bench as refbench
    algo = refbench.SDAOptimizer()
    np.random.seed(1234 + i)
    algo.prepare(job, getattr(gbf, fname), int(dim))
    algo._maxcall = BUDGET
    try:
        with np.errstate(all="ignore"):
            algo.optimize()
    except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
        pass
    ok = bool(algo.success)
    return (job, i, ok, int(algo.fcall_success) if ok else BUDGET,
            float(algo.fsuccess) if algo.fsuccess is not None else np.nan, int(algo._k.N))


def merge(results, jobs):
    path = os.path.join(ROOT, "tests", "golden", "suite_stats_ref_full.npz")
    z = np.load(path, allow_pickle=True)
    rows = {str(r["job"]): r for r in z["rows"]}
    for job in jobs:
        rs = sorted(r for r in results if r[0] == job)
        if [r[1] for r in rs] != list(range(NB_RUNS)):
            continue                                   # incomplete job: not recorded
        rows[job] = dict(job=job, functor=job.rsplit("_", 1)[0], dim=rs[0][5],
                         success=np.array([r[2] for r in rs], np.uint8),
                         ncall=np.array([r[3] for r in rs], np.int32),
                         fvalue=np.array([r[4] for r in rs]))
        print("%-16s N=%-4d success %5.1f%%  median ncall %9.1f" % (
            job, rs[0][5], 100.0 * rows[job]["success"].mean(), np.median(rows[job]["ncall"])), flush=True)
    out = np.array([rows[k] for k in sorted(rows)], dtype=object)
    np.savez_compressed(path + ".tmp.npz", rows=out, budget=int(z["budget"]), nb_runs=int(z["nb_runs"]),
                        numpy_version=str(z["numpy_version"]), wall_s=float(z["wall_s"]))
    os.replace(path + ".tmp.npz", path)


if __name__ == "__main__":
    t0 = time.time()
    nproc, jobs = int(sys.argv[1]), sys.argv[2:]
    args = [(j, i) for j in jobs for i in range(NB_RUNS)]
    results, done = [], set()
    with multiprocessing.Pool(nproc) as pool:
        for r in pool.imap_unordered(run_one, args, chunksize=1):
            results.append(r)
            n = sum(1 for q in results if q[0] == r[0])
            if n == NB_RUNS and r[0] not in done:      # a job is complete: record it right away
                done.add(r[0])
                merge(results, [r[0]])
                print("   (%.0f s)" % (time.time() - t0), flush=True)
