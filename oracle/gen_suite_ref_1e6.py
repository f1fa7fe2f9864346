"""Reference statistics of the suite at the REAL budget (1e6 evaluations per
run, bench.py:35-38) for the jobs tests/golden/suite_stats_ref_full.npz (budget 200,000) does not
hold or cannot discriminate: every partial-success job, every never-succeed job, and the n-D jobs at
D = 70..100.  Same protocol as oracle/gen_suite_ref_full.py: the live reference's SDAOptimizer
(benchmark/bench.py:155-216, :256-333) imported read-only from /root/reference, run i seeded with
`np.random.seed(1234 + i)`; parallel over the runs.  NB_RUNS = 100 except the D >= 70 jobs, which
get NB_RUNS_ND = 20 runs each (a D = 100 run is ~2 minutes of one core); the fixture records the
number of runs per job.  Build container only.

    python oracle/gen_suite_ref_1e6.py n_procs [minutes]   -> tests/golden/suite_stats_ref_1e6.npz
"""
import multiprocessing
import os
import sys
import time
import warnings

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")           # one core per worker process

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

BUDGET = 1000000
NB_RUNS = 100
NB_RUNS_ND = 20
OUT = os.path.join(ROOT, "tests", "golden", "suite_stats_ref_1e6.npz")

# cheapest first, so that a bounded run records as many whole jobs as possible
PLAN = (
    ["Whitley", "Rana", "Ratkowsky01", "Griewank", "BiggsExp05", "EggHolder", "Trefethen",
     "SineEnvelope", "Mishra03", "XinSheYang01", "Stochastic"],
    ["PowerSum", "NewFunction01", "NewFunction02", "DeVilliersGlasser02", "Thurber", "Cola"],
    ["Bukin06", "CrownedCross", "Csendes", "Damavandi", "Deceptive", "Meyer", "Mishra04", "Mishra06",
     "Shubert03", "Weierstrass", "Zimmerman"],
    ["%s_%d" % (f, d) for d in (70, 80, 90, 100) for f in ("Rosenbrock", "Rastrigin", "Ackley01")],
    ["Rosenbrock_%d" % d for d in (20, 30, 40, 50, 60)],
)


def n_runs(job):
    tail = job.rsplit("_", 1)[-1]
    return NB_RUNS_ND if tail.isdigit() and int(tail) >= 70 else NB_RUNS


def run_one(arg):
    job, i = arg
    from oracle import ref_import
    ref_import.load()
    import sdaopt.benchmark.go_benchmark_functions as gbf
    from sdaopt.benchmark import bench as refbench
    fname, _, dim = job.rpartition("_")
    nd = dim.isdigit() and hasattr(gbf, fname)
    klass = getattr(gbf, fname if nd else job)
    algo = refbench.SDAOptimizer()
    np.random.seed(1234 + i)
    algo.prepare(job, klass, int(dim) if nd else None)
    algo._maxcall = BUDGET
    t0 = time.time()
    try:
        with np.errstate(all="ignore"):
            algo.optimize()
    except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
        pass
    except Exception:                        # the reference raises on all-NaN starts: a failed run
        algo._success = False
        algo._fsuccess = np.nan
    ok = bool(algo.success)
    return (job, i, ok, int(algo.fcall_success) if ok else BUDGET,
            float(algo.fsuccess) if algo.fsuccess is not None else np.nan, int(algo._k.N),
            (fname if nd else job), int(algo.nbcall), time.time() - t0)


def save(rows, t0):
    out = np.array([rows[k] for k in sorted(rows)], dtype=object)
    np.savez_compressed(OUT + ".tmp.npz", rows=out, budget=BUDGET, numpy_version=np.__version__,
                        wall_s=time.time() - t0)
    os.replace(OUT + ".tmp.npz", OUT)


if __name__ == "__main__":
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9
    rows = {}
    if os.path.exists(OUT):
        rows = {str(r["job"]): r for r in np.load(OUT, allow_pickle=True)["rows"]}
    jobs = [j for grp in PLAN for j in grp if j not in rows]
    args = [(j, i) for j in jobs for i in range(n_runs(j))]
    t0 = time.time()
    partial = {}
    with multiprocessing.Pool(nproc) as pool:
        for r in pool.imap_unordered(run_one, args, chunksize=1):
            partial.setdefault(r[0], []).append(r)
            rs = partial[r[0]]
            print("  %s run %d: %s ncall %d (%.1f s)" % (r[0], r[1], "hit" if r[2] else "miss", r[3], r[8]),
                  file=sys.stderr, flush=True)
            if len(rs) == n_runs(r[0]):
                rs.sort()
                rows[r[0]] = dict(job=r[0], functor=r[6], dim=r[5], nb_runs=len(rs),
                                  success=np.array([q[2] for q in rs], np.uint8),
                                  ncall=np.array([q[3] for q in rs], np.int32),
                                  fvalue=np.array([q[4] for q in rs]),
                                  nbcall=np.array([q[7] for q in rs], np.int64),
                                  cpu_s=np.array([q[8] for q in rs]))
                print("%-22s N=%-4d runs %3d success %5.1f%%  median ncall %9.1f  cpu %.0f s  (t = %.0f s)" % (
                    r[0], r[5], len(rs), 100.0 * rows[r[0]]["success"].mean(),
                    np.median(rows[r[0]]["ncall"]), rows[r[0]]["cpu_s"].sum(), time.time() - t0), flush=True)
                save(rows, t0)
            if time.time() - t0 > 60.0 * minutes:
                print("wall budget reached with %d jobs recorded" % len(rows), flush=True)
                pool.terminate()
                break
    print("wrote", len(rows), "jobs in %.0f s" % (time.time() - t0))
