"""Reference statistics of the suite: every job the GPU suite runs at a
dimension <= 10 (the 193 default-dimension jobs and the _5/_10 n-D jobs; the reference needs ~1 s
per 3,000 evaluations at D >= 5, so the D = 20..100 jobs would take days on the build host's
cores), NB_RUNS = 100, run i seeded 1234 + i, through the reference's own
SDAOptimizer (benchmark/bench.py:155-216, :256-333) imported read-only from /root/reference.
Budget 200,000 evaluations per run instead of 1e6 to bound the CPU time (a failing run costs 13 s at
1e6); the GPU side of the comparison uses the same budget.  Build container only.

    python oracle/gen_suite_ref_full.py [n_procs]      -> tests/golden/suite_stats_ref_full.npz
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


def run_job(job):
    name, fname, dim = job
    from oracle import ref_import
    ref_import.load()
    import sdaopt.benchmark.go_benchmark_functions as gbf
    from sdaopt.benchmark import bench as refbench
    klass = getattr(gbf, fname)
    nd = name != fname                     # '<name>_<dim>' jobs pass the dimension (bench.py:108-113)
    succ = np.zeros(NB_RUNS, np.uint8)
    ncall = np.zeros(NB_RUNS, np.int32)
    fval = np.zeros(NB_RUNS)
    t0 = time.time()
    for i in range(NB_RUNS):
        algo = refbench.SDAOptimizer()
        np.random.seed(1234 + i)
        algo.prepare(name, klass, dim if nd else None)
        algo._maxcall = BUDGET
        try:
            with np.errstate(all="ignore"):
                algo.optimize()
        except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
            pass
        except Exception as e:             # the reference raises on all-NaN starts; count as failure
            algo.success = False
            algo.fcall_success = BUDGET
            algo.fsuccess = np.nan
        succ[i] = bool(algo.success)
        ncall[i] = int(algo.fcall_success) if algo.success else BUDGET
        fval[i] = float(algo.fsuccess) if algo.fsuccess is not None else np.nan
    return name, fname, int(algo._k.N), succ, ncall, fval, time.time() - t0


def save(out, t0):
    rows = np.array([dict(job=k, **v) for k, v in sorted(out.items())], dtype=object)
    path = os.path.join(ROOT, "tests", "golden", "suite_stats_ref_full.npz")
    np.savez_compressed(path + ".tmp.npz", rows=rows, budget=BUDGET, nb_runs=NB_RUNS,
                        numpy_version=np.__version__, wall_s=time.time() - t0)
    os.replace(path + ".tmp.npz", path)


if __name__ == "__main__":
    from sdaopt_b200.benchmark.bench import build_job_list
    jobs = [j for j in build_job_list() if j[2] <= 10]
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 1e9     # wall budget: stop handing out jobs
    # cheapest first (median evaluations of the GPU suite run), so a bounded run covers most jobs
    cost = {}
    csv_path = os.path.join(ROOT, "profiles", "r1_suite_full_results.csv")
    if os.path.exists(csv_path):
        import csv
        for r in list(csv.reader(open(csv_path), delimiter=",", quotechar="|"))[1:]:
            cost[r[0]] = min(float(r[9]), BUDGET) * (1.0 if float(r[2].strip("%")) == 100.0 else 3.0)
    jobs.sort(key=lambda j: cost.get(j[0], 1e9))
    out = {}
    t0 = time.time()
    with multiprocessing.Pool(nproc) as pool:
        it = pool.imap_unordered(run_job, jobs, chunksize=1)
        for name, fname, n, succ, ncall, fval, dt in it:
            out[name] = dict(functor=fname, dim=n, success=succ, ncall=ncall, fvalue=fval)
            print("%-28s N=%-4d success %5.1f%% median ncall %9.1f  (%.0f s, total %.0f s)" % (
                name, n, 100.0 * succ.mean(), np.median(ncall), dt, time.time() - t0), flush=True)
            save(out, t0)
            if time.time() - t0 > 60.0 * minutes:
                print("wall budget reached with %d of %d jobs" % (len(out), len(jobs)), flush=True)
                pool.terminate()
                break
    save(out, t0)
    print("wrote", len(out), "jobs in %.0f s" % (time.time() - t0))
