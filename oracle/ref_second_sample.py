"""A SECOND, independent sample of the reference's suite runs for a few jobs (seeds 5000 + i instead
of 1234 + i): how far two samples of 100 runs of the reference itself lie apart on a heavy-tailed
job.  Build container only.   python oracle/ref_second_sample.py n_procs job [job ...]"""
import multiprocessing
import os
import sys

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import gen_suite_ref_1e6 as g  # noqa: E402


def run_one(arg):
    job, i = arg
    import sdaopt.benchmark.go_benchmark_functions as gbf
    from sdaopt.benchmark import bench as refbench
    algo = refbench.SDAOptimizer()
    np.random.seed(5000 + i)
    algo.prepare(job, getattr(gbf, job), None)
    try:
        with np.errstate(all="ignore"):
            algo.optimize()
    except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
        pass
    return job, i, bool(algo.success), int(algo.fcall_success) if algo.success else g.BUDGET


if __name__ == "__main__":
    from oracle import ref_import
    ref_import.load()
    nproc, jobs = int(sys.argv[1]), sys.argv[2:]
    with multiprocessing.Pool(nproc) as pool:
        res = pool.map(run_one, [(j, i) for j in jobs for i in range(100)], chunksize=1)
    z = np.load(g.OUT, allow_pickle=True)
    first = {str(r["job"]): r for r in z["rows"]}
    from scipy.stats import mannwhitneyu, ks_2samp
    for j in jobs:
        n2 = np.array([r[3] for r in res if r[0] == j])
        n1 = first[j]["ncall"]
        print("%-14s first sample: success %3d%% median %8.0f | second: success %3d%% median %8.0f | MWU p %.3g KS p %.3g"
              % (j, 100 * first[j]["success"].mean(), np.median(n1), 100 * np.mean([r[2] for r in res if r[0] == j]),
                 np.median(n2), mannwhitneyu(n1, n2).pvalue, ks_2samp(n1, n2).pvalue), flush=True)
        print("   quantiles 10/25/50/75/90 first ", np.percentile(n1, [10, 25, 50, 75, 90]).astype(int))
        print("   quantiles 10/25/50/75/90 second", np.percentile(n2, [10, 25, 50, 75, 90]).astype(int), flush=True)
