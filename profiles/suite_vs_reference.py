"""GPU suite runs against the reference's own statistics, job by job (north star: success rate and
ncall-to-global of NB_RUNS = 100 runs inside the reference's confidence limits).
    python profiles/suite_vs_reference.py [tests/golden/suite_stats_ref_full.npz]
Reference arrays: oracle/gen_suite_ref_full.py (live reference, seeds 1234+i, budget 200,000).
GPU: 100 Philox chains per job, keys 1234+i, same budget.  Two independent samples per job ->
Fisher's exact test on the success counts, Mann-Whitney and Kolmogorov-Smirnov on ncall."""
import os
import sys
import time

import numpy as np
from scipy.stats import fisher_exact, ks_2samp, mannwhitneyu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdaopt_b200 as sb  # noqa: E402


def compare(path):
    z = np.load(path, allow_pickle=True)
    budget, n = int(z["budget"]), int(z["nb_runs"])
    seeds = np.arange(n, dtype=np.uint64) + np.uint64(1234)
    out = []
    for r in z["rows"]:
        f = sb.DeviceFunctor(r["functor"], dimensions=int(r["dim"]))
        g = sb.sda_batch(f, f.bounds, n, seeds=seeds, maxiter=int(1e6), max_calls=budget,
                         fglob=f.fglob, tol=1e-8)
        ks, kr = int(g.hit.sum()), int(r["success"].sum())
        p_succ = fisher_exact([[ks, n - ks], [kr, n - kr]])[1]
        gn = np.where(g.hit != 0, g.ncall_hit, budget)
        if ks == 0 and kr == 0:
            p_mw = p_ks = 1.0
        else:
            p_mw = mannwhitneyu(gn, r["ncall"], alternative="two-sided").pvalue
            p_ks = ks_2samp(gn, r["ncall"]).pvalue
        out.append(dict(job=r["job"], dim=int(r["dim"]), gpu_success=ks, ref_success=kr, p_success=p_succ,
                        gpu_median=float(np.median(gn)), ref_median=float(np.median(r["ncall"])),
                        p_mw=p_mw, p_ks=p_ks))
    return out


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "suite_stats_ref_full.npz")
    t0 = time.time()
    rows = compare(path)
    print("%-28s %3s %9s %9s %9s %11s %11s %9s %9s" % ("job", "D", "gpu succ", "ref succ", "p(succ)",
                                                      "gpu median", "ref median", "p(MWU)", "p(KS)"))
    bad = 0
    for r in rows:
        flag = "" if min(r["p_success"], r["p_mw"], r["p_ks"]) > 1e-4 else "  <-- differs"
        bad += bool(flag)
        print("%-28s %3d %8d%% %8d%% %9.2g %11.1f %11.1f %9.2g %9.2g%s" % (
            r["job"], r["dim"], r["gpu_success"], r["ref_success"], r["p_success"], r["gpu_median"],
            r["ref_median"], r["p_mw"], r["p_ks"], flag))
    print("%d jobs, %d differ at p < 1e-4, %.0f s" % (len(rows), bad, time.time() - t0))
