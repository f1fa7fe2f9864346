"""GPU suite runs against the reference's own statistics, job by job (north star: success rate and
ncall-to-global of NB_RUNS = 100 runs inside the reference's confidence limits).
    python profiles/suite_vs_reference.py [suite_stats_ref_full.npz | suite_stats_ref_1e6.npz]
Reference arrays: oracle/gen_suite_ref_full.py / gen_suite_ref_nd.py (budget 200,000) and
oracle/gen_suite_ref_1e6.py (budget 1e6) -- the live reference, seeds 1234+i.  GPU: 100 Philox chains
per job, keys 1234+i, same budget.  Two independent samples per job -> Fisher's exact test on the
success counts, Mann-Whitney and Kolmogorov-Smirnov on ncall (tests/parity_metrics.py, the same code
the GPU test asserts on)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_metrics as pm  # noqa: E402
import sdaopt_b200 as sb  # noqa: E402

if __name__ == "__main__":
    fixture = sys.argv[1] if len(sys.argv) > 1 else "suite_stats_ref_full.npz"
    t0 = time.time()
    rows = pm.suite_compare(sb, os.path.basename(fixture))
    s = pm.suite_summary(rows)
    print("%-28s %3s %4s %9s %9s %9s %11s %11s %9s %9s" % ("job", "D", "runs", "gpu succ", "ref succ", "p(succ)",
                                                          "gpu median", "ref median", "p(MWU)", "p(KS)"))
    fmt = lambda p: "    -    " if p is None else "%9.2g" % p
    for r in rows:
        ps = [p for p in (r["p_success"], r["p_mw"], r["p_ks"]) if p is not None]
        flag = "" if min(ps) >= s["bonferroni_threshold"] else "  <-- differs"
        print("%-28s %3d %4d %8.0f%% %8.0f%% %s %11.1f %11.1f %s %s%s" % (
            r["job"], r["dim"], r["ref_runs"], r["gpu_success"], 100.0 * r["ref_success"] / r["ref_runs"],
            fmt(r["p_success"]), r["gpu_median"], r["ref_median"], fmt(r["p_mw"]), fmt(r["p_ks"]), flag))
    print("summary:", s)
    print("%.0f s" % (time.time() - t0))
