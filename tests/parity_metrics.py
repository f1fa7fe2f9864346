"""Parity measurements shared by tests/test_gpu_parity.py (which asserts them against the pinned
values in tests/golden/gpu_pins.json) and profiles/record_parity_pins.py (which records the pins on
a B200).  Everything here calls the CUDA path through the C ABI and compares with the live-reference
fixtures and the CPU oracle."""
import json
import os

import numpy as np

from conftest import GOLD, load_golden, tape_from_seed

PINS_PATH = os.path.join(GOLD, "gpu_pins.json")

# production-RNG chain cases: (functor, dim, maxiter).  D = 1000 and Schwefel26 are BASELINE config 3.
PHILOX_CASES = [("Rastrigin", 50, 30), ("Ackley01", 10, 80), ("Schwefel26", 10, 80), ("Rosenbrock", 2, 150),
                ("Griewank", 7, 80), ("Rastrigin", 100, 8), ("Exponential", 33, 40),
                ("Schwefel26", 100, 8), ("Rastrigin", 1000, 3), ("Ackley01", 1000, 3),
                ("Schwefel26", 1000, 3)]
# BASELINE config 4: atoms, maxiter (pure SA), constant-density box 0.7 (N/6)^(1/3)
SUTTON_CHEN_CASES = [(6, 40), (38, 3), (128, 3)]


def load_pins():
    if not os.path.exists(PINS_PATH):
        return None
    with open(PINS_PATH) as fh:
        return json.load(fh)


def functor(sb, name, params=None):
    return sb.DeviceFunctor(name, params=params if params is not None and len(params) else ())


def rel_dev(e, ref):
    """|e - ref| / max(|ref|, 1): the trajectory energies inherit an ABSOLUTE x error (see the module
    docstring of test_gpu_parity.py), so values below 1 are compared absolutely."""
    e, ref = np.asarray(e, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if e.size == 0:
        return 0.0
    return float(np.max(np.abs(e - ref) / np.maximum(np.abs(ref), 1.0)))


def trajectory_metrics(sb, oracle, fname):
    """Replay of a reference run on its own uniform tape: index of the first decision that differs
    (n_evals if none), largest energy deviation before it, integer outputs, and the same against the
    C oracle."""
    g = load_golden(fname)
    params = g["params"] if len(g["params"]) else None
    f = functor(sb, str(g["functor"]), params)
    n_evals = len(g["energies"])
    tape = tape_from_seed(int(g["seed"]), int(g["tape_len"]) + 4096)
    x0 = g["x0"] if "x0" in g.files else None
    r = sb.sda_batch(f, g["bounds"], 1, x0=x0, tape=tape, maxiter=int(g["maxiter"]), pure_sa=True,
                     trace_len=n_evals)
    dec, e = r.trace_d[0], r.trace_e[0]
    same = dec == g["decisions"]
    first_diff = int(np.argmin(same)) if not same.all() else n_evals
    P = oracle.Problem(str(g["functor"]), g["bounds"], params=params)
    o = oracle.run(P, 1, x0=x0, tape=tape, maxiter=int(g["maxiter"]), pure_sa=True, trace_len=n_evals,
                   tables=(g["temp_table"], g["sigmax_table"]))
    same_o = o["trace_d"][0] == dec
    first_diff_oracle = int(np.argmin(same_o)) if not same_o.all() else n_evals
    return dict(n_evals=n_evals, first_diff=first_diff, first_diff_oracle=first_diff_oracle,
                max_edev=rel_dev(e[:first_diff], g["energies"][:first_diff]),
                status=int(r.status[0]), nit_ok=bool(r.nit[0] == int(g["nit"])),
                ncall_ok=bool(r.ncall[0] == int(g["ncall"])),
                tape_ok=bool(r.tape_used[0] == int(g["tape_len"])),
                x_dev=float(np.max(np.abs(r.x[0] - g["x"]))),
                fun=float(r.fun[0]), fun_ref=float(g["fun"]))


def _chains_vs_oracle(sb, oracle, f, oname, bounds, dim, maxiter, n):
    seeds = np.arange(1000, 1000 + n, dtype=np.uint64) * np.uint64(2654435761)
    n_evals = 2 * dim * (maxiter - 1)
    r = sb.sda_batch(f, bounds, n, seeds=seeds, maxiter=maxiter, pure_sa=True, trace_len=n_evals)
    tables = sb._sda.temperature_tables(maxiter)
    o = oracle.run(oracle.Problem(oname, bounds), n, seeds=seeds, maxiter=maxiter, pure_sa=True,
                   trace_len=n_evals, tables=tables, n_threads=min(8, n))
    same = r.trace_d == o["trace_d"]
    good = same.all(axis=1)
    # energy deviation over the identical prefix of every chain
    first = np.where(good, n_evals, np.argmin(same, axis=1))
    dev = max(rel_dev(r.trace_e[c, :first[c]], o["trace_e"][c, :first[c]]) for c in range(n))
    return dict(counts_ok=bool((r.ncall == o["ncall"]).all() and (r.nit == o["nit"]).all()),
                status_ok=bool((r.status == 0).all()), exact_chains=float(good.mean()),
                frac_same=float(same.mean()), max_edev=float(dev),
                fun_dev=rel_dev(r.fun[good], o["fun"][good]))


def philox_metrics(sb, oracle, name, dim, maxiter, n=None):
    """n chains on the production RNG against the oracle on the same Philox keys."""
    f = sb.DeviceFunctor(name, dimensions=dim)
    n = n or (24 if dim < 1000 else 8)
    return _chains_vs_oracle(sb, oracle, f, name, f.bounds, dim, maxiter, n)


def sutton_chen_box(n_atoms):
    L = 0.7 * (n_atoms / 6.0) ** (1.0 / 3.0)        # suttonchen.py:56-57 at constant density
    return [(-L, L)] * (3 * n_atoms)


def sutton_chen_metrics(sb, oracle, n_atoms, maxiter, n=8):
    return _chains_vs_oracle(sb, oracle, sb.SuttonChen(n_atoms), "SuttonChen", sutton_chen_box(n_atoms),
                             3 * n_atoms, maxiter, n)


def lbfgsb_metrics(sb, oracle, fname):
    """ObjectiveFunWrapper.local_search against the SciPy golden cases: success flag, minimum and
    call count per case, and the call count of the C oracle."""
    rows = load_golden(fname)["rows"]
    out = []
    for r in rows:
        owf = sb.ObjectiveFunWrapper(r["bounds"], r["functor"])
        e, x = owf.local_search(r["x0"])
        if str(r["functor"]) in oracle.FN:
            ok, xo, fo, nfev_o = oracle.local_search(oracle.Problem(r["functor"], r["bounds"]), r["x0"])
        else:                                   # catalogue part 3: no C restatement of the objective
            ok, fo, nfev_o = bool(r["success"]), float(r["f"]), -1
        out.append(dict(functor=str(r["functor"]), dim=int(r["dim"]), success=x is not None,
                        success_ref=bool(r["success"]), f=float(e), f_ref=float(r["f"]),
                        x_dev=0.0 if x is None else float(np.max(np.abs(x - r["x"]))),
                        x_scale=float(1 + np.abs(r["x"]).max()),
                        nfev=int(owf.nb_fun_call), nfev_ref=int(r["nfev"]), nfev_oracle=int(nfev_o),
                        success_oracle=bool(ok), f_oracle=float(fo)))
    return out


# ---- suite statistics against the live reference's own runs ---------------------------------------
def suite_compare(sb, fixture, chunk=64):
    """GPU suite runs (Philox keys 1234 + i, NB_RUNS = 100) against the reference's recorded runs of
    the same jobs (tests/golden/suite_stats_ref_*.npz: seeds 1234 + i through the reference's own
    SDAOptimizer).  The jobs of a chunk are in flight together (BatchPool).  Per job: Fisher's exact
    test on the success counts, Mann-Whitney and Kolmogorov-Smirnov on ncall-to-global (failures at
    the budget), the GPU success count against the reference's Wilson 95 % interval."""
    from scipy.stats import fisher_exact, ks_2samp, mannwhitneyu
    from sdaopt_b200.device import sda_batch_many
    z = load_golden(fixture)
    budget = int(z["budget"])
    rows = list(z["rows"])
    n = 100
    seeds = np.arange(n, dtype=np.uint64) + np.uint64(1234)
    out = []
    for c0 in range(0, len(rows), chunk):
        part = rows[c0:c0 + chunk]
        specs = []
        for r in part:
            f = sb.DeviceFunctor(str(r["functor"]), dimensions=int(r["dim"]))
            specs.append(dict(functor=f, bounds=f.bounds, n_chains=n, seeds=seeds, max_calls=budget,
                              fglob=f.fglob, tol=1e-8))
        for r, g in zip(part, sda_batch_many(specs, maxiter=int(1e6))):
            nr = int(len(r["success"]))
            ks, kr = int(g.hit.sum()), int(np.sum(r["success"]))
            gn = np.where(g.hit != 0, g.ncall_hit, budget)
            rn = np.asarray(r["ncall"], dtype=np.int64)
            row = dict(job=str(r["job"]), dim=int(r["dim"]), ref_runs=nr, gpu_success=ks, ref_success=kr,
                       p_success=float(fisher_exact([[ks, n - ks], [kr, nr - kr]])[1]),
                       gpu_median=float(np.median(gn)), ref_median=float(np.median(rn)))
            lo, hi = wilson(kr, nr)
            row["inside_wilson95"] = bool(lo - 1e-9 <= ks / n <= hi + 1e-9)
            if ks == 0 and kr == 0 or (np.ptp(gn) == 0 and np.ptp(rn) == 0 and gn[0] == rn[0]):
                row["p_mw"] = row["p_ks"] = None           # both samples constant and equal: no test
            else:
                row["p_mw"] = float(mannwhitneyu(gn, rn, alternative="two-sided").pvalue)
                row["p_ks"] = float(ks_2samp(gn, rn).pvalue)
            out.append(row)
    return out


def wilson(k, n, z=1.96):
    p = k / n
    d = 1 + z * z / n
    c = p + z * z / (2 * n)
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n))
    return (c - h) / d, (c + h) / d


def suite_summary(rows, alpha=0.05):
    """Family-wise verdict over the jobs.  (1) Bonferroni at family-wise level alpha over every test
    that was actually made: no p below alpha / n_tests.  (2) Calibration: the per-test rejections at
    0.05 must be what identical distributions produce, i.e. their count must not exceed the 99.9 %
    quantile of Binomial(n_tests, 0.05) (the three tests of a job are correlated, which only makes
    this stricter in n_jobs terms).  (3) Location: geometric mean over the jobs of the ratio of the
    median ncall (GPU / reference) within 5 % of 1.  (4) Success rate inside the reference's Wilson
    95 % interval for at least 90 % of the jobs."""
    from scipy.stats import binom
    ps = [r[k] for r in rows for k in ("p_success", "p_mw", "p_ks") if r[k] is not None]
    n_tests = len(ps)
    thr = alpha / max(n_tests, 1)
    worst = sorted(((min(p for p in (r["p_success"], r["p_mw"], r["p_ks"]) if p is not None), r["job"])
                    for r in rows))[:5]
    n_rej = int(sum(p < 0.05 for p in ps))
    ratios = [r["gpu_median"] / r["ref_median"] for r in rows if r["ref_median"] > 0 and r["gpu_median"] > 0]
    return dict(jobs=len(rows), tests=n_tests, bonferroni_threshold=thr, min_p=float(min(ps)),
                family_wise_ok=bool(min(ps) >= thr), worst=worst,
                rejections_at_05=n_rej, rejections_allowed=int(binom.ppf(0.999, n_tests, 0.05)),
                geo_mean_median_ratio=float(np.exp(np.mean(np.log(ratios)))),
                inside_wilson95=float(np.mean([r["inside_wilson95"] for r in rows])))
