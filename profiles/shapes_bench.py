"""BASELINE configs 3 and 4 at their stated chain counts, timed on the device.

    python profiles/shapes_bench.py [--quick] [--only rastrigin,ackley,schwefel,sutton] > gpurun_out/shapes.jsonl

Config 3: 65,536 independent chains on Rastrigin / Ackley01 / Schwefel26 at D = 10, 100, 1000 with the
L-BFGS-B local search on.  Config 4: Sutton-Chen clusters, N = 38 .. 512 atoms (3N-D), 8,192 chains,
box of constant density 0.7 (N/6)^(1/3).  maxiter is reduced for the large shapes (a
D = 1000 run at maxiter = 1000 is ~10^12 evaluations per batch) and stated per row; the rates do not
depend on it beyond the first steps.

Method: inputs resident in HBM (DeviceBatch), one warm-up launch, then `reps` launches bracketed by CUDA
events on the launching stream with a device synchronise on both sides, fresh Philox keys per launch;
clocks and throttle reasons sampled with nvidia-smi during the timed region.  Per row: evaluations/s
(every objective call the chains made, annealing + L-BFGS-B, i.e. the reference's `ncall`), chain
steps/s, and the algorithmic FP64 rate in the units of SURVEY 8(d) against the DFMA peak measured in
the same process (sda_fp64_peak):
    Rastrigin   37 D per evaluation            Ackley01   36 D + 68        Schwefel26   47 D
    Sutton-Chen 30 N(N-1)/2 + 12 N             sampler    190 per visit sample, (D+1)/2 samples per
                                                           annealing evaluation
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sdaopt_b200 as sb  # noqa: E402
from sdaopt_b200 import _lib  # noqa: E402
from sdaopt_b200.device import DeviceBatch  # noqa: E402
from bench import ClockSampler  # noqa: E402


def objective_flops(name, dim):
    if name == "Rastrigin":
        return 37.0 * dim
    if name == "Ackley01":
        return 36.0 * dim + 68.0
    if name == "Schwefel26":
        return 47.0 * dim
    n = dim // 3
    return 30.0 * n * (n - 1) / 2.0 + 12.0 * n


def run_shape(name, dim, n_chains, maxiter, pure_sa, bounds, reps, peak, warm=True):
    f = sb.DeviceFunctor(name, dimensions=dim)
    b = DeviceBatch(f, bounds if bounds is not None else f.bounds, n_chains, maxiter=maxiter, pure_sa=pure_sa)
    if warm:                       # (a multi-second launch does not need one: bench.py's shapes leg skips it)
        b.set_seeds(np.arange(n_chains) + 17)
        b.launch()
        torch.cuda.synchronize()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    ms, evals, ls_evals = [], 0.0, 0.0
    for r in range(reps):
        b.set_seeds(np.arange(n_chains) + 1_000_003 * (r + 1))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.launch()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        evals += float(b.out["ncall"].sum().item())
        ls_evals += float(b.out["ncall_ls"].sum().item())
        assert int((b.out["status"] != 0).sum().item()) == 0
    clocks = sampler.stop()
    sec = sum(ms) * 1e-3
    chain_evals = evals - ls_evals
    obj = objective_flops(name, dim)
    flops = evals * obj + chain_evals * 190.0 * (dim + 1) / 2.0
    row = {"functor": name, "dim": dim, "chains": n_chains, "maxiter": maxiter, "local_search": not pure_sa,
           "ms_per_launch": float(np.mean(ms)), "launches_per_batch": int(_lib.lib().sda_last_launch_count()),
           "evals_per_s": evals / sec, "chain_steps_per_s": chain_evals / sec,
           "ls_evals_fraction": ls_evals / max(evals, 1.0),
           "algorithmic_tflops": flops / sec / 1e12, "frac_of_dfma_peak": flops / sec / 1e12 / peak,
           "clocks": clocks}
    if name == "SuttonChen":
        n = dim // 3
        row["atoms"] = n
        row["pair_terms_per_s"] = evals / sec * n * (n - 1)          # ordered pairs, as round 1 reported them
        row["unordered_pairs_per_s"] = evals / sec * n * (n - 1) / 2.0
    return row


def quick_shapes():
    """The two large shapes bench.py times as well: config 3 at D = 1000 and config 4 at N = 512."""
    L = 0.7 * (512 / 6.0) ** (1.0 / 3.0)
    return [("Rastrigin", 1000, 65536, 3, False, None),
            ("SuttonChen", 1536, 8192, 2, True, [(-L, L)] * 1536)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="the two shapes bench.py also runs")
    ap.add_argument("--only", default="rastrigin,ackley,schwefel,sutton")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    _lib.require_cuda()
    peak, pms = C.c_double(), C.c_double()
    _lib.check(_lib.lib().sda_fp64_peak(C.byref(peak), C.byref(pms), torch.cuda.current_device()))
    only = a.only.split(",")
    shapes = []
    names = {"rastrigin": "Rastrigin", "ackley": "Ackley01", "schwefel": "Schwefel26"}
    for key, name in names.items():
        if key not in only:
            continue
        for dim, maxiter in ((10, 1000), (100, 120), (1000, 3)):
            if a.quick and not (name == "Rastrigin" and dim == 1000):
                continue
            shapes.append((name, dim, 65536, maxiter, False, None))
    if "sutton" in only:
        for atoms, maxiter, pure in ((38, 40, False), (38, 40, True), (64, 20, True), (128, 8, True), (256, 3, True),
                                     (512, 2, True)):
            if a.quick and atoms != 512:
                continue
            L = 0.7 * (atoms / 6.0) ** (1.0 / 3.0)
            shapes.append(("SuttonChen", 3 * atoms, 8192, maxiter, pure, [(-L, L)] * (3 * atoms)))
    for name, dim, n, maxiter, pure, bounds in shapes:
        row = run_shape(name, dim, n, maxiter, pure, bounds, a.reps, peak.value)
        row["dfma_peak_tflops"] = peak.value
        print(json.dumps(row), flush=True)
        extra = "" if name != "SuttonChen" else "  %7.1f G pair terms/s" % (row["pair_terms_per_s"] / 1e9)
        print("%-11s D=%-5d chains %6d maxiter %4d LS %-3s %9.1f ms  %8.1f M evals/s  %6.2f TFLOP/s = %.3f of DFMA peak%s" % (
            name, dim, n, maxiter, "off" if pure else "on", row["ms_per_launch"], row["evals_per_s"] / 1e6,
            row["algorithmic_tflops"], row["frac_of_dfma_peak"], extra), file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
