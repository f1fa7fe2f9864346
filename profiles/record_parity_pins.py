"""Record today's parity figures of the CUDA path on a B200 -> gpurun_out/gpu_pins.json, to be
committed as tests/golden/gpu_pins.json: per trajectory fixture the index of the first decision that
differs from the live reference and the largest energy deviation before it; per production-RNG case
the fraction of chains whose whole decision sequence equals the oracle's; per L-BFGS-B golden file
the number of cases with SciPy's exact call count.  tests/test_gpu_parity.py asserts these values
(>= for the counts, <= 4x for the deviations), so a regression from, say, 92,085 identical
evaluations to 4,001 fails.    python profiles/record_parity_pins.py"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_metrics as pm  # noqa: E402
import sdaopt_b200 as sb  # noqa: E402
from oracle import sda_oracle as oracle  # noqa: E402

oracle.build()
pins = {"traj": {}, "philox": {}, "sutton_chen": {}, "lbfgsb": {}}
for path in sorted(glob.glob(os.path.join(pm.GOLD, "traj_*.npz"))):
    name = os.path.basename(path)
    m = pm.trajectory_metrics(sb, oracle, name)
    pins["traj"][name] = m
    print(name, m, flush=True)
for name, dim, maxiter in pm.PHILOX_CASES:
    m = pm.philox_metrics(sb, oracle, name, dim, maxiter)
    pins["philox"]["%s-%d-%d" % (name, dim, maxiter)] = m
    print(name, dim, maxiter, m, flush=True)
for n_atoms, maxiter in pm.SUTTON_CHEN_CASES:
    m = pm.sutton_chen_metrics(sb, oracle, n_atoms, maxiter)
    pins["sutton_chen"]["%d-%d" % (n_atoms, maxiter)] = m
    print("SuttonChen", n_atoms, maxiter, m, flush=True)
for fname in ("lbfgsb_scipy.npz", "lbfgsb_scipy_large.npz", "lbfgsb_scipy_cat.npz"):
    rows = pm.lbfgsb_metrics(sb, oracle, fname)
    pins["lbfgsb"][fname] = dict(cases=len(rows), same_nfev_as_scipy=sum(r["nfev"] == r["nfev_ref"] for r in rows),
                                 same_nfev_as_oracle=sum(r["nfev"] == r["nfev_oracle"] for r in rows),
                                 success_match=sum(r["success"] == r["success_ref"] for r in rows), rows=rows)
    print(fname, {k: v for k, v in pins["lbfgsb"][fname].items() if k != "rows"}, flush=True)
    for r in rows:
        print("   ", r, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_pins.json"), "w") as fh:
    json.dump(pins, fh, indent=1, sort_keys=True)
print("wrote gpurun_out/gpu_pins.json")
