#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE Python reference (build container only).

    python oracle/gen_golden.py            # re-executes itself with numpy's AVX-512 loops disabled

Everything written here comes from running the unmodified reference under /root/reference
(oracle/ref_import.py) -- the C oracle and the CUDA path are checked against these files; they are
never used to produce them.

Two numpy configurations are recorded:
  * "libm"  : NPY_DISABLE_CPU_FEATURES=AVX512* so that numpy's scalar/array exp and log go through
              glibc libm.  In this configuration the reference is bit-reproducible by a C program,
              and the oracle is pinned bit-for-bit (energies, decisions, x) on these fixtures.
  * "as-run": numpy's default dispatch on the build host (AVX-512 exp/log differ from libm by 1 ulp
              in ~4.6 % / 0.1 % of arguments).  Only summary numbers are stored, to document that
              the reference's trajectory itself depends on the host's numpy dispatch.
Uniform tapes are NOT stored: they are numpy RandomState(seed).random_sample() streams, regenerated
by the tests from the seed (MT19937 is part of numpy on every box); only their length is stored.
"""
import os
import sys
import warnings

_AVX512 = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX512_KNL AVX512_KNM"

if __name__ == "__main__" and os.environ.get("SDA_GOLDEN_CHILD") is None:
    env = dict(os.environ, SDA_GOLDEN_CHILD="libm", NPY_DISABLE_CPU_FEATURES=_AVX512,
               PYTHONDONTWRITEBYTECODE="1")
    rc = os.spawnve(os.P_WAIT, sys.executable, [sys.executable] + sys.argv, env)
    if rc == 0:
        env2 = dict(os.environ, SDA_GOLDEN_CHILD="asrun", PYTHONDONTWRITEBYTECODE="1")
        rc = os.spawnve(os.P_WAIT, sys.executable, [sys.executable] + sys.argv, env2)
    sys.exit(rc)

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import scipy  # noqa: E402

from oracle import ref_import  # noqa: E402

ref = ref_import.load()
from sdaopt import _sda as R  # noqa: E402
import sdaopt.benchmark.go_benchmark_functions as gbf  # noqa: E402
from sdaopt.benchmark import suttonchen  # noqa: E402
from sdaopt.benchmark import bench as refbench  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
MODE = os.environ["SDA_GOLDEN_CHILD"]

DEC_IMPROVE, DEC_NEW_BEST, DEC_METRO, DEC_REJECT = 1, 2, 3, 4


def k1_tables(maxiter, T0=5230., qv=2.62, T_restart=0.1):
    """T_i and sigmax(T_i) with the reference's own expressions (_sda.py:396-402, :76-86)."""
    t1 = np.exp((qv - 1) * np.log(2.0)) - 1.0
    Ts, Ss = [], []
    for i in range(int(maxiter)):
        s = float(i) + 2.0
        t2 = np.exp((qv - 1) * np.log(s)) - 1.0
        temperature = T0 * t1 / t2
        factor1 = np.exp(np.log(temperature) / (qv - 1.0))
        factor2 = np.exp((4.0 - qv) * np.log(qv - 1.0))
        factor3 = np.exp((2.0 - qv) * np.log(2.0) / (qv - 1.0))
        factor4 = np.sqrt(np.pi) * factor1 * factor2 / (factor3 * (3.0 - qv))
        factor5 = 1.0 / (qv - 1.0) - 0.5
        d1 = 2.0 - factor5
        factor6 = np.pi * (1.0 - factor5) / np.sin(np.pi * (1.0 - factor5)) / np.exp(
            R.gammaln(d1))
        Ts.append(temperature)
        Ss.append(np.exp(-(qv - 1.0) * np.log(factor6 / factor4) / (3.0 - qv)))
        if temperature < T_restart:
            break
    return np.array(Ts), np.array(Ss)


def modified_rastrigin(x):
    """README.md:61-71 without the bookkeeping globals."""
    shift = 3.14159
    return np.sum((x - shift) ** 2 - 10 * np.cos(2 * np.pi * (x - shift))) + 10 * np.size(x)


def record_run(func, x0, bounds, seed, maxiter, pure_sa):
    """Run the reference SDARunner, recording energies / decisions of every counted evaluation
    made by the Markov chain (not by the local search)."""
    rs = ref_import.RecordingRandomState(seed)
    holder = {}
    evals = []

    def f(x):
        gr = holder.get("gr")
        if gr is not None:
            evals.append((float(func(x)), gr.es.current_energy, gr.es.ebest, holder["in_ls"]))
            return evals[-1][0]
        return func(x)

    gr = R.SDARunner(f, x0, bounds, rs, None, maxsteps=maxiter, pure_sa=pure_sa)
    holder["in_ls"] = False
    holder["gr"] = gr
    orig_ls = gr.owf.local_search

    def ls(x, maxlsiter=None):
        holder["in_ls"] = True
        try:
            return orig_ls(x)
        finally:
            holder["in_ls"] = False
    gr.owf.local_search = ls
    gr.search()
    res = gr.result
    chain = [(i, e) for i, e in enumerate(evals) if not e[3]]
    energies = np.array([e[0] for _, e in chain])
    dec = np.zeros(len(chain), np.uint8)
    for n, (i, (e, ecur, ebest, _)) in enumerate(chain):
        if e < ecur:
            dec[n] = DEC_NEW_BEST if e < ebest else DEC_IMPROVE
        else:
            nxt_cur = evals[i + 1][1] if i + 1 < len(evals) else gr.es.current_energy
            dec[n] = DEC_METRO if (nxt_cur == e and e != ecur) else DEC_REJECT
    return dict(energies=energies, decisions=dec, x=np.array(res.x), fun=float(res.fun),
                nit=int(res.nit), ncall=int(res.ncall), tape_len=len(rs.tape),
                n_ls_evals=int(sum(1 for e in evals if e[3])))


def gen_trajectories():
    out = {}
    # BASELINE config 2 (README.md:41-80): x0 given, seed=RandomState(1234), pure_sa
    D = 50
    np.random.seed(1234)
    lower = np.array([-5.12] * D)
    upper = np.array([10.24] * D)
    x_init = lower + np.random.rand(D) * (upper - lower)
    bounds = list(zip(lower, upper))
    for maxiter in (50, 1000):
        r = record_run(modified_rastrigin, x_init, bounds, 1234, maxiter, True)
        T, S = k1_tables(maxiter)
        np.savez_compressed(os.path.join(GOLD, "traj_readme_rastrigin50_it%d.npz" % maxiter),
                            functor="ShiftedRastrigin", params=[3.14159], bounds=np.array(bounds),
                            x0=x_init, seed=1234, maxiter=maxiter, temp_table=T, sigmax_table=S,
                            **r)
        out[maxiter] = r
        print("traj README D=50 maxiter", maxiter, "fun", r["fun"], "ncall", r["ncall"], "tape",
              r["tape_len"])
    # catalogue functions, x0=None (init draws from the tape), short pure-SA chains
    for name, D, maxiter, seed in [("Rastrigin", 10, 120, 11), ("Ackley01", 10, 120, 12),
                                   ("Schwefel26", 10, 120, 13), ("Rosenbrock", 2, 300, 14),
                                   ("Schwefel01", 5, 150, 15), ("Exponential", 10, 100, 16),
                                   ("Griewank", 7, 100, 17), ("Sphere", 33, 40, 18),
                                   ("Rastrigin", 100, 12, 19)]:
        k = getattr(gbf, name)(dimensions=D)
        r = record_run(k.fun, None, k.bounds, seed, maxiter, True)
        T, S = k1_tables(maxiter)
        np.savez_compressed(os.path.join(GOLD, "traj_%s_d%d.npz" % (name, D)), functor=name,
                            params=[], bounds=np.array(k.bounds, dtype=float), seed=seed,
                            maxiter=maxiter, temp_table=T, sigmax_table=S, **r)
        print("traj", name, D, "fun", r["fun"], "ncall", r["ncall"], "tape", r["tape_len"])


def gen_ls_runs():
    """Whole sda() runs with local search on (SURVEY App. C.4 anchors)."""
    rows = []
    for name, D, seed in [("Rastrigin", 2, 1234), ("Rosenbrock", 2, 1234), ("Rastrigin", 10, 1234),
                          ("Ackley01", 10, 1234), ("Schwefel26", 10, 1234), ("Rastrigin", 8, 77),
                          ("Schwefel01", 5, 5), ("Exponential", 10, 6)]:
        k = getattr(gbf, name)(dimensions=D)
        r = record_run(k.fun, None, k.bounds, seed, 1000, False)
        rows.append(dict(functor=name, dim=D, seed=seed, fun=r["fun"], nit=r["nit"],
                         ncall=r["ncall"], tape_len=r["tape_len"], x=r["x"],
                         n_chain_evals=len(r["energies"]), n_ls_evals=r["n_ls_evals"],
                         bounds=np.array(k.bounds, dtype=float)))
        print("lsrun", name, D, seed, "fun", r["fun"], "ncall", r["ncall"], "tape", r["tape_len"])
    np.savez_compressed(os.path.join(GOLD, "ls_runs.npz"), rows=np.array(rows, dtype=object))


def gen_lbfgsb():
    """SciPy L-BFGS-B through the reference's ObjectiveFunWrapper.local_search (_sda.py:347-351):
    the fun_and_grad request points, final x/f, success, nb_fun_call."""
    rows = []
    for name, D in [("Rastrigin", 10), ("Rastrigin", 50), ("Ackley01", 10), ("Schwefel26", 10),
                    ("Rosenbrock", 10), ("Rosenbrock", 2), ("Schwefel01", 5), ("Exponential", 10),
                    ("Sphere", 5)]:
        k = getattr(gbf, name)(dimensions=D)
        lo = np.array([b[0] for b in k.bounds], dtype=float)
        up = np.array([b[1] for b in k.bounds], dtype=float)
        rs = np.random.RandomState(7)
        for t in range(6):
            x0 = lo + rs.random_sample(D) * (up - lo)
            if t % 3 == 0:
                x0 = np.clip(np.array(k.global_optimum[0], dtype=float) + rs.normal(size=D) * 0.3,
                             lo, up)
            if t == 5:
                x0[0] = lo[0]
                x0[-1] = up[-1]          # start on the box
            pts = []

            def f(x, _p=pts, _k=k):
                _p.append(np.array(x))
                return _k.fun(x)
            owf = R.ObjectiveFunWrapper(k.bounds, f)
            e, x = owf.local_search(x0)
            rows.append(dict(functor=name, dim=D, bounds=np.array(k.bounds, dtype=float), x0=x0,
                             success=x is not None, f=float(e),
                             x=np.array(x0 if x is None else x), nfev=owf.nb_fun_call,
                             request_points=np.array(pts[::2 * D + 1])))
    np.savez_compressed(os.path.join(GOLD, "lbfgsb_scipy.npz"), rows=np.array(rows, dtype=object),
                        scipy_version=scipy.__version__)
    print("lbfgsb cases", len(rows))


def gen_lbfgsb_large():
    """The same at the dimensions of BASELINE config 3 (D = 100 and D = 1000), where the device
    keeps the L-BFGS-B workspace in global memory instead of shared memory.  Only the first request
    points are stored (a D = 1000 case makes a few hundred requests of 1000 doubles each)."""
    rows = []
    for name, D, n_cases in [("Rastrigin", 100, 3), ("Ackley01", 100, 3), ("Schwefel26", 100, 3),
                             ("Rastrigin", 1000, 2), ("Ackley01", 1000, 2), ("Schwefel26", 1000, 2)]:
        k = getattr(gbf, name)(dimensions=D)
        lo = np.array([b[0] for b in k.bounds], dtype=float)
        up = np.array([b[1] for b in k.bounds], dtype=float)
        rs = np.random.RandomState(70 + D)
        for t in range(n_cases):
            x0 = lo + rs.random_sample(D) * (up - lo)
            if t == 0:
                x0 = np.clip(np.array(k.global_optimum[0], dtype=float) + rs.normal(size=D) * 0.3,
                             lo, up)
            if t == 2:
                x0[0] = lo[0]
                x0[-1] = up[-1]          # start on the box
            pts = []

            def f(x, _p=pts, _k=k):
                _p.append(np.array(x))
                return _k.fun(x)
            owf = R.ObjectiveFunWrapper(k.bounds, f)
            e, x = owf.local_search(x0)
            req = pts[::2 * D + 1]
            rows.append(dict(functor=name, dim=D, bounds=np.array(k.bounds, dtype=float), x0=x0,
                             success=x is not None, f=float(e),
                             x=np.array(x0 if x is None else x), nfev=owf.nb_fun_call,
                             n_requests=len(req), request_points=np.array(req[:4])))
            print("lbfgsb large", name, D, t, "success", x is not None, "f", float(e), "nfev",
                  owf.nb_fun_call, flush=True)
    np.savez_compressed(os.path.join(GOLD, "lbfgsb_scipy_large.npz"), rows=np.array(rows, dtype=object),
                        scipy_version=scipy.__version__)


def gen_lbfgsb_catalogue():
    """The same for the catalogue classes whose suite statistics hinge on the local search's last
    digits (success means f <= fglob + 1e-8 and L-BFGS-B stops on a 2.2e-9 decrease): starts near the
    global optimum and random starts in the box."""
    rows = []
    for name in ["Whitley", "Rana", "SineEnvelope", "Mishra03", "Bukin06", "Thurber", "Cola", "PowerSum",
                 "DeVilliersGlasser02", "NewFunction01", "Trefethen", "EggHolder"]:
        k = getattr(gbf, name)()
        D = k.N
        lo = np.array([b[0] for b in k.bounds], dtype=float)
        up = np.array([b[1] for b in k.bounds], dtype=float)
        rs = np.random.RandomState(500 + len(name))
        for t in range(10):
            x0 = lo + rs.random_sample(D) * (up - lo)
            if t < 6:
                scale = [1e-1, 1e-2, 1e-3, 1e-4, 3e-2, 3e-3][t]
                x0 = np.clip(np.array(k.global_optimum[0], dtype=float) * (1.0 + rs.normal(size=D) * scale)
                             + rs.normal(size=D) * scale * 1e-2, lo, up)
            owf = R.ObjectiveFunWrapper(k.bounds, k.fun)
            with np.errstate(all="ignore"):
                e, x = owf.local_search(x0)
            rows.append(dict(functor=name, dim=D, bounds=np.array(k.bounds, dtype=float), x0=x0,
                             success=x is not None, f=float(e), x=np.array(x0 if x is None else x),
                             nfev=owf.nb_fun_call, fglob=float(k.fglob)))
            print("lbfgsb cat", name, t, "success", x is not None, "f", float(e), "nfev", owf.nb_fun_call, flush=True)
    np.savez_compressed(os.path.join(GOLD, "lbfgsb_scipy_cat.npz"), rows=np.array(rows, dtype=object),
                        scipy_version=scipy.__version__)


def gen_sutton_chen_traj():
    """BASELINE config 4 at the size the reference defines (suttonchen.py:55-62: 6 atoms, bounds
    +-0.7): the reference's own run on a recorded tape, pure SA -- energies and decisions of every
    evaluation -- plus a 13-atom cluster in the constant-density box 0.7 (N/6)^(1/3)."""
    for n_atoms, maxiter, seed in [(6, 80, 21), (13, 30, 22)]:
        L = 0.7 * (n_atoms / 6.0) ** (1.0 / 3.0)
        bounds = [(-L, L)] * (3 * n_atoms)
        r = record_run(suttonchen.sutton_chen, None, bounds, seed, maxiter, True)
        T, S = k1_tables(maxiter)
        np.savez_compressed(os.path.join(GOLD, "traj_SuttonChen_n%d.npz" % n_atoms), functor="SuttonChen",
                            params=[], bounds=np.array(bounds, dtype=float), seed=seed, maxiter=maxiter,
                            temp_table=T, sigmax_table=S, **r)
        print("traj SuttonChen", n_atoms, "fun", r["fun"], "ncall", r["ncall"], "tape", r["tape_len"])


def gen_maxfun_anchor():
    """maxfun only breaks the inner loop, the run still makes maxiter attempts."""
    rows = []
    for name, D, maxfun, maxiter, seed in [("Rastrigin", 2, 100, 200, 1234), ("Rastrigin", 5, 1000, 150, 3),
                                           ("Rosenbrock", 2, 100, 200, 1234)]:
        k = getattr(gbf, name)(dimensions=D)
        ret = R.sda(k.fun, None, k.bounds, maxiter=maxiter, maxfun=maxfun, seed=seed, pure_sa=True)
        ret_ls = R.sda(k.fun, None, k.bounds, maxiter=maxiter, maxfun=maxfun, seed=seed)
        rows.append(dict(functor=name, dim=D, maxfun=maxfun, maxiter=maxiter, seed=seed,
                         nit=int(ret.nit), ncall=int(ret.ncall), fun=float(ret.fun),
                         nit_ls=int(ret_ls.nit), ncall_ls=int(ret_ls.ncall), fun_ls=float(ret_ls.fun),
                         bounds=np.array(k.bounds, dtype=float)))
        print("maxfun", name, D, maxfun, maxiter, "pure_sa nit/ncall", ret.nit, ret.ncall,
              "with LS", ret_ls.nit, ret_ls.ncall)
    np.savez_compressed(os.path.join(GOLD, "maxfun_anchor.npz"), rows=np.array(rows, dtype=object))


def gen_objective_points():
    rows = []
    rs = np.random.RandomState(2024)
    for name, dims in [("Rastrigin", (2, 10, 50, 100, 1000)), ("Rosenbrock", (2, 10, 100)),
                       ("Ackley01", (2, 10, 100, 1000)), ("Schwefel01", (2, 5, 100)),
                       ("Schwefel26", (2, 10, 100, 1000)), ("Exponential", (2, 10, 100)),
                       ("Sphere", (2, 33)), ("Griewank", (2, 7, 50))]:
        for D in dims:
            k = getattr(gbf, name)(dimensions=D)
            lo = np.array([b[0] for b in k.bounds], dtype=float)
            up = np.array([b[1] for b in k.bounds], dtype=float)
            X = lo + rs.random_sample((16, D)) * (up - lo)
            X[0] = np.array(k.global_optimum[0], dtype=float)
            X[1] = lo
            X[2] = up
            F = np.array([k.fun(x) for x in X])
            rows.append(dict(functor=name, dim=D, params=np.zeros(0), X=X, F=F, fglob=k.fglob,
                             bounds=np.array(k.bounds, dtype=float)))
    # catalogue widening: default dimension, plus D = 7 and 40 where the class is n-D
    from sdaopt_b200.functors import _CATALOGUE2
    for name, (n_default, nd, _box, _f) in sorted(_CATALOGUE2.items()):
        for D in ([n_default, 7, 40] if nd else [n_default]):
            k = getattr(gbf, name)(dimensions=D) if nd else getattr(gbf, name)()
            lo = np.array([b[0] for b in k.bounds], dtype=float)
            up = np.array([b[1] for b in k.bounds], dtype=float)
            X = lo + rs.random_sample((16, k.N)) * (up - lo)
            if k.global_optimum is not None and len(k.global_optimum[0]) == k.N:
                X[0] = np.array(k.global_optimum[0], dtype=float)
            X[1] = lo
            X[2] = up
            F = np.array([float(k.fun(x)) for x in X])
            rows.append(dict(functor=name, dim=k.N, params=np.zeros(0), X=X, F=F, fglob=k.fglob,
                             bounds=np.array(k.bounds, dtype=float)))
    D = 50
    X = -5.12 + rs.random_sample((16, D)) * 15.36
    rows.append(dict(functor="ShiftedRastrigin", dim=D, params=np.array([3.14159]), X=X,
                     F=np.array([modified_rastrigin(x) for x in X]), fglob=0.0,
                     bounds=np.array([(-5.12, 10.24)] * D)))
    # Sutton-Chen: the 6-atom known answer (suttonchen.py:43-52) + random clusters (App. C.5)
    xg = np.array([-0.2900181566, -0.2176381343, -0.2842129662, -0.4342237494, -0.1257555181,
                   -0.9118061294, 0.1138133902, 0.2919699092, -0.3023945653, -0.0303922026,
                   0.3838525254, -0.9299877285, -0.5060632976, 0.3614667414, -0.4868774344,
                   0.1856529384, -0.1952523504, -0.7273232603])
    for N in (6, 13, 38):
        X = np.array([np.random.RandomState(N + i).uniform(-2, 2, 3 * N) for i in range(4)])
        if N == 6:
            X[0] = xg
        rows.append(dict(functor="SuttonChen", dim=3 * N, params=np.zeros(0), X=X,
                         F=np.array([suttonchen.sutton_chen(x) for x in X]), fglob=np.nan,
                         bounds=np.array([(-2., 2.)] * (3 * N))))
    np.savez_compressed(os.path.join(GOLD, "objective_points.npz"),
                        rows=np.array(rows, dtype=object))
    print("objective cases", len(rows))


def gen_objective_points_cat3():
    """Catalogue part 3 (functor ids 60..): points and values from the live reference classes.  No C
    restatement exists for these; the device functors are compared with these values directly.
    `cond` is the change of the reference value when every coordinate moves by one ulp -- the
    resolution below which a differently ordered evaluation cannot be expected to agree."""
    from sdaopt_b200.functors import _CATALOGUE3
    rows = []
    rs = np.random.RandomState(2025)
    for name, (n_default, nd, _box, _f) in sorted(_CATALOGUE3.items()):
        dims = [n_default, 7, 40] if nd else [n_default]
        if name.startswith("PermFunction"):
            dims = [2, 7]                 # j ** k overflows the reference's int64 beyond N ~ 15
        if name == "LennardJones":
            dims = [6, 9, 39]
        for D in dims:
            k = getattr(gbf, name)(dimensions=D) if nd else getattr(gbf, name)()
            lo = np.array([b[0] for b in k.bounds], dtype=float)
            up = np.array([b[1] for b in k.bounds], dtype=float)
            X = lo + rs.random_sample((16, k.N)) * (up - lo)
            if k.global_optimum is not None and len(k.global_optimum[0]) == k.N:
                X[0] = np.array(k.global_optimum[0], dtype=float)
            X[1] = lo
            X[2] = up
            with np.errstate(all="ignore"):
                F = np.array([float(k.fun(x)) for x in X])
                Fu = np.array([float(k.fun(np.nextafter(x, np.inf))) for x in X])
                Fd = np.array([float(k.fun(np.nextafter(x, -np.inf))) for x in X])
            cond = np.fmax(np.abs(Fu - F), np.abs(Fd - F))
            rows.append(dict(functor=name, dim=k.N, params=np.zeros(0), X=X, F=F, cond=cond,
                             fglob=k.fglob, bounds=np.array(k.bounds, dtype=float)))
    np.savez_compressed(os.path.join(GOLD, "objective_points_cat3.npz"),
                        rows=np.array(rows, dtype=object))
    print("catalogue-3 objective cases", len(rows))


def gen_sampler():
    """VisitingDistribution unit fixtures (test_sda.py:51-79)."""
    rs = ref_import.RecordingRandomState(1234)
    lo = np.array([-5.12, -5.12])
    up = np.array([5.12, 5.12])
    vd = R.VisitingDistribution(lo, up, 2.62, rs)
    out = {}
    for T in (5230., 100., 0.1):
        n0 = len(rs.tape)
        vals = np.array([vd.visit_fn(T) for _ in range(2000)])
        out["visit_fn_T%g" % T] = vals
        out["visit_fn_T%g_tape" % T] = np.array([n0, len(rs.tape)])
    rs2 = ref_import.RecordingRandomState(1234)
    vd2 = R.VisitingDistribution(lo, up, 2.62, rs2)
    out["visiting_step0"] = vd2.visiting(np.zeros(2), 0, 5230.)
    out["visiting_step0_tape"] = len(rs2.tape)
    out["visiting_step2"] = vd2.visiting(np.zeros(2), 2, 5230.)
    out["visiting_step2_tape"] = len(rs2.tape)
    T, S = k1_tables(1400)
    out["temp_table"] = T
    out["sigmax_table"] = S
    np.savez_compressed(os.path.join(GOLD, "sampler.npz"), **out)
    print("sampler fixtures: restart index", len(T) - 1)


def gen_suite():
    """Suite semantics (bench.py:155-216, :256-333): seeds 1234+i, prepare() draws N uniforms, then
    sda(seed=None) continues on the global stream; success / ncall of first hit."""
    rows = []
    for name, dim in [("Rastrigin", None), ("Rosenbrock", None), ("Rastrigin", 10),
                      ("Ackley01", 10), ("Schwefel01", 5), ("Easom", None), ("Sphere", None)]:
        klass = getattr(gbf, name)
        for i in range(5):
            algo = refbench.SDAOptimizer()
            np.random.seed(1234 + i)
            algo.prepare(name, klass, dim)
            algo._maxcall = 200000        # bounded budget for the fixture (bench.py:37 uses 1e6)
            try:
                algo.optimize()
            except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
                pass
            k = algo._k
            rows.append(dict(functor=name, dim=k.N, seed=1234 + i, success=bool(algo.success),
                             ncall=int(algo.fcall_success), fvalue=float(algo.fsuccess),
                             nbcall=int(algo.nbcall), fglob=float(k.fglob), max_calls=200000,
                             bounds=np.array(k.bounds, dtype=float)))
        print("suite", name, dim, [(r["success"], r["ncall"]) for r in rows[-5:]])
    np.savez_compressed(os.path.join(GOLD, "suite_runs.npz"), rows=np.array(rows, dtype=object))


def gen_suite_stats():
    """The north star's statistical criterion needs the REFERENCE's own distribution: NB_RUNS = 100
    suite-mode runs (bench.py:170-216 semantics, seeds 1234+i) of a few jobs; success / ncall /
    fvalue arrays as BenchUnit would hold them."""
    rows = []
    for name, dim in [("Rastrigin", None), ("Rastrigin", 5), ("Rastrigin", 10), ("Ackley01", None),
                      ("Ackley01", 10), ("Rosenbrock", None), ("Schwefel01", 5), ("Sphere", None),
                      ("Easom", None), ("Exponential", 10), ("Beale", None), ("HimmelBlau", None),
                      ("Zacharov", None), ("StyblinskiTang", None), ("Levy13", None)]:
        klass = getattr(gbf, name)
        succ, ncall, fval = [], [], []
        for i in range(100):
            algo = refbench.SDAOptimizer()
            np.random.seed(1234 + i)
            algo.prepare(name, klass, dim)
            algo._maxcall = 200000
            try:
                algo.optimize()
            except (refbench.OptimumFoundException, refbench.OptimumNotFoundException):
                pass
            succ.append(bool(algo.success)); ncall.append(int(algo.fcall_success)); fval.append(float(algo.fsuccess))
        k = algo._k
        rows.append(dict(functor=name, dim=k.N, success=np.array(succ), ncall=np.array(ncall),
                         fvalue=np.array(fval), fglob=float(k.fglob), max_calls=200000,
                         bounds=np.array(k.bounds, dtype=float)))
        print("suite-stats", name, k.N, "success", np.mean(succ), "median ncall", np.median(ncall), flush=True)
    np.savez_compressed(os.path.join(GOLD, "suite_stats_ref.npz"), rows=np.array(rows, dtype=object))


def gen_asrun_summary():
    D = 50
    np.random.seed(1234)
    lower = np.array([-5.12] * D)
    upper = np.array([10.24] * D)
    x_init = lower + np.random.rand(D) * (upper - lower)
    r = record_run(modified_rastrigin, x_init, list(zip(lower, upper)), 1234, 1000, True)
    from numpy._core._multiarray_umath import __cpu_features__ as feats
    np.savez_compressed(os.path.join(GOLD, "asrun_readme_rastrigin50_it1000.npz"),
                        fun=r["fun"], ncall=r["ncall"], tape_len=r["tape_len"], x=r["x"],
                        energies_head=r["energies"][:5000], decisions=r["decisions"],
                        numpy_version=np.__version__,
                        cpu_features=[k for k, v in feats.items() if v])
    print("as-run README D=50: fun", r["fun"], "tape", r["tape_len"])


if __name__ == "__main__":
    if MODE == "libm" and os.environ.get("SDA_GOLDEN_ONLY") == "objectives":
        gen_objective_points()
    elif MODE == "libm" and os.environ.get("SDA_GOLDEN_ONLY") == "objectives_cat3":
        gen_objective_points_cat3()
    elif MODE == "libm" and os.environ.get("SDA_GOLDEN_ONLY") == "suite_stats":
        gen_suite_stats()
    elif MODE == "libm" and os.environ.get("SDA_GOLDEN_ONLY") == "lbfgsb_cat":
        gen_lbfgsb_catalogue()
    elif MODE == "libm" and os.environ.get("SDA_GOLDEN_ONLY") == "round2":
        gen_lbfgsb_large()
        gen_sutton_chen_traj()
        gen_maxfun_anchor()
    elif os.environ.get("SDA_GOLDEN_ONLY"):
        pass
    elif MODE == "libm":
        gen_sampler()
        gen_objective_points()
        gen_objective_points_cat3()
        gen_trajectories()
        gen_ls_runs()
        gen_lbfgsb()
        gen_lbfgsb_large()
        gen_lbfgsb_catalogue()
        gen_sutton_chen_traj()
        gen_maxfun_anchor()
        gen_suite()
        gen_suite_stats()
    else:
        gen_asrun_summary()
