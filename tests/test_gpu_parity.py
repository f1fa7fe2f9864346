"""GPU: the CUDA path, called through the C ABI (libsdaopt_b200.so), against the CPU oracle and
the golden fixtures recorded from the live reference.

Tolerances (DESIGN.md "Parity"): decisions and draw counts are compared exactly; objective values
at identical points to 1e-12 relative (floor 1e-12 * 10 D for the cancelling `10 D + sum` forms);
trajectory energies to TRAJ_RTOL = 1e-4 relative -- den = exp(4.26 log|y|) inherits the absolute
rounding error of its argument (|arg| up to ~40, so ~1e-14 relative in den), and a proposal
x + visit with |visit| up to 1e8 turns that into an absolute error up to ~1e-6 in x before the
wrap; CUDA's, glibc's and numpy's exp/log differ by 1 ulp on a few percent of arguments, and the
reference itself moves by that much between numpy builds (tests/golden/asrun_*)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from conftest import GOLD, load_golden, tape_from_seed
import parity_metrics as pm

pytestmark = pytest.mark.gpu

TRAJ_RTOL = 1e-4       # ceiling only; the asserted tolerance is 4x the deviation recorded in gpu_pins.json
TRAJ = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "traj_*.npz")))
# tests/golden/gpu_pins.json: what the CUDA path achieved on a B200 when the pins were recorded
# (profiles/record_parity_pins.py) -- first-divergence indices, identical-chain fractions, energy
# deviations, L-BFGS-B call-count matches.  The tests assert these, not a blanket tolerance.
PINS = pm.load_pins()


def _pin(group, key):
    assert PINS is not None, "tests/golden/gpu_pins.json is missing: run profiles/record_parity_pins.py on a B200"
    return PINS[group][key]


@pytest.fixture(scope="module")
def sb():
    import sdaopt_b200
    from sdaopt_b200 import _lib
    assert _lib.lib().sda_device_count() > 0, "CUDA device required: no CPU fallback"
    return sdaopt_b200


def _functor(sb, name, params):
    return sb.DeviceFunctor(name, params=params if params is not None and len(params) else ())


def _energy_close(e, ref, dim, rel):
    # rel applies to max(|ref|, 1) when it is the trajectory tolerance: the x error inherited from
    # earlier accepted moves is absolute, so energies below 1 are compared absolutely
    floor = 1e-12 * 10.0 * dim if rel < 1e-9 else rel
    return np.abs(e - ref) <= np.maximum(rel * np.abs(ref), floor)


# ---- K3: objectives -----------------------------------------------------------------------------
def test_objectives_match_reference_and_oracle(sb, oracle):
    rows = load_golden("objective_points.npz")["rows"]
    for r in rows:
        f = _functor(sb, r["functor"], r["params"])
        F = f(r["X"])
        tol = 1e-12 if r["functor"] != "SuttonChen" else 1e-11
        ok = _energy_close(F, r["F"], r["dim"], tol)
        assert ok.all(), (r["functor"], r["dim"], F[~ok], r["F"][~ok])
        P = oracle.Problem(r["functor"], r["bounds"], params=r["params"] if len(r["params"]) else None)
        Fo = np.array([P.eval(x) for x in r["X"]])
        assert _energy_close(F, Fo, r["dim"], tol).all()


def test_catalogue3_objectives_match_reference(sb):
    """functor ids 60.. against values of the live reference classes
    (oracle/gen_golden.py::gen_objective_points_cat3).  Tolerance: 1e-12 relative, or 32x the
    change the reference itself shows under a one-ulp move of x (ill-conditioned points), or the
    cancellation floor of test_objectives_match_reference_and_oracle."""
    rows = load_golden("objective_points_cat3.npz")["rows"]
    assert len({r["functor"] for r in rows}) == 136
    bad = []
    for r in rows:
        f = _functor(sb, r["functor"], r["params"])
        F = f(r["X"])
        ref, cond = r["F"], np.nan_to_num(r["cond"], nan=0.0, posinf=0.0)
        tol = np.maximum(np.maximum(1e-12 * np.abs(ref), 32 * cond), 1e-12 * 10.0 * r["dim"])
        same_nan = np.isnan(F) & np.isnan(ref)
        same_inf = np.isinf(F) & np.isinf(ref) & (np.sign(F) == np.sign(ref))
        with np.errstate(invalid="ignore"):
            ok = same_nan | same_inf | (np.abs(F - ref) <= tol)
        if not ok.all():
            bad.append((r["functor"], r["dim"], F[~ok][:3], ref[~ok][:3]))
    assert not bad, bad


def test_rng_in_objective_classes(sb):
    """Stochastic (go_funcs_S.py:1274-1280) and XinSheYang01 (go_funcs_X.py:47-51) draw U[0,1) factors
    from NumPy's global generator inside fun(); the device functors take them from Philox noise
    addressed by (point, call number) (sda_objectives_cat.cuh; sda_eval uses call number 0, chains
    their evaluation counter -- the in-chain statistics are checked against the live reference by
    test_suite_statistics_match_reference_whole_catalogue).  Checked here: (i) values equal the
    integer restatement oracle/rng_objectives.py to 1e-12; (ii) f(global optimum) == fglob exactly, as in the
    reference where every term vanishes; (iii) the factors seen at 20,000 distinct points follow the
    reference's law U[0,1) (Kolmogorov-Smirnov) and neighbouring points (1 ulp apart) get
    uncorrelated factors -- what the finite-difference gradient of the reference observes."""
    from scipy.stats import kstest
    from oracle import rng_objectives as ro
    rng = np.random.RandomState(7)
    for name in ("Stochastic", "XinSheYang01"):
        for dim in (2, 5, 10):
            f = sb.DeviceFunctor(name, dimensions=dim)
            X = rng.uniform(-5.0, 5.0, size=(24, dim))
            F = f(X)
            ref = np.array([ro.fun(name, x) for x in X])
            assert np.all(np.abs(F - ref) <= 1e-12 * np.maximum(np.abs(ref), 1.0)), (name, dim, F, ref)
            opt = 1.0 / np.arange(1, dim + 1) if name == "Stochastic" else np.zeros(dim)
            assert f(opt) == 0.0 == f.fglob
    n = 20000
    f = sb.DeviceFunctor("Stochastic", dimensions=2)
    X = np.column_stack([rng.uniform(-5.0, 0.0, n), np.full(n, 0.5)])   # second term vanishes
    eps0 = f(X) / np.abs(X[:, 0] - 1.0)
    assert eps0.min() >= 0.0 and eps0.max() < 1.0 + 1e-12
    assert kstest(eps0, "uniform").pvalue > 1e-3
    Xn = X.copy()
    Xn[:, 0] = np.nextafter(X[:, 0], 0.0)
    eps0n = f(Xn) / np.abs(Xn[:, 0] - 1.0)
    assert abs(np.corrcoef(eps0, eps0n)[0, 1]) < 0.03
    X1 = np.column_stack([np.ones(n), rng.uniform(1.0, 5.0, n)])        # first term vanishes
    eps1 = f(X1) / np.abs(X1[:, 1] - 0.5)
    assert kstest(eps1, "uniform").pvalue > 1e-3
    g = sb.DeviceFunctor("XinSheYang01", dimensions=2)
    Y = np.column_stack([rng.uniform(0.5, 5.0, n), np.zeros(n)])
    assert kstest(g(Y) / Y[:, 0], "uniform").pvalue > 1e-3


def test_device_math_accuracy(sb):
    """csrc/sda_math.cuh against x87 long double: log_pos and exp_any within 1.5 ulp, cos_2pi within
    4e-16 absolute of cos(2 pi x) (exact-period reduction; what `x^2 - 10 cos(2 pi x)` needs)."""
    from sdaopt_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(5)

    def probe(kind, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        _lib.check(L.sda_math_probe_host(kind, x.ctypes.data, x.size, out.ctypes.data, 0))
        return out

    ld = np.longdouble
    assert np.finfo(ld).nmant >= 63, "needs an extended-precision long double as the referee"
    # cos(2 pi x): the Rastrigin boxes, the README box, wide arguments, exact special points
    x = np.concatenate([rs.uniform(-5.12, 5.12, 200000), rs.uniform(-5.12, 10.24, 200000) - 3.14159,
                        rs.uniform(-600, 600, 100000), rs.uniform(-1e9, 1e9, 20000),
                        np.arange(-8, 9) * 0.125, [0.0, -0.0, 0.25, 0.75, 1e-300, 2e15, -2e15]])
    two_pi = ld(2) * ld("3.14159265358979323846264338327950288")
    ref = np.cos(two_pi * (x.astype(ld) - np.rint(x).astype(ld))).astype(np.float64)
    got = probe(0, x)
    assert np.abs(got - ref).max() <= 4e-16
    assert probe(0, np.array([0.0, 1.0, -3.0, 0.5, 2.5]))[:].tolist() == [1.0, 1.0, 1.0, -1.0, -1.0]
    assert np.isnan(probe(0, np.array([np.inf, np.nan]))).all()
    # log: s in (0, 1) as the polar method feeds it, |y| over the whole normal range
    x = np.concatenate([rs.random_sample(300000), np.exp(rs.uniform(-700, 700, 300000)),
                        1.0 + rs.uniform(-1e-3, 1e-3, 50000), [1.0, 0.5, 2.0, 2.2250738585072014e-308, 1.7e308]])
    ref_ld = np.log(x.astype(ld))
    got = probe(1, x)
    ulp = np.spacing(np.abs(ref_ld.astype(np.float64)))
    err = np.abs(got.astype(ld) - ref_ld).astype(np.float64) / np.maximum(ulp, 5e-324)
    assert err[x != 1.0].max() <= 1.5 and got[x == 1.0].tolist() == [0.0]
    sp = probe(1, np.array([0.0, -1.0, np.inf, 5e-324]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and sp[3] == pytest.approx(np.log(5e-324))
    # exp: the sampler's range -c ln|y| with c = 4.26, and the edges
    x = np.concatenate([rs.uniform(-708, 709, 400000), rs.uniform(-2, 2, 200000), [0.0, 709.5, 710.0, -708.5, -800.0]])
    ref_ld = np.exp(x.astype(ld))
    got = probe(2, x)
    fin = (x > -708.0) & (x < 709.0)
    ulp = np.spacing(ref_ld[fin].astype(np.float64))
    err = np.abs(got[fin].astype(ld) - ref_ld[fin]).astype(np.float64) / ulp
    assert err.max() <= 1.5
    assert got[x == 710.0][0] == np.inf and got[x == 709.5][0] == np.inf     # documented: +inf from 709 on
    assert got[x == -800.0][0] == 0.0 and got[x == 0.0][0] == 1.0
    assert np.isnan(probe(2, np.array([np.nan])))[0]
    # division / square root of normal operands (sampler: -2 ln(s) / s with s >= 2^-106, its root)
    x = np.concatenate([np.exp(rs.uniform(-73, 73, 300000)), rs.uniform(1.7, 2.42, 100000),
                        rs.random_sample(100000) + 1e-30])
    for kind, ref_ld in ((3, np.sqrt(x.astype(ld))), (4, ld("1.2345678901234567") / x.astype(ld))):
        got = probe(kind, x)
        ulp = np.spacing(ref_ld.astype(np.float64))
        err = np.abs(got.astype(ld) - ref_ld).astype(np.float64) / ulp
        assert err.max() <= 1.0, (kind, err.max())
        assert (got != ref_ld.astype(np.float64)).mean() < 0.1       # faithful; mostly the rounded value


def test_sutton_chen_known_answer_and_sizes(sb, oracle):
    xg = load_golden("objective_points.npz")["rows"][-3]["X"][0]
    assert sb.SuttonChen(6)(xg) == pytest.approx(-1163.9639632, abs=5e-8)      # suttonchen.py:43-52
    for n_atoms in (38, 128, 512):
        X = np.random.RandomState(n_atoms).uniform(-2, 2, (3, 3 * n_atoms)) * (n_atoms / 6) ** (1 / 3)
        F = sb.SuttonChen(n_atoms)(X)
        P = oracle.Problem("SuttonChen", [(-20, 20)] * (3 * n_atoms))
        Fo = np.array([P.eval(x) for x in X])
        np.testing.assert_allclose(F, Fo, rtol=1e-11)


# ---- K2: visiting distribution --------------------------------------------------------------------
def test_visit_fn_against_reference(sb):
    from sdaopt_b200 import _lib
    g = load_golden("sampler.npz")
    tape = tape_from_seed(1234, 40000)
    for T in (5230., 100., 0.1):
        vals = g["visit_fn_T%g" % T]
        a, b = g["visit_fn_T%g_tape" % T]
        i = {5230.: 0}.get(T)
        sig = float(g["sigmax_table"][0]) if T == 5230. else sb._sda.sigmax(T, 2.62)
        out = np.empty(len(vals))
        used = C.c_int64()
        t = np.ascontiguousarray(tape[a:b])
        _lib.check(_lib.lib().sda_visit_fn_host(2.62, sig, t.ctypes.data, t.size, out.ctypes.data,
                                                 len(vals), C.byref(used), 0))
        assert used.value == b - a                      # identical polar rejections
        np.testing.assert_allclose(out, vals, rtol=5e-13)


def test_visiting_stepping_reference_unit_tests(sb):
    """test_sda.py:51-64 and SURVEY App. C.1 anchors, through the mirrored class."""
    from scipy._lib._util import check_random_state
    g = load_golden("sampler.npz")
    lower, upper = np.array([-5.12] * 2), np.array([5.12] * 2)
    rs = check_random_state(1234)
    vd = sb.VisitingDistribution(lower, upper, 2.62, rs)
    x_step_low = vd.visiting(np.zeros(2), 0, 5230)
    assert np.all(x_step_low != 0)
    np.testing.assert_allclose(x_step_low, g["visiting_step0"], rtol=1e-13)
    x_step_high = vd.visiting(np.zeros(2), 2, 5230)
    assert x_step_high[0] != 0 and x_step_high[1] == 0
    np.testing.assert_allclose(x_step_high, g["visiting_step2"], rtol=1e-13)
    # the RandomState advanced by exactly what the reference consumed
    ref_rs = np.random.RandomState(1234)
    ref_rs.random_sample(int(g["visiting_step2_tape"]))
    assert rs.random_sample() == ref_rs.random_sample()


def test_visiting_dist_high_temperature_and_gaussian(sb):
    """test_sda.py:66-79 and :93-103 (statistical), via the device sampler on a tape."""
    from sdaopt_b200 import _lib
    tape = tape_from_seed(1234, 20000)
    out = np.empty(5000)
    used = C.c_int64()
    sig = sb._sda.sigmax(5230., 2.62)
    _lib.check(_lib.lib().sda_visit_fn_host(2.62, sig, tape.ctypes.data, tape.size, out.ctypes.data,
                                             5000, C.byref(used), 0))
    assert out.min() < 1e-10 and out.max() > 1e10
    # gaussian_fn(1) == visit kernel with qv = 1, sigmax = 1
    _lib.check(_lib.lib().sda_visit_fn_host(1.0, 1.0, tape.ctypes.data, tape.size, out.ctypes.data,
                                             5000, C.byref(used), 0))
    assert abs(out.mean()) < 0.05 and abs(np.median(out)) < 0.05 and abs(out.std() - 1) < 0.05
    from scipy._lib._util import check_random_state
    vd = sb.VisitingDistribution(np.array([-5.12] * 2), np.array([5.12] * 2), 2.62,
                                 check_random_state(1234))
    v = np.array([vd.gaussian_fn(1) for _ in range(50)])
    np.testing.assert_allclose(v, out[:50], rtol=1e-15)


# ---- trajectory replay: the reference's RandomState draws injected ---------------------------------
@pytest.mark.parametrize("fname", TRAJ)
def test_trajectory_replay_against_reference(sb, oracle, fname):
    """North star, first criterion: with the reference's RandomState draws injected and local search
    off, the chain reproduces the reference's accept/reject decisions.  Asserted per fixture: the
    decision sequence is identical up to AT LEAST the pinned index (every short fixture: all of it;
    the 99,900-evaluation README run: 92,085 -- see the module docstring for why a 1-ulp exp/log
    difference can flip a near-tie late in a long run, and tests/golden/asrun_* for the reference
    doing the same between numpy builds), draw count / nit / ncall exact, and energies before that
    index within 4x the recorded deviation (which is the propagated x error, not rounding of f)."""
    m = pm.trajectory_metrics(sb, oracle, fname)
    pin = _pin("traj", fname)
    assert m["status"] == 0 and m["nit_ok"] and m["ncall_ok"]
    assert m["first_diff"] >= pin["first_diff"], (m["first_diff"], pin["first_diff"])
    assert m["first_diff"] >= min(m["n_evals"], 4000)
    # Sutton-Chen: the r^-9 pair terms turn the same ~1e-7 absolute x error into ~1e-3 relative in the
    # energy of a compressed cluster (pins: 3.6e-3 at 6 atoms); every other fixture stays below 1e-4
    ceiling = 5e-2 if "SuttonChen" in fname else TRAJ_RTOL
    assert m["max_edev"] <= min(ceiling, max(4.0 * pin["max_edev"], 1e-13)), (m["max_edev"], pin["max_edev"])
    if m["first_diff"] == m["n_evals"]:
        assert m["tape_ok"] and m["x_dev"] <= 2e-5
        assert m["fun"] == pytest.approx(m["fun_ref"], rel=1e-3, abs=1e-9)
    # ... and the C oracle on the same tape agrees with the GPU at least as far
    assert m["first_diff_oracle"] >= m["first_diff"] or m["first_diff_oracle"] >= pin["first_diff_oracle"]


# ---- production RNG: the oracle implements the same Philox addressing -------------------------------
@pytest.mark.parametrize("name,dim,maxiter", pm.PHILOX_CASES)
def test_philox_chains_match_oracle(sb, oracle, name, dim, maxiter):
    """Chains on the production RNG against the C oracle on the same Philox keys, incl. BASELINE
    config 3's D = 1000 shapes: evaluation and iteration counts exact, the fraction of chains whose
    whole decision sequence is identical at least the pinned one, energies on the identical prefixes
    within 4x the recorded deviation."""
    m = pm.philox_metrics(sb, oracle, name, dim, maxiter)
    pin = _pin("philox", "%s-%d-%d" % (name, dim, maxiter))
    assert m["counts_ok"] and m["status_ok"]
    assert m["exact_chains"] >= pin["exact_chains"] - 1e-12, (m["exact_chains"], pin["exact_chains"])
    assert m["exact_chains"] >= 0.75
    assert m["max_edev"] <= min(TRAJ_RTOL, max(4.0 * pin["max_edev"], 1e-13)), (m["max_edev"], pin["max_edev"])
    assert m["fun_dev"] <= 1e-3


@pytest.mark.parametrize("n_atoms,maxiter", pm.SUTTON_CHEN_CASES)
def test_sutton_chen_chains_match_oracle(sb, oracle, n_atoms, maxiter):
    """BASELINE config 4 at chain level: Sutton-Chen clusters (6, 38, 128 atoms; box of constant
    density) annealed on the production RNG against the oracle on the same keys.  The device pair sum
    uses rsqrt and another pair order than suttonchen.py:28-40, so energies agree to ~1e-11, and a
    decision can only flip on a tie at that level."""
    m = pm.sutton_chen_metrics(sb, oracle, n_atoms, maxiter)
    pin = _pin("sutton_chen", "%d-%d" % (n_atoms, maxiter))
    assert m["counts_ok"] and m["status_ok"]
    assert m["exact_chains"] >= pin["exact_chains"] - 1e-12, (m["exact_chains"], pin["exact_chains"])
    assert m["exact_chains"] >= 0.75
    assert m["max_edev"] <= min(5e-2, max(4.0 * pin["max_edev"], 1e-13)), (m["max_edev"], pin["max_edev"])


def test_maxfun_is_a_soft_limit(sb):
    """`maxfun` only breaks the inner loop (_sda.py:413-418), the while loop restarts
    at i = 0 and the run still makes maxiter attempts.  tests/golden/maxfun_anchor.npz holds the live
    reference's (nit, ncall): pure-SA call counts are data independent, so they must be equal on either
    RNG; with local search on the reference's run is replayed from its own uniforms."""
    for r in load_golden("maxfun_anchor.npz")["rows"]:
        f = sb.DeviceFunctor(r["functor"], dimensions=int(r["dim"]))
        kw = dict(maxiter=int(r["maxiter"]), maxfun=float(r["maxfun"]), seed=int(r["seed"]))
        for rng in ("philox", "tape"):
            ret = sb.sda(f, None, r["bounds"], pure_sa=True, rng=rng, **kw)
            assert ret.nit == int(r["nit"]) == int(r["maxiter"]) and ret.ncall == int(r["ncall"])
            assert ret.ncall > r["maxfun"]                   # the limit did not stop the run
        # local search on, the reference's own draws injected: once ncall >= maxfun the chain breaks
        # out before MarkovChain.local_search (:413-414), so only the first steps search
        ret = sb.sda(f, None, r["bounds"], rng="tape", **kw)
        assert ret.nit == int(r["nit_ls"])
        assert abs(ret.ncall - int(r["ncall_ls"])) <= 0.1 * int(r["ncall_ls"]), (ret.ncall, int(r["ncall_ls"]))


# ---- K6/K7: local search ------------------------------------------------------------------------------
@pytest.mark.parametrize("fname", ["lbfgsb_scipy.npz", "lbfgsb_scipy_large.npz"])
def test_local_search_against_scipy_and_oracle(sb, oracle, fname):
    """ObjectiveFunWrapper.local_search against SciPy's own L-BFGS-B runs recorded through the
    reference (oracle/gen_golden.py): 54 cases at D <= 50 and 15 at D = 100 / 1000, where the
    device keeps the L-BFGS-B workspace in global memory.  Per case: same success flag, same
    minimum; the number of cases with SciPy's exact call count is at least the pinned one."""
    rows = pm.lbfgsb_metrics(sb, oracle, fname)
    pin = _pin("lbfgsb", fname)
    for r in rows:
        assert r["success"] == r["success_ref"], (r["functor"], r["dim"])
        if not r["success"]:
            continue
        if r["functor"] in ("Rastrigin", "Sphere", "Exponential", "Schwefel01", "Schwefel26", "Ackley01"):
            # (absolute floor grows with D: Schwefel26 is 418.98 D minus a sum of the same size)
            assert r["f"] == pytest.approx(r["f_ref"], rel=1e-6, abs=1e-8 * max(1.0, r["dim"] / 10.0)), (r["functor"], r["dim"])
            assert r["x_dev"] <= 2e-5 * r["x_scale"], (r["functor"], r["dim"], r["x_dev"])
    same = sum(r["nfev"] == r["nfev_ref"] for r in rows)
    assert same >= pin["same_nfev_as_scipy"], (same, pin["same_nfev_as_scipy"])
    assert same >= 0.6 * len(rows)


def test_local_search_on_catalogue_functions_against_scipy(sb, oracle):
    """The same against SciPy on the catalogue classes whose suite statistics hinge on the last digits
    of the local search (Whitley, Rana, Bukin06, Thurber, Cola, PowerSum, DeVilliersGlasser02 ...:
    tests/golden/lbfgsb_scipy_cat.npz, 120 cases, starts near the optimum and random starts).  These
    are ill-conditioned on purpose (Bukin06's sqrt kink, DeVilliersGlasser02's exp/sin fit), so a
    one-ulp difference in f can send the line search another way: asserted are the pinned counts of
    cases with SciPy's success flag and with SciPy's exact call count, and -- where the call count is
    SciPy's -- the same minimum."""
    rows = pm.lbfgsb_metrics(sb, oracle, "lbfgsb_scipy_cat.npz")
    pin = _pin("lbfgsb", "lbfgsb_scipy_cat.npz")
    ok = sum(r["success"] == r["success_ref"] for r in rows)
    same = sum(r["nfev"] == r["nfev_ref"] for r in rows)
    assert ok >= pin["success_match"] and ok >= 0.95 * len(rows), (ok, pin["success_match"])
    assert same >= pin["same_nfev_as_scipy"] and same >= 0.7 * len(rows), (same, pin["same_nfev_as_scipy"])
    for r in rows:
        if r["success"] and r["success_ref"] and r["nfev"] == r["nfev_ref"]:
            # (ill-conditioned on purpose: Bukin06's kink and PowerSum's flat valley each have one case that
            # ends 1.5e-4 / 2e-3 away with SciPy's exact call count)
            assert r["f"] == pytest.approx(r["f_ref"], rel=1e-2, abs=1e-8), (r["functor"], r["f"], r["f_ref"])


def test_local_search_zero_plateau_trap(sb):
    """f rounds to 0 while |g| > gtol: SciPy's line search keeps shrinking the step until the
    trial point stops changing, ScalarFunction then serves the cache, and the run ends ABNORMAL
    after 28 requests = 476 calls (measured with scipy 1.18.1; oracle/lbfgsb_oracle.h)."""
    owf = sb.ObjectiveFunWrapper([(-5.12, 5.12)] * 8, "Rastrigin")
    e, x = owf.local_search(np.array([6e-9, 0, 0, 0, 0, 0, 0, 0.]))
    assert x is None and e == 1e16 and owf.nb_fun_call == 476


# ---- whole runs: the reference's own tests (sdaopt/tests/test_sda.py) ----------------------------------
def test_low_dim(sb):
    ret = sb.sda(sb.Rastrigin(), None, [(-5.12, 5.12)] * 2, seed=1234)
    np.testing.assert_allclose(ret.fun, 0., atol=1e-12)
    assert set(ret.keys()) == {"x", "fun", "nit", "ncall"} and ret.nit == 1000


def test_high_dim(sb):
    ret = sb.sda(sb.Rastrigin(), None, [(-5.12, 5.12)] * 8)
    np.testing.assert_allclose(ret.fun, 0., atol=1e-12)


def test_reproduce(sb):
    res = [sb.sda(sb.Rastrigin(), None, [(-5.12, 5.12)] * 2, seed=1234) for _ in range(3)]
    np.testing.assert_equal(res[0].x, res[1].x)
    np.testing.assert_equal(res[0].x, res[2].x)


def test_max_reinit_and_reset(sb):
    with pytest.raises(ValueError):
        sb.sda(sb.Constant(2e16), None, [(-5.12, 5.12)] * 2)
    from scipy._lib._util import check_random_state
    owf = sb.ObjectiveFunWrapper([(-5.12, 5.12)] * 2, sb.Constant(2e16))
    es = sb.EnergyState(np.array([-5.12] * 2), np.array([5.12] * 2))
    with pytest.raises(ValueError):
        es.reset(owf, check_random_state(None))


def test_bounds_integrity(sb):
    with pytest.raises(ValueError):
        sb.sda(sb.Rastrigin(), None, [(-5.12, 5.12), (1, 0), (5.12, 5.12)])
    with pytest.raises(ValueError):
        sb.sda(sb.Rastrigin(), np.zeros(3), [(-5.12, 5.12)] * 2)
    with pytest.raises(TypeError):
        sb.sda(lambda x: 0.0, None, [(-5.12, 5.12)] * 2)


def test_tape_mode_replays_reference_run_with_local_search(sb):
    """rng='tape': the RandomState uniforms themselves drive the run (SURVEY App. C.4 anchors)."""
    rows = load_golden("ls_runs.npz")["rows"]
    exact = 0
    for r in rows[:5]:
        f = sb.DeviceFunctor(r["functor"], dimensions=r["dim"])
        ret = sb.sda(f, None, r["bounds"], seed=int(r["seed"]), rng="tape")
        assert ret.nit == 1000
        assert ret.fun == pytest.approx(r["fun"], abs=1e-8)
        assert abs(ret.ncall - r["ncall"]) <= 0.10 * r["ncall"]
        exact += int(ret.ncall == r["ncall"])
    assert exact >= 2, exact


def test_readme_examples(sb):
    """README.md:28-39 and :41-80 (BASELINE configs 1 and 2 with local search)."""
    ret = sb.sda(sb.Rosenbrock(), None, [(-30, 30)] * 2, seed=1234)
    assert ret.fun < 1e-8 and np.allclose(ret.x, 1.0, atol=1e-3) and ret.nit == 1000
    np.random.seed(1234)
    lower, upper = np.array([-5.12] * 50), np.array([10.24] * 50)
    x_init = lower + np.random.rand(50) * (upper - lower)
    ret = sb.sda(sb.ShiftedRastrigin(3.14159), x_init, list(zip(lower, upper)), seed=1234)
    assert ret.fun < 1e-8 and np.allclose(ret.x, 3.14159, atol=1e-4)
    assert 99900 < ret.ncall < 200000


# ---- statistical parity with local search on (north star, second criterion) ---------------------------
_wilson = pm.wilson


@pytest.mark.parametrize("name,dim", [("Rastrigin", 2), ("Rastrigin", 10), ("Ackley01", 10),
                                      ("Rosenbrock", 2), ("Schwefel01", 5)])
def test_suite_statistics_match_oracle(sb, oracle, name, dim):
    """Suite semantics (bench.py:292-321) over NB_RUNS=100 seeds: success rate inside the oracle's
    Wilson 95 % interval and median ncall-to-global inside its order-statistic interval."""
    f = sb.DeviceFunctor(name, dimensions=dim)
    n = 100
    seeds = np.arange(n, dtype=np.uint64) + np.uint64(1234)
    kw = dict(maxiter=int(1e6), max_calls=100000, fglob=f.fglob, tol=1e-8)
    r = sb.sda_batch(f, f.bounds, n, seeds=seeds, **kw)
    o = oracle.run(oracle.Problem(name, f.bounds), n, seeds=seeds, n_threads=8,
                   tables=sb._sda.temperature_tables(int(1e6)), **kw)
    lo, hi = _wilson(int(o["hit"].sum()), n)
    assert lo - 1e-9 <= r.hit.mean() <= hi + 1e-9 or r.hit.sum() == o["hit"].sum()
    so = np.sort(o["ncall_hit"])
    # distribution-free 95 % CI of the median for n = 100: order statistics 40 and 61
    assert so[39] <= np.median(r.ncall_hit) <= so[60], (np.median(r.ncall_hit), so[39], so[60])
    # the same seeds give the same first hits unless L-BFGS-B noise moved a chain
    assert np.mean(r.ncall_hit == o["ncall_hit"]) > 0.5


# ---- full-size properties (BASELINE config 3 shape) ------------------------------------------------------
def test_full_size_batch_properties(sb):
    """65,536 Rastrigin D=50 chains (short schedule): every chain reports nit == maxiter, exactly
    2 D (maxiter - 1) annealing calls, a best point inside the box whose energy is reproduced by
    an independent evaluation, and two launches with the same seeds are bit-identical."""
    f = sb.Rastrigin(50)
    n, maxiter = 65536, 6
    r1 = sb.sda_batch(f, f.bounds, n, maxiter=maxiter, pure_sa=True)
    r2 = sb.sda_batch(f, f.bounds, n, maxiter=maxiter, pure_sa=True)
    assert (r1.status == 0).all() and (r1.nit == maxiter).all()
    assert (r1.ncall == 2 * 50 * (maxiter - 1)).all()
    assert (r1.x >= -5.12).all() and (r1.x <= 5.12).all()
    np.testing.assert_array_equal(r1.x, r2.x)
    np.testing.assert_array_equal(r1.fun, r2.fun)
    F = f(r1.x[::64])
    np.testing.assert_allclose(F, r1.fun[::64], rtol=1e-13)
    assert len(np.unique(r1.fun)) > 0.99 * n           # independent streams


# ---- suite runner (config 5 slice): one launch per job, BenchUnit pickles + results.csv ----------------
def test_suite_runner_writes_reference_format(sb, oracle, tmp_path):
    import csv
    from sdaopt_b200.benchmark import Benchmarker, report
    from sdaopt_b200.benchmark import benchunit
    folder = str(tmp_path / "DATA")
    units = Benchmarker(20, folder, names=["Rastrigin", "Sphere"], max_calls=100000).run()
    assert len(units) == 2 + 11                                  # Rastrigin is in the n-D selection
    by_name = {u.name: u for u in units}
    assert os.path.exists(os.path.join(folder, "Rastrigin_10-SDA-20.data"))
    u = benchunit.load(os.path.join(folder, "Rastrigin-SDA-20.data"))
    assert u.success.all() and 100 <= u.medall <= 3000          # App. C.7: 166..694 on the reference
    assert by_name["Sphere"].success.all() and by_name["Sphere"].medall < 100
    assert by_name["Rastrigin_10"].success.mean() >= 0.9 and 2000 < by_name["Rastrigin_10"].med < 9000
    # the same seeds through the oracle: same first hits on most runs
    f = sb.Rastrigin(2)
    o = oracle.run(oracle.Problem("Rastrigin", f.bounds), 20, seeds=np.arange(20, dtype=np.uint64) + np.uint64(1234),
                   maxiter=int(1e6), max_calls=100000, fglob=0.0, tol=1e-8,
                   tables=sb._sda.temperature_tables(int(1e6)))
    assert np.mean(u.values()["ncall"] == o["ncall_hit"]) >= 0.6
    out = os.path.join(folder, "results.csv")
    report(folder, kind="csv", path=out)
    rows = list(csv.reader(open(out), delimiter=",", quotechar="|"))
    assert len(rows) == 14 and rows[0][0] == "Function name"
    # a second run skips existing files (bench.py:164-167)
    assert Benchmarker(20, folder, names=["Rastrigin", "Sphere"], max_calls=100000).run() == []


# ---- user-supplied CUDA device functor --------------------------------------------------
USER_SRC = r"""
// shifted, scaled Rastrigin written by a "user": params = {shift, amplitude}
template <class X>
__device__ double sda_user_objective(const X &x, int dim, const double *params)
{
    double s = 0.0;
    for (int k = 0; k < dim; ++k) {
        const double d = x[k] - params[0];
        s += d * d - params[1] * cos(6.283185307179586 * d);
    }
    return s + params[1] * dim;
}
"""


def test_user_cuda_functor(sb):
    f = sb.functors.compile_user_functor(USER_SRC, params=[0.5, 10.0], name="my_rastrigin")
    X = np.random.RandomState(3).uniform(-4, 5, (32, 6))
    want = np.sum((X - 0.5) ** 2 - 10.0 * np.cos(2 * np.pi * (X - 0.5)), axis=1) + 10.0 * 6
    np.testing.assert_allclose(f(X), want, rtol=1e-12, atol=1e-11)
    ret = sb.sda(f, None, [(-4.62, 5.62)] * 6, seed=7)
    assert ret.fun < 1e-10 and np.allclose(ret.x, 0.5, atol=1e-5) and ret.nit == 1000
    # the default library does not know the functor id
    from sdaopt_b200 import _lib
    assert _lib.lib().sda_has_user_functor() == 0 and _lib.lib(f.library_path).sda_has_user_functor() == 1


# ---- the north star's statistical criterion against the REFERENCE's own runs ---------------------------
def test_suite_statistics_match_reference(sb):
    """tests/golden/suite_stats_ref.npz holds NB_RUNS = 100 suite-mode runs of the live reference
    (seeds 1234+i, bench.py:170-216) for 15 jobs.  The GPU runs use Philox keys, so the two samples
    are independent draws of size 100 from (if the port is right) the same distribution: success rate
    inside the reference's Wilson 95 % interval; two-sample Mann-Whitney and Kolmogorov-Smirnov tests on
    ncall-to-global that do not reject at 1e-3; and the GPU median inside the reference's 95 %
    order-statistic interval (ranks 40..61 of 100) for most jobs -- for two independent samples that
    happens ~83 % of the time per job even for identical distributions (Levy13, IQR 324..1153: oracle
    tape vs Philox medians 608 vs 634 at n = 1000, p = 0.28), hence a count, not a per-job assert."""
    from scipy.stats import mannwhitneyu, ks_2samp
    rows = load_golden("suite_stats_ref.npz")["rows"]
    inside95 = 0
    for r in rows:
        f = sb.DeviceFunctor(r["functor"], dimensions=r["dim"])
        n = 100
        seeds = np.arange(n, dtype=np.uint64) + np.uint64(1234)
        g = sb.sda_batch(f, r["bounds"], n, seeds=seeds, maxiter=int(1e6), max_calls=int(r["max_calls"]),
                         fglob=r["fglob"], tol=1e-8)
        lo, hi = _wilson(int(r["success"].sum()), n)
        assert lo - 1e-9 <= g.hit.mean() <= hi + 1e-9, (r["functor"], r["dim"], g.hit.mean(), lo, hi)
        so = np.sort(r["ncall"])
        med = np.median(g.ncall_hit)
        inside95 += int(so[39] <= med <= so[60])
        p = mannwhitneyu(g.ncall_hit, r["ncall"], alternative="two-sided").pvalue
        pk = ks_2samp(g.ncall_hit, r["ncall"]).pvalue
        assert p > 1e-3 and pk > 1e-3, (r["functor"], r["dim"], p, pk, med, np.median(r["ncall"]))
    assert inside95 >= 0.6 * len(rows), inside95


@pytest.mark.parametrize("fixture,min_jobs", [("suite_stats_ref_full.npz", 200), ("suite_stats_ref_1e6.npz", 20)])
def test_suite_statistics_match_reference_whole_catalogue(sb, fixture, min_jobs):
    """The north star's second criterion over the catalogue, against the REFERENCE's own runs.
    suite_stats_ref_full.npz: 219 jobs at a 200,000-evaluation budget (oracle/gen_suite_ref_full.py,
    gen_suite_ref_nd.py).  suite_stats_ref_1e6.npz: the discriminating jobs at the suite's real budget
    of 1e6 -- every partial-success job, every never-succeed job, the n-D jobs at D = 70..100 (20
    reference runs each; oracle/gen_suite_ref_1e6.py).  Two independent samples per job (the GPU
    uses Philox keys), so the criterion is statistical and family-wise: see pm.suite_summary."""
    rows = pm.suite_compare(sb, fixture)
    s = pm.suite_summary(rows)
    assert s["jobs"] >= min_jobs
    if fixture == "suite_stats_ref_full.npz":
        assert {"Stochastic", "XinSheYang01"} <= {r["job"] for r in rows}
    assert s["family_wise_ok"], s
    assert s["rejections_at_05"] <= s["rejections_allowed"], s
    assert 0.95 <= s["geo_mean_median_ratio"] <= 1.05, s
    assert s["inside_wilson95"] >= 0.90, s
    # jobs the reference never solves are not solved here either, and vice versa
    for r in rows:
        if r["ref_runs"] >= 100 and r["ref_success"] == 0:
            assert r["gpu_success"] <= 5, r
        if r["ref_runs"] >= 100 and r["ref_success"] == r["ref_runs"]:
            assert r["gpu_success"] >= 95, r


# ---- the two execution paths of sda_batch_run ---------------------------------------------------------
def test_phased_and_persistent_paths_are_bit_identical(sb):
    """Large Philox batches run as bounded rounds of {annealing launch, local-search launch}; the
    single persistent launch is the other path.  Same chain code, same draws: identical bits."""
    from sdaopt_b200 import _lib
    f = sb.Rastrigin(20)
    out = {}
    old = os.environ.get("SDA_PHASED")
    try:
        for mode in ("0", "1"):
            os.environ["SDA_PHASED"] = mode
            out[mode] = sb.sda_batch(f, f.bounds, 1500, maxiter=90)
            n = _lib.lib().sda_last_launch_count()
            assert (n == 1) if mode == "0" else (n > 1)
    finally:
        if old is None:
            os.environ.pop("SDA_PHASED", None)
        else:
            os.environ["SDA_PHASED"] = old
    for k in ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls", "status"):
        np.testing.assert_array_equal(out["0"][k], out["1"][k])
    assert (out["1"].n_ls > 0).all() and (out["1"].status == 0).all()


def test_block_synchronous_local_search_launch_is_bit_identical(sb):
    """SDA_LS_SYNC=1: the local-search launch of the phased path with its warps in lock step on the
    phase clock (one CTA of 16 warps per SM).  Only the schedule changes, not one bit of a chain."""
    f = sb.Rastrigin(20)
    out = {}
    keep = {k: os.environ.get(k) for k in ("SDA_PHASED", "SDA_LS_SYNC")}
    try:
        os.environ["SDA_PHASED"] = "1"
        for mode in ("0", "1"):
            os.environ["SDA_LS_SYNC"] = mode
            out[mode] = sb.sda_batch(f, f.bounds, 3000, maxiter=60)
    finally:
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for k in ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls", "status"):
        np.testing.assert_array_equal(out["0"][k], out["1"][k])
    assert (out["1"].n_ls > 0).all() and (out["1"].status == 0).all()



def test_memoised_gradient_terms_are_bit_identical(sb):
    """Separable objectives (Rastrigin, Ackley01, Schwefel26 ...): the 2 D perturbed evaluations of a
    finite-difference gradient reuse the base point's terms (sda_objectives.cuh: separable_term).
    SDA_LS_MEMO=0 evaluates every perturbed point in full; both must give the same bits everywhere --
    minimum, point, call counts, first hits."""
    keep = os.environ.get("SDA_LS_MEMO")
    try:
        for name, dim, n, kw in (("Rastrigin", 50, 600, dict(maxiter=40)), ("Ackley01", 10, 300, dict(maxiter=60)),
                                 ("Schwefel26", 100, 100, dict(maxiter=12)), ("Exponential", 7, 200, dict(maxiter=30)),
                                 ("Rastrigin", 5, 100, dict(maxiter=int(1e6), max_calls=30000, fglob=0.0, tol=1e-8))):
            f = sb.DeviceFunctor(name, dimensions=dim)
            seeds = np.arange(n, dtype=np.uint64) + np.uint64(77)
            out = {}
            for mode in ("0", "1"):
                os.environ["SDA_LS_MEMO"] = mode
                out[mode] = sb.sda_batch(f, f.bounds, n, seeds=seeds, **kw)
            for k in ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls", "status", "hit", "ncall_hit", "f_hit"):
                np.testing.assert_array_equal(out["0"][k], out["1"][k], err_msg="%s D=%d %s" % (name, dim, k))
            assert (out["1"].n_ls > 0).all()
    finally:
        if keep is None:
            os.environ.pop("SDA_LS_MEMO", None)
        else:
            os.environ["SDA_LS_MEMO"] = keep


def test_results_do_not_depend_on_batch_size_at_fixed_group_width(sb):
    """A chain is a pure function of (problem, config, Philox key, lanes per chain).  The library picks
    the lanes per chain from the dimension AND the batch size (small batches get wider groups), and the
    partial sums of an objective follow the lane count -- so across batch sizes results are guaranteed
    identical only at a fixed width (SDA_GROUP), which is what a sharded run that must reproduce a
    single-GPU run bit for bit sets.  Pinned here: same keys, batches of 7 / 300 / 5,000 chains."""
    keep = os.environ.get("SDA_GROUP")
    try:
        os.environ["SDA_GROUP"] = "8"
        f = sb.Rastrigin(50)
        seeds = np.arange(5000, dtype=np.uint64) + np.uint64(99)
        big = sb.sda_batch(f, f.bounds, 5000, seeds=seeds, maxiter=12)
        for n in (7, 300):
            small = sb.sda_batch(f, f.bounds, n, seeds=seeds[:n], maxiter=12)
            for k in ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls"):
                np.testing.assert_array_equal(small[k], big[k][:n], err_msg="%d %s" % (n, k))
    finally:
        if keep is None:
            os.environ.pop("SDA_GROUP", None)
        else:
            os.environ["SDA_GROUP"] = keep


def test_concurrent_host_threads_with_different_problems(sb):
    """sda_batch_run from two host threads at once (ctypes releases the GIL), different dimensions and
    batch sizes: the workspace layout is per call (no file-scope state), so each thread gets exactly
    what it gets alone."""
    import threading
    jobs = [(sb.Rastrigin(50), 20000, dict(maxiter=10)), (sb.Schwefel26(7), 333, dict(maxiter=50)),
            (sb.Ackley01(100), 900, dict(maxiter=6)), (sb.Rosenbrock(2), 64, dict(maxiter=200))]
    alone = [sb.sda_batch(f, f.bounds, n, **kw) for f, n, kw in jobs]
    for _ in range(3):
        got = [None] * len(jobs)

        def work(i):
            f, n, kw = jobs[i]
            got[i] = sb.sda_batch(f, f.bounds, n, **kw)
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
        [t.start() for t in th]
        [t.join() for t in th]
        for a, b in zip(alone, got):
            for k in ("x", "fun", "nit", "ncall", "ncall_ls", "n_ls", "status"):
                np.testing.assert_array_equal(a[k], b[k])
