"""Host-side mirror of the reference interface ``sdaopt/_sda.py`` over the CUDA library.

Same names, argument meaning and error behaviour as the reference (file:line citations are to
/root/reference/sdaopt/_sda.py); all arithmetic of the hot path runs in libsdaopt_b200.so:

* ``sda(func, x0, bounds, maxiter, minimizer_kwargs, initial_temp, visit, accept, maxfun, seed,
  pure_sa)`` -> ``OptimizeResult{x, fun, nit, ncall}``                                 (:431-589)
* ``SDARunner`` (:354-428), ``VisitingDistribution`` (:20-107), ``EnergyState`` (:110-166),
  ``ObjectiveFunWrapper`` (:276-351) -- the unit-test surface of ``sdaopt/__init__.py:9-13``
* ``sda_batch(...)``: the batched form -- many independent runs in one persistent launch.
"""
import ctypes as C

import numpy as np
from scipy.optimize import OptimizeResult
from scipy.special import gammaln
from scipy._lib._util import check_random_state

from . import _lib
from . import functors as _functors

__all__ = ['sda', 'sda_batch']
BIG_VALUE = 1e16


# ------------------------------------------------------------------------------------------------
# K1: temperature schedule and sigmax(T) -- evaluated on the host with the reference's own
# expressions so that the constants the kernels consume are bit-identical to the reference's.
# ------------------------------------------------------------------------------------------------
def temperature_tables(maxiter, initial_temp=5230., qv=2.62, temperature_restart=0.1):
    """T_i (:396-402) and sigmax(T_i) (:76-86) for i = 0 .. first index with T_i < restart."""
    t1 = np.exp((qv - 1) * np.log(2.0)) - 1.0
    factor2 = np.exp((4.0 - qv) * np.log(qv - 1.0))
    factor3 = np.exp((2.0 - qv) * np.log(2.0) / (qv - 1.0))
    factor5 = 1.0 / (qv - 1.0) - 0.5
    d1 = 2.0 - factor5
    factor6 = np.pi * (1.0 - factor5) / np.sin(np.pi * (1.0 - factor5)) / np.exp(gammaln(d1))
    temps, sigs = [], []
    for i in range(int(max(maxiter, 1))):
        s = float(i) + 2.0
        t2 = np.exp((qv - 1) * np.log(s)) - 1.0
        temperature = initial_temp * t1 / t2
        temps.append(temperature)
        sigs.append(sigmax(temperature, qv, factor2, factor3, factor6))
        if temperature < temperature_restart:
            break
    return np.array(temps, dtype=np.float64), np.array(sigs, dtype=np.float64)


def sigmax(temperature, qv, factor2=None, factor3=None, factor6=None):
    """The temperature-only constant of visit_fn (:76-86)."""
    factor1 = np.exp(np.log(temperature) / (qv - 1.0))
    if factor2 is None:
        factor2 = np.exp((4.0 - qv) * np.log(qv - 1.0))
        factor3 = np.exp((2.0 - qv) * np.log(2.0) / (qv - 1.0))
        factor5 = 1.0 / (qv - 1.0) - 0.5
        d1 = 2.0 - factor5
        factor6 = np.pi * (1.0 - factor5) / np.sin(np.pi * (1.0 - factor5)) / np.exp(gammaln(d1))
    factor4 = np.sqrt(np.pi) * factor1 * factor2 / (factor3 * (3.0 - qv))
    return float(np.exp(-(qv - 1.0) * np.log(factor6 / factor4) / (3.0 - qv)))


def _split_bounds(bounds):
    lu = list(zip(*bounds))
    return np.array(lu[0], dtype=np.float64), np.array(lu[1], dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data


# ------------------------------------------------------------------------------------------------
# batched runs
# ------------------------------------------------------------------------------------------------
class BatchResult(dict):
    """Per-chain arrays: x[C,D], fun, nit, ncall, ncall_ls, n_ls, status, hit, ncall_hit, f_hit
    (+ trace_e, trace_d, tape_used when requested)."""
    __getattr__ = dict.__getitem__


def _make_config(maxiter, initial_temp, visit, accept, maxfun, pure_sa, rng_mode, max_calls, fglob,
                 tol, trace_len, tables):
    temps, sigs = tables
    cfg = _lib.SdaConfig(int(maxiter), float(initial_temp), float(visit), float(accept),
                         float(maxfun), int(bool(pure_sa)), 0.1, int(rng_mode), int(max_calls),
                         float(fglob), float(tol), int(trace_len), temps.ctypes.data,
                         sigs.ctypes.data, int(temps.size))
    return cfg


def sda_batch(func, bounds, n_chains, x0=None, seeds=None, tape=None, maxiter=1000,
              initial_temp=5230., visit=2.62, accept=-5.0, maxfun=1e7, pure_sa=False,
              max_calls=0, fglob=float("nan"), tol=1e-8, trace_len=0, device=0):
    """Run ``n_chains`` independent ``sda`` runs of one problem in a single launch.

    ``seeds[C]`` are the per-chain Philox keys (default: chain index).  With ``tape[C, L]`` the
    chains consume those uniforms in the reference's draw order instead (trajectory replay).
    ``max_calls``/``fglob``/``tol`` switch on the suite's first-hit monitor
    (benchmark/bench.py:292-321).  Returns a :class:`BatchResult` of numpy arrays.
    """
    functor = _functors.resolve(func)
    lower, upper = _split_bounds(bounds)
    dim = lower.size
    n_chains = int(n_chains)
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        if x0.size != n_chains * dim:
            raise ValueError('Bounds size does not match x0')                 # :360-361
    if not np.all(lower < upper):
        raise ValueError('Bounds are note consistent min < max')              # :366-367
    if seeds is not None:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        if seeds.size != n_chains:
            raise ValueError("seeds must have one entry per chain")
    tape_len = 0
    if tape is not None:
        tape = np.ascontiguousarray(tape, dtype=np.float64).reshape(n_chains, -1)
        tape_len = tape.shape[1]
    problem, keep = _functors.make_problem(functor, lower, upper)
    _lib.require_cuda()
    tables = temperature_tables(maxiter, initial_temp, visit, 0.1)
    cfg = _make_config(maxiter, initial_temp, visit, accept, maxfun, pure_sa,
                       _lib.RNG_TAPE if tape is not None else _lib.RNG_PHILOX, max_calls, fglob,
                       tol, trace_len, tables)
    tl = max(int(trace_len), 0)
    out = BatchResult(
        x=np.zeros((n_chains, dim)), fun=np.zeros(n_chains), nit=np.zeros(n_chains, np.int64),
        ncall=np.zeros(n_chains, np.int64), ncall_ls=np.zeros(n_chains, np.int64),
        n_ls=np.zeros(n_chains, np.int64), status=np.zeros(n_chains, np.int32),
        hit=np.zeros(n_chains, np.int32), ncall_hit=np.zeros(n_chains, np.int64),
        f_hit=np.zeros(n_chains), trace_e=np.zeros((n_chains, max(tl, 1))),
        trace_d=np.zeros((n_chains, max(tl, 1)), np.uint8), tape_used=np.zeros(n_chains, np.int64))
    res = _lib.SdaResult(*[out[k].ctypes.data for k in _lib.RESULT_FIELDS])
    rc = _lib.lib(functor.library_path).sda_batch_run_host(C.byref(problem), C.byref(cfg), n_chains, _ptr(x0),
                                       _ptr(seeds), _ptr(tape), tape_len, C.byref(res), int(device))
    _lib.check(rc)
    out["trace_e"] = out["trace_e"][:, :tl]
    out["trace_d"] = out["trace_d"][:, :tl]
    return out


def _raise_chain_errors(status):
    if np.any(status & _lib.CHAIN_REINIT_FAILED):
        raise ValueError('Stopping algorithm because function '                # :151-156
                         'create NaN or (+/-) inifinity values even with '
                         'trying new random parameters')


# ------------------------------------------------------------------------------------------------
# the reference's classes
# ------------------------------------------------------------------------------------------------
class _TapeFeeder(object):
    """Hands a device kernel the next uniforms of a RandomState and then advances the real
    generator by exactly the number the kernel consumed (the draw count is data dependent)."""

    def __init__(self, rs):
        self.rs = rs

    def run(self, n_guess, fn):
        while True:
            state = self.rs.get_state()
            tape = self.rs.random_sample(int(n_guess))
            self.rs.set_state(state)
            used, value = fn(tape)
            if used >= 0:
                if used:
                    self.rs.random_sample(int(used))
                return value, tape[:used]
            n_guess *= 4


class VisitingDistribution(object):
    """Distorted Cauchy-Lorentz visiting distribution (:20-107), sampled by the CUDA kernels of
    the annealing path from the uniforms of ``rs`` (consumed in the reference's order)."""
    tail_limit = 1.e8
    min_visit_bound = 1.e-10

    def __init__(self, lb, ub, qv, rs, device=0):
        self.qv = qv
        self.rs = rs
        self.lower = np.ascontiguousarray(lb, dtype=np.float64)
        self.upper = np.ascontiguousarray(ub, dtype=np.float64)
        self.b_range = self.upper - self.lower
        self.x_gauss = None
        self.s_gauss = 0
        self.root_gauss = None
        self._device = device
        self._feed = _TapeFeeder(rs)
        self._problem, self._keep = _functors.make_problem(
            _functors.DeviceFunctor("Sphere"), self.lower, self.upper)

    def visiting(self, x, step, temperature):
        """:40-73"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        dim = x.size
        sig = sigmax(temperature, self.qv)
        _lib.require_cuda()

        def fn(tape):
            xv = np.empty(dim)
            used = C.c_int64()
            _lib.check(_lib.lib().sda_visiting_host(
                C.byref(self._problem), float(self.qv), x.ctypes.data, int(step),
                float(temperature), sig, tape.ctypes.data, tape.size, xv.ctypes.data,
                C.byref(used), self._device))
            return used.value, xv
        value, _ = self._feed.run(8 * dim + 64, fn)
        return value

    def _visit_fn(self, qv, sig):
        _lib.require_cuda()

        def fn(tape):
            out = np.empty(1)
            used = C.c_int64()
            _lib.check(_lib.lib().sda_visit_fn_host(float(qv), float(sig), tape.ctypes.data,
                                                     tape.size, out.ctypes.data, 1, C.byref(used),
                                                     self._device))
            return used.value, float(out[0])
        value, used_tape = self._feed.run(64, fn)
        # the accepted polar pair is the last two uniforms consumed (:96-104)
        self.x_gauss = used_tape[-2] * 2.0 - 1.0
        y_gauss = used_tape[-1] * 2.0 - 1.0
        self.s_gauss = self.x_gauss ** 2 + y_gauss ** 2
        self.root_gauss = np.sqrt(-2.0 / self.s_gauss * np.log(self.s_gauss))
        return value

    def visit_fn(self, temperature):
        """:75-91"""
        return self._visit_fn(self.qv, sigmax(temperature, self.qv))

    def gaussian_fn(self, axis):
        """:93-107.  With qv = 1 and sigmax = 1 the visit kernel returns root * y_gauss itself."""
        if axis == 1:
            return self._visit_fn(1.0, 1.0)
        return self.root_gauss * self.x_gauss


class ObjectiveFunWrapper(object):
    """Objective + call counter + default local minimiser (:276-351) over a device functor."""

    def __init__(self, bounds, func, **kwargs):
        self.func = _functors.resolve(func)
        self.nb_fun_call = 0
        self.kwargs = kwargs
        # the kernel runs exactly the reference's default minimiser (method L-BFGS-B, jac =
        # self.gradient with eps = 1e-6, bounds = the search box, option dict :300-311); a kwarg that
        # would change any of that cannot be honoured on the device, so it is refused, not ignored
        unsupported = set(kwargs) - {'method'}
        if unsupported or kwargs.get('method', 'L-BFGS-B') != 'L-BFGS-B':
            raise NotImplementedError(
                "sdaopt_b200 runs the reference's default local search (L-BFGS-B with the option "
                "dict of _sda.py:300-311) on the GPU; minimizer_kwargs %r are not supported"
                % sorted(kwargs))
        self.lower, self.upper = _split_bounds(bounds)
        self.ls_max_iter = min(max(self.lower.size * 6, 100), 1000)               # :291-296
        self.reps = 1.e-6                                                         # :315
        self.fun_args = ()

    def func_wrapper(self, x):
        """:321-323"""
        self.nb_fun_call += 1
        return self.func(np.asarray(x, dtype=np.float64))

    def local_search(self, x, maxlsiter=None, device=0):
        """:347-351 -> (e, x) or (BIG_VALUE, None) when the minimiser did not converge."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(1, -1)
        problem, keep = _functors.make_problem(self.func, self.lower, self.upper)
        xo = np.empty_like(x)
        fo = np.empty(1)
        ok = np.zeros(1, np.int32)
        nfev = np.zeros(1, np.int64)
        _lib.require_cuda()
        _lib.check(_lib.lib(self.func.library_path).sda_local_search_host(C.byref(problem), x.ctypes.data, 1,
                                                     xo.ctypes.data, fo.ctypes.data, ok.ctypes.data,
                                                     nfev.ctypes.data, device))
        self.nb_fun_call += int(nfev[0])
        if not ok[0]:
            return BIG_VALUE, None
        return float(fo[0]), xo[0]


class EnergyState(object):
    """The reference's ``EnergyState`` (:110-166) as a host handle on device state.

    ``reset`` runs the (re)initialisation rule on the GPU (``sda_energy_reset_host``: the same
    ``energy_reset`` device function the annealing kernels call at chain start and on re-annealing)
    with the uniforms of ``rs`` injected in the reference's draw order, then advances ``rs`` by the
    number of uniforms the kernel consumed.  Attributes mirror the reference's: ``current_location``,
    ``current_energy``, ``xbest``, ``ebest`` (the best is kept across resets, :163-165)."""
    MAX_REINIT_COUNT = 1000

    def __init__(self, lower, upper):
        self.lower = np.ascontiguousarray(lower, dtype=np.float64)
        self.upper = np.ascontiguousarray(upper, dtype=np.float64)
        self.current_location = self.current_energy = self.xbest = self.ebest = None

    def reset(self, owf, rs, x0=None, device=0):
        functor = owf.func if isinstance(owf, ObjectiveFunWrapper) else _functors.resolve(owf)
        dim = self.lower.size
        start = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        if start is not None and start.size != dim:
            raise ValueError('Bounds size does not match x0')
        problem, keep = _functors.make_problem(functor, self.lower, self.upper)
        have_best = self.ebest is not None
        _lib.require_cuda()

        def fn(tape):
            loc, best = np.empty(dim), np.zeros(dim)
            if have_best:
                best[:] = self.xbest
            e, eb = C.c_double(), C.c_double(self.ebest if have_best else 0.0)
            status, used = C.c_int32(), C.c_int64()
            _lib.check(_lib.lib(functor.library_path).sda_energy_reset_host(
                C.byref(problem), _ptr(start), tape.ctypes.data, tape.size, int(have_best),
                loc.ctypes.data, C.byref(e), best.ctypes.data, C.byref(eb), C.byref(status),
                C.byref(used), int(device)))
            return used.value, (loc, e.value, best, eb.value, status.value)

        (loc, e, best, eb, status), _ = _TapeFeeder(rs).run(4 * dim + 16, fn)
        _raise_chain_errors(np.array([status]))
        self.current_location, self.current_energy = loc, e
        self.xbest, self.ebest = best, eb


class SDARunner(object):
    """:354-428.  ``search()`` runs the whole annealing + local-search loop on the GPU."""
    MAX_REINIT_COUNT = 1000

    def __init__(self, fun, x0, bounds, seed=None, minimizer_kwargs=None,
                 temperature_start=5230, qv=2.62, qa=-5.0,
                 maxfun=1e7, maxsteps=500, pure_sa=False, rng='philox', device=0):
        if x0 is not None and not len(x0) == len(bounds):
            raise ValueError('Bounds size does not match x0')
        lower, upper = _split_bounds(bounds)
        if not np.all(lower < upper):
            raise ValueError('Bounds are note consistent min < max')
        if minimizer_kwargs is None:
            minimizer_kwargs = dict()
        self.owf = ObjectiveFunWrapper(bounds, fun, **minimizer_kwargs)
        self.rs = check_random_state(seed)
        self.es = EnergyState(lower, upper)
        self._x0 = None if x0 is None else np.array(x0, dtype=np.float64)
        self._bounds = bounds
        self.maxfun = maxfun
        self.maxsteps = maxsteps
        self.temperature_start = temperature_start
        self.temperature_restart = 0.1
        self.qv = qv
        self.qa = qa
        self.pure_sa = pure_sa
        if rng not in ('philox', 'tape'):
            raise ValueError("rng must be 'philox' or 'tape'")
        self._rng = rng
        self._device = device
        self._iter = 0

    def _tape_guess(self):
        d = self.es.lower.size
        steps = max(int(self.maxsteps), 1)
        per_step = 2.7 * (d * d + d) + 6 * d + 16
        return int(per_step * steps * 1.15 + 4096 + 1100 * d)

    def search(self):
        """:393-418 (plus the initial EnergyState.reset of :375-376) in one device call."""
        common = dict(x0=self._x0, maxiter=self.maxsteps, initial_temp=self.temperature_start,
                      visit=self.qv, accept=self.qa, maxfun=self.maxfun, pure_sa=self.pure_sa,
                      device=self._device)
        if self._rng == 'tape':
            # the reference's own RandomState draws, injected (exact draw protocol)
            n = self._tape_guess()
            while True:
                state = self.rs.get_state()
                tape = self.rs.random_sample(n)
                self.rs.set_state(state)
                r = sda_batch(self.owf.func, self._bounds, 1, tape=tape, **common)
                if not (r.status[0] & _lib.CHAIN_TAPE_EXHAUSTED):
                    self.rs.random_sample(int(r.tape_used[0]))
                    break
                n *= 2
        else:
            # Philox key drawn from the RandomState: seeded runs are reproducible (test_reproduce)
            key = np.array([int(self.rs.randint(0, 2 ** 31 - 1)) * (2 ** 31)
                            + int(self.rs.randint(0, 2 ** 31 - 1))], dtype=np.uint64)
            r = sda_batch(self.owf.func, self._bounds, 1, seeds=key, **common)
        _raise_chain_errors(r.status)
        self.es.xbest = r.x[0].copy()
        self.es.ebest = float(r.fun[0])
        self._iter = int(r.nit[0])
        self.owf.nb_fun_call = int(r.ncall[0])
        self._last = r

    @property
    def result(self):
        """ The OptimizeResult (:420-428) """
        res = OptimizeResult()
        res.x = self.es.xbest
        res.fun = self.es.ebest
        res.nit = self._iter
        res.ncall = self.owf.nb_fun_call
        return res


def sda(func, x0, bounds, maxiter=1000, minimizer_kwargs=None,
        initial_temp=5230., visit=2.62, accept=-5.0, maxfun=1e7, seed=None,
        pure_sa=False, rng='philox', device=0):
    """
    Find the global minimum of a function using the Simulated Dual Annealing algorithm on a B200.

    Same signature, defaults and result fields as the reference ``sda`` (_sda.py:431-589):
    ``OptimizeResult`` with ``x``, ``fun``, ``nit``, ``ncall``.  ``func`` must be a device
    objective (see :mod:`sdaopt_b200.functors`); ``seed`` follows ``check_random_state``.
    Extra keywords: ``rng='philox'`` (production, one counter-based stream per run keyed from
    ``seed``) or ``rng='tape'`` (the ``RandomState`` uniforms themselves are injected, which
    replays the reference's draw protocol); ``device`` = CUDA ordinal.
    """
    gr = SDARunner(func, x0, bounds, seed, minimizer_kwargs,
                   temperature_start=initial_temp, qv=visit, qa=accept,
                   maxfun=maxfun, maxsteps=maxiter, pure_sa=pure_sa, rng=rng, device=device)
    gr.search()
    return gr.result
