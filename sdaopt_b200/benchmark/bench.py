"""The SDA benchmark runs on the GPU.

Reference semantics kept (sdaopt/benchmark/bench.py): job list = every catalogue class at its default
dimension plus the n-D selection at D in {5, 10, 20, ..., 100} (:48-55, :103-118); NB_RUNS runs per
job, run i seeded 1234 + i (:171); `sda(func, x0=None, bounds, maxiter=1e6, pure_sa=False)` (:328-333);
a run succeeds at the first evaluation with f <= fglob + 1e-8 before the 1,000,000-th call, failures
record ncall = 1e6 (:292-321, :269).  One job = ONE launch of NB_RUNS chains (first-hit monitor inside
the kernel); jobs are dealt to the ranks by expected cost and run concurrently on CUDA streams.

Only objectives with a registered device functor can run; the rest of the 195-class catalogue is
reported as `not ported`.
"""
import logging
import os
import time

import numpy as np

from .. import functors
from .. import sharding
from .benchunit import BenchUnit

logger = logging.getLogger(__name__)

DEFAULT_TOL = 1.e-8          # bench.py:35
MAX_FN_CALL = 1e6            # bench.py:36
MAX_IT = int(1e6)            # bench.py:38
DIMENSIONS = [5] + list(range(10, 110, 10))                                   # bench.py:48-50
N_DIM_FUNC_SELECTION = ['Ackley01', 'Exponential', 'Rastrigin', 'Rosenbrock', 'Schwefel01']  # :52-55
# catalogue classes with a device functor, at the default dimension of their class (N = 2)
CATALOGUE = sorted(['Ackley01', 'Exponential', 'Griewank', 'Rastrigin', 'Rosenbrock', 'Schwefel01',
                    'Schwefel26', 'Sphere'] + list(functors._CATALOGUE_ALL))


def build_job_list(names=None):
    """[(job name, functor name, dim)] in the reference's order (bench.py:103-118)."""
    jobs = []
    for name in sorted(names or CATALOGUE):
        if name in N_DIM_FUNC_SELECTION:
            for dim in DIMENSIONS:
                jobs.append(('{0}_{1}'.format(name, dim), name, dim))
        jobs.append((name, name, functors.DeviceFunctor(name, dimensions=None if name in functors._CATALOGUE_ALL else 2).N))
    return jobs


def expected_cost(job):
    """Static cost model for sharding: evaluations grow with D (annealing 2D per step, gradient 2D
    per request) -- enough to spread the heavy n-D jobs over the ranks."""
    return float(job[2]) ** 2


class Benchmarker(object):
    """`Benchmarker(nbruns, folder).run()` as in bench.py:84-153, SDA only."""

    def __init__(self, nbruns, folder, names=None, max_calls=MAX_FN_CALL, rank=0, world=1,
                 device=0):
        self.nbruns = int(nbruns)
        self.folder = folder
        self.jobs = build_job_list(names)
        self.max_calls = int(max_calls)
        self.rank, self.world, self.device = rank, world, device

    def my_jobs(self):
        idx = sharding.shard_by_cost([expected_cost(j) for j in self.jobs], self.world, self.rank)
        return [self.jobs[i] for i in idx]

    def run(self):
        import torch
        from ..device import DeviceBatch
        os.makedirs(self.folder, exist_ok=True)
        todo = []
        for job in self.my_jobs():
            bu = BenchUnit(self.nbruns, job[0], 'SDA', max_fn_call=self.max_calls)
            if os.path.exists(os.path.join(self.folder, bu.filename)):
                logger.info('File {} already existing, skipping...'.format(bu.filename))
                continue
            todo.append((job, bu))
        t0 = time.time()
        torch.cuda.set_device(self.device)
        launched = []
        for job, bu in todo:
            f = functors.DeviceFunctor(job[1], dimensions=job[2])
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                batch = DeviceBatch(f, f.bounds, self.nbruns, maxiter=MAX_IT, max_calls=self.max_calls,
                                    fglob=f.fglob, tol=DEFAULT_TOL, device=self.device)
                batch.set_seeds(1234 + np.arange(self.nbruns))            # bench.py:171
                batch.launch()
            launched.append((job, bu, batch, stream))
        for job, bu, batch, stream in launched:
            stream.synchronize()
            with torch.cuda.stream(stream):
                r = batch.fetch()
            dt = time.time() - t0
            self.record(bu, r, dt)
            bu.write(self.folder)
            logger.info('Func: {0} - Algo: SDA -> success {1:.0f}% median ncall {2}'.format(
                job[0], 100 * r.hit.mean(), bu.medall))
        return [bu for _, bu in todo]

    def record(self, bu, r, seconds):
        """bench.py:184-188 / :211-215 for every run of the job."""
        ok = r.hit.astype(bool)
        bu._values['success'][:] = ok
        bu._values['ncall'][:] = np.where(ok, r.ncall_hit, self.max_calls)
        bu._values['fvalue'][:] = np.where(ok, r.f_hit, np.inf)
        bu._values['ncall_max'][:] = np.where(ok, r.ncall_hit, self.max_calls)
        # the chains of a job run concurrently: the per-run time is the job's wall time / nbruns
        bu._values['time'][:] = np.where(ok, seconds / max(self.nbruns, 1), np.inf)
