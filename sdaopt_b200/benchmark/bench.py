"""The SDA benchmark runs on the GPU.

Reference semantics kept (sdaopt/benchmark/bench.py): job list = every catalogue class at its default
dimension plus the n-D selection at D in {5, 10, 20, ..., 100} (:48-55, :103-118); NB_RUNS runs per
job, run i seeded 1234 + i (:171); `sda(func, x0=None, bounds, maxiter=1e6, pure_sa=False)` (:328-333);
a run succeeds at the first evaluation with f <= fglob + 1e-8 before the 1,000,000-th call, failures
record ncall = 1e6 (:292-321, :269).  One job = ONE launch of its runs (first-hit monitor inside the
kernel); all jobs are in flight at once on CUDA streams; the runs of every job are sharded over the ranks.

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
    """Launch-ordering cost of a job: measured mean evaluations per run (suite_costs.py; unknown jobs
    count as budget-long) times a per-evaluation weight that grows with the dimension."""
    from .suite_costs import MEAN_EVALS
    return float(MEAN_EVALS.get(job[0], MAX_FN_CALL)) * (8.0 + float(job[2]))


class Benchmarker(object):
    """`Benchmarker(nbruns, folder).run()` as in bench.py:84-153, SDA only.

    The work list is the (job, run) pairs.  Every rank takes a contiguous slice of the RUNS of every
    job (sharding.shard_range: run i keeps its seed 1234 + i whatever the number of ranks), so the
    ranks carry equal shares of every job whatever its cost, launches all its jobs at once through one
    BatchPool (long jobs first, nothing between the launches synchronises), and the per-run records
    are gathered on rank 0 (sharding.gather_ragged: NCCL on GPUs), which writes the BenchUnit files."""

    def __init__(self, nbruns, folder, names=None, max_calls=MAX_FN_CALL, rank=0, world=1,
                 device=0):
        self.nbruns = int(nbruns)
        self.folder = folder
        self.jobs = build_job_list(names)
        self.max_calls = int(max_calls)
        self.rank, self.world, self.device = rank, world, device
        self.stats = {}

    def run(self, write=True):
        import torch
        from ..device import BatchPool
        if write:
            os.makedirs(self.folder, exist_ok=True)
        todo = []
        for job in self.jobs:
            bu = BenchUnit(self.nbruns, job[0], 'SDA', max_fn_call=self.max_calls)
            if write and os.path.exists(os.path.join(self.folder, bu.filename)):
                logger.info('File {} already existing, skipping...'.format(bu.filename))
                continue
            todo.append((job, bu))
        if not todo:
            return []
        todo.sort(key=lambda jb: -expected_cost(jb[0]))              # long jobs first
        begin, end = sharding.shard_range(self.nbruns, self.world, self.rank)
        t0 = time.time()
        torch.cuda.set_device(self.device)
        dev = torch.device('cuda', self.device)
        rec = torch.zeros((0, 6), dtype=torch.float64, device=dev)
        slowest = (None, 0.0)
        launches = 0
        if end > begin:
            specs = []
            for job, _ in todo:
                f = functors.DeviceFunctor(job[1], dimensions=job[2])
                specs.append(dict(functor=f, bounds=f.bounds, n_chains=end - begin,
                                  seeds=sharding.chain_keys(1234, begin, end),       # bench.py:171
                                  max_calls=self.max_calls, fglob=f.fglob, tol=DEFAULT_TOL))
            pool = BatchPool(specs, maxiter=MAX_IT, device=self.device)
            t_alloc = time.time() - t0
            pool.launch()
            t_enqueue = time.time() - t0 - t_alloc
            rows = []
            for n, r, ms in pool.results():
                rows.append(sharding.suite_records(n, begin, r.hit, r.ncall_hit, r.f_hit, ms))
                if ms > slowest[1]:
                    slowest = (todo[n][0][0], ms)
                logger.debug('rank %d: %s done at %.1f ms', self.rank, todo[n][0][0], ms)
            rec = torch.from_numpy(np.concatenate(rows, axis=0)).to(dev)
            launches = pool.launches
            self.stats.update(alloc_s=t_alloc, enqueue_s=t_enqueue, h2d_bytes=pool.h2d_bytes(),
                              d2h_bytes=pool.d2h_bytes())
        full = sharding.gather_ragged(rec)                            # rank 0: every (job, run) record
        self.stats.update(wall_s=time.time() - t0, jobs=len(todo), runs=self.nbruns, n_gpus=self.world,
                          slowest_job=slowest[0], slowest_job_ms=slowest[1], gpu_launches=launches)
        if full is None:
            return []
        per_job = sharding.assemble_suite_records(full.cpu().numpy(), len(todo), self.nbruns)
        for (job, bu), (ok, ncall_hit, f_hit, ms) in zip(todo, per_job):
            self.record(bu, ok, ncall_hit, f_hit, ms * 1e-3)
            if write:
                bu.write(self.folder)
            logger.info('Func: {0} - Algo: SDA -> success {1:.0f}% median ncall {2} (device time {3:.2f} s)'.format(
                job[0], 100 * bu.success.mean(), bu.medall, ms * 1e-3))
        self.stats['wall_s'] = time.time() - t0
        return [bu for _, bu in sorted(todo, key=lambda jb: jb[0][0])]

    def record(self, bu, ok, ncall_hit, f_hit, seconds):
        """bench.py:184-188 / :211-215 for every run of the job."""
        bu._values['success'][:] = ok
        bu._values['ncall'][:] = np.where(ok, ncall_hit, self.max_calls)
        bu._values['fvalue'][:] = np.where(ok, f_hit, np.inf)
        bu._values['ncall_max'][:] = np.where(ok, ncall_hit, self.max_calls)
        # the chains of a job run concurrently: the per-run time is the job's device time / nbruns
        bu._values['time'][:] = np.where(ok, seconds / max(self.nbruns, 1), np.inf)
