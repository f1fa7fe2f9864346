"""Sharding of independent chains over ranks and the end-of-run gather.

A chain (function, dimension, run index) never reads another chain's state
(sdaopt/benchmark/job.py:46-47), so the work list is dealt to the ranks with no data-path
collective; the only communication is one gather of fixed-size per-chain records to rank 0
(replaces the pickles of benchmark/benchunit.py:156-161).  Backend-agnostic: NCCL on GPUs, gloo in
the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD_FIELDS = ("fun", "ncall", "nit", "status", "hit", "ncall_hit", "f_hit")


def shard_range(n_total, world, rank):
    """Contiguous, balanced [begin, end) of the rank's chains (sizes differ by at most one)."""
    return n_total * rank // world, n_total * (rank + 1) // world


def shard_by_cost(costs, world, rank):
    """Indices for `rank` after dealing the items round-robin in decreasing expected cost
    (suite jobs: dimension and known-hard functions first), so every GPU gets ~equal cost."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    return np.sort(order[rank::world])


def chain_keys(base_seed, begin, end):
    """Philox keys of the chains [begin, end): key = base_seed + global chain index, so a chain's draws
    do not depend on the number of ranks (nor do its results, as long as every rank's batch gets the
    same lanes per chain: the library derives them from the dimension and the batch size, SDA_GROUP
    pins them)."""
    return (np.arange(begin, end, dtype=np.uint64) + np.uint64(base_seed))


def pack_records(x, cols):
    """[C, D + len(cols)] float64 tensor: x* followed by the scalar per-chain records."""
    parts = [x] + [c.to(torch.float64).reshape(-1, 1) for c in cols]
    return torch.cat(parts, dim=1).contiguous()


def gather_records(rec, dst=0):
    """Gather equally-shaped record tensors on `dst`; returns the stacked [world*C, W] tensor on
    dst and None elsewhere.  world == 1 (or no process group) is the identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    world, rank = dist.get_world_size(), dist.get_rank()
    bufs = [torch.empty_like(rec) for _ in range(world)] if rank == dst else None
    dist.gather(rec, bufs, dst=dst)
    return torch.cat(bufs, dim=0) if rank == dst else None


def gather_ragged(rec, dst=0):
    """Same for shards of different length (n_total not divisible by world): pads to the longest
    shard, gathers, and trims with the gathered lengths."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    world, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    lens = [int(v.item()) for v in lens]
    pad = torch.zeros((max(lens), rec.shape[1]), dtype=rec.dtype, device=rec.device)
    pad[:rec.shape[0]] = rec
    full = gather_records(pad, dst)
    if full is None:
        return None
    full = full.reshape(world, max(lens), rec.shape[1])
    return torch.cat([full[r, :lens[r]] for r in range(world)], dim=0)


# ---- the suite's work list: (job, run) pairs, the runs of every job split over the ranks -----------
SUITE_RECORD_COLS = ("job", "run", "hit", "ncall_hit", "f_hit", "device_ms")


def suite_records(job_index, run_begin, hit, ncall_hit, f_hit, device_ms):
    """[n_runs, 6] float64 rows of one job's runs [run_begin, run_begin + n) on this rank."""
    n = len(hit)
    k = np.arange(run_begin, run_begin + n, dtype=np.float64)
    return np.stack([np.full(n, float(job_index)), k, np.asarray(hit, dtype=np.float64),
                     np.asarray(ncall_hit, dtype=np.float64), np.asarray(f_hit, dtype=np.float64),
                     np.full(n, float(device_ms))], axis=1)


def assemble_suite_records(full, n_jobs, nbruns):
    """Gathered rows of all ranks -> per job (hit[nbruns], ncall_hit, f_hit, device_ms max over
    ranks), runs in index order.  Raises if a (job, run) pair is missing or duplicated."""
    full = np.asarray(full, dtype=np.float64)
    out = []
    for j in range(n_jobs):
        rows = full[full[:, 0] == j]
        rows = rows[np.argsort(rows[:, 1], kind="stable")]
        if rows.shape[0] != nbruns or not np.array_equal(rows[:, 1], np.arange(nbruns, dtype=np.float64)):
            raise ValueError("suite records of job %d are incomplete: %d of %d runs" % (j, rows.shape[0], nbruns))
        out.append((rows[:, 2] != 0, rows[:, 3], rows[:, 4], float(rows[:, 5].max())))
    return out
