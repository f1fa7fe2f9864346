"""CPU, world_size 2, gloo: the N>1 host logic -- chain sharding, rank-independent Philox keys and
the end-of-run gather of per-chain records (the path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdaopt_b200 import sharding

D = 5


def _fake_records(keys):
    """Deterministic stand-in for a rank's kernel output: a pure function of the chain key, as the
    real per-chain results are (counter-based RNG)."""
    k = torch.from_numpy(keys.astype(np.int64)).to(torch.float64)
    x = torch.stack([torch.sin(k * (i + 1)) for i in range(D)], dim=1)
    return sharding.pack_records(x, [k * 0.5, k.to(torch.int64) * 3, torch.full_like(k, 1000)])


def _worker(rank, world, port, n_total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(n_total, world, rank)
    keys = sharding.chain_keys(1234, b, e)
    rec = _fake_records(keys)
    full = sharding.gather_ragged(rec, dst=0)
    if rank == 0:
        torch.save(full, out_path)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def _suite_worker(rank, world, port, n_jobs, nbruns, out_path):
    """The suite's sharding as Benchmarker.run does it: every rank takes a slice of the runs of every
    job, makes its records, rank 0 assembles them per job."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.shard_range(nbruns, world, rank)
    rows = []
    for j in range(n_jobs):
        keys = sharding.chain_keys(1234, b, e).astype(np.float64)
        rows.append(sharding.suite_records(j, b, (keys + j) % 3 != 0, keys * 10 + j, keys * 0.25 - j, 100.0 * rank + j))
    rec = torch.from_numpy(np.concatenate(rows, axis=0)) if rows and e > b else torch.zeros((0, 6), dtype=torch.float64)
    full = sharding.gather_ragged(rec, dst=0)
    if rank == 0:
        torch.save(full, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_suite_run_sharding_assembles_every_job_run_pair(tmp_path):
    n_jobs, nbruns = 7, 25                 # ragged: 12 + 13 runs per job
    out = str(tmp_path / "suite.pt")
    mp.spawn(_suite_worker, args=(2, _free_port(), n_jobs, nbruns, out), nprocs=2, join=True)
    per_job = sharding.assemble_suite_records(torch.load(out).numpy(), n_jobs, nbruns)
    keys = sharding.chain_keys(1234, 0, nbruns).astype(np.float64)
    for j, (ok, ncall_hit, f_hit, ms) in enumerate(per_job):
        np.testing.assert_array_equal(ok, (keys + j) % 3 != 0)        # run i keeps seed 1234 + i on any world size
        np.testing.assert_array_equal(ncall_hit, keys * 10 + j)
        np.testing.assert_array_equal(f_hit, keys * 0.25 - j)
        assert ms == 100.0 + j                                         # max over the ranks
    import pytest
    with pytest.raises(ValueError):
        sharding.assemble_suite_records(torch.load(out).numpy()[1:], n_jobs, nbruns)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gather_equals_single_rank(tmp_path):
    n_total = 37                      # ragged: 18 + 19
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = torch.load(out)
    want = _fake_records(sharding.chain_keys(1234, 0, n_total))
    assert got.shape == (n_total, D + 3)
    torch.testing.assert_close(got, want, rtol=0, atol=0)


def test_shard_helpers():
    ranges = [sharding.shard_range(65536 * 3 + 1, 8, r) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == 65536 * 3 + 1
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(7))
    sizes = [e - b for b, e in ranges]
    assert max(sizes) - min(sizes) <= 1
    costs = np.array([1e6, 10, 1e6, 500, 20, 1e6, 1e6, 30, 40, 50], dtype=float)
    parts = [sharding.shard_by_cost(costs, 4, r) for r in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(10))
    loads = [costs[p].sum() for p in parts]
    assert max(loads) / min(loads) < 1.01          # the four 1e6 jobs land on four ranks
    k = sharding.chain_keys(7, 10, 14)
    assert k.dtype == np.uint64 and k.tolist() == [17, 18, 19, 20]
    assert sharding.gather_records(torch.ones(3, 2)) is not None     # world == 1: identity
