"""Device-resident batches: PyTorch owns the HBM buffers and the stream, the C ABI does the work.

``DeviceBatch`` keeps a problem's inputs (x0, seeds), outputs and scratch resident on the GPU so
that ``launch()`` is exactly one persistent kernel on torch's current stream -- the form the
benchmark's ``value`` leg times.  ``run_from_host()`` is the end-to-end form: pinned host inputs are
copied in, the batch runs, and the per-chain records are copied back.
"""
import ctypes as C
import functools
import time

import numpy as np
import torch

from . import _lib
from . import functors as _functors
from ._sda import temperature_tables, _split_bounds, BatchResult


@functools.lru_cache(maxsize=16)
def cached_tables(maxiter, initial_temp, visit, temperature_restart):
    """K1 tables per (maxiter, T0, qv, T_restart): the suite asks for maxiter = 1e6 once per job."""
    return temperature_tables(maxiter, initial_temp, visit, temperature_restart)


def _align(n, a=256):
    return (int(n) + a - 1) // a * a


class BatchPool(object):
    """Many independent problems (suite jobs, benchmark shapes) in flight at once.

    ``specs`` is a list of dicts ``{functor, bounds, n_chains, seeds, max_calls, fglob, tol}``.  All
    device memory (inputs, per-chain records, scratch) is ONE allocation and all pinned host memory
    is ONE allocation, both made before the first launch: page-locked and device allocations are
    implicit synchronisation points, and made per job (as the first suite runner did) they serialise
    the launches -- that was 50 of its 58 seconds.  ``launch()`` stages every job's inputs with a
    single H2D copy and enqueues each job on its own stream in the given order (put the long jobs
    first), followed by an asynchronous copy of its records into pinned memory and an event;
    ``results()`` yields (index, BatchResult, seconds since launch) in launch order."""

    OUT8 = ("fun", "nit", "ncall", "ncall_ls", "n_ls", "ncall_hit", "f_hit")
    OUT4 = ("status", "hit")

    def __init__(self, specs, maxiter=1000, initial_temp=5230., visit=2.62, accept=-5.0, maxfun=1e7,
                 pure_sa=False, device=None, max_streams=512):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.specs = specs
        self._tables = cached_tables(int(maxiter), float(initial_temp), float(visit), 0.1)
        self.jobs = []
        in_off = out_off = ws_off = 0
        for sp in specs:
            functor = _functors.resolve(sp["functor"])
            lower, upper = _split_bounds(sp["bounds"])
            if not np.all(lower < upper):
                raise ValueError('Bounds are note consistent min < max')
            Cn, D, P = int(sp["n_chains"]), int(lower.size), int(functor.params.size)
            j = dict(functor=functor, lower=lower, upper=upper, C=Cn, D=D, P=P,
                     seeds=np.ascontiguousarray(sp["seeds"], dtype=np.uint64))
            if j["seeds"].size != Cn:
                raise ValueError("seeds must have one entry per chain")
            j["cfg"] = _lib.SdaConfig(int(maxiter), float(initial_temp), float(visit), float(accept),
                                      float(maxfun), int(bool(pure_sa)), 0.1, _lib.RNG_PHILOX,
                                      int(sp.get("max_calls", 0)), float(sp.get("fglob", float("nan"))),
                                      float(sp.get("tol", 1e-8)), 0, self._tables[0].ctypes.data,
                                      self._tables[1].ctypes.data, int(self._tables[0].size))
            # inputs: lower[D] upper[D] params[P] seeds[C]
            j["in_off"] = in_off
            j["in_parts"] = [("lower", 0, D), ("upper", _align(8 * D), D), ("params", 2 * _align(8 * D), P),
                             ("seeds", 2 * _align(8 * D) + _align(8 * P), Cn)]
            in_off += 2 * _align(8 * D) + _align(8 * P) + _align(8 * Cn)
            # records: x[C*D] then the 8-byte fields then the 4-byte fields, one contiguous block
            j["out_off"] = out_off
            parts, o = [("x", 0, Cn * D, 8)], _align(8 * Cn * D)
            for k in self.OUT8:
                parts.append((k, o, Cn, 8)); o += _align(8 * Cn)
            for k in self.OUT4:
                parts.append((k, o, Cn, 4)); o += _align(4 * Cn)
            j["out_parts"], j["out_bytes"] = parts, o
            out_off += o
            probe = _lib.SdaProblem(functor.functor_id, D, None, None, None, P)
            j["ws_bytes"] = int(_lib.lib(functor.library_path).sda_workspace_bytes(
                C.byref(probe), C.byref(j["cfg"]), Cn))
            j["ws_off"] = ws_off
            ws_off += _align(j["ws_bytes"])
            self.jobs.append(j)
        self.in_bytes, self.out_bytes, self.ws_bytes = in_off, out_off, ws_off
        self.dev = torch.empty(in_off + out_off + ws_off, dtype=torch.uint8, device=self.device)
        self.host = torch.empty(in_off + out_off, dtype=torch.uint8).pin_memory()
        hb = self.host.numpy()
        base = self.dev.data_ptr()
        for j in self.jobs:
            src = dict(lower=j["lower"], upper=j["upper"], params=j["functor"].params, seeds=j["seeds"])
            ptr = {}
            for name, off, n in j["in_parts"]:
                a = j["in_off"] + off
                hb[a:a + 8 * n] = np.ascontiguousarray(src[name]).view(np.uint8).ravel()
                ptr[name] = base + a
            j["problem"] = _lib.SdaProblem(j["functor"].functor_id, j["D"], ptr["lower"], ptr["upper"],
                                           ptr["params"] if j["P"] else None, j["P"])
            j["seeds_ptr"] = ptr["seeds"]
            ob = base + self.in_bytes + j["out_off"]
            fields = {name: ob + off for name, off, _, _ in j["out_parts"]}
            j["result"] = _lib.SdaResult(*[fields.get(k) for k in _lib.RESULT_FIELDS])
            j["ws_ptr"] = base + self.in_bytes + self.out_bytes + j["ws_off"]
        n_streams = max(1, min(int(max_streams), len(self.jobs)))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self.start = torch.cuda.Event(enable_timing=True)
        self.events = [torch.cuda.Event(enable_timing=True) for _ in self.jobs]
        self.launches = 0
        self.t_launch = None

    def h2d_bytes(self):
        return self.in_bytes

    def d2h_bytes(self):
        return self.out_bytes

    def launch(self):
        """Enqueue everything; returns at once (nothing here synchronises).

        Two passes.  First every kernel, then -- behind ALL kernels -- every job's copy of its records to
        pinned memory and its event, cheapest job first.  The streams of a process share a few hardware
        queues (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default, 32 at most; this package asks for 32), and a
        queue issues its commands in order: with `kernel A, copy A, kernel B` in one queue, `copy A`
        waits for kernel A (same stream) and kernel B -- another stream, no dependency -- waits behind
        it.  With the copies behind all kernels no kernel waits for another job, and the records of the
        cheap jobs reach the host early (suite: median job done at 22 s instead of 29 s; the total is the
        longest chain either way)."""
        main = torch.cuda.current_stream(self.device)
        self.t_launch = time.perf_counter()
        self.dev[:self.in_bytes].copy_(self.host[:self.in_bytes], non_blocking=True)
        self.start.record(main)
        for n, j in enumerate(self.jobs):
            st = self.streams[n % len(self.streams)]
            if n < len(self.streams):
                st.wait_event(self.start)
            L = _lib.lib(j["functor"].library_path)
            _lib.check(L.sda_batch_run(C.byref(j["problem"]), C.byref(j["cfg"]), j["C"], None,
                                       j["seeds_ptr"], None, 0, C.byref(j["result"]), j["ws_ptr"],
                                       j["ws_bytes"], C.c_void_p(st.cuda_stream)))
            self.launches += int(L.sda_last_launch_count())
        for n in reversed(range(len(self.jobs))):          # specs come longest first: copy back cheapest first
            j = self.jobs[n]
            st = self.streams[n % len(self.streams)]
            a = self.in_bytes + j["out_off"]
            with torch.cuda.stream(st):
                self.host[a:a + j["out_bytes"]].copy_(self.dev[a:a + j["out_bytes"]], non_blocking=True)
            self.events[n].record(st)

    def result(self, n):
        """Wait for job n; (BatchResult, milliseconds on the device clock since launch())."""
        self.events[n].synchronize()
        j = self.jobs[n]
        hb = self.host.numpy()
        a = self.in_bytes + j["out_off"]
        out = {}
        for name, off, cnt, size in j["out_parts"]:
            dt = np.float64 if name in ("x", "fun", "f_hit") else (np.int64 if size == 8 else np.int32)
            out[name] = hb[a + off:a + off + size * cnt].view(dt).copy()
        out["x"] = out["x"].reshape(j["C"], j["D"])
        return BatchResult(out), self.start.elapsed_time(self.events[n])

    def results(self):
        for n in reversed(range(len(self.jobs))):          # the order the copies were enqueued in
            r, ms = self.result(n)
            yield n, r, ms


def sda_batch_many(specs, device=None, **kwargs):
    """Run a list of independent problems concurrently (see :class:`BatchPool`); returns the list of
    BatchResult in the order of ``specs``."""
    pool = BatchPool(specs, device=device, **kwargs)
    pool.launch()
    out = [None] * len(specs)
    for n, r, _ in pool.results():
        out[n] = r
    return out


class DeviceBatch(object):
    def __init__(self, func, bounds, n_chains, maxiter=1000, initial_temp=5230., visit=2.62,
                 accept=-5.0, maxfun=1e7, pure_sa=False, max_calls=0, fglob=float("nan"),
                 tol=1e-8, use_x0=False, device=None):
        _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.functor = _functors.resolve(func)
        lower, upper = _split_bounds(bounds)
        if not np.all(lower < upper):
            raise ValueError('Bounds are note consistent min < max')
        self.dim = int(lower.size)
        self.n_chains = int(n_chains)
        f64, i64 = torch.float64, torch.int64
        dev = self.device
        self.lower = torch.from_numpy(lower).to(dev)
        self.upper = torch.from_numpy(upper).to(dev)
        self.params = torch.from_numpy(self.functor.params.copy()).to(dev)
        self._tables = cached_tables(int(maxiter), float(initial_temp), float(visit), 0.1)
        self.cfg = _lib.SdaConfig(int(maxiter), float(initial_temp), float(visit), float(accept),
                                  float(maxfun), int(bool(pure_sa)), 0.1, _lib.RNG_PHILOX,
                                  int(max_calls), float(fglob), float(tol), 0,
                                  self._tables[0].ctypes.data, self._tables[1].ctypes.data,
                                  int(self._tables[0].size))
        self.problem = _lib.SdaProblem(self.functor.functor_id, self.dim, self.lower.data_ptr(),
                                       self.upper.data_ptr(),
                                       self.params.data_ptr() if self.params.numel() else None,
                                       int(self.params.numel()))
        Cn, D = self.n_chains, self.dim
        self.seeds = torch.arange(Cn, dtype=i64, device=dev)
        self.x0 = torch.empty((Cn, D), dtype=f64, device=dev) if use_x0 else None
        self.out = dict(
            x=torch.empty((Cn, D), dtype=f64, device=dev), fun=torch.empty(Cn, dtype=f64, device=dev),
            nit=torch.empty(Cn, dtype=i64, device=dev), ncall=torch.empty(Cn, dtype=i64, device=dev),
            ncall_ls=torch.empty(Cn, dtype=i64, device=dev), n_ls=torch.empty(Cn, dtype=i64, device=dev),
            status=torch.empty(Cn, dtype=torch.int32, device=dev),
            hit=torch.empty(Cn, dtype=torch.int32, device=dev),
            ncall_hit=torch.empty(Cn, dtype=i64, device=dev), f_hit=torch.empty(Cn, dtype=f64, device=dev))
        self.result = _lib.SdaResult(*[self.out[k].data_ptr() if k in self.out else None
                                       for k in _lib.RESULT_FIELDS])
        self._L = _lib.lib(self.functor.library_path)
        nbytes = self._L.sda_workspace_bytes(C.byref(self.problem), C.byref(self.cfg), Cn)
        self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        # pinned host mirrors for the end-to-end form
        self.h_seeds = torch.empty(Cn, dtype=i64).pin_memory()
        self.h_x0 = torch.empty((Cn, D), dtype=f64).pin_memory() if use_x0 else None
        self.h_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in self.out.items()}

    def set_seeds(self, seeds):
        self.seeds.copy_(torch.as_tensor(np.asarray(seeds, dtype=np.int64)))

    def launch(self):
        """One persistent kernel on torch's current stream (asynchronous)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._L.sda_batch_run(
            C.byref(self.problem), C.byref(self.cfg), self.n_chains,
            self.x0.data_ptr() if self.x0 is not None else None, self.seeds.data_ptr(), None, 0,
            C.byref(self.result), self.workspace.data_ptr(), self.workspace.numel(),
            C.c_void_p(stream))
        _lib.check(rc)

    def h2d_bytes(self):
        return self.h_seeds.numel() * 8 + (self.h_x0.numel() * 8 if self.h_x0 is not None else 0)

    def d2h_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.h_out.values())

    def run_from_host(self):
        """pinned host inputs -> device, run, per-chain records -> pinned host.  Synchronous."""
        self.seeds.copy_(self.h_seeds, non_blocking=True)
        if self.h_x0 is not None:
            self.x0.copy_(self.h_x0, non_blocking=True)
        self.launch()
        for k, v in self.out.items():
            self.h_out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return BatchResult({k: v.numpy() for k, v in self.h_out.items()})

    def fetch(self):
        torch.cuda.current_stream(self.device).synchronize()
        return BatchResult({k: v.cpu().numpy() for k, v in self.out.items()})
