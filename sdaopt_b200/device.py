"""Device-resident batches: PyTorch owns the HBM buffers and the stream, the C ABI does the work.

``DeviceBatch`` keeps a problem's inputs (x0, seeds), outputs and scratch resident on the GPU so
that ``launch()`` is exactly one persistent kernel on torch's current stream -- the form the
benchmark's ``value`` leg times.  ``run_from_host()`` is the end-to-end form: pinned host inputs are
copied in, the batch runs, and the per-chain records are copied back.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import functors as _functors
from ._sda import temperature_tables, _split_bounds, BatchResult


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
        self._tables = temperature_tables(maxiter, initial_temp, visit, 0.1)
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
