"""Stand-alone FP64 peak probe for MEASURED_PEAKS.json (which holds HBM and bf16 figures but no FP64
one): the DFMA microbenchmark of libsdaopt_b200.so (sda_fp64_peak: 8 independent FMA chains per thread,
8 CTAs of 256 threads per SM, 65,536 trips, best of 3, CUDA events) with the clocks sampled while it runs.

    python profiles/fp64_peak_probe.py  ->  one JSON line: {"fp64_tflops": ..., "clocks": {...}, ...}
Nominal: 148 SMs x 64 FP64 lanes x 2 flop x 1.965 GHz = 37.2 TFLOP/s."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdaopt_b200 import _lib  # noqa: E402
from bench import ClockSampler  # noqa: E402

_lib.require_cuda()
sampler = ClockSampler(0)
sampler.start()
vals = []
for _ in range(20):                      # ~0.5 s under load so that nvidia-smi sees it
    t, ms = C.c_double(), C.c_double()
    _lib.check(_lib.lib().sda_fp64_peak(C.byref(t), C.byref(ms), 0))
    vals.append((t.value, ms.value))
clocks = sampler.stop()
best = max(v[0] for v in vals)
print(json.dumps({"fp64_tflops": best, "fp64_tflops_median": sorted(v[0] for v in vals)[len(vals) // 2],
                  "kernel_ms": min(v[1] for v in vals), "nominal_tflops": 148 * 64 * 2 * 1.965e-3,
                  "how": "sda_dfma_kernel: 8 FMA chains/thread, 8 x 256-thread CTAs per SM, 65,536 trips, "
                         "best of 60 launches, CUDA events", "clocks": clocks}))
