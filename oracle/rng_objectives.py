"""CPU restatement of the two catalogue objectives that draw random numbers inside ``fun()``.
TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

Reference: ``Stochastic.fun`` (sdaopt/benchmark/go_benchmark_functions/go_funcs_S.py:1274-1280)
returns ``sum(rnd * abs(x - 1/i))`` with ``rnd = uniform(0, 1, N)`` and ``XinSheYang01.fun``
(go_funcs_X.py:47-51) returns ``sum(np.random.random(N) * abs(x) ** i)``; both draw from NumPy's
GLOBAL generator at every call, so a value at a point cannot be pinned by a golden vector.  The CUDA
path (sdaopt_b200/csrc/sda_objectives_cat.cuh) replaces the draw by Philox4x32-10 noise addressed by
the point and the call number: h = hash64(x) + salt * 0x9E3779B97F4A7C15, counter (h lo, h hi, i, 0),
key ("SDA1", per-class tag), the first two output words turned into numpy's 53-bit
``random_sample()`` double.  Inside a chain ``salt`` is the chain's evaluation counter, so every call
draws fresh factors as in the reference; ``sda_eval`` uses salt = 0.  This file restates exactly that,
in integers, so the device values can be compared term by term; the *distribution* of the factors is
checked against the reference's own ``uniform`` law in tests/test_gpu_parity.py, and the in-chain
behaviour by the suite statistics against the live reference (tests/golden/suite_stats_ref_full.npz).
"""
import numpy as np

from . import sda_oracle as orc

TAGS = {"Stochastic": 0x53544F43, "XinSheYang01": 0x58535931}
_M64 = (1 << 64) - 1


def _mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def point_hash(x, salt=0):
    bits = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
    h = (salt * 0x9E3779B97F4A7C15) & _M64
    for k, b in enumerate(bits):
        h = (h + _mix64((int(b) + (k + 1) * 0x9E3779B97F4A7C15) & _M64)) & _M64
    return h


def noise(x, name, salt=0):
    """eps[N] in [0, 1): the factors the device functor multiplies the terms of f(x) with."""
    h = point_hash(x, salt)
    eps = np.empty(len(x))
    for k in range(len(x)):
        o = orc.philox4x32_10((h & 0xFFFFFFFF, h >> 32, k, 0), (0x53444131, TAGS[name]))
        eps[k] = ((int(o[0]) >> 5) * 67108864.0 + (int(o[1]) >> 6)) / 9007199254740992.0
    return eps


def fun(name, x, salt=0):
    x = np.asarray(x, dtype=np.float64)
    i = np.arange(1, len(x) + 1)
    eps = noise(x, name, salt)
    if name == "Stochastic":
        return float(np.sum(eps * np.abs(x - 1.0 / i)))
    return float(np.sum(eps * np.abs(x) ** i.astype(np.float64)))
