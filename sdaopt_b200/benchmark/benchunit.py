"""Per (function, algorithm) result record, format-compatible with the reference's
``sdaopt/benchmark/benchunit.py``: five arrays of length nbruns keyed by METRICS (:13-19), the
pickle file ``<name>-<algo>-<nbruns>.data`` (:152-161) and the summary statistics the report uses
(:66-149).  Pickles are written so that the reference's own ``benchstore.report`` unpickles them as
its ``BenchUnit`` (same module path, same attribute names)."""
import os
import pickle
import sys
import types

import numpy as np

METRICS = ['success', 'ncall', 'fvalue', 'time', 'ncall_max']
_REF_MODULE = 'sdaopt.benchmark.benchunit'


class BenchUnit(object):
    def __init__(self, nbruns, name, algo, max_fn_call=1e6):
        self._name = name
        self._algo = algo
        self._nbruns = nbruns
        fill = dict(success=False, ncall=max_fn_call, ncall_max=max_fn_call, fvalue=1e12,
                    time=np.nan)                                    # defaults of benchunit.py:29-37
        self._values = {k: np.full(nbruns, fill[k], dtype=np.float64) for k in METRICS}

    def __str__(self):
        return '{0}-{1}'.format(self._name, self._algo)

    def update(self, key, index, value):
        self._values[key][index] = value

    def values(self):
        return self._values

    name = property(lambda self: self._name)
    algo = property(lambda self: self._algo)
    success = property(lambda self: self._values['success'])

    def _stat(self, key, fn, digits, only_success=True):
        v = self._values[key]
        if only_success:
            v = v[np.where(self.success)]
        if not v.size:
            return np.inf
        return np.round(fn(v), digits)

    # statistics over the successful runs (benchunit.py:66-133)
    best = property(lambda self: self._stat('ncall', np.amin, 6))
    worst = property(lambda self: self._stat('ncall', np.amax, 6))
    mean = property(lambda self: self._stat('ncall', np.mean, 6))
    med = property(lambda self: self._stat('ncall', np.median, 6))
    std = property(lambda self: self._stat('ncall', np.std, 6))
    medall = property(lambda self: self._stat('ncall', np.median, 6, only_success=False))
    lowest = property(lambda self: self._stat('fvalue', np.min, 8))
    highest = property(lambda self: self._stat('fvalue', np.max, 2))
    time = property(lambda self: self._stat('time', np.mean, 6))

    @property
    def filename(self):
        return '{0}-{1}-{2}.data'.format(self._name, self._algo, self._nbruns)

    def write(self, folder=None):
        folder = folder or os.getcwd()
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, self.filename), 'wb') as f:
            f.write(dumps_as_reference(self))


def dumps_as_reference(bu):
    """Pickle `bu` under the reference's class path so that either package can load the file."""
    own_module, own_qual = BenchUnit.__module__, BenchUnit.__qualname__
    saved = {k: sys.modules.get(k) for k in ('sdaopt', 'sdaopt.benchmark', _REF_MODULE)}
    try:
        if saved[_REF_MODULE] is None or not hasattr(saved[_REF_MODULE], 'BenchUnit'):
            for k in saved:
                if saved[k] is None:
                    sys.modules[k] = types.ModuleType(k)
            sys.modules[_REF_MODULE].BenchUnit = BenchUnit
            BenchUnit.__module__ = _REF_MODULE
            return pickle.dumps(bu, protocol=2)
        # the reference is importable in this process: hand it an instance of its own class
        ref = saved[_REF_MODULE].BenchUnit.__new__(saved[_REF_MODULE].BenchUnit)
        ref.__dict__.update(bu.__dict__)
        return pickle.dumps(ref, protocol=2)
    finally:
        BenchUnit.__module__, BenchUnit.__qualname__ = own_module, own_qual
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == _REF_MODULE and name == 'BenchUnit':
            return BenchUnit
        return super().find_class(module, name)


def load(path):
    """Read a `.data` file written by this package or by the reference."""
    with open(path, 'rb') as f:
        return _Unpickler(f).load()
