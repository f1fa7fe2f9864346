"""CPU: the suite's record and report formats -- BenchUnit arrays, pickle name,
results.csv columns -- and the job list of the GPU runner."""
import csv
import os
import pickle
import sys
import types

import numpy as np

from sdaopt_b200.benchmark import BenchUnit, METRICS, report, build_job_list
from sdaopt_b200.benchmark import benchunit as bmod


def _unit():
    bu = BenchUnit(5, 'Rastrigin_10', 'SDA')
    for i, (ok, n, f) in enumerate([(1, 4268, 1e-9), (1, 4719, 2e-9), (0, 1e6, np.inf),
                                    (1, 5420, 0.0), (1, 3676, 5e-9)]):
        bu.update('success', i, ok)
        bu.update('ncall', i, n)
        bu.update('fvalue', i, f)
        bu.update('time', i, 0.25)
        bu.update('ncall_max', i, n)
    return bu


def test_benchunit_defaults_and_statistics():
    fresh = BenchUnit(3, 'X', 'SDA')
    assert list(fresh.values().keys()) == METRICS == ['success', 'ncall', 'fvalue', 'time', 'ncall_max']
    assert (fresh.values()['ncall'] == 1e6).all() and (fresh.values()['fvalue'] == 1e12).all()
    assert not fresh.success.any() and fresh.best == np.inf
    bu = _unit()
    assert bu.filename == 'Rastrigin_10-SDA-5.data' and str(bu) == 'Rastrigin_10-SDA'
    assert bu.best == 3676 and bu.worst == 5420 and bu.mean == 4520.75
    assert bu.med == 4493.5 and bu.medall == 4719 and bu.lowest == 0.0
    assert bu.std == np.round(np.std([4268, 4719, 5420, 3676]), 6)


def test_pickle_is_loadable_as_the_reference_class(tmp_path):
    bu = _unit()
    bu.write(str(tmp_path))
    path = os.path.join(str(tmp_path), bu.filename)
    raw = open(path, 'rb').read()
    assert b'sdaopt.benchmark.benchunit' in raw and b'BenchUnit' in raw
    # a stand-in for the reference package: plain pickle.load must resolve the class there
    class RefBenchUnit(object):
        pass
    mods = {k: types.ModuleType(k) for k in ('sdaopt', 'sdaopt.benchmark', 'sdaopt.benchmark.benchunit')}
    mods['sdaopt.benchmark.benchunit'].BenchUnit = RefBenchUnit
    RefBenchUnit.__module__ = 'sdaopt.benchmark.benchunit'
    RefBenchUnit.__qualname__ = RefBenchUnit.__name__ = 'BenchUnit'
    sys.modules.update(mods)
    try:
        got = pickle.load(open(path, 'rb'))
    finally:
        for k in mods:
            sys.modules.pop(k, None)
    assert type(got) is RefBenchUnit
    assert got._name == 'Rastrigin_10' and got._nbruns == 5
    np.testing.assert_array_equal(got._values['ncall'], bu.values()['ncall'])
    # and our own loader reads it back as our class
    mine = bmod.load(path)
    assert isinstance(mine, BenchUnit) and mine.mean == bu.mean
    assert BenchUnit.__module__ == 'sdaopt_b200.benchmark.benchunit'


def test_report_csv_columns(tmp_path):
    _unit().write(str(tmp_path))
    out = os.path.join(str(tmp_path), 'results.csv')
    report(str(tmp_path), kind='csv', path=out)
    rows = list(csv.reader(open(out), delimiter=',', quotechar='|'))
    assert rows[0] == ['Function name', 'Algorithm', 'Success Rate', 'Best', 'Average', 'Worst', 'Std',
                       'fvalue (mean)', 'Mean time (ms)', 'Median All']       # benchstore.py:37-41
    assert rows[1][:3] == ['Rastrigin_10', 'SDA', '80.0%'] and rows[1][3] == '3676.0'
    assert float(rows[1][8]) == 250.0
    raw = report(str(tmp_path))
    assert list(raw['SDA'].keys()) == ['Rastrigin_10']


def test_job_list_follows_reference_rules():
    jobs = build_job_list()
    names = [j[0] for j in jobs]
    # n-D selection at D = 5, 10, 20 ... 100 followed by the default-dimension job (bench.py:103-118)
    assert names[0] == 'AMGM'
    assert names[1:13] == ['Ackley01_5', 'Ackley01_10', 'Ackley01_20', 'Ackley01_30', 'Ackley01_40',
                          'Ackley01_50', 'Ackley01_60', 'Ackley01_70', 'Ackley01_80', 'Ackley01_90',
                          'Ackley01_100', 'Ackley01']
    from sdaopt_b200 import functors
    # all 250 jobs of the reference's list (195 classes + 5 n-D functions x 11 dimensions)
    assert len(jobs) == 8 + len(functors._CATALOGUE_ALL) + 5 * 11 == 250
    assert {'Stochastic', 'XinSheYang01'} <= set(names)
    assert ('Trid', 'Trid', 6) in jobs and ('McCormick', 'McCormick', 2) in jobs
    assert ('Griewank', 'Griewank', 2) in jobs and not any(n.startswith('Griewank_') for n in names)
