"""``report(folder, kind, path)``: the CSV / raw summary of a results folder, with the columns of the
reference's ``sdaopt/benchmark/benchstore.py:37-59`` so that the two outputs can be diffed."""
import csv
import os

import numpy as np

from . import benchunit

HEADER = ['Function name', 'Algorithm', 'Success Rate', 'Best', 'Average', 'Worst', 'Std',
          'fvalue (mean)', 'Mean time (ms)', 'Median All']


def _row(bu):
    et = bu.time
    et = np.inf if et < 0 else et * 1000
    ok = bu.success
    rate = np.round(np.where(ok)[0].size * 100 / ok.size, 1)
    return [bu.name, bu.algo, "{0}%".format(rate), bu.best, bu.mean, bu.worst,
            '{0} {1}'.format('(+/-)', bu.std), bu.lowest, round(et, 6), bu.medall]


def report(folder, kind="raw", path=None):
    files = sorted(x for x in os.listdir(folder) if x.endswith('.data'))
    units = [benchunit.load(os.path.join(folder, f)) for f in files]
    if kind == 'csv':
        with open(path, 'w') as csvfile:
            w = csv.writer(csvfile, delimiter=',', quotechar='|', quoting=csv.QUOTE_MINIMAL)
            w.writerow(HEADER)
            for bu in units:
                w.writerow(_row(bu))
        return path
    res = {}
    for bu in units:
        res.setdefault(bu.algo, {})[bu.name] = bu.values()
    return res
