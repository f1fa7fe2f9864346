"""GPU suite runner: the SDA slice of the reference's benchmark harness (sdaopt/benchmark/),
i.e. what a benchmark *run* is (bench.py:155-216, :256-333) and the record / report formats
(benchunit.py, benchstore.py) -- SURVEY.md section 8, rows a8, d (config 5) and f-1."""
from .benchunit import BenchUnit, METRICS            # noqa: F401
from .benchstore import report                       # noqa: F401
from .bench import Benchmarker, build_job_list       # noqa: F401
