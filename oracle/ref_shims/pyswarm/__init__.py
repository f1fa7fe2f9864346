"""Empty stand-in so that the reference's benchmark/bench.py:25 imports (oracle tooling only)."""


def pso(*a, **k):
    raise NotImplementedError("pyswarm is not installed; stub for importing the reference")
