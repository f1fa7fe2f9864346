"""Empty stand-in so that the reference's benchmark/bench.py:26 imports (oracle tooling only)."""
