"""Import the live Python reference read-only (build container only; /root/reference does not
exist on the GPU box).  Used by oracle/gen_golden.py and by nothing at test/bench run time."""
import os
import sys

import numpy as np

REF = os.environ.get("SDAOPT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "sdaopt"))


def load():
    """Returns the reference ``sdaopt`` package (with its benchmark sub-package importable)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF)
    sys.dont_write_bytecode = True
    shims = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
    for p in (shims, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "NAN"):
        np.NAN = np.nan          # benchunit.py:37 uses the alias removed in NumPy 2
    import sdaopt
    return sdaopt


class RecordingRandomState(np.random.RandomState):
    """RandomState that records every uniform handed out by random_sample()."""

    def __init__(self, seed=None):
        super().__init__(seed)
        self.tape = []

    def random_sample(self, size=None):
        v = super().random_sample(size)
        if size is None:
            self.tape.append(float(v))
        else:
            self.tape.extend(np.asarray(v).ravel().tolist())
        return v
