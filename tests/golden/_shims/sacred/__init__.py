"""Import-only stand-in: the reference's exp_utils imports sacred at module scope."""


class Experiment:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("sacred is not installed; this is an import shim")
