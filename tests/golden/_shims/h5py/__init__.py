"""Import-only stand-in: the reference's exp_utils imports h5py at module scope."""


class File:  # pragma: no cover - never instantiated by the sampler path
    def __init__(self, *a, **k):
        raise RuntimeError("h5py is not installed; this is an import shim")
