from . import api  # noqa: F401


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("pyro is not installed; this is an import shim")


class NUTS(_Unavailable):
    pass


class HMC(_Unavailable):
    pass
