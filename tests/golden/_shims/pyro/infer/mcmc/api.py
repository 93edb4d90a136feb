class MCMC:
    def __init__(self, *a, **k):
        raise RuntimeError("pyro is not installed; this is an import shim")
