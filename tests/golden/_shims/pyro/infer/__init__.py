from . import mcmc  # noqa: F401
