"""Import-only stand-in: experiments/train_bnn.py:14-15 imports pyro's NUTS / HMC / MCMC at module
scope and uses them only for inference="HMC" (pyro's own sampler, not this path)."""
from . import infer  # noqa: F401
