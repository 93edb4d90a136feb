"""Drop-in for the reference's `bnn_priors.mcmc` (mcmc/__init__.py:1-3)."""
from .hmc import HMC
from .sgld import SGLD
from .verlet_sgld import VerletSGLD

__all__ = ["HMC", "SGLD", "VerletSGLD"]
