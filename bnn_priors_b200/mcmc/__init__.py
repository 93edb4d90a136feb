"""Drop-in for the reference's `bnn_priors.mcmc` package (its __init__ exports the same
three names): SG-MCMC samplers whose transitions run as one sm_100a kernel launch."""
from .sgld import SGLD
from .verlet_sgld import VerletSGLD
from .hmc import HMC

__all__ = ["HMC", "SGLD", "VerletSGLD"]
