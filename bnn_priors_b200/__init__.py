"""B200-native SG-MCMC sampling engine behind the sampler API of ratschlab/bnn_priors.

    from bnn_priors_b200 import mcmc          # SGLD, VerletSGLD, HMC
    from bnn_priors_b200.prior_fusion import fuse_prior

The only implementation of the path is the sm_100a CUDA library built by
`python -m bnn_priors_b200.build`; importing the samplers without it (or using
them on CPU tensors) raises.
"""
__version__ = "0.1.0"
