from torch.distributions import MultivariateNormal  # noqa: F401
