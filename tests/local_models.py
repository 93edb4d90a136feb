"""Tiny stand-ins for the reference's model / prior classes, for the GPU tests
(the GPU box has no reference checkout).  They follow the reference's
INTERFACES -- `Prior` modules holding the parameter as `.p` with constant
`loc` / `scale` / `df` buffers and a `_dist` (prior/base.py:17-76), models with
`log_prior()`, `log_likelihood()` and `split_potential_and_acc()`
(models/base.py:25-77,187-191) -- written from scratch.  Test infrastructure only.
"""
import math

import torch
import torch.distributions as td
import torch.nn as nn
import torch.nn.functional as F


class Prior(nn.Module):
    _dist = None

    def __init__(self, shape, **kwargs):
        super().__init__()
        self.kwargs_keys = list(kwargs)
        for k, v in kwargs.items():
            if isinstance(v, nn.Module):          # a hyper-prior: the value is v() (prior/base.py:11-14,37-38)
                self.add_module(k, v)
            elif isinstance(v, nn.Parameter):
                self.register_parameter(k, v)
            else:
                self.register_buffer(k, torch.as_tensor(v, dtype=torch.float32))
        with torch.no_grad():
            self.p = nn.Parameter(self._sample_value(torch.Size(shape)))

    def _sample_value(self, shape):
        return self._dist_obj().sample(shape)

    def _dist_obj(self):
        vals = {k: getattr(self, k) for k in self.kwargs_keys}
        return self._dist(**{k: (v() if isinstance(v, nn.Module) else v) for k, v in vals.items()})

    def log_prob(self):
        return self._dist_obj().log_prob(self.p).sum()

    def forward(self):
        return self.p


class Normal(Prior):
    _dist = td.Normal

    def __init__(self, shape, loc=0., scale=1.):
        super().__init__(shape, loc=loc, scale=scale)


class Laplace(Prior):
    _dist = td.Laplace

    def __init__(self, shape, loc=0., scale=1.):
        super().__init__(shape, loc=loc, scale=scale)


class StudentT(Prior):
    _dist = td.StudentT

    def __init__(self, shape, loc=0., scale=1., df=3.):
        super().__init__(shape, df=df, loc=loc, scale=scale)


class Cauchy(Prior):
    _dist = td.Cauchy

    def __init__(self, shape, loc=0., scale=1.):
        super().__init__(shape, loc=loc, scale=scale)


class GeneralizedNormal(td.Distribution):
    "density of prior/distributions.py:14-79 (log_prob only; sampling through a Laplace proxy)"
    arg_constraints = {}

    def __init__(self, loc, scale, beta):
        self.loc, self.scale, self.beta = loc, scale, beta
        super().__init__(validate_args=False)

    def sample(self, sample_shape=torch.Size()):
        return td.Laplace(self.loc, self.scale).sample(sample_shape)

    def log_prob(self, value):
        return (-torch.log(2 * self.scale) - torch.lgamma(1 / self.beta) + torch.log(self.beta)
                - torch.pow(torch.abs(value - self.loc) / self.scale, self.beta))


class GenNorm(Prior):
    _dist = GeneralizedNormal

    def __init__(self, shape, loc=0., scale=1., beta=0.5):
        super().__init__(shape, loc=loc, scale=scale, beta=beta)


class LogNormal(Prior):
    "prior/loc_scale.py:86-92: the parameter lives in log space, the layer sees exp(p)"
    _dist = td.Normal

    def __init__(self, shape, loc=0., scale=1.):
        super().__init__(shape, loc=loc, scale=scale)

    def forward(self):
        return self.p.exp()

    def log_prob(self):
        return super().log_prob() - self.p.sum()


class Uniform(Prior):
    "prior/transformed.py:12-47: a Gaussian variable pushed through its CDF; constant density of p"
    _dist = td.Uniform

    def __init__(self, shape, low, high):
        super().__init__(shape, low=low, high=high)
        with torch.no_grad():
            self.p.copy_(torch.randn(shape))

    def forward(self):
        return self.low + (self.high - self.low) * td.Normal(0., 1.).cdf(self.p)

    def log_prob(self):
        return -torch.log(self.high - self.low) * self.p.numel()


class DoubleGamma(Prior):
    "prior/transformed.py:83-96"

    def __init__(self, shape, loc=0., scale=1., concentration=1.5):
        super().__init__(shape, loc=loc, scale=scale, concentration=concentration)

    def _dist(self, loc, scale, concentration):
        return td.Gamma(concentration, 1 / scale)

    def _dist_obj(self):
        return self._dist(self.loc, self.scale, self.concentration)

    def log_prob(self):
        return (self._dist_obj().log_prob((self.p - self.loc).abs()) - math.log(2)).sum()


class Improper(Normal):
    "log_prob is overridden to zero (prior/loc_scale.py:94-97): fused as a prior without gradient"

    def log_prob(self):
        return 0. * self.p.sum()


class PositiveImproper(Improper):
    "prior/loc_scale.py:100-103: an improper prior on softplus(p)"

    def forward(self):
        return F.softplus(self.p)


def _inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Gamma(Prior):
    "prior/transformed.py:50-63: a Gamma density on softplus(p)"
    _dist = td.Gamma

    def __init__(self, shape, concentration, rate):
        super().__init__(shape, concentration=concentration, rate=rate)

    def _sample_value(self, shape):
        return _inv_softplus(super()._sample_value(shape))

    def forward(self):
        return F.softplus(self.p)

    def log_prob(self):
        return self._dist_obj().log_prob(self()).sum()


class HalfCauchy(Prior):
    "prior/transformed.py:66-80"
    _dist = td.HalfCauchy

    def __init__(self, shape, scale=1., multiplier=1.):
        super().__init__(shape, scale=scale)
        self.multiplier = multiplier

    def _sample_value(self, shape):
        return _inv_softplus(super()._sample_value(shape))

    def forward(self):
        return F.softplus(self.p) * self.multiplier

    def log_prob(self):
        return self._dist_obj().log_prob(self()).sum()


def hierarchical(base, hyper, **extra):
    """The pattern of prior/hierarchical.py and prior/empirical_bayes.py: `base` (Normal /
    Laplace / StudentT) whose scale is a scalar prior module, initialised so that the scale
    starts at the nominal value.  hyper in {"gamma", "uniform", "horseshoe", "empirical"}."""
    def make(shape, loc=0., scale=1.):
        if hyper == "gamma":
            sp = Gamma([], concentration=scale, rate=extra.get("rate", 1.))
            init = _inv_softplus(torch.tensor(float(scale)))
        elif hyper == "uniform":
            sp = Uniform([], 0., scale * 2.)
            init = torch.tensor(0.)
        elif hyper == "horseshoe":
            sp = HalfCauchy([], scale=extra.get("hyperscale", 1.), multiplier=scale)
            init = _inv_softplus(torch.tensor(1.))
        else:
            sp = PositiveImproper([], 0., 1.)
            init = _inv_softplus(torch.tensor(float(scale)))
        with torch.no_grad():
            sp.p.copy_(init)
        kw = {"df": extra["df"]} if "df" in extra else {}
        return base(shape, loc, sp, **kw)
    return make


class LearnedScaleNormal(Prior):
    "a prior the kernel must NOT fuse: the scale is a Parameter (prior/empirical_bayes.py:24-29)"
    _dist = td.Normal

    def __init__(self, shape, loc=0., scale=1.):
        super().__init__(shape, loc=loc, scale=nn.Parameter(torch.tensor(float(scale))))


class PriorLinear(nn.Module):
    "models/layers.py:5-18: a Linear layer whose weight and bias are Prior modules"

    def __init__(self, weight_prior, bias_prior):
        super().__init__()
        self.weight_prior, self.bias_prior = weight_prior, bias_prior

    def forward(self, x):
        return F.linear(x, self.weight_prior(), self.bias_prior())


class TinyClassifier(nn.Module):
    """din -> width -> width -> dout MLP with priors, the shape of the reference's
    ClassificationDenseNet (models/dense_nets.py:48-71) incl. std/sqrt(fan_in) scales."""

    def __init__(self, din, dout, width, prior_w=Normal, prior_b=Normal, w_kw=None, extra_bn=False):
        super().__init__()
        w_kw = w_kw or {}
        dims = [din, width, width, dout]
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers.append(PriorLinear(prior_w((b, a), 0., math.sqrt(2.) / math.sqrt(a), **w_kw),
                                      prior_b((b,), 0., 1.)))
            if b != dout:
                if extra_bn:
                    layers.append(nn.BatchNorm1d(b))      # parameters without a prior
                layers.append(nn.ReLU())
        self.net = nn.Sequential(*layers)

    def priors(self):
        return [m for m in self.modules() if isinstance(m, Prior)]

    def log_prior(self):
        return sum(m.log_prob() for m in self.priors())

    def forward(self, x):
        "p(y | x, params) (models/base.py:37-40,179-180)"
        return td.Categorical(logits=self.net(x))

    def log_likelihood_avg(self, x, y):
        return td.Categorical(logits=self.net(x)).log_prob(y).sum() / x.shape[0]

    def split_potential_and_acc(self, x, y, eff_num_data):
        loss = -self.log_likelihood_avg(x, y)
        log_prior = self.log_prior()
        potential_avg = loss - log_prior / eff_num_data
        return loss, log_prior, potential_avg


class TinyRegressor(TinyClassifier):
    "the shape of the reference's DenseNet (models/dense_nets.py:27-45): Normal(net(x), noise_std)"

    def __init__(self, din, dout, width, noise_std=1.0):
        super().__init__(din, dout, width)
        self.noise_std = noise_std

    def forward(self, x):
        return td.Normal(self.net(x), self.noise_std)


class GaussianTarget:
    """N data points from N(mean, std^2 I) in D dims with a flat prior: the posterior
    over the location is N(xbar, std^2/N) -- the target of the reference's
    distribution-preservation tests (testing/test_sgld.py:13-59)."""

    def __init__(self, N, D, mean, std, device, seed=0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.N, self.D, self.std = N, D, std
        self.x = (torch.randn(N, D, generator=g) * std + mean).to(device)
        self.xbar = self.x.mean(0)
        self.post_std = std / math.sqrt(N)

    def potential_avg(self, theta):
        "-(1/N) log p(x | theta) up to a constant"
        return (0.5 * ((self.x - theta) / self.std) ** 2).sum() / self.N
