"""prior_fusion's recognition of prior modules against the reference's OWN classes
(build container only; the GPU box has no reference checkout): every hierarchical /
empirical-Bayes prior with a sampled scale is recognised and its closed forms reproduce
the class's log_prob and autograd gradients; priors with further sampled hyper-parameters
(df, beta) or other densities are left to autograd."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "bnn_priors")),
                                reason="no reference checkout here")


@pytest.fixture(scope="module")
def ref_prior():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(HERE, "golden", "_shims"))
    sys.path.insert(0, REFERENCE)
    try:
        from bnn_priors import prior
        yield prior
    finally:
        sys.path.remove(REFERENCE)
        sys.path.remove(os.path.join(HERE, "golden", "_shims"))


def test_reference_hierarchical_classes_are_recognised_and_verified(ref_prior):
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200 import prior_fusion as PF
    torch.manual_seed(0)
    cases = [("NormalGamma", dict(scale=0.8, rate=1.5), N.PRIOR_NORMAL, N.PRIOR_HYPER_GAMMA),
             ("NormalUniform", dict(scale=0.6), N.PRIOR_NORMAL, N.PRIOR_HYPER_UNIFORM),
             ("Horseshoe", dict(scale=0.5, hyperscale=2.0), N.PRIOR_NORMAL, N.PRIOR_HYPER_HALFCAUCHY),
             ("LaplaceGamma", dict(scale=1.3, rate=0.7), N.PRIOR_LAPLACE, N.PRIOR_HYPER_GAMMA),
             ("LaplaceUniform", dict(scale=0.4), N.PRIOR_LAPLACE, N.PRIOR_HYPER_UNIFORM),
             ("StudentTGamma", dict(scale=0.9, df=2), N.PRIOR_STUDENT_T, N.PRIOR_HYPER_GAMMA),
             ("StudentTUniform", dict(scale=0.7, df=5), N.PRIOR_STUDENT_T, N.PRIOR_HYPER_UNIFORM),
             ("NormalEmpirical", dict(scale=0.3), N.PRIOR_NORMAL, N.PRIOR_HYPER_IMPROPER),
             ("LaplaceEmpirical", dict(scale=1.1), N.PRIOR_LAPLACE, N.PRIOR_HYPER_IMPROPER)]
    for name, kw, kind, hkind in cases:
        m = getattr(ref_prior, name)(torch.Size([30, 7]), 0.1, **kw)
        spec = PF.describe_hier_prior(m)
        assert spec is not None, name
        assert spec[0] == kind and spec[4] == hkind and spec[3] is m.scale, name
        assert PF.matches_hier_module(m, spec), name
        assert PF.describe_prior(m) is None                      # not a constant-scale prior
        with torch.no_grad():
            m.scale.p.add_(0.7)
        assert PF.matches_hier_module(m, spec), name


def test_reference_priors_with_other_sampled_hyperparameters_stay_in_autograd(ref_prior):
    from bnn_priors_b200 import prior_fusion as PF
    for name, kw in (("StudentTEmpirical", dict(scale=0.5, df=3.)), ("GenNormEmpirical", dict(scale=0.5, beta=0.7)),
                     ("GenNormUniform", dict(scale=0.5, beta=1.0))):
        m = getattr(ref_prior, name)(torch.Size([10]), 0., **kw)
        assert PF.describe_hier_prior(m) is None and PF.describe_prior(m) is None, name
    # constant-scale classes are the business of describe_prior
    for name in ("Normal", "Laplace", "StudentT"):
        m = getattr(ref_prior, name)(torch.Size([10]), 0., 1.)
        assert PF.describe_hier_prior(m) is None and PF.describe_prior(m) is not None
    # the scalar priors themselves
    assert PF.describe_hyper(ref_prior.Gamma([], 0.8, 1.5))[0] == 10
    assert PF.describe_hyper(ref_prior.Gamma([3], 0.8, 1.5)) is None          # not a scalar
    assert PF.describe_hyper(ref_prior.Normal([], 0., 1.)) is None
