"""Every prior kind of the kernel against the REFERENCE's own numbers: the golden
vectors in tests/golden/priors.npz hold, for parameter values incl. the density's kink
and a far tail, `Prior.log_prob()` and its autograd gradient as computed by the
reference classes (prior/loc_scale.py, prior/transformed.py) -- see make_golden.py."""
import json
import math
import os

import numpy as np
import pytest
import torch

import local_models as LM
from replay import GOLDEN_DIR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

Z = np.load(os.path.join(GOLDEN_DIR, "priors.npz"))
META = json.loads(bytes(Z["meta"]).decode())


@pytest.mark.parametrize("case", META, ids=[m["name"] for m in META])
def test_kernel_prior_matches_reference_log_prob_and_gradient(case):
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200 import mcmc
    p_ref, g_ref = Z[case["name"] + "_p"], Z[case["name"] + "_grad"]
    # two more tensors around it, so that the segment is neither first nor last
    params = [torch.nn.Parameter(torch.randn(37, device=DEV)), torch.nn.Parameter(torch.tensor(p_ref, device=DEV)),
              torch.nn.Parameter(torch.randn(4100, device=DEV))]
    opt = mcmc.SGLD(params, lr=1.0, num_data=1.0, momentum=0.0, temperature=0.0)
    (fg,) = opt.flat_groups
    fg.set_prior(1, case["kind"], case["loc"], case["scale"], case.get("df", 3.0))
    fg.prior_fused = True
    for p in params:
        p.grad = torch.zeros_like(p)
    # log-prior of the current parameters: a read-only launch
    fg.sync_views(True)
    fg.reduce_now(1.0)
    lp = float(fg.fetch()[1, N.S_LOG_PRIOR])
    assert math.isclose(lp, case["log_prob"], rel_tol=5e-6, abs_tol=1e-5), (lp, case["log_prob"])
    assert float(fg.fetch()[0, N.S_LOG_PRIOR]) == 0.0 and float(fg.fetch()[2, N.S_LOG_PRIOR]) == 0.0
    # gradient: lr = N = 1, no momentum, no noise, zero likelihood gradient  =>  p' = p + dlogp/dp
    opt.step(calc_metrics=False)
    moved = params[1].detach().cpu().numpy().astype(np.float64) - p_ref.astype(np.float64)
    tol = 1e-5 * np.abs(g_ref) + 2.5e-7 * np.maximum(np.abs(p_ref), np.abs(p_ref + g_ref)) + 1e-30
    assert np.all(np.abs(moved - g_ref) <= tol), float(np.max(np.abs(moved - g_ref) / tol))
    # the neighbours (no prior, zero gradient) did not move
    assert float(fg.fetch()[1, N.S_NONFINITE]) == 0.0


@pytest.mark.parametrize("prior_w,w_kw", [(LM.Cauchy, None), (LM.GenNorm, dict(beta=1.5)), (LM.LogNormal, None),
                                          (LM.Uniform, "uniform"), (LM.Improper, None)])
def test_fuse_prior_recognises_and_follows_autograd(prior_w, w_kw):
    """A model with the prior in autograd vs the same model with the prior fused: same
    trajectory, same log_prior, for the N4 kinds that give stable dynamics."""
    from bnn_priors_b200 import mcmc
    from bnn_priors_b200.prior_fusion import fuse_prior

    def build():
        torch.manual_seed(7)
        if w_kw == "uniform":
            mk = lambda shape, loc, scale: LM.Uniform(shape, -1.5, 2.0)      # noqa: E731
        elif prior_w is LM.LogNormal:
            mk = lambda shape, loc, scale: LM.LogNormal(shape, -2.0, 0.3)   # noqa: E731
        else:
            mk = lambda shape, loc, scale: prior_w(shape, loc, scale, **(w_kw or {}))   # noqa: E731
        model = LM.TinyClassifier(12, 3, 8, prior_w=mk).to(DEV)
        opt = mcmc.VerletSGLD(list(model.parameters()), lr=5e-3, num_data=64.0, momentum=0.9, temperature=1.0, seed=3)
        return model, opt

    ma, oa = build()
    mb, ob = build()
    mb.load_state_dict(ma.state_dict())
    fp = fuse_prior(mb, ob, grad_max=1e6)
    assert len(fp.fused_modules) == 6 and not fp.other_modules
    x = torch.rand(64, 12, device=DEV)
    y = torch.randint(0, 3, (64,), device=DEV)
    gen_a, gen_b = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)

    def noise(opt, gen):
        opt.set_replay_noise([torch.randn(p.shape, generator=gen) for p in opt.param_groups[0]["params"]])

    noise(oa, gen_a); noise(ob, gen_b)
    oa.sample_momentum(); ob.sample_momentum()
    for it in range(8):
        vals = []
        for model, opt in ((ma, oa), (mb, ob)):
            opt.zero_grad()
            loss, log_prior, potential = model.split_potential_and_acc(x, y, 64.0)
            potential.backward()
            vals.append((float(loss), float(log_prior)))
        assert vals[0][0] == pytest.approx(vals[1][0], rel=2e-5, abs=1e-6)
        assert vals[0][1] == pytest.approx(vals[1][1], rel=5e-6, abs=1e-4), (it, vals)
        noise(oa, gen_a); noise(ob, gen_b)
        oa.step(calc_metrics=False); ob.step(calc_metrics=False)
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6), it


def test_describe_prior_for_all_kinds_and_lookalikes():
    from bnn_priors_b200.prior_fusion import describe_prior, matches_module
    assert describe_prior(LM.Cauchy((3,), 0.1, 2.0)) == (4, pytest.approx(0.1), 2.0, 3.0)
    assert describe_prior(LM.GenNorm((3,), 0.0, 1.5, beta=0.75)) == (5, 0.0, 1.5, 0.75)
    assert describe_prior(LM.LogNormal((3,), -1.0, 0.5)) == (6, -1.0, 0.5, 3.0)
    assert describe_prior(LM.Uniform((3,), -2.0, 3.0)) == (7, -2.0, 5.0, 3.0)
    assert describe_prior(LM.Improper((3,), 0., 1.)) == (8, 0.0, 1.0, 3.0)
    assert describe_prior(LM.DoubleGamma((3,), 0.2, 2.0, concentration=1.7)) == (9, pytest.approx(0.2), 2.0,
                                                                                 pytest.approx(1.7))
    assert describe_prior(LM.LearnedScaleNormal((3,), 0., 1.)) is None
    for m in (LM.Cauchy((50,), 0.1, 2.0), LM.GenNorm((50,), 0.0, 1.5, beta=0.75), LM.LogNormal((50,), -1.0, 0.5),
              LM.Uniform((50,), -2.0, 3.0), LM.Improper((50,), 0., 1.), LM.DoubleGamma((50,), 0.2, 2.0, concentration=1.7),
              LM.Normal((50,), 1.0, 0.3), LM.Laplace((50,), 1.0, 0.3), LM.StudentT((50,), 0.0, 0.3, df=4.0)):
        assert matches_module(m, describe_prior(m)), type(m).__name__

    class DoubleGamma(LM.Normal):           # same name as a known prior, another density
        def log_prob(self):
            return -(self.p ** 4).sum()
    fake = DoubleGamma((50,), 0.0, 1.0)
    fake.kwargs_keys = ["loc", "scale", "concentration"]
    fake.register_buffer("concentration", torch.tensor(1.5))
    spec = describe_prior(fake)
    assert spec is not None and not matches_module(fake, spec)      # fuse_prior would leave it to autograd
