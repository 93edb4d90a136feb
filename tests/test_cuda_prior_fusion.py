"""fuse_prior(): the kernel's closed-form prior gradient / log-prior against
torch.distributions + autograd on a model shaped like the reference's
(tests/local_models.py), driven the way the reference runners drive it
(inference.py:215-223, inference_reject.py:18-33)."""
import math

import pytest
import torch

import local_models as LM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(prior_w, w_kw=None, extra_bn=False, seed=0, sampler="VerletSGLD", **hp):
    from bnn_priors_b200 import mcmc
    torch.manual_seed(seed)
    model = LM.TinyClassifier(20, 4, 16, prior_w=prior_w, w_kw=w_kw, extra_bn=extra_bn).to(DEV)
    x = torch.rand(96, 20, device=DEV)
    y = torch.randint(0, 4, (96,), device=DEV)
    hp = dict(dict(lr=0.02, num_data=96.0, momentum=0.9, temperature=1.0), **hp)
    opt = getattr(mcmc, sampler)(list(model.parameters()), **hp, seed=seed)
    return model, opt, x, y


def _noise(opt, gen):
    z = [torch.randn(p.shape, generator=gen) for g in opt.param_groups for p in g["params"]]
    opt.set_replay_noise(z)


def _runner_step(model, opt, x, y, n, grad_max=1e6):
    "inference.py:215-220"
    opt.zero_grad()
    loss, log_prior, potential = model.split_potential_and_acc(x, y, n)
    potential.backward()
    for p in opt.param_groups[0]["params"]:
        p.grad.clamp_(min=-grad_max, max=grad_max)
    return float(loss), float(log_prior), float(potential)


@pytest.mark.parametrize("prior_w,w_kw", [(LM.Normal, None), (LM.Laplace, None), (LM.StudentT, dict(df=3.0)),
                                          (LM.StudentT, dict(df=7.5))])
def test_fused_prior_follows_the_autograd_prior(prior_w, w_kw):
    from bnn_priors_b200.prior_fusion import fuse_prior
    ma, oa, x, y = _setup(prior_w, w_kw, extra_bn=True)
    mb, ob, _, _ = _setup(prior_w, w_kw, extra_bn=True)
    mb.load_state_dict(ma.state_dict())
    fp = fuse_prior(mb, ob, grad_max=1e6)
    assert len(fp.fused_modules) == 6 and not fp.other_modules
    n = 96.0
    ga, gb = torch.Generator().manual_seed(1), torch.Generator().manual_seed(1)
    _noise(oa, ga); _noise(ob, gb)
    oa.sample_momentum(); ob.sample_momentum()
    for it in range(12):
        la = _runner_step(ma, oa, x, y, n)
        lb = _runner_step(mb, ob, x, y, n)
        # loss identical, log_prior from the kernel's reduction vs torch.distributions
        assert la[0] == pytest.approx(lb[0], rel=2e-5, abs=1e-6)
        assert la[1] == pytest.approx(lb[1], rel=5e-6, abs=1e-4), (it, la, lb)
        assert la[2] == pytest.approx(lb[2], rel=2e-5, abs=1e-6)
        _noise(oa, ga); _noise(ob, gb)
        name = "initial_step" if it == 0 else ("final_step" if it == 11 else "step")
        kw = dict(save_state=True) if it == 0 else {}
        getattr(oa, name)(calc_metrics=True, **kw)
        getattr(ob, name)(calc_metrics=True, **kw)
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6), it
            sa, sb = oa.state[pa], ob.state[pb]
            for k in ("est_temperature", "est_config_temp", "delta_energy", "prev_new_momentum_delta"):
                assert sa[k] == pytest.approx(sb[k], rel=1e-4, abs=1e-4), (it, k)
    da, db = oa.delta_energy(0.1, 0.2), ob.delta_energy(0.1, 0.2)
    assert da == pytest.approx(db, rel=1e-5, abs=1e-4)
    # p.grad of a fused tensor is the likelihood gradient only; BatchNorm tensors have no prior
    kinds = [int(k) for k in ob.flat_groups[0].table["prior_kind"]]
    assert kinds.count(0) == 4 and len(kinds) == 10
    fp.unfuse()
    assert "log_prior" not in mb.__dict__ and not ob.flat_groups[0].prior_fused


def test_log_prior_is_recomputed_when_parameters_change():
    from bnn_priors_b200.prior_fusion import fuse_prior
    model, opt, x, y = _setup(LM.StudentT, dict(df=4.0))
    fp = fuse_prior(model, opt)
    want = sum(float(m.log_prob()) for m in model.priors())
    assert float(model.log_prior()) == pytest.approx(want, rel=2e-6)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(1.5)                       # somebody else writes the parameters (he_initialize, load_samples)
    want2 = sum(float(m.log_prob()) for m in model.priors())
    assert want2 != pytest.approx(want, rel=1e-3)
    assert float(model.log_prior()) == pytest.approx(want2, rel=2e-6)
    # the value can be backpropagated like the reference's (inference_reject.py:20-22)
    opt.zero_grad()
    lp = model.log_prior()
    (lp / -96.0).backward()
    assert float(opt.flat_groups[0].G.abs().sum()) == 0.0
    # after a rejected proposal the cached value is dropped
    opt.sample_momentum()
    _runner_step(model, opt, x, y, 96.0)
    before = float(model.log_prior())
    opt.initial_step(save_state=True)
    assert float(model.log_prior()) != pytest.approx(before, rel=1e-9)
    _runner_step(model, opt, x, y, 96.0)
    opt.final_step()
    real = torch.rand
    torch.rand = lambda *a, **k: torch.tensor(0.5)
    try:
        rejected, _ = opt.maybe_reject(1e9)
    finally:
        torch.rand = real
    assert rejected
    assert float(model.log_prior()) == pytest.approx(before, rel=2e-6)
    fp.unfuse()


def test_unsupported_priors_stay_in_autograd():
    from bnn_priors_b200.prior_fusion import describe_prior, fuse_prior
    assert describe_prior(LM.Improper((3,), 0., 1.)) == (8, 0.0, 1.0, 3.0)      # log_prob == 0: no gradient, no value
    assert describe_prior(LM.LearnedScaleNormal((3,), 0., 1.)) is None
    assert describe_prior(LM.Normal((3,), 0.5, 2.0)) == (1, 0.5, 2.0, 3.0)
    assert describe_prior(LM.StudentT((3,), 0., 2.0, 5.0)) == (3, 0.0, 2.0, 5.0)
    from bnn_priors_b200 import mcmc
    torch.manual_seed(3)
    model = LM.TinyClassifier(10, 3, 8, prior_w=LM.Laplace, prior_b=LM.LearnedScaleNormal).to(DEV)
    opt = mcmc.SGLD(list(model.parameters()), lr=1e-2, num_data=50.0, momentum=0.9)
    fp = fuse_prior(model, opt)
    assert len(fp.fused_modules) == 3 and len(fp.other_modules) == 3
    want = sum(float(m.log_prob()) for m in model.priors())
    lp = model.log_prior()
    assert float(lp) == pytest.approx(want, rel=2e-6)
    opt.zero_grad()
    (lp / -50.0).backward()                     # the learned scales still get their gradient
    scales = [m.scale for m in fp.other_modules]
    assert all(s.grad is not None and float(s.grad.abs()) > 0 for s in scales)


def test_grad_clamp_is_applied_to_the_fused_sum():
    from bnn_priors_b200.prior_fusion import fuse_prior
    ma, oa, x, y = _setup(LM.Normal, sampler="SGLD", temperature=0.0)
    mb, ob, _, _ = _setup(LM.Normal, sampler="SGLD", temperature=0.0)
    mb.load_state_dict(ma.state_dict())
    gmax = 1e-3                                  # small enough to bite
    fuse_prior(mb, ob, grad_max=gmax)
    oa.sample_momentum(); ob.sample_momentum()
    for _ in range(3):
        _runner_step(ma, oa, x, y, 96.0, grad_max=gmax)
        # (the runner's own clamp acts on the likelihood part only; the exact
        # counterpart of the reference's clamp of the SUM is the one in the kernel)
        _runner_step(mb, ob, x, y, 96.0, grad_max=1e30)
        oa.step(); ob.step()
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6)


def test_reject_runner_cycle_with_exact_gradients():
    """the call sequence of VerletSGLDRunnerReject (inference_reject.py:18-33,
    57-59, 88-91, 119-127, 156) -- full-data gradients ACCUMULATED over minibatches into the
    flat G array, prior term backpropagated once -- with the prior in autograd (A) and fused
    (B): same potentials, same delta_energy, same accept/reject decisions, same parameters."""
    from bnn_priors_b200.prior_fusion import fuse_prior
    n = 96.0
    ma, oa, x, y = _setup(LM.StudentT, dict(df=3.0), extra_bn=True)
    mb, ob, _, _ = _setup(LM.StudentT, dict(df=3.0), extra_bn=True)
    mb.load_state_dict(ma.state_dict())
    fuse_prior(mb, ob, grad_max=1e6)
    batches = [(x[i:i + 32], y[i:i + 32]) for i in range(0, 96, 32)]

    def exact(model, opt):                      # inference_reject.py:18-33
        opt.zero_grad()
        log_prior = model.log_prior()
        log_norm_prior = log_prior / -n
        log_norm_prior.backward()
        loss = 0.
        for xb, yb in batches:
            this_loss = -torch.distributions.Categorical(logits=model.net(xb)).log_prob(yb).sum() / n
            this_loss.backward()                # accumulates into p.grad = views of the flat G
            loss = loss + this_loss
        return float(loss + log_norm_prior)

    ga, gb = torch.Generator().manual_seed(2), torch.Generator().manual_seed(2)
    ua, ub = exact(ma, oa), exact(mb, ob)
    assert ua == pytest.approx(ub, rel=2e-6)
    _noise(oa, ga); _noise(ob, gb)
    oa.sample_momentum(); ob.sample_momentum()
    _noise(oa, ga); _noise(ob, gb)
    oa.initial_step(calc_metrics=True, save_state=True); ob.initial_step(calc_metrics=True, save_state=True)
    decisions = []
    for epoch in range(4):
        for xb, yb in batches:                  # the minibatch steps of an epoch (:88-91)
            _runner_step(ma, oa, xb, yb, n); _runner_step(mb, ob, xb, yb, n)
            _noise(oa, ga); _noise(ob, gb)
            oa.step(calc_metrics=False); ob.step(calc_metrics=False)
        va, vb = exact(ma, oa), exact(mb, ob)   # (:119)
        assert va == pytest.approx(vb, rel=5e-6)
        _noise(oa, ga); _noise(ob, gb)
        oa.final_step(calc_metrics=True); ob.final_step(calc_metrics=True)
        da, db = oa.delta_energy(ua, va), ob.delta_energy(ub, vb)
        assert da == pytest.approx(db, rel=1e-4, abs=2e-3), (epoch, da, db)
        u = [0.9, 1e-4, 0.5, 0.2][epoch]
        real = torch.rand
        torch.rand = lambda *a, **k: torch.tensor(u)
        try:
            ra, rb = oa.maybe_reject(da), ob.maybe_reject(db)
        finally:
            torch.rand = real
        assert ra[0] == rb[0]
        decisions.append(ra[0])
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=5e-5, atol=5e-6), epoch
        # store_metrics (inference.py:262-279): every per-tensor scalar comes from one device read
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            for k in ("preconditioner", "est_temperature", "est_config_temp"):
                assert oa.state[pa][k] == pytest.approx(ob.state[pb][k], rel=1e-4, abs=1e-5), (epoch, k)
        if ra[0]:                               # a rejected proposal restores the old potential (:127-139)
            ua, ub = exact(ma, oa), exact(mb, ob)
        else:
            ua, ub = va, vb
        _noise(oa, ga); _noise(ob, gb)
        oa.initial_step(calc_metrics=False, save_state=True); ob.initial_step(calc_metrics=False, save_state=True)
    assert len(decisions) == 4
