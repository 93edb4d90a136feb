"""Hierarchical priors fused into the sampler (second half): Normal /
Laplace / StudentT whose scale is a sampled scalar with a Gamma / Uniform / HalfCauchy /
improper hyper-prior (prior/hierarchical.py, prior/empirical_bayes.py).

(a) tests/golden/hier_priors.npz holds, from the unmodified reference classes, the total
    log density and its autograd gradients w.r.t. the weights and the hyper-parameter;
    the pre-pass + step of the kernel must reproduce them;
(b) a model with such priors in autograd vs the same model fused: same trajectory;
(c) bookkeeping: freshness of the pre-pass, rejection (rollback), unfuse."""
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

Z = np.load(os.path.join(GOLDEN_DIR, "hier_priors.npz"))
META = json.loads(bytes(Z["meta"]).decode())


@pytest.mark.parametrize("case", META, ids=[m["name"] for m in META])
def test_kernel_matches_reference_log_prob_and_gradients(case):
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200 import mcmc
    p_ref, g_ref = Z[case["name"] + "_p"], Z[case["name"] + "_grad_p"]
    # neighbours on both sides; the weights span more than one chunk boundary pattern (513 elements)
    params = [torch.nn.Parameter(torch.randn(37, device=DEV)), torch.nn.Parameter(torch.tensor(p_ref, device=DEV)),
              torch.nn.Parameter(torch.tensor(case["u"], device=DEV)), torch.nn.Parameter(torch.randn(4100, device=DEV))]
    opt = mcmc.SGLD(params, lr=1.0, num_data=1.0, momentum=0.0, temperature=0.0)
    (fg,) = opt.flat_groups
    fg.set_prior(1, case["kind"], case["loc"], 123.0, case["df"])        # the scale in the table is stale on purpose
    fg.set_hyper_link(1, 2, case["hyper_kind"], case["hyper_a"], case["hyper_b"])
    fg.prior_fused = True
    for p in params:
        p.grad = torch.zeros_like(p)
    fg.sync_views(True)
    assert not fg.hyper_fresh()
    fg.hyper_prepass(1.0)
    assert fg.hyper_fresh() and fg.log_prior_fresh()
    st = fg.fetch()
    # log density: weights given s(u), and the scale's own density (both summed by model.log_prior)
    assert math.isclose(st[1, N.S_LOG_PRIOR], case["log_prob_weights"], rel_tol=5e-6, abs_tol=1e-4)
    assert math.isclose(st[2, N.S_LOG_PRIOR], case["log_prob_hyper"], rel_tol=5e-6, abs_tol=1e-5)
    assert st[0, N.S_LOG_PRIOR] == 0.0 and st[3, N.S_LOG_PRIOR] == 0.0
    # -(1/N) dlog p/du with N = 1
    assert math.isclose(-st[2, N.S_HYPER], case["grad_u"], rel_tol=2e-5, abs_tol=1e-4), (st[2, N.S_HYPER], case["grad_u"])
    # the scale the step will use is now in the device copy of the segment table
    table = np.frombuffer(fg.table_dev.cpu().numpy().tobytes(), dtype=N.SEGMENT_DTYPE)
    assert math.isclose(float(table["prior_scale"][1]), case["scale"], rel_tol=2e-6)
    # gradient through a step: lr = N = 1, no momentum, no noise, zero likelihood gradient => x' = x + dlogp/dx
    post = case["kind"] in (N.PRIOR_NORMAL, N.PRIOR_LAPLACE)          # the step refreshes the hyper state itself
    assert fg.hyper_post_ok == post
    launches = fg.launches
    opt.step(calc_metrics=False)
    # the pre-pass was fresh: one launch (the epilogue of a BNNP_F_HYPER_POST step stays pending)
    assert fg.launches == launches + 1
    assert fg.hyper_fresh() == post
    moved = params[1].detach().cpu().numpy().astype(np.float64) - p_ref.astype(np.float64)
    tol = 2e-5 * np.abs(g_ref) + 2.5e-7 * np.maximum(np.abs(p_ref), np.abs(p_ref + g_ref)) + 1e-30
    assert np.all(np.abs(moved - g_ref) <= tol), float(np.max(np.abs(moved - g_ref) / tol))
    du = float(params[2].detach().double()) - float(np.float32(case["u"]))
    assert abs(du - case["grad_u"]) <= 2e-5 * abs(case["grad_u"]) + 2.5e-7 * max(abs(case["u"]), abs(case["u"] + case["grad_u"]))
    assert float(fg.fetch()[1, N.S_NONFINITE]) == 0.0
    if post:
        # what the step's epilogue left for the NEW parameters == what a pre-pass computes for them
        after = fg.fetch().copy()
        scale_after = float(np.frombuffer(fg.table_dev.cpu().numpy().tobytes(), dtype=N.SEGMENT_DTYPE)["prior_scale"][1])
        fg.hyper_prepass(1.0)
        again = fg.fetch()
        scale_again = float(np.frombuffer(fg.table_dev.cpu().numpy().tobytes(), dtype=N.SEGMENT_DTYPE)["prior_scale"][1])
        u_now = float(params[2].detach())
        if math.isfinite(u_now) and np.isfinite(again[:3, N.S_LOG_PRIOR]).all():
            assert scale_after == scale_again
            for row, col in ((1, N.S_LOG_PRIOR), (2, N.S_LOG_PRIOR), (2, N.S_HYPER), (1, N.S_HYPER)):
                assert math.isclose(after[row, col], again[row, col], rel_tol=2e-5, abs_tol=1e-6), (row, col, after[row, col], again[row, col])
    # a second step: pre-pass + epilogue + step where the scale statistic needs the new scale (StudentT),
    # the step alone otherwise
    launches = fg.launches
    opt.step(calc_metrics=False)
    assert fg.launches == launches + (1 if post else 3)
    if post:
        # ... and a third one carries the second one's pending epilogue (BNNP_F_HYPER_CHAIN): still one launch
        opt.step(calc_metrics=False)
        assert fg.launches == launches + 2


CASES = [(LM.Normal, "gamma", {}), (LM.Normal, "uniform", {}), (LM.Normal, "horseshoe", dict(hyperscale=2.0)),
         (LM.Laplace, "gamma", dict(rate=0.7)), (LM.Laplace, "empirical", {}), (LM.StudentT, "uniform", dict(df=5.)),
         (LM.StudentT, "gamma", dict(df=2.)), (LM.Normal, "empirical", {})]


@pytest.mark.parametrize("base,hyper,extra", CASES, ids=[f"{b.__name__}-{h}" for b, h, _ in CASES])
@pytest.mark.parametrize("sampler", ["VerletSGLD", "SGLD"])
def test_fused_hierarchical_prior_follows_autograd(base, hyper, extra, sampler):
    """Prior in autograd (what the reference does) vs prior fused, same noise: same losses,
    same log_prior, same parameters AND hyper-parameters after every step."""
    from bnn_priors_b200 import mcmc
    from bnn_priors_b200.prior_fusion import fuse_prior

    def build():
        torch.manual_seed(11)
        model = LM.TinyClassifier(12, 3, 8, prior_w=LM.hierarchical(base, hyper, **extra)).to(DEV)
        opt = getattr(mcmc, sampler)(list(model.parameters()), lr=2e-3, num_data=64.0, momentum=0.9,
                                     temperature=1.0, seed=3)
        return model, opt

    ma, oa = build()
    mb, ob = build()
    mb.load_state_dict(ma.state_dict())
    fp = fuse_prior(mb, ob, grad_max=1e6)
    # 3 weight priors + their 3 scale priors + 3 bias priors, nothing left in autograd
    assert len(fp.fused_modules) == 9 and not fp.other_modules
    (fg,) = ob.flat_groups
    assert len(fg.hyper_links) == 3
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
            # the rest of the reference runner's _model_potential_and_grad (inference.py:219-220): every
            # parameter must HAVE a gradient here, also a fused hyper-parameter that autograd no longer reaches
            for p in opt.param_groups[0]["params"]:
                p.grad.clamp_(min=-1e6, max=1e6)
            vals.append((float(loss), float(log_prior)))
        assert vals[0][0] == pytest.approx(vals[1][0], rel=2e-5, abs=1e-6)
        assert vals[0][1] == pytest.approx(vals[1][1], rel=5e-6, abs=1e-4), (it, vals)
        noise(oa, gen_a); noise(ob, gen_b)
        if sampler == "VerletSGLD" and it == 0:
            oa.initial_step(save_state=True, calc_metrics=True); ob.initial_step(save_state=True, calc_metrics=True)
        else:
            oa.step(calc_metrics=(it % 2 == 0)); ob.step(calc_metrics=(it % 2 == 0))
        for (na, pa), (nb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6), (it, na)
        if it % 2 == 0:         # diagnostics of the hyper-parameters too (dot products with the fused gradient)
            for pa, pb in zip(ma.parameters(), mb.parameters()):
                assert oa.state[pa]["est_config_temp"] == pytest.approx(ob.state[pb]["est_config_temp"], rel=2e-4, abs=1e-5)
    if sampler == "VerletSGLD":
        # the M-H bookkeeping sees the same gradient (point energy uses dot(grad, grad) incl. the prior part)
        for model, opt in ((ma, oa), (mb, ob)):
            opt.zero_grad()
            _, _, potential = model.split_potential_and_acc(x, y, 64.0)
            potential.backward()
        noise(oa, gen_a); noise(ob, gen_b)
        oa.final_step(); ob.final_step()
        da, db = oa.delta_energy(1.0, 1.1), ob.delta_energy(1.0, 1.1)
        assert da == pytest.approx(db, rel=1e-4, abs=1e-3)
        # rejection: both go back to the snapshot, and the fused log-prior follows
        torch.manual_seed(0); ra, _ = oa.maybe_reject(1e9)
        torch.manual_seed(0); rb, _ = ob.maybe_reject(1e9)
        assert ra and rb
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-5, atol=2e-6)
        assert float(ma.log_prior()) == pytest.approx(float(mb.log_prior()), rel=5e-6, abs=1e-4)
    fp.unfuse()
    assert not fg.has_hyper and not fg.prior_fused
    assert float(mb.log_prior()) == pytest.approx(float(ma.log_prior()), rel=5e-6, abs=1e-4)


@pytest.mark.parametrize("sampler", ["VerletSGLD", "SGLD", "HMC"])
def test_chained_epilogue_gives_the_bits_of_the_finalised_one(sampler):
    """Steps with sampled scales (Normal / Laplace) chain: the BNNP_F_HYPER_POST epilogue of a step rides on
    the next launch (BNNP_F_HYPER_CHAIN), whose CTAs derive the new scales and hyper gradients from the pending
    records themselves.  Same bits as applying every epilogue with bnnp_finalize before the next launch."""
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200 import mcmc

    def build():
        g = torch.Generator(device=DEV).manual_seed(7)
        shapes = [(300,), (), (5000,), (), (40, 130), (), (17,)]
        params = [torch.nn.Parameter(torch.randn(s, device=DEV, generator=g) * 0.3 if len(s) else
                                     torch.tensor(0.2, device=DEV)) for s in shapes]
        hp = dict(lr=2e-3, num_data=50.0) if sampler == "HMC" else dict(lr=2e-3, num_data=50.0, momentum=0.9, temperature=1.0)
        if sampler == "HMC":
            hp["raise_on_nan"] = False
        opt = getattr(mcmc, sampler)(params, **hp, seed=5)
        (fg,) = opt.flat_groups
        fg.set_prior(0, N.PRIOR_NORMAL, 0.0, 1.0, 3.0)
        fg.set_hyper_link(0, 1, N.PRIOR_HYPER_GAMMA, 1.5, 0.7)
        fg.set_prior(2, N.PRIOR_LAPLACE, 0.1, 1.0, 3.0)
        fg.set_hyper_link(2, 3, N.PRIOR_HYPER_IMPROPER, 0.0, 1.0)
        fg.set_prior(4, N.PRIOR_NORMAL, 0.0, 1.0, 3.0)
        fg.set_hyper_link(4, 5, N.PRIOR_HYPER_HALFCAUCHY, 1.0, 2.0)
        fg.set_prior(6, N.PRIOR_STUDENT_T, 0.0, 0.5, 4.0)         # constant hyper-parameters
        fg.prior_fused = True
        return opt, params, fg

    oa, pa, fa = build()
    ob, pb, fb = build()
    gg = torch.Generator(device=DEV).manual_seed(9)
    oa.sample_momentum(); ob.sample_momentum()
    chained = 0
    bufs = {id(o): [torch.empty_like(p) for p in ps] for o, ps in ((oa, pa), (ob, pb))}    # gradients that stay put
    for it in range(9):
        grads = [torch.randn(p.shape, device=DEV, generator=gg) * 0.05 for p in pa]
        for o, ps in ((oa, pa), (ob, pb)):
            o.zero_grad()
            for p, t, buf in zip(ps, grads, bufs[id(o)]):
                buf.copy_(t)
                p.grad = buf
        la = fa.launches
        if it == 0 and sampler != "SGLD":
            oa.initial_step(save_state=True, calc_metrics=False); ob.initial_step(save_state=True, calc_metrics=False)
        else:
            oa.step(calc_metrics=(it == 4)); ob.step(calc_metrics=(it == 4))
        if it >= 2:
            assert fa.launches == la + 1              # one launch per step: nothing is finalised in between
            chained += 1
        fb.flush_pending()                            # chain b: every epilogue applied before the next launch
        if it == 5:
            oa.sample_momentum(keep=0.5); ob.sample_momentum(keep=0.5)    # a launch without a prior carries it too
            fb.flush_pending()
        for p, q in zip(pa, pb):
            assert torch.equal(p, q), (it, p.shape)
    assert chained >= 6
    sa, sb = fa.fetch().copy(), fb.fetch().copy()
    for col in (N.S_LOG_PRIOR, N.S_HYPER, N.S_SUM_GG, N.S_SQ_MEAN):
        assert (sa[:, col] == sb[:, col]).all(), col
    ta = np.frombuffer(fa.table_dev.cpu().numpy().tobytes(), dtype=N.SEGMENT_DTYPE)["prior_scale"]
    tb = np.frombuffer(fb.table_dev.cpu().numpy().tobytes(), dtype=N.SEGMENT_DTYPE)["prior_scale"]
    assert (ta == tb).all()


def test_lookalikes_are_left_to_autograd():
    from bnn_priors_b200 import prior_fusion as PF
    # a vector-valued scale prior, a learnable df, a weight prior that is not Normal/Laplace/StudentT
    m = LM.Normal((5, 4), 0., LM.Gamma([3], 1.0, 1.0))
    assert PF.describe_hier_prior(m) is None
    m = LM.StudentT((5,), 0., LM.Gamma([], 1.0, 1.0), df=LM.PositiveImproper([], 0., 1.))
    assert PF.describe_hier_prior(m) is None
    m = LM.Cauchy((5,), 0., LM.Gamma([], 1.0, 1.0))
    assert PF.describe_hier_prior(m) is None

    class Gamma(LM.Gamma):                       # same name, another density of the scale
        def log_prob(self):
            return -(self() ** 2).sum()
    m = LM.Normal((50,), 0., Gamma([], 1.0, 1.0))
    spec = PF.describe_hier_prior(m)
    assert spec is None or not PF.matches_hier_module(m, spec)


def test_launch_refuses_a_pending_prepass_epilogue():
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200 import mcmc
    params = [torch.nn.Parameter(torch.randn(100, device=DEV)), torch.nn.Parameter(torch.tensor(0.3, device=DEV))]
    opt = mcmc.SGLD(params, lr=1e-3, num_data=10.0, momentum=0.0, temperature=0.0)
    (fg,) = opt.flat_groups
    fg.set_prior(0, N.PRIOR_NORMAL, 0.0, 1.0, 3.0)
    fg.set_hyper_link(0, 1, N.PRIOR_HYPER_GAMMA, 1.0, 1.0)
    fg.prior_fused = True
    fg.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_P | N.F_LOG_PRIOR | N.F_HYPER, N.NOISE_NONE, cm=1.0, inv_num_data=0.1)
    with pytest.raises(N.BnnpError, match="bnnp_finalize first"):
        fg.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_G, N.NOISE_NONE, cm=1.0)
    fg.flush_pending()
    fg.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_G, N.NOISE_NONE, cm=1.0)
    with pytest.raises(ValueError, match="exactly one element"):
        fg.set_hyper_link(1, 0, N.PRIOR_HYPER_GAMMA, 1.0, 1.0)
