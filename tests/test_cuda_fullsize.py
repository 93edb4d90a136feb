"""Size-independent properties at BASELINE's full size: the 25,124,842-parameter
`vwidth_resnet18 width=96` segment table (54 tensors, tests/golden/model_shapes.json).
The oracle would need minutes here, so parity is asserted through invariants:
determinism, snapshot/rollback identity, leapfrog reversibility, the SGD known
answer, fp64 dot products recomputed by torch, zero padding."""
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
TENSORS = json.load(open(os.path.join(HERE, "golden", "model_shapes.json")))["vwidth_resnet18_w96_cifar10_gaussian"]["tensors"]


def make(sampler, seed=0, **hp):
    from bnn_priors_b200 import mcmc
    g = torch.Generator(device=DEV).manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(tuple(t["shape"]), device=DEV, generator=g) * t["scale"]) for t in TENSORS]
    opt = getattr(mcmc, sampler)(params, **hp, seed=seed)
    (fg,) = opt.flat_groups
    assert fg.n_params == 25124842 and fg.nseg == 54
    for p, v in zip(params, fg.g_views):
        p.grad = v
    new_grads(fg, 1000 + seed)
    return opt, params, fg


def new_grads(fg, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    noise = torch.empty_like(fg.G).normal_(0.0, 1e-3, generator=g)
    for p, o, n in zip(fg.params, fg.off, fg.numel):          # keep the padding zero, like autograd does
        fg.G[o:o + n].copy_(noise[o:o + n])


def padding_is_zero(fg, flat):
    pad = torch.ones(fg.total, dtype=torch.bool, device=flat.device)
    for o, n in zip(fg.off, fg.numel):
        pad[o:o + n] = False
    return int(pad.sum()) > 0 and not bool(flat[pad].any())


HP = dict(lr=5e-4, num_data=50000.0, momentum=0.994, temperature=1.0)


def test_two_runs_are_bitwise_identical_and_padding_stays_zero():
    runs = []
    for _ in range(2):
        opt, params, fg = make("VerletSGLD", seed=11, **HP)
        opt.sample_momentum()
        opt.initial_step(save_state=True, calc_metrics=True)
        for i in range(3):
            new_grads(fg, 100 + i)
            opt.step(calc_metrics=(i == 1))
        new_grads(fg, 200)
        opt.final_step()
        de = opt.delta_energy(1.0, 1.0001)
        runs.append((fg.P.clone(), fg.M.clone(), fg.state_dev.clone(), de))
        assert padding_is_zero(fg, fg.P) and padding_is_zero(fg, fg.M) and padding_is_zero(fg, fg.prev_p)
        del opt, params, fg
    a, b = runs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[2], b[2]) and a[3] == b[3]                 # the reductions are order-deterministic


def test_snapshot_and_rollback_are_exact():
    opt, params, fg = make("VerletSGLD", seed=12, **HP)
    opt.sample_momentum()
    p0, g0, m0 = fg.P.clone(), fg.G.clone(), fg.M.clone()
    opt.initial_step(save_state=True, calc_metrics=False)
    assert torch.equal(fg.prev_p, p0) and torch.equal(fg.prev_g, g0) and torch.equal(fg.prev_m, m0)
    for i in range(2):
        new_grads(fg, 300 + i)
        opt.step(calc_metrics=False)
    assert not torch.equal(fg.P, p0)
    real = torch.rand
    torch.rand = lambda *a, **k: torch.tensor(0.5)
    try:
        rejected, log_acc = opt.maybe_reject(1e12)
    finally:
        torch.rand = real
    assert rejected and log_acc == -1e12
    assert torch.equal(fg.P, p0) and torch.equal(fg.G, g0) and torch.equal(fg.M, m0)
    assert all(p.data_ptr() == v.data_ptr() for p, v in zip(params, fg.p_views))
    assert opt.maybe_reject(-1e12) == (False, 1e12)
    for g in opt.param_groups:
        g["temperature"] = 0.0
    assert opt.maybe_reject(5.0) == (False, 0.)                     # verlet_sgld.py:55-56


def test_reductions_match_fp64_dots_at_full_size():
    opt, params, fg = make("VerletSGLD", seed=13, **HP)
    opt.sample_momentum()
    P, G, M = fg.P.double(), fg.G.double(), fg.M.double()
    opt.initial_step(save_state=False, calc_metrics=True)
    n = HP["num_data"]
    curv = n ** 2 * (HP["lr"] / n) / 8
    for p, o, k in zip(params, fg.off, fg.numel):
        st = opt.state[p]
        mm = float((M[o:o + k] ** 2).sum())
        pg = float((P[o:o + k] * G[o:o + k]).sum())
        gg = float((G[o:o + k] ** 2).sum())
        assert st["est_temperature"] == pytest.approx(mm / k, rel=2e-6)
        pg_abs = float((P[o:o + k] * G[o:o + k]).abs().sum())
        assert st["est_config_temp"] == pytest.approx(pg * n / k, rel=1e-5, abs=3e-7 * pg_abs * n / k)
        assert st["delta_energy"] == pytest.approx(-curv * gg, rel=2e-6)
        m_new = opt.state[p]["momentum_buffer"].double().reshape(-1)
        c_gm = -.5 * math.sqrt(HP["lr"] * n)
        gm = float((G[o:o + k] * m_new).sum())
        gm_abs = float((G[o:o + k] * m_new).abs().sum())
        assert st["prev_new_momentum_delta"] == pytest.approx(c_gm * gm, rel=1e-5, abs=3e-7 * abs(c_gm) * gm_abs)


def test_hmc_leapfrog_is_reversible_at_full_size():
    """testing/test_hmc.py:17-65 in fp32: initial, L steps, final; negate the
    momentum; the same again brings back the start.  Gradient of a fixed quadratic."""
    opt, params, fg = make("HMC", seed=14, lr=1e-4, num_data=100.0, raise_on_nan=False)

    def grad():
        torch.mul(fg.P, 0.25, out=fg.G)      # U = .125 |theta|^2 per data point

    opt.sample_momentum()
    p0, m0 = fg.P.clone(), fg.M.clone()

    def trajectory(L=6):
        grad(); opt.initial_step(save_state=False, calc_metrics=False)
        for _ in range(L):
            grad(); opt.step(calc_metrics=False)
        p_before_final = fg.P.clone()
        grad(); opt.final_step(calc_metrics=False)
        assert torch.equal(fg.P, p_before_final)                    # final_step does not move p (:42)

    trajectory()
    assert not torch.allclose(fg.P, p0, rtol=1e-3, atol=1e-5)
    fg.M.neg_()
    trajectory()
    assert torch.allclose(fg.P, p0, rtol=1e-4, atol=2e-6)
    assert torch.allclose(-fg.M, m0, rtol=1e-4, atol=2e-5)


def test_sgd_known_answer_at_full_size():
    "testing/test_sgld.py:61-80 on the 25M-parameter chain"
    opt, params, fg = make("SGLD", seed=15, lr=0.1, num_data=1.0, momentum=0.9, temperature=0.0)
    ref = [p.detach().clone().requires_grad_() for p in params]
    sgd = torch.optim.SGD(ref, lr=0.1, momentum=0.9)
    opt.sample_momentum()
    for i in range(3):
        new_grads(fg, 400 + i)
        for r, v in zip(ref, fg.g_views):
            r.grad = v.clone()
        opt.step(calc_metrics=False)
        sgd.step()
    for p, r in zip(params, ref):
        assert torch.allclose(p, r, rtol=1e-5, atol=1e-7)
    # the moving average of g^2 is kept through its mean (sgld.py:153-154,173)
    opt2, _, fg2 = make("SGLD", seed=15, lr=0.1, num_data=1.0, momentum=0.9, temperature=0.0)
    sq = [torch.ones(k, device=DEV, dtype=torch.float64) for k in fg2.numel]
    opt2.sample_momentum()
    for i in range(3):
        new_grads(fg2, 400 + i)
        for s, o, k in zip(sq, fg2.off, fg2.numel):
            s.mul_(0.99).addcmul_(fg2.G[o:o + k].double(), fg2.G[o:o + k].double(), value=0.01)
        opt2.step(calc_metrics=False)
    opt2.update_preconditioner()
    means = [float(s.mean()) + 1e-8 for s in sq]
    lo = min(means)
    for p, mval in zip(fg2.params, means):
        assert opt2.state[p]["preconditioner"] == pytest.approx((mval / lo) ** (-1 / 4), rel=1e-6)


def test_flat_index_beyond_int32_known_answer():
    """Maximum sizes: one tensor of 2^31 + 12,293 floats (8.6 GB; flat indices, Philox counters and
    chunk bases beyond int32) between two small ones.  Known answers: SGD equivalence (the reference's
    test_sgd_equivalence, testing/test_sgld.py:61-80: momentum 0, temperature 0 => p' = p - lr g) at
    the far end of the big tensor and on the tensor after it; dot(g, g) against torch in float64;
    the in-kernel noise at flat index > 2^31 against its specification."""
    free, _ = torch.cuda.mem_get_info()
    n_big = 2 ** 31 + 12293
    if free < 5 * 4 * n_big:
        pytest.skip("needs ~45 GB of free device memory")
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200 import mcmc
    from oracle import sgmcmc_oracle as O
    g = torch.Generator(device=DEV).manual_seed(1)
    params = [torch.nn.Parameter(torch.randn(100, device=DEV, generator=g)),
              torch.nn.Parameter(torch.empty(n_big, device=DEV).normal_(0, 1, generator=g)),
              torch.nn.Parameter(torch.randn(777, device=DEV, generator=g))]
    lr, n_data = 0.25, 4.0
    opt = mcmc.SGLD(params, lr=lr, num_data=n_data, momentum=0.0, temperature=0.0, seed=5)
    (fg,) = opt.flat_groups
    assert fg.off[2] > 2 ** 31 and fg.nchunks > 2 ** 19
    for p, v in zip(params, fg.g_views):
        p.grad = v
    tail = slice(n_big - 5000, n_big)
    fg.g_views[1][tail].normal_(0, 1, generator=g)
    fg.g_views[1][:5000].normal_(0, 1, generator=g)
    fg.g_views[2].normal_(0, 1, generator=g)
    before_tail, before_last = params[1].detach()[tail].clone(), params[2].detach().clone()
    g_tail, g_last = fg.g_views[1][tail].clone(), fg.g_views[2].clone()
    gg = float((fg.g_views[1][tail].double() ** 2).sum() + (fg.g_views[1][:5000].double() ** 2).sum())
    opt.step(calc_metrics=False)
    # p' = p + h * (-hn * g) with h = sqrt(lr/N), hn = sqrt(lr N)  =>  p - lr g
    assert torch.allclose(params[1].detach()[tail], before_tail - lr * g_tail, rtol=1e-6, atol=1e-7)
    assert torch.allclose(params[2].detach(), before_last - lr * g_last, rtol=1e-6, atol=1e-7)
    assert math.isclose(float(fg.fetch()[1, N.S_SUM_GG]), gg, rel_tol=1e-6)
    # noise beyond 2^31: sample_momentum writes eps * sqrt(T); compare the last quads with the specification
    opt.param_groups[0]["temperature"] = 1.0
    call = fg.call
    opt.sample_momentum()
    m = opt.state[params[1]]["momentum_buffer"].reshape(-1)
    q0 = (fg.off[1] + n_big - 64) // 4 * 4
    want = O.philox_normal_segment(fg.key, call, q0, 64)
    got = fg.M[q0:q0 + 64].cpu().numpy()
    valid = min(64, fg.off[1] + n_big - q0)
    d = np.abs(got[:valid] - want[:valid])        # fast-math Box-Muller: same bounds as test_cuda_ops.py
    assert d.max() < 2e-3 and d.mean() < 2e-5, (d.max(), d.mean())
    assert m.shape[0] == n_big
    del opt, params, fg
    torch.cuda.empty_cache()
