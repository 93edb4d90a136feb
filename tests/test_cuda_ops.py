"""CUDA samplers vs the oracle on seeded ragged inputs, the in-kernel Philox stream
vs its specification, and the error behaviour of the reference's sampler API.
All calls go through bnn_priors_b200.mcmc -> libbnnp.so."""
import math

import numpy as np
import pytest
import torch

from oracle import sgmcmc_oracle as O
from replay import _rel

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
# every boundary of the layout: < 1 quad, quad tail, 128-byte padding, chunk tail, > 1 chunk
RAGGED = [(1,), (3,), (4,), (5,), (31,), (33,), (4095,), (4096,), (4097,), (3, 5, 7), (2, 4100), (12289,)]
TRAJ_TOL = 5e-6      # fp32 elementwise work, a handful of steps, relative to the tensor's RMS


def _mcmc():
    from bnn_priors_b200 import mcmc
    return mcmc


class Pair:
    """The same chain in the CUDA sampler and in the oracle, driven in lock step
    with identical gradients and replayed noise."""

    def __init__(self, kind, shapes, hp, seed=0, precond=True):
        self.rng = np.random.default_rng(seed)
        self.shapes, self.kind = shapes, kind
        p0 = [self.rng.standard_normal(s).astype(np.float32) for s in shapes]
        self.params = [torch.nn.Parameter(torch.tensor(a, device=DEV)) for a in p0]
        self.opt = getattr(_mcmc(), kind)(self.params, **hp, seed=seed)
        g = dict(hp)
        g.pop("raise_on_nan", None)
        if kind == "HMC":
            g.update(momentum=1.0, temperature=1.0)
        self.chain = O.Chain(p0, O.Group(**g), dot_dtype=np.float64)
        if precond:
            vals = self.rng.uniform(0.2, 1.3, len(shapes))
            for p, seg, v in zip(self.params, self.chain.segs, vals):
                self.opt.state[p]["preconditioner"] = float(v)
                seg.preconditioner = float(v)
        self.n_grad = 0

    def grads(self, scale=0.3):
        self.n_grad += 1
        for p, seg, s in zip(self.params, self.chain.segs, self.shapes):
            g = (self.rng.standard_normal(s) * scale).astype(np.float32)
            seg.g = g.reshape(-1).copy()
            t = torch.tensor(g, device=DEV)
            if p.grad is None or self.n_grad % 2 == 0:
                p.grad = t                 # foreign tensor -> adopted into the flat G
            else:
                p.grad.copy_(t)            # in place into the flat view

    def noise(self):
        z = [self.rng.standard_normal(s).astype(np.float32) for s in self.shapes]
        self.opt.set_replay_noise([torch.tensor(a) for a in z])
        return lambda i, n: z[i].reshape(-1)

    def sample_momentum(self, keep=0.0):
        nz = self.noise()
        self.opt.sample_momentum(keep=keep)
        O.sample_momentum(self.chain, nz, keep=keep)

    def step(self, name, **kw):
        nz = self.noise()
        getattr(self.opt, name)(**kw)
        phase = {"initial_step": O.PHASE_INITIAL, "step": O.PHASE_MID, "final_step": O.PHASE_FINAL}[name]
        cm = kw.get("calc_metrics", True)
        if self.kind == "SGLD":
            O.sgld_step(self.chain, nz, calc_metrics=cm, is_final=(name == "final_step"))
        elif self.kind == "VerletSGLD":
            O.verlet_step(self.chain, nz, phase=phase, calc_metrics=cm,
                          save_state=kw.get("save_state", name == "initial_step"))
        else:
            O.hmc_step(self.chain, phase=phase, calc_metrics=cm,
                       save_state=kw.get("save_state", name == "initial_step"))

    def check(self, scalars=("est_temperature", "est_config_temp"), tol=TRAJ_TOL):
        for p, seg in zip(self.params, self.chain.segs):
            assert _rel(p.detach().cpu().numpy().reshape(-1), seg.p) < tol
            m = self.opt.state[p].get("momentum_buffer")
            if m is not None and seg.m is not None:
                assert _rel(m.cpu().numpy().reshape(-1), seg.m) < tol
            st = self.opt.state[p]
            for k in scalars:
                want = getattr(seg, k)
                if want is None or (isinstance(want, float) and math.isnan(want)):
                    continue
                got = st[k]
                scale = 1e-3
                if k in ("delta_energy", "prev_new_momentum_delta") and seg.m is not None:
                    # running sums of signed terms c_gm * g.m: judge against one term's size
                    g64, m64 = seg.g.astype(np.float64), seg.m.astype(np.float64)
                    scale = abs(self.chain.group.derived.get("bhn", 1.0)) * seg.preconditioner * \
                        math.sqrt(float(g64 @ g64) * float(m64 @ m64)) + 1e-3
                assert abs(got - want) <= 2e-5 * max(abs(want), scale), (k, got, want, seg.p.size)


HP = dict(lr=2e-3, num_data=40.0, momentum=0.9, temperature=0.7)


def test_sgld_matches_oracle_on_ragged_segments():
    pr = Pair("SGLD", RAGGED, HP, seed=1)
    pr.sample_momentum()
    for i in range(5):
        pr.grads()
        pr.step("step", calc_metrics=(i % 2 == 0))
        pr.check()
    pr.opt.update_preconditioner()
    O.update_preconditioner(pr.chain)
    for p, seg in zip(pr.params, pr.chain.segs):
        assert abs(pr.opt.state[p]["preconditioner"] - seg.preconditioner) < 1e-6 * seg.preconditioner
    pr.sample_momentum(keep=0.3)
    pr.grads()
    pr.step("initial_step")
    pr.grads()
    pr.step("final_step")
    pr.check()
    assert pr.opt.delta_energy(0.0, 1.0) == math.inf        # sgld.py:54-55


def test_sgld_without_momentum_and_descent_phase():
    hp = dict(HP, momentum=0.0)
    pr = Pair("SGLD", RAGGED[:8], hp, seed=2)
    for i in range(3):
        pr.grads()
        pr.step("step", calc_metrics=True)
        pr.check()
    assert all("momentum_buffer" not in dict.keys(pr.opt.state[p]) for p in pr.params)   # sgld.py:132-134
    for g in pr.opt.param_groups:
        g["temperature"] = 0.0                                 # no noise is drawn: sgld.py:141
    pr.chain.group.temperature = 0.0
    pr.grads()
    pr.opt.step(calc_metrics=True)
    O.sgld_step(pr.chain, None, calc_metrics=True)
    pr.check()


@pytest.mark.parametrize("momentum", [0.9, 0.0])
def test_verlet_matches_oracle_on_ragged_segments(momentum):
    pr = Pair("VerletSGLD", RAGGED, dict(HP, momentum=momentum), seed=3)
    pr.sample_momentum()
    u0, u1 = 0.37, 0.52
    for cycle in range(3):
        pr.grads()
        pr.step("initial_step", save_state=True, calc_metrics=(cycle == 0))
        for i in range(3):
            pr.grads()
            pr.step("step", calc_metrics=(i == 1))
            pr.check(scalars=("est_temperature", "est_config_temp", "delta_energy", "prev_new_momentum_delta"))
            # the runner asks for delta_energy on metrics steps too (inference_reject.py:93-108)
            a, b = pr.opt.delta_energy(u0, u1), O.verlet_delta_energy(pr.chain, u0, u1)
            assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (a, b)
        pr.grads()
        pr.step("final_step")
        pr.check(scalars=("est_temperature", "est_config_temp", "delta_energy", "prev_new_momentum_delta"))
        a, b = pr.opt.delta_energy(u0, u1), O.verlet_delta_energy(pr.chain, u0, u1)
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (a, b)
        # force rejects and an accept through the energy difference handed in
        forced = 1e6 if cycle != 1 else -1e6
        real = torch.rand
        torch.rand = lambda *x, **k: torch.tensor(0.5)
        try:
            rej_a, la = pr.opt.maybe_reject(forced)
        finally:
            torch.rand = real
        rej_b, lb = O.maybe_reject(pr.chain, forced, 0.5)
        assert rej_a == rej_b == (cycle != 1) and la == lb
        pr.check(scalars=())
        for p, seg in zip(pr.params, pr.chain.segs):     # p.grad is restored too (verlet_sgld.py:65)
            assert _rel(p.grad.cpu().numpy().reshape(-1), seg.g) < 1e-7


def test_hmc_matches_oracle_on_ragged_segments():
    pr = Pair("HMC", RAGGED, dict(lr=1e-2, num_data=25.0), seed=4)
    for cycle in range(2):
        pr.sample_momentum()
        pr.grads()
        pr.step("initial_step", save_state=True, calc_metrics=True)
        pr.check(scalars=("est_temperature", "est_config_temp", "delta_energy"))
        for i in range(4):
            pr.grads()
            pr.step("step", calc_metrics=(i % 2 == 0))
            pr.check(scalars=("est_temperature", "est_config_temp", "delta_energy"))
        pr.grads()
        pr.step("final_step", calc_metrics=(cycle == 0))
        pr.check(scalars=("est_temperature", "est_config_temp", "delta_energy"))
        a, b = pr.opt.delta_energy(1.0, 0.9), O.hmc_delta_energy(pr.chain, 1.0, 0.9)
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (a, b)
    # delta_energy after the user flips the momentum in place (testing/test_hmc.py:49-50):
    for p in pr.params:
        pr.opt.state[p]["momentum_buffer"].neg_()
    a2 = pr.opt.delta_energy(1.0, 0.9)
    assert abs(a2 - a) <= 1e-9 * max(1.0, abs(a))


def test_philox_stream_matches_its_specification():
    """sample_momentum / a Verlet step with the production noise source reproduce
    oracle.philox_normal_segment(key, call, flat offset, numel) -- the counter layout
    (element quad, launch counter) and the Box-Muller transform are as specified."""
    from bnn_priors_b200 import _native as N
    shapes = [(5,), (4097,), (33,), (2, 3000)]
    params = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    opt = _mcmc().VerletSGLD(params, lr=1e-2, num_data=10.0, momentum=0.5, temperature=4.0, seed=77, chain=3)
    (fg,) = opt.flat_groups
    key = N.philox_key(77, (3 << 16) + 0)
    assert key == O.philox_key(77, (3 << 16) + 0) == tuple(fg.key)
    opt.sample_momentum()                       # launch counter 0: m = sqrt(T) * z
    worst, mean = 0.0, []
    for p, off in zip(params, fg.off):
        want = 2.0 * O.philox_normal_segment(key, 0, off, p.numel())
        got = opt.state[p]["momentum_buffer"].cpu().numpy().reshape(-1)
        d = np.abs(got - want)
        worst = max(worst, d.max())
        mean.append(d.mean())
    assert worst < 2e-3 and max(mean) < 4e-6, (worst, mean)
    assert float(fg.M.abs().sum()) == pytest.approx(
        sum(float(opt.state[p]["momentum_buffer"].abs().sum()) for p in params), rel=1e-6)   # padding stays zero
    # launch counter 1: a step with zero gradient and zero old momentum weight is pure noise
    for p in params:
        p.grad = torch.zeros_like(p)
        opt.state[p]["momentum_buffer"].zero_()
    opt.step(calc_metrics=False)
    ns = opt.param_groups[0]["noise_std"]
    for p, off in zip(params, fg.off):
        want = np.float32(ns) * O.philox_normal_segment(key, 1, off, p.numel())
        got = opt.state[p]["momentum_buffer"].cpu().numpy().reshape(-1)
        assert np.abs(got - want).max() < 2e-3 * ns and np.abs(got - want).mean() < 4e-6 * ns
    # another chain id gives another stream
    params2 = [torch.nn.Parameter(torch.zeros(s, device=DEV)) for s in shapes]
    opt2 = _mcmc().VerletSGLD(params2, lr=1e-2, num_data=10.0, momentum=0.5, temperature=4.0, seed=77, chain=4)
    opt2.sample_momentum()
    a = opt2.flat_groups[0].M
    opt.sample_momentum()
    assert abs(float(torch.corrcoef(torch.stack([a, fg.M]))[0, 1])) < 0.05


def test_philox_normals_are_standard_normal():
    import scipy.stats
    n = 1 << 22
    p = torch.nn.Parameter(torch.zeros(n, device=DEV))
    opt = _mcmc().SGLD([p], lr=1e-2, num_data=1.0, momentum=0.9, temperature=1.0, seed=5)
    opt.sample_momentum()
    z1 = opt.state[p]["momentum_buffer"].double().cpu().numpy().copy()
    opt.sample_momentum()
    z2 = opt.state[p]["momentum_buffer"].double().cpu().numpy()
    assert np.isfinite(z1).all()
    assert abs(z1.mean()) < 4 / math.sqrt(n) and abs(z1.var() - 1) < 6 * math.sqrt(2 / n)
    assert abs(scipy.stats.kurtosis(z1)) < 0.02 and abs(scipy.stats.skew(z1)) < 0.01
    assert scipy.stats.kstest(z1[::64], "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z1, z2)[0, 1]) < 5 / math.sqrt(n)             # launches decorrelate
    assert abs(np.corrcoef(z1[:-1], z1[1:])[0, 1]) < 5 / math.sqrt(n)    # neighbours decorrelate
    assert abs(np.corrcoef(z1[:-4:4], z1[4::4])[0, 1]) < 10 / math.sqrt(n)   # neighbouring quads
    assert np.abs(z1).max() > 4.5                                        # tails reach out


# ---------------------------------------------------------------------------------
# error behaviour of the reference API (SURVEY 8b "Error conventions")
# ---------------------------------------------------------------------------------
def _params(n=3, size=100):
    return [torch.nn.Parameter(torch.randn(size, device=DEV)) for _ in range(n)]


def test_errors_match_the_reference():
    mcmc = _mcmc()
    with pytest.raises(AssertionError):
        mcmc.SGLD(_params(), lr=-1.0, num_data=1)
    ps = _params()
    opt = mcmc.SGLD(ps, lr=1e-2, num_data=1, momentum=0.9)
    with pytest.raises(AssertionError):
        opt.sample_momentum(keep=1.5)
    with pytest.raises(AssertionError):
        opt.step(save_state=True)                          # sgld.py:74
    with pytest.raises(RuntimeError, match="No gradient for parameter with shape"):
        opt.step()
    for p in ps:
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError, match="forgot to call `sample_momentum`"):
        opt.step()
    opt.sample_momentum()
    opt.step()
    # raise_on_nan (sgld.py:102-104)
    ps = _params()
    opt = mcmc.HMC(ps, lr=1e-2, num_data=1)                # HMC defaults to raise_on_nan=True
    opt.sample_momentum()
    for p in ps:
        p.grad = torch.zeros_like(p)
    ps[1].grad[7] = float("inf")
    with pytest.raises(ValueError, match="is not finite"):
        opt.initial_step()
    ps[1].grad[7] = float("nan")
    with pytest.raises(ValueError, match="is not finite"):
        opt.step()
    ps[1].grad[7] = 0.0
    opt.step()
    opt.param_groups[0]["temperature"] = 0.5
    with pytest.raises(AssertionError):                     # hmc.py:39
        opt.step()
    # groups that disagree (verlet_sgld.py:30-31, 53-54)
    a, b = _params(1), _params(1)
    opt = mcmc.VerletSGLD([dict(params=a), dict(params=b, num_data=7)], lr=1e-2, num_data=3, momentum=0.9)
    with pytest.raises(AssertionError, match="num_data"):
        opt.delta_energy(0., 0.)
    opt = mcmc.VerletSGLD([dict(params=_params(1)), dict(params=_params(1), temperature=2.)], lr=1e-2, num_data=3)
    with pytest.raises(AssertionError, match="temperature"):
        opt.maybe_reject(0.)


def test_raise_on_no_grad_false_skips_the_tensor():
    mcmc = _mcmc()
    ps = _params(3, 5000)
    before = [p.detach().clone() for p in ps]
    opt = mcmc.SGLD(ps, lr=1e-2, num_data=1, momentum=0.9, raise_on_no_grad=False)
    opt.sample_momentum()
    m1 = opt.state[ps[1]]["momentum_buffer"].clone()
    ps[0].grad = torch.ones_like(ps[0])
    ps[2].grad = torch.ones_like(ps[2])
    opt.step()
    assert torch.equal(ps[1].detach(), before[1])                       # untouched (sgld.py:96-101)
    assert torch.equal(opt.state[ps[1]]["momentum_buffer"], m1)
    assert not torch.equal(ps[0].detach(), before[0]) and not torch.equal(ps[2].detach(), before[2])


def test_metrics_are_only_refreshed_when_asked():
    mcmc = _mcmc()
    ps = _params(2, 3000)
    opt = mcmc.VerletSGLD(ps, lr=1e-2, num_data=5, momentum=0.9)
    opt.sample_momentum()
    for p in ps:
        p.grad = torch.randn_like(p)
    assert "est_temperature" not in opt.state[ps[0]]
    opt.initial_step(calc_metrics=True)
    t0 = opt.state[ps[0]]["est_temperature"]
    m_old = opt.state[ps[0]]["momentum_buffer"].double()
    opt.step(calc_metrics=False)
    assert opt.state[ps[0]]["est_temperature"] == t0                     # stale on purpose
    m_now = opt.state[ps[0]]["momentum_buffer"].double()
    opt.step(calc_metrics=True)
    want = float((m_now * m_now).sum()) / ps[0].numel()                 # momentum BEFORE this step
    assert opt.state[ps[0]]["est_temperature"] == pytest.approx(want, rel=1e-6)
    assert set(opt.param_groups[0]) >= {"params", "lr", "num_data", "momentum", "temperature", "rmsprop_alpha",
                                        "rmsprop_eps", "b^2h^2", "bh", "bhn", "mom_decay", "grad_v", "noise_std"}
    assert all(a is b for a, b in zip(opt.state.keys(), ps))             # state order = param order
    assert isinstance(opt.state[ps[0]]["preconditioner"], float)
    sq = opt.state[ps[0]]["square_avg"]
    assert sq.shape == ps[0].shape and float(sq.mean()) != 1.0


def test_views_survive_zero_grad_load_state_dict_and_schedulers():
    mcmc = _mcmc()
    lin = torch.nn.Linear(64, 32).to(DEV)
    ps = list(lin.parameters())
    opt = mcmc.SGLD(ps, lr=1e-3, num_data=1, momentum=0.9, temperature=0.0)
    (fg,) = opt.flat_groups
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.5 ** s)       # inference.py:96-101
    opt.sample_momentum()
    x = torch.randn(8, 64, device=DEV)
    for it in range(3):
        opt.zero_grad(set_to_none=False)
        assert all(p.grad is v for p, v in zip(ps, fg.g_views)) and float(fg.G.abs().sum()) == 0.0
        lin(x).pow(2).sum().backward()                        # autograd accumulates into the flat views
        assert all(p.grad is v for p, v in zip(ps, fg.g_views)) and float(fg.G.abs().sum()) > 0.0
        lin.load_state_dict({k: v.clone() for k, v in lin.state_dict().items()})   # inference.py:199-207
        assert all(p.data_ptr() == v.data_ptr() for p, v in zip(ps, fg.p_views))
        opt.step(calc_metrics=False)
        sched.step()
        assert opt.param_groups[0]["lr"] == pytest.approx(1e-3 * 0.5 ** (it + 1))
    # a parameter whose storage was swapped (Prior.sample(), prior/base.py:68-70) is re-adopted
    with torch.no_grad():
        ps[0].data = torch.full_like(ps[0], 3.0)
    opt.zero_grad()
    lin(x).sum().backward()
    opt.step(calc_metrics=False)
    assert ps[0].data_ptr() == fg.p_views[0].data_ptr()
    assert abs(float(ps[0].detach().mean()) - 3.0) < 0.1


def test_default_zero_grad_drops_gradients_and_step_copies_them_in():
    """torch >= 2 semantics (what the reference's runner gets from `optimizer.zero_grad()`):
    p.grad = None, backward() stores fresh tensors, the step reads them in place (gradient pointer
    table); same trajectory as accumulating into the views."""
    mcmc = _mcmc()
    torch.manual_seed(1)
    nets = [torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Tanh(), torch.nn.Linear(17, 3)).to(DEV) for _ in range(2)]
    nets[1].load_state_dict(nets[0].state_dict())
    opts = [mcmc.VerletSGLD(list(n.parameters()), lr=1e-2, num_data=10, momentum=0.9, temperature=1.0, seed=4) for n in nets]
    x = torch.randn(8, 33, device=DEV)
    for o in opts:
        o.sample_momentum()
    for it in range(4):
        for k, (n, o) in enumerate(zip(nets, opts)):
            o.zero_grad(set_to_none=(k == 0))
            if k == 0:
                assert all(p.grad is None for p in n.parameters())
            n(x).pow(2).sum().backward()
            if k == 0:
                (fg,) = o.flat_groups
                assert all(p.grad is not v for p, v in zip(fg.params, fg.g_views))
            o.step(calc_metrics=(it == 2))
        for p0, p1 in zip(nets[0].parameters(), nets[1].parameters()):
            assert torch.equal(p0, p1)
        # the gradients autograd handed over were read where they lie: nothing was copied into G
        assert opts[0].flat_groups[0].copies == 0
    # gradient accumulation over several backward calls without zero_grad (inference_reject.py:18-33)
    opts[0].zero_grad()
    for _ in range(3):
        nets[0](x).pow(2).sum().backward()
    opts[1].zero_grad(set_to_none=False)
    for _ in range(3):
        nets[1](x).pow(2).sum().backward()
    for o in opts:
        o.step(calc_metrics=False)
    for p0, p1 in zip(nets[0].parameters(), nets[1].parameters()):
        assert torch.allclose(p0, p1, rtol=1e-6, atol=1e-7)
    # after a rejection p.grad is the restored gradient (the flat view)
    o = opts[0]
    o.zero_grad()
    nets[0](x).pow(2).sum().backward()
    o.initial_step(save_state=True)
    o.zero_grad()
    nets[0](x).pow(2).sum().backward()
    o.final_step()
    torch.manual_seed(0)
    rejected, _ = o.maybe_reject(1e9)
    (fg,) = o.flat_groups
    assert rejected and all(p.grad is v for p, v in zip(fg.params, fg.g_views))
    assert torch.equal(fg.G, fg.prev_g)


def test_optimizer_protocol_state_dict_and_late_param_groups():
    mcmc = _mcmc()
    ps = _params(3, 50)
    opt = mcmc.VerletSGLD(ps, lr=1e-2, num_data=10, momentum=0.9, temperature=1.0)
    opt.sample_momentum()
    for p in ps:
        p.grad = torch.ones_like(p)
    opt.initial_step(save_state=True)
    sd = opt.state_dict()                                   # torch.optim.Optimizer protocol
    assert len(sd["param_groups"]) == 1 and sd["param_groups"][0]["lr"] == 1e-2
    assert len(sd["state"]) == 3
    st = sd["state"][0]
    assert torch.equal(st["momentum_buffer"], opt.state[ps[0]]["momentum_buffer"]) and "delta_energy" in st
    with pytest.raises(NotImplementedError, match="at construction"):
        opt.add_param_group(dict(params=_params(1, 5)))


def test_two_param_groups():
    mcmc = _mcmc()
    a, b = _params(2, 3000), _params(2, 70)
    pa, pb = [p.detach().clone() for p in a], [p.detach().clone() for p in b]
    opt = mcmc.SGLD([dict(params=a), dict(params=b, lr=0.0)], lr=1e-2, num_data=10, momentum=0.9, temperature=0.0)
    opt.sample_momentum()
    for p in a + b:
        p.grad = torch.ones_like(p)
    opt.step()
    assert all(not torch.equal(p.detach(), q) for p, q in zip(a, pa))
    assert all(torch.equal(p.detach(), q) for p, q in zip(b, pb))         # lr = 0 -> h = 0
    assert len(opt.flat_groups) == 2


def test_sgd_equivalence():
    """testing/test_sgld.py:61-80: temperature 0, num_data 1 == torch.optim.SGD with momentum."""
    mcmc = _mcmc()
    torch.manual_seed(0)
    net1 = torch.nn.Sequential(torch.nn.Linear(20, 30), torch.nn.Tanh(), torch.nn.Linear(30, 4)).to(DEV)
    net2 = torch.nn.Sequential(torch.nn.Linear(20, 30), torch.nn.Tanh(), torch.nn.Linear(30, 4)).to(DEV)
    net2.load_state_dict(net1.state_dict())
    lr = 0.1
    sgld = mcmc.SGLD(net1.parameters(), lr=lr, num_data=1, momentum=0.9, temperature=0.)
    sgd = torch.optim.SGD(net2.parameters(), lr=lr, momentum=0.9)
    sgld.sample_momentum()
    x, y = torch.randn(16, 20, device=DEV), torch.randn(16, 4, device=DEV)
    for _ in range(4):
        for net, opt in ((net1, sgld), (net2, sgd)):
            opt.zero_grad()
            (net(x) - y).pow(2).mean().backward()
            opt.step()
        for p1, p2 in zip(net1.parameters(), net2.parameters()):
            assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-6)
