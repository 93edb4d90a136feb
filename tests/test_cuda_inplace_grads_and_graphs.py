"""Round-2 plumbing of the sampler step, on the GPU:

  * gradients are read IN PLACE through the per-segment pointer table (BnnpLaunch.seg_grad):
    after the runner's `zero_grad()` + `backward()` a step launches no copy, and gives the same
    bits as the same step with the gradients in the flat G array;
  * capturable mode (BnnpLaunch.ctl + bnnp_advance): same bits as the normal mode, and a
    forward / backward / step iteration recorded in a CUDA graph replays to the same bits as the
    eager loop -- new noise, alternating direction, folded sums and a changed learning rate included;
  * state_dict / load_state_dict round trip, the non-finite pre-check, torch's step hooks.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
SHAPES = [(257,), (40, 30), (5,), (4097,), (3, 4100), (1,)]
HP = dict(lr=2e-3, num_data=40.0, momentum=0.9, temperature=0.7)


def _mcmc():
    from bnn_priors_b200 import mcmc
    return mcmc


def _make(kind="VerletSGLD", seed=3, shapes=SHAPES, **kw):
    g = torch.Generator(device=DEV).manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(s, device=DEV, generator=g)) for s in shapes]
    hp = dict(HP)
    if kind == "HMC":
        hp = dict(lr=HP["lr"], num_data=HP["num_data"], raise_on_nan=False)
    hp.update(kw)
    opt = getattr(_mcmc(), kind)(params, **hp, seed=seed)
    return opt, params


def _grads(shapes, n, seed=11):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [[torch.randn(s, device=DEV, generator=g) * 0.3 for s in shapes] for _ in range(n)]


def _run(opt, params, grad_seq, mode, calls=None):
    """drive `opt` through initial_step / step... / final_step with the given gradients, handed over
    as tensors of their own ("foreign", what autograd does after zero_grad()) or copied into the
    flat views ("views")"""
    opt.sample_momentum()
    names = calls or (["initial_step"] + ["step"] * (len(grad_seq) - 2) + ["final_step"])
    for name, gs in zip(names, grad_seq):
        if mode == "foreign":
            opt.zero_grad()
            for p, g in zip(params, gs):
                p.grad = g.clone()
        else:
            opt.zero_grad(set_to_none=False)
            for p, g in zip(params, gs):
                p.grad.copy_(g)
        getattr(opt, name)(calc_metrics=(name != "step"))
    return [p.detach().clone() for p in params], [opt.state[p]["momentum_buffer"].clone() for p in params]


@pytest.mark.parametrize("kind", ["SGLD", "VerletSGLD", "HMC"])
def test_foreign_gradients_are_read_in_place_and_give_the_same_bits(kind):
    grads = _grads(SHAPES, 6)
    a, pa = _make(kind)
    b, pb = _make(kind)
    ra = _run(a, pa, grads, "views")
    rb = _run(b, pb, grads, "foreign")
    for x, y in zip(ra[0] + ra[1], rb[0] + rb[1]):
        assert torch.equal(x, y)
    (fg,) = b.flat_groups
    assert fg.copies == 0                      # nothing was copied into G
    for p, q in zip(pa, pb):
        for k in ("est_temperature", "est_config_temp"):
            assert a.state[p][k] == b.state[q][k]
    if kind != "SGLD":
        assert a.delta_energy(1.0, 1.5) == b.delta_energy(1.0, 1.5)


def test_pointer_table_is_rewritten_only_when_an_address_changes():
    opt, params = _make("SGLD")
    (fg,) = opt.flat_groups
    opt.sample_momentum()
    bufs = [torch.randn_like(p) for p in params]
    for p, g in zip(params, bufs):
        p.grad = g
    opt.step(calc_metrics=False)
    writes = fg.table_writes
    for _ in range(5):                         # new tensor OBJECTS at the same addresses (what the caching
        opt.zero_grad()                        # allocator gives a steady training loop)
        for p, g in zip(params, bufs):
            p.grad = g.view_as(g)
        opt.step(calc_metrics=False)
    assert fg.table_writes == writes and fg.copies == 0
    params[2].grad = torch.randn_like(params[2])   # one gradient moved
    opt.step(calc_metrics=False)
    assert fg.table_writes == writes + 1 and fg.copies == 0


def test_a_gradient_that_cannot_be_read_in_place_is_copied():
    opt, params = _make("SGLD")
    ref, rparams = _make("SGLD")
    (fg,) = opt.flat_groups
    gs = _grads(SHAPES, 1)[0]
    for o, ps in ((opt, params), (ref, rparams)):
        o.sample_momentum()
    for p, g in zip(rparams, gs):
        p.grad = g.clone()
    big = torch.zeros(gs[1].numel() + 1, device=DEV)
    for i, (p, g) in enumerate(zip(params, gs)):
        if i == 1:
            big[1:].copy_(g.reshape(-1))
            p.grad = big[1:].view_as(g)        # 4-byte aligned only
        elif i == 4:
            p.grad = g.t().contiguous().t()    # not contiguous
        else:
            p.grad = g.clone()
    opt.step(calc_metrics=True)
    ref.step(calc_metrics=True)
    assert fg.copies == 2
    for p, q in zip(params, rparams):
        assert torch.equal(p, q)
    opt.step(calc_metrics=True)                # same tensors, same versions: not copied again
    assert fg.copies == 2
    params[1].grad.mul_(2.0)
    opt.step(calc_metrics=True)
    assert fg.copies == 3


def test_freed_gradients_are_never_read():
    """model.log_prior() right after zero_grad() (inference_reject.py:19-20) must not touch the
    gradient pointers, and a step with a missing gradient skips that tensor."""
    opt, params = _make("VerletSGLD", raise_on_no_grad=False)
    (fg,) = opt.flat_groups
    from bnn_priors_b200 import _native as N
    for i in range(fg.nseg):
        fg.set_prior(i, N.PRIOR_NORMAL, 0.0, 1.0, 3.0)
    fg.prior_fused = True
    opt.sample_momentum()
    for p in params:
        p.grad = torch.randn_like(p)
    opt.initial_step()
    opt.zero_grad()
    torch.cuda.empty_cache()
    fg.reduce_log_prior(1.0 / 40.0)
    lp = float(fg.fetch()[:, N.S_LOG_PRIOR].sum())
    want = sum(float(torch.distributions.Normal(0., 1.).log_prob(p.detach().double()).sum()) for p in params)
    assert abs(lp - want) < 1e-5 * abs(want)
    before = [p.detach().clone() for p in params]
    for p in params[:3]:
        p.grad = torch.randn_like(p)
    opt.step()
    for i, (p, b) in enumerate(zip(params, before)):
        assert torch.equal(p, b) == (i >= 3)


@pytest.mark.parametrize("kind", ["SGLD", "VerletSGLD", "HMC"])
def test_capturable_mode_gives_the_same_bits(kind):
    grads = _grads(SHAPES, 7)
    a, pa = _make(kind)
    b, pb = _make(kind, capturable=True)
    ra = _run(a, pa, grads, "foreign")
    rb = _run(b, pb, grads, "foreign")
    for x, y in zip(ra[0] + ra[1], rb[0] + rb[1]):
        assert torch.equal(x, y)
    for p, q in zip(pa, pb):
        for k in ("est_temperature", "est_config_temp"):
            assert a.state[p][k] == b.state[q][k]
    if kind != "SGLD":
        assert a.delta_energy(1.0, 1.5) == b.delta_energy(1.0, 1.5)
        assert [a.state[p]["delta_energy"] for p in pa] == [b.state[q]["delta_energy"] for q in pb]


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(20, 64)
        self.b = torch.nn.Linear(64, 64)
        self.c = torch.nn.Linear(64, 4)

    def forward(self, x):
        return self.c(torch.relu(self.b(torch.relu(self.a(x)))))


@pytest.mark.parametrize("kind", ["SGLD", "VerletSGLD"])
def test_a_captured_forward_backward_step_iteration_replays_to_the_eager_bits(kind):
    """A whole iteration -- zero_grad, forward, backward, sampler step -- in ONE CUDA graph."""
    torch.manual_seed(0)
    x = torch.randn(256, 20, device=DEV)
    y = torch.randint(0, 4, (256,), device=DEV)
    net0 = _Net().to(DEV)
    nets = [copy.deepcopy(net0) for _ in range(2)]
    hp = dict(lr=1e-3, num_data=256.0, momentum=0.9, temperature=1.0)
    opts = [getattr(_mcmc(), kind)(list(n.parameters()), **hp, seed=5, capturable=True) for n in nets]
    lrs = [1e-3, 1e-3, 7e-4, 7e-4, 7e-4, 2e-4, 2e-4, 2e-4]

    def iteration(net, opt):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(x), y)
        loss.backward()
        opt.step(calc_metrics=True)
        return loss

    for net, opt in zip(nets, opts):
        opt.sample_momentum()
        if kind == "VerletSGLD":
            opt.zero_grad()
            torch.nn.functional.cross_entropy(net(x), y).backward()
            opt.initial_step()
    # eager
    for lr in lrs:
        opts[0].param_groups[0]["lr"] = lr
        iteration(nets[0], opts[0])
    # captured: warm up on a side stream (torch's recipe), capture one iteration, replay
    net, opt = nets[1], opts[1]
    n_warm = 2
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for lr in lrs[:n_warm]:
            opt.param_groups[0]["lr"] = lr
            iteration(net, opt)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    opt.param_groups[0]["lr"] = lrs[n_warm]
    opt.sync_hyperparameters()
    with torch.cuda.graph(graph):              # recording only: nothing runs yet
        static_loss = iteration(net, opt)
    for lr in lrs[n_warm:]:
        opt.param_groups[0]["lr"] = lr
        opt.sync_hyperparameters()             # a changed lr reaches the graph through the control block
        graph.replay()
    torch.cuda.synchronize()
    for p, q in zip(nets[0].parameters(), nets[1].parameters()):
        assert torch.equal(p, q)
        assert torch.equal(opts[0].state[p]["momentum_buffer"], opts[1].state[q]["momentum_buffer"])
        assert opts[0].state[p]["est_temperature"] == opts[1].state[q]["est_temperature"]
    assert torch.isfinite(static_loss)
    if kind == "VerletSGLD":
        assert opts[0].delta_energy(0.0, 0.1) == opts[1].delta_energy(0.0, 0.1)


def test_a_stale_coefficient_is_not_baked_into_a_graph():
    opt, params = _make("SGLD", capturable=True)
    opt.sample_momentum()
    for p in params:
        p.grad = torch.randn_like(p)
    opt.step(calc_metrics=False)
    opt.param_groups[0]["lr"] = 1e-4            # changed behind the sampler's back, no eager step since
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="sync_hyperparameters"):
        with torch.cuda.graph(torch.cuda.CUDAGraph()):
            opt.step(calc_metrics=False)
    torch.cuda.synchronize()
    opt.sync_hyperparameters()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step(calc_metrics=False)
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p).all() for p in params)


def test_raise_on_nan_refuses_to_be_captured():
    opt, params = _make("HMC", capturable=True, raise_on_nan=True)
    opt.sample_momentum()
    for p in params:
        p.grad = torch.randn_like(p)
    opt.initial_step(save_state=False)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="raise_on_nan=False"):
        with torch.cuda.graph(torch.cuda.CUDAGraph()):
            opt.step(calc_metrics=False)
    torch.cuda.synchronize()
    opt.step(calc_metrics=False)                 # eager steps are unaffected
    assert all(torch.isfinite(p).all() for p in params)


def test_state_dict_round_trip():
    grads = _grads(SHAPES, 4)
    a, pa = _make("VerletSGLD")
    _run(a, pa, grads, "foreign", calls=["initial_step", "step", "step", "step"])
    a.update_preconditioner()
    sd = copy.deepcopy(a.state_dict())
    b, pb = _make("VerletSGLD", seed=3)
    with torch.no_grad():
        for p, q in zip(pa, pb):
            q.copy_(p)
    b.load_state_dict(sd)
    for p, q in zip(pa, pb):
        sa, sb = a.state[p], b.state[q]
        assert torch.equal(sa["momentum_buffer"], sb["momentum_buffer"])
        assert torch.equal(sa["prev_parameter"], sb["prev_parameter"])
        for k in ("preconditioner", "delta_energy", "prev_new_momentum_delta"):
            assert sa[k] == sb[k], k
        assert abs(float(sa["square_avg"].mean()) - float(sb["square_avg"].mean())) <= 1e-6 * float(sa["square_avg"].mean())
    # and both continue identically (same Philox key; align the counters)
    b.flat_groups[0].call = a.flat_groups[0].call
    b.flat_groups[0]._parity = a.flat_groups[0]._parity
    g = _grads(SHAPES, 1, seed=99)[0]
    for o, ps in ((a, pa), (b, pb)):
        for p, t in zip(ps, g):
            p.grad = t.clone()
        o.step()
    for p, q in zip(pa, pb):
        assert torch.equal(p, q)


@pytest.mark.parametrize("kind", ["SGLD", "HMC"])
def test_non_finite_gradient_raises_before_anything_is_touched(kind):
    opt, params = _make(kind, raise_on_nan=True)
    opt.sample_momentum()
    for p in params:
        p.grad = torch.randn_like(p)
    opt.initial_step()
    before = [p.detach().clone() for p in params]
    mom = [opt.state[p]["momentum_buffer"].clone() for p in params]
    sq = [float(opt.state[p]["square_avg"].mean()) for p in params]
    params[3].grad[17] = float("inf")
    with pytest.raises(ValueError, match="is not finite"):
        opt.step()
    for p, b, m, s in zip(params, before, mom, sq):
        assert torch.equal(p, b) and torch.equal(opt.state[p]["momentum_buffer"], m)
        assert float(opt.state[p]["square_avg"].mean()) == s
    params[3].grad[17] = 0.0
    opt.step()                                   # the sampler carries on
    assert all(torch.isfinite(p).all() for p in params)


def test_step_hooks_still_fire():
    opt, params = _make("SGLD")
    opt.sample_momentum()
    for p in params:
        p.grad = torch.randn_like(p)
    seen = []
    h1 = opt.register_step_pre_hook(lambda o, a, k: seen.append("pre"))
    h2 = opt.register_step_post_hook(lambda o, a, k: seen.append("post"))
    opt.step(calc_metrics=False)
    assert seen == ["pre", "post"]
    h1.remove()
    h2.remove()
    opt.step(calc_metrics=False)
    assert seen == ["pre", "post"]
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda i: 0.5 ** i)
    opt.step(calc_metrics=False)
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - HP["lr"] * 0.5) < 1e-15
