"""The optional host-side helper (csrc/bnnp_host.cpp, a torch C++ extension built in-tree by
`bnn_priors_b200.build.build_host`): the per-step scan of the parameters' gradients over the ATen objects.
Its answers must be the ones the Python scan in mcmc/_flat.py gives.  CPU tensors are enough here (it only
looks at pointers, versions and definedness); the GPU suite runs the samplers with and without it."""
import pytest
import torch

from bnn_priors_b200 import _native as N


@pytest.fixture(scope="module")
def hm():
    m = N.host_module()
    if m is None:
        pytest.skip("bnnp_host.so not built (python -m bnn_priors_b200.build)")
    return m


def test_scan_answers(hm):
    ps = [torch.nn.Parameter(torch.randn(s)) for s in ((3, 4), (5,), (), (7, 2))]
    sc = hm.GradScanner(ps, [p.data_ptr() for p in ps])
    assert sc.scan() == 1                                  # no gradients yet
    for p in ps:
        p.grad = torch.randn_like(p)
    assert sc.scan() == 1                                  # gradients the table does not know
    sc.set_table([p.grad.data_ptr() for p in ps])
    assert sc.scan() == 0
    keep = [p.grad for p in ps]
    for p, g in zip(ps, keep):
        p.grad = g.view_as(g)                              # new tensor objects, same addresses
    assert sc.scan() == 0
    ps[1].grad = torch.randn_like(ps[1])                   # one gradient moved
    assert sc.scan() == 1
    ps[1].grad = keep[1]
    assert sc.scan() == 0
    sq = torch.nn.Parameter(torch.randn(4, 4))
    sq.grad = torch.randn(4, 4)
    sc2 = hm.GradScanner([sq], [sq.data_ptr()])
    sc2.set_table([sq.grad.data_ptr()])
    assert sc2.scan() == 0
    sq.grad = sq.grad.t()                                  # same first byte, another layout
    assert sc2.scan() == 1
    ps[2].data = torch.tensor(1.0)                         # a parameter's storage was swapped
    assert sc.scan() == 2
    ps[3].grad = None
    assert sc.scan() == 1


def test_freshness_signature_and_versions(hm):
    ps = [torch.nn.Parameter(torch.randn(6)) for _ in range(3)]
    for p in ps:
        p.grad = torch.randn_like(p)
    sc = hm.GradScanner(ps, [p.data_ptr() for p in ps])
    assert not sc.fresh()
    assert sc.capture() and sc.fresh()
    ps[0].grad.mul_(2.0)                                   # modified in place
    assert not sc.fresh()
    assert sc.capture() and sc.fresh()
    old = ps[1].grad
    ps[1].grad = old.clone()                               # another tensor with the same values
    assert not sc.fresh()
    ps[1].grad = old
    assert sc.fresh()
    sc.drop()
    assert not sc.fresh()
    v = sc.params_version()
    assert v == sum(p._version for p in ps)
    with torch.no_grad():
        ps[2].add_(1.0)
    assert sc.params_version() == v + 1 == sum(p._version for p in ps)


def test_drop_grads_is_zero_grad_with_exceptions(hm):
    ps = [torch.nn.Parameter(torch.randn(4)) for _ in range(5)]
    for p in ps:
        p.grad = torch.randn_like(p)
    kept = ps[3].grad
    sc = hm.GradScanner(ps, [p.data_ptr() for p in ps])
    sc.capture()
    sc.drop_grads([3])
    assert [p.grad is None for p in ps] == [True, True, True, False, True] and ps[3].grad is kept
    assert not sc.fresh()
    torch.randn(4, requires_grad=True).sum().backward()    # autograd is unimpressed
    (ps[0] * 2).sum().backward()
    assert torch.equal(ps[0].grad, torch.full((4,), 2.0))
