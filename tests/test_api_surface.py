"""The drop-in classes expose the reference's public surface: same class
hierarchy, same public method names and signatures (tests/golden/api_signatures.json,
recorded from the reference by tests/golden/make_api.py).  Engine-only additions
must be keyword-only."""
import inspect
import json
import os

import pytest
import torch

from bnn_priors_b200 import mcmc

HERE = os.path.dirname(os.path.abspath(__file__))
API = json.load(open(os.path.join(HERE, "golden", "api_signatures.json")))
PUBLIC = ("__init__", "sample_momentum", "step", "initial_step", "final_step", "delta_energy",
          "maybe_reject", "update_preconditioner", "_point_energy", "_update_group_fn",
          "_preconditioner_default")


@pytest.mark.parametrize("cname", ["SGLD", "VerletSGLD", "HMC"])
def test_signatures(cname):
    cls = getattr(mcmc, cname)
    ref = API[cname]
    assert [b.__name__ for b in cls.__mro__[1:] if b is not object] == ref["bases"]
    assert issubclass(cls, torch.optim.Optimizer)
    for m, want in ref["methods"].items():
        if m not in PUBLIC:
            continue
        got = [(n, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default))
               for n, p in inspect.signature(getattr(cls, m)).parameters.items()]
        extra = got[len(want):]
        assert [list(g) for g in got[:len(want)]] == want, (cname, m)
        assert all(kind == "KEYWORD_ONLY" for _, kind, _ in extra), (cname, m, extra)


def test_module_exports():
    assert sorted(mcmc.__all__) == ["HMC", "SGLD", "VerletSGLD"]
    from bnn_priors_b200.mcmc import sgld
    assert callable(sgld.dot)


def test_live_reference_signatures_if_present():
    """In the build container, compare against the reference itself as well."""
    import sys
    if not os.path.isdir("/root/reference/bnn_priors"):
        pytest.skip("no reference checkout on this machine")
    sys.path.insert(0, os.path.join(HERE, "golden", "_shims"))
    sys.path.insert(0, "/root/reference")
    try:
        from bnn_priors import mcmc as ref
    finally:
        sys.path.remove("/root/reference")
    for cname in ("SGLD", "VerletSGLD", "HMC"):
        for m in ("step", "initial_step", "final_step", "sample_momentum", "delta_energy"):
            a = list(inspect.signature(getattr(getattr(ref, cname), m)).parameters)
            b = list(inspect.signature(getattr(getattr(mcmc, cname), m)).parameters)
            assert a == b[:len(a)], (cname, m, a, b)
