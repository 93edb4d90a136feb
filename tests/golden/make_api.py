#!/usr/bin/env python
"""Record the public surface of the reference's bnn_priors.mcmc (class names, method
names, signatures, base classes) into tests/golden/api_signatures.json, so that the
drop-in classes can be checked against it where /root/reference does not exist.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_api.py
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("BNNP_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REFERENCE)

from bnn_priors import mcmc  # noqa: E402


def sig(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([name, p.kind.name, d])
    return out


api = {}
for cname in ("SGLD", "VerletSGLD", "HMC"):
    cls = getattr(mcmc, cname)
    methods = {}
    for m, fn in inspect.getmembers(cls, predicate=inspect.isfunction):
        if m.startswith("__") and m != "__init__":
            continue
        if not any(m in k.__dict__ for k in cls.__mro__ if k.__module__.startswith("bnn_priors")):
            continue
        methods[m] = sig(fn)
    api[cname] = dict(bases=[b.__name__ for b in cls.__mro__[1:] if b is not object], methods=methods)
with open(os.path.join(HERE, "api_signatures.json"), "w") as f:
    json.dump(api, f, indent=1, sort_keys=True)
print({k: sorted(v["methods"]) for k, v in api.items()})
