#!/usr/bin/env python
"""Record the compact golden traces at the REAL segment tables (tests/real_tables.py) from the
UNMODIFIED reference samplers (CPU, fp32).  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_real_golden.py [case ...]

Inputs are seeds (initial parameters, likelihood gradients, N(0,1) tensors); the prior's share of
every gradient is what the reference's own prior densities give through autograd at the reference's
current parameters.  Outputs are fingerprints (strided samples + moments), scalars, delta energies
and decisions.  Nothing here is on the product path.
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, os.environ.get("BNNP_REFERENCE", "/root/reference"))

import real_tables as RT  # noqa: E402
from bnn_priors import mcmc as ref_mcmc  # noqa: E402

SCALAR_KEYS = ("preconditioner", "est_temperature", "est_config_temp", "delta_energy", "prev_new_momentum_delta")


def scalars(opt, params):
    out = {}
    for k in SCALAR_KEYS:
        vals = [opt.state[p].get(k) for p in params]
        out[k] = [None if v is None else float(v) for v in vals]
    out["square_avg_mean"] = [float(opt.state[p]["square_avg"].double().mean()) if "square_avg" in opt.state[p] else None
                              for p in params]
    return out


def record(name):
    inp = RT.Inputs(name)
    params = [torch.nn.Parameter(t.clone()) for t in inp.p0()]
    ctor = dict(inp.ctor)
    if inp.sampler == "HMC":
        ctor["raise_on_nan"] = True
    opt = getattr(ref_mcmc, inp.sampler)(params, **ctor)
    lr0, num_data = inp.ctor["lr"], inp.ctor["num_data"]
    arrays, events = {}, []
    nfp = 0
    last_de = None
    real_randn_like, real_rand = torch.randn_like, torch.rand
    for k, spec in enumerate(inp.script):
        ev = {key: v for key, v in spec.items()}
        op = ev["op"]
        for g in opt.param_groups:
            g["lr"] = lr0 * ev.get("lr_scale", 1.0)
        if op in RT.R.STEP_OPS:
            lik = inp.lik_grad(ev["grad"])
            pg = inp.prior_grad([p.detach() for p in params], num_data)
            for p, a, b in zip(params, lik, pg):
                p.grad = a + b
            with torch.no_grad():
                ev["pg_abs"] = [float((p.detach() * p.grad).abs().double().sum()) for p in params]
        noise = iter(inp.noise(k))
        used = []

        def randn_like(t, *a, **kw):
            z = next(noise)
            assert z.shape == t.shape
            used.append(1)
            return z.clone()
        torch.randn_like = randn_like
        if op == "maybe_reject":
            torch.rand = lambda *a, **kw: torch.tensor(ev["u"], dtype=torch.float32)
        try:
            if op == "delta_energy":
                # the synthetic gradients are not the gradient of any potential, so the sampler's part
                # of the energy difference is arbitrary: the potential difference handed over is chosen
                # such that the total lands at the script's target (one acceptance, one rejection)
                u0 = ev.pop("u0")
                target = ev.pop("target") * inp.ctor.get("temperature", 1.0)
                de0 = opt.delta_energy(u0, u0)
                ev["potentials"] = [u0, u0 - (de0 - target) / num_data]
                out = opt.delta_energy(*ev["potentials"])
                ev["out"] = float(out)
                last_de = float(out)
            elif op == "maybe_reject":
                ev["arg"] = last_de
                rej, lap = opt.maybe_reject(last_de)
                ev["out"] = [bool(rej), float(lap)]
            else:
                getattr(opt, op)(**ev.get("kwargs", {}))
        finally:
            torch.randn_like, torch.rand = real_randn_like, real_rand
        ev["noise"] = bool(used)
        if op not in ("delta_energy", "update_preconditioner") and not ev.pop("skip_fp", False):
            ps, pm = RT.fingerprint([p.detach() for p in params])
            arrays[f"ps{nfp}"], arrays[f"pm{nfp}"] = ps, pm
            moms = [opt.state[p].get("momentum_buffer") for p in params]
            if all(m is not None for m in moms):
                ms, mm = RT.fingerprint(moms)
                arrays[f"ms{nfp}"], arrays[f"mm{nfp}"] = ms, mm
            ev["fp"] = nfp
            nfp += 1
        if "fp" in ev or op in ("delta_energy", "update_preconditioner"):
            ev["scalars"] = scalars(opt, params)
        else:
            ev.pop("pg_abs", None)
        events.append(ev)
    decisions = [e["out"][0] for e in events if e["op"] == "maybe_reject"]
    meta = dict(case=name, table=inp.tag, sampler=inp.sampler, ctor=inp.ctor, events=events,
                n_params=int(sum(inp.sizes)), tensors=len(inp.sizes))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"{name}: {len(events)} events, {sum(inp.sizes)} params in {len(inp.sizes)} tensors, decisions {decisions}, "
          f"{os.path.getsize(path) / 1024:.0f} KiB")
    assert all(math.isfinite(e["out"]) for e in events if e["op"] == "delta_energy")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for name in (sys.argv[1:] or list(RT.CASES)):
        record(name)
