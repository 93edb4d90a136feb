#!/usr/bin/env python
"""A short, fixed sequence of launches of one step-kernel variant for ncu (GPU box only).

    ncu ... python tools/ncu_target.py <case> [n]

cases: sgld | sgld_metrics | verlet | verlet_fused | verlet_save | hmc | sgld_foreign (the gradient read in
place from tensors of their own, as after the runner's zero_grad() + backward())."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "sgld"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
smp, fused, call = {
    "verlet_hier": ("VerletSGLD", True, None),
    "sgld": ("SGLD", False, lambda o: o.step(calc_metrics=False)),
    "sgld_foreign": ("SGLD", False, lambda o: o.step(calc_metrics=False)),
    "sgld_metrics": ("SGLD", False, lambda o: o.step(calc_metrics=True)),
    "verlet": ("VerletSGLD", False, lambda o: o.step(calc_metrics=False)),
    "verlet_fused": ("VerletSGLD", True, lambda o: o.step(calc_metrics=False)),
    "verlet_save": ("VerletSGLD", False, lambda o: o.initial_step(save_state=True, calc_metrics=False)),
    "hmc": ("HMC", False, lambda o: o.step(calc_metrics=False)),
}[case]
if case == "verlet_hier":
    # every prior-carrying weight tensor gets a sampled scale (NormalGamma): step + epilogue launch per step
    from bnn_priors_b200 import _native as N, mcmc
    g = torch.Generator(device=dev).manual_seed(0)
    params, links = [], []
    for t in bench.load_tensors():
        params.append(torch.nn.Parameter(torch.randn(tuple(t["shape"]), device=dev, generator=g) * (t["scale"] if t["kind"] else 1.0)))
        if t["kind"] and len(t["shape"]) > 1:
            links.append((len(params) - 1, len(params), t))
            params.append(torch.nn.Parameter(torch.tensor(0.1, device=dev)))
    opt = mcmc.VerletSGLD(params, **bench.HP, seed=0)
    (fg,) = opt.flat_groups
    for w, h, t in links:
        fg.set_prior(w, N.PRIOR_NORMAL, 0.0, t["scale"], 3.0)
        fg.set_hyper_link(w, h, N.PRIOR_HYPER_GAMMA, 1.0, 1.0)
    fg.prior_fused = True
    for p, v in zip(params, fg.g_views):
        p.grad = v
        v.normal_(0.0, 1e-3, generator=g)
    opt.sample_momentum()
    for _ in range(n):
        opt.step(calc_metrics=False)
    torch.cuda.synchronize()
    print(case, "launches", fg.launches)
    sys.exit(0)
opt, params, fg = bench.make_chain(dev, 0, smp, fused_prior=fused)
if case == "sgld_foreign":
    bufs = [torch.randn_like(p) * 1e-3 for p in params]
    for _ in range(3):
        opt.zero_grad()
        for p, g in zip(params, bufs):
            p.grad = g.view_as(g)
        call(opt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off: only the loop below is listed
    for _ in range(n):
        opt.zero_grad()
        for p, g in zip(params, bufs):
            p.grad = g.view_as(g)
        call(opt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    assert fg.copies == 0
else:
    call(opt)
    for _ in range(n - 1):
        fg.relaunch()
torch.cuda.synchronize()
print(case, "launches", fg.launches)
