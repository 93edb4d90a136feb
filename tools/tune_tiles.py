#!/usr/bin/env python
"""Kernel time of the main transitions for one build of libbnnp.so (BNNP_LIB picks
it): used to choose BNNP_THREADS / BNNP_UNROLL / BNNP_MIN_CTAS.  GPU box only.

    for f in bnn_priors_b200/_lib/tune/*.so; do BNNP_LIB=$PWD/$f python tools/tune_tiles.py; done
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
K = int(os.environ.get("TUNE_STEPS", "100"))
out = {"lib": os.path.basename(os.environ.get("BNNP_LIB", "default"))}
cases = (("sgld", "SGLD", False, lambda o: o.step(calc_metrics=False)),
         ("sgld_metrics", "SGLD", False, lambda o: o.step(calc_metrics=True)),
         ("verlet", "VerletSGLD", False, lambda o: o.step(calc_metrics=False)),
         ("verlet_fused", "VerletSGLD", True, lambda o: o.step(calc_metrics=False)),
         ("verlet_save", "VerletSGLD", False, lambda o: o.initial_step(save_state=True, calc_metrics=False)),
         ("hmc", "HMC", False, lambda o: o.step(calc_metrics=False)))
only = os.environ.get("TUNE_CASES")
for name, smp, fused, call in cases:
    if only and name not in only.split(","):
        continue
    opt, params, fg = bench.make_chain(dev, 0, smp, fused_prior=fused)
    call(opt)
    for _ in range(200):
        fg.relaunch()
    ms = sorted(bench.timed_gpu(fg.relaunch, K, dev, False) / K for _ in range(5))
    out[name] = [round(ms[0] * 1e3, 2), round(ms[2] * 1e3, 2)]      # min, median of 5
    del opt, params, fg
    torch.cuda.empty_cache()
if not only or "verlet_hier" in only.split(","):
    # hierarchical priors: every weight tensor with a prior gets a sampled scale
    # (NormalGamma); a step = read-only pre-pass over P + its epilogue + the step launch
    from bnn_priors_b200 import _native as N, mcmc
    tensors = bench.load_tensors()
    g = torch.Generator(device=dev).manual_seed(0)
    params, links = [], []
    for t in tensors:
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
    step = lambda: opt.step(calc_metrics=False)  # noqa: E731
    for _ in range(50):
        step()
    ms = sorted(bench.timed_gpu(step, K, dev, False) / K for _ in range(5))
    out["verlet_hier"] = [round(ms[0] * 1e3, 2), round(ms[2] * 1e3, 2)]
    out["verlet_hier_links"] = len(links)
    pre = lambda: fg.hyper_prepass(1.0 / bench.HP["num_data"])  # noqa: E731
    ms = sorted(bench.timed_gpu(pre, K, dev, False) / K for _ in range(5))
    out["hier_prepass_plus_epilogue"] = [round(ms[0] * 1e3, 2), round(ms[2] * 1e3, 2)]
    del opt, params, fg
    torch.cuda.empty_cache()
if not only or "rollback" in only.split(","):
    # the restore after a rejected proposal (verlet_sgld.py:63-69): P, G, M <- prev_*  (24 B/param)
    opt, params, fg = bench.make_chain(dev, 0, "VerletSGLD")
    opt.initial_step(save_state=True, calc_metrics=False)
    for _ in range(20):
        fg.rollback()
    ms = sorted(bench.timed_gpu(fg.rollback, K, dev, False) / K for _ in range(5))
    out["rollback"] = [round(ms[0] * 1e3, 2), round(ms[2] * 1e3, 2)]
    del opt, params, fg
    torch.cuda.empty_cache()
if not only or "probe" in only.split(","):
    # ceiling of the access pattern: read p, g, m / write p, m and nothing else
    from bnn_priors_b200 import _native as N
    n = 25124864
    P, G, M = (torch.zeros(n, device=dev) for _ in range(3))
    lib = N.lib()
    s = torch.cuda.current_stream(dev).cuda_stream
    probe = lambda: N.check(lib.bnnp_probe_stream(P.data_ptr(), G.data_ptr(), M.data_ptr(), n, s), "probe")  # noqa: E731
    for _ in range(200):
        probe()
    ms = sorted(bench.timed_gpu(probe, K, dev, False) / K for _ in range(5))
    out["probe_stream"] = [round(ms[0] * 1e3, 2), round(ms[2] * 1e3, 2)]
print(json.dumps(out), flush=True)
