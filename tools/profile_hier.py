#!/usr/bin/env python
"""Where the time of a step with hierarchical priors goes (25M-parameter chain, 21 sampled
scales): device time of the pre-pass kernel, its epilogue launch and the step kernel (raw
C-ABI relaunches, CUDA events), and host time of the Python calls around them.  GPU box only."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bnn_priors_b200 import _native as N, mcmc  # noqa: E402

dev = torch.device("cuda", 0)
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
inv_n = 1.0 / bench.HP["num_data"]
K = 200
out = {}


def dev_us(fn):
    for _ in range(20):
        fn()
    return bench.timed_gpu(fn, K, dev, False) / K * 1e3


def host_us(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        fn()
    dt = (time.perf_counter() - t0) / K * 1e6
    torch.cuda.synchronize()
    return dt


# the pre-pass kernel alone: relaunch its argument block (flush in between, as bnnp_launch demands)
fg.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_P | N.F_LOG_PRIOR | N.F_HYPER, N.NOISE_NONE, cm=1.0, inv_num_data=inv_n)
fg.flush_pending()


def prepass_only():
    fg._pending = None          # measurement only: drop the epilogue instead of applying it
    fg.relaunch()


out["prepass_kernel_us"] = dev_us(prepass_only)
fg.flush_pending()


def finalize_only():
    fg._pending = pend
    fg.flush_pending()


fg.relaunch()
pend = fg._pending
fg.flush_pending()
out["epilogue_launch_us"] = dev_us(finalize_only)
opt.step(calc_metrics=False)


def step_only():
    fg._pending = None          # measurement only (a BNNP_F_HYPER_POST epilogue must not ride on the next launch)
    fg.relaunch()


out["step_kernel_us"] = dev_us(step_only)
fg._pending = None
out["host_prepass_call_us"] = host_us(lambda: fg.hyper_prepass(inv_n))
out["host_step_call_us"] = host_us(lambda: opt.step(calc_metrics=False))
out["host_p_version_us"] = host_us(fg._p_version)
out["host_sync_views_us"] = host_us(lambda: fg.sync_views(True))
out["device_step_call_us"] = dev_us(lambda: opt.step(calc_metrics=False))
print(json.dumps(out))
