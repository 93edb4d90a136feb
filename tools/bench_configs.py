#!/usr/bin/env python
"""Application-level context for the small BASELINE configs (2-5): wall time of one training
iteration = minibatch forward/backward in stock torch autograd + one sampler transition, with

  ours         bnn_priors_b200.mcmc.VerletSGLD, Gaussian prior fused into the kernel
  eager        the reference's per-tensor torch op sequence for the same transition
               (oracle/sgmcmc_torch.py: ~9 kernels and 2 `.item()` syncs per tensor, 4 with
               diagnostics), prior gradient added with one extra op per tensor -- cheaper than the
               autograd graph through Prior.log_prob that the reference really runs

on (a) the 784-50-50-10 MLP of `classificationdensenet` (6 tensors, 42,310 parameters, synthetic
MNIST, bs 128) and (b) a ResNet-20 with BatchNorm of the size of `googleresnet` (synthetic
CIFAR-10, bs 128).  Diagnostics every 10th step like experiments/train_bnn.py:81 (metrics_skip).
This is NOT the bench metric (bench.py is); it shows what share of an iteration the sampler is.
GPU box only.    python tools/bench_configs.py
"""
import json
import math
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
from bnn_priors_b200 import _native as N, mcmc  # noqa: E402
from models import ResNet20  # noqa: E402
from oracle import sgmcmc_torch as OT  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True


def mlp():
    return nn.Sequential(nn.Linear(784, 50), nn.ReLU(), nn.Linear(50, 50), nn.ReLU(), nn.Linear(50, 10))


HP = dict(lr=1e-3, num_data=50000.0, momentum=0.994, temperature=1.0)
ITERS, WARM = 200, 30


def run(name, make_model, x, y):
    res = {}
    # ---------------------------------------------------------------- ours
    torch.manual_seed(0)
    model = make_model().to(dev)
    params = list(model.parameters())
    opt = mcmc.VerletSGLD(params, **HP, seed=0)
    (fg,) = opt.flat_groups
    for i, p in enumerate(params):
        if p.dim() > 1:                                        # weights: N(0, 2/fan_in)-style Gaussian prior, fused
            fg.set_prior(i, N.PRIOR_NORMAL, 0.0, math.sqrt(2.0 / p[0].numel()), 3.0)
    fg.prior_fused = True
    opt.sample_momentum()

    def it_ours(i):
        opt.zero_grad()
        loss = F.cross_entropy(model(x), y)
        loss.backward()
        opt.step(calc_metrics=(i % 10 == 0))
        if i % 10 == 0:
            _ = opt.state[params[0]]["est_temperature"]      # the runner reads the diagnostics (inference.py:262-294)

    # ---------------------------------------------------------------- eager restatement of the reference
    torch.manual_seed(0)
    model_e = make_model().to(dev)
    params_e = list(model_e.parameters())
    scales = [math.sqrt(2.0 / p[0].numel()) if p.dim() > 1 else None for p in params_e]

    class Chain(OT.TorchVerletChain):
        pass
    ch = Chain([p.detach() for p in params_e], **HP)
    ch.p = [p.data for p in params_e]                          # update the model's tensors in place
    ch.sample_momentum()

    def it_eager(i):
        for p in params_e:
            p.grad = None
        loss = F.cross_entropy(model_e(x), y)
        loss.backward()
        with torch.no_grad():
            for k, p in enumerate(params_e):
                if scales[k] is not None:                      # -(1/N) dlog N(p; 0, s)/dp = p / (N s^2)
                    p.grad.add_(p, alpha=1.0 / (HP["num_data"] * scales[k] ** 2))
                ch.g[k] = p.grad
        ch.step(calc_metrics=(i % 10 == 0))

    def it_fwd_bwd_only(i):
        for p in params_e:
            p.grad = None
        F.cross_entropy(model_e(x), y).backward()

    for label, fn in (("forward_backward_only", it_fwd_bwd_only), ("ours", it_ours), ("eager", it_eager)):
        for i in range(WARM):
            fn(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(ITERS):
            fn(i)
        torch.cuda.synchronize()
        res[label + "_ms_per_iteration"] = round((time.perf_counter() - t0) / ITERS * 1e3, 4)
    res["tensors"], res["params"] = len(params), sum(p.numel() for p in params)
    fb = res["forward_backward_only_ms_per_iteration"]
    res["sampler_share_ours"] = round(1 - fb / res["ours_ms_per_iteration"], 3)
    res["sampler_share_eager"] = round(1 - fb / res["eager_ms_per_iteration"], 3)
    res["iteration_speedup"] = round(res["eager_ms_per_iteration"] / res["ours_ms_per_iteration"], 2)
    return res


out = {}
g = torch.Generator(device=dev).manual_seed(0)
out["mlp_784_50_50_10_mnist_bs128"] = run("mlp", mlp, torch.rand(128, 784, device=dev, generator=g),
                                          torch.randint(0, 10, (128,), device=dev, generator=g))
out["resnet20_bn_cifar10_bs128"] = run("resnet20", ResNet20, torch.randn(128, 3, 32, 32, device=dev, generator=g),
                                       torch.randint(0, 10, (128,), device=dev, generator=g))
print(json.dumps(out, indent=1))
