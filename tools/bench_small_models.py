#!/usr/bin/env python
"""Per-step latency of the sampler API on the small BASELINE configs (launch-bound:
the whole chain fits in L2), host-side cost included.  GPU box only.

    python tools/bench_small_models.py
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
K = 2000
for tag, smp in (("classificationdensenet_mnist_gaussian", "VerletSGLD"),
                 ("classificationconvnet_mnist_laplace", "VerletSGLD"),
                 ("googleresnet_cifar10_studentt", "VerletSGLD"),
                 ("googleresnet_cifar10_gaussian", "HMC")):
    for fused in (False, True):
        opt, params, fg = bench.make_chain(dev, 0, smp, tag=tag, fused_prior=fused)
        step = lambda: opt.step(calc_metrics=False)   # noqa: E731
        for _ in range(200):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        host_us = (time.perf_counter() - t0) / K * 1e6          # enqueue cost (host only)
        torch.cuda.synchronize()
        wall_us = (time.perf_counter() - t0) / K * 1e6
        ms = bench.timed_gpu(fg.relaunch, K, dev, False) / K
        t0 = time.perf_counter()
        for _ in range(200):
            opt.step(calc_metrics=True)
            _ = opt.state[params[0]]["est_temperature"]
        metrics_us = (time.perf_counter() - t0) / 200 * 1e6
        out[f"{tag}/{smp}{'+fused' if fused else ''}"] = dict(
            params=fg.n_params, tensors=fg.nseg, ctas=fg.nchunks,
            api_host_us_per_step=round(host_us, 1), api_wall_us_per_step=round(wall_us, 1),
            kernel_back_to_back_us=round(ms * 1e3, 2), metrics_step_with_readback_us=round(metrics_us, 1))
        del opt, params, fg
print(json.dumps(out, indent=1))
