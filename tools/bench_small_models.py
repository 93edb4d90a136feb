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
        # the reference runner's loop: zero_grad() drops the gradients, backward() hands over tensors of
        # their own (new objects, the addresses the allocator re-uses); host time of zero_grad + step only
        bufs = [torch.randn_like(p) * 1e-3 for p in params]
        t_zero = t_step = 0.0
        c0 = fg.copies
        for it in range(K + 100):
            a = time.perf_counter()
            opt.zero_grad()
            b = time.perf_counter()
            for p, g in zip(params, bufs):
                p.grad = g.view_as(g)
            c = time.perf_counter()
            opt.step(calc_metrics=False)
            d = time.perf_counter()
            if it >= 100:
                t_zero += b - a
                t_step += d - c
        torch.cuda.synchronize()
        rec = dict(
            params=fg.n_params, tensors=fg.nseg, ctas=fg.nchunks,
            api_host_us_per_step=round(host_us, 1), api_wall_us_per_step=round(wall_us, 1),
            kernel_back_to_back_us=round(ms * 1e3, 2), metrics_step_with_readback_us=round(metrics_us, 1),
            runner_loop_zero_grad_host_us=round(t_zero / K * 1e6, 1), runner_loop_step_host_us=round(t_step / K * 1e6, 1),
            runner_loop_gradient_copies=fg.copies - c0)
        del opt, params, fg, bufs
        # capturable mode: one step recorded in a CUDA graph, replayed
        opt, params, fg = bench.make_chain(dev, 0, smp, tag=tag, fused_prior=fused, capturable=True)
        for _ in range(5):
            opt.step(calc_metrics=False)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            opt.step(calc_metrics=False)
        for _ in range(50):
            graph.replay()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(K):
            graph.replay()
        rec["graph_replay_host_us"] = round((time.perf_counter() - t0) / K * 1e6, 1)
        torch.cuda.synchronize()
        rec["graph_replay_wall_us"] = round((time.perf_counter() - t0) / K * 1e6, 1)
        out[f"{tag}/{smp}{'+fused' if fused else ''}"] = rec
        del opt, params, fg, graph
print(json.dumps(out, indent=1))
