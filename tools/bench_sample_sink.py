#!/usr/bin/env python
"""Time to store one posterior sample: the reference's way (exp_utils.py:426-431:
`.cpu().detach().unsqueeze(0).numpy()` per state_dict entry, synchronous) vs
FlatSampleSaver (one D2D snapshot + one async D2H).  GPU box only.

    python tools/bench_sample_sink.py
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bnn_priors_b200.sample_sink import FlatSampleSaver  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
for tag in ("googleresnet_cifar10_gaussian", "vwidth_resnet18_w96_cifar10_gaussian"):
    opt, params, fg = bench.make_chain(dev, 0, "VerletSGLD", tag=tag)
    tensors = bench.load_tensors(tag)
    sd = {t["name"]: p.detach() for t, p in zip(tensors, params)}
    n = 12
    # reference style: blocks the host until every tensor has arrived
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        d = {k: v.cpu().detach().unsqueeze(0).numpy() for k, v in sd.items()}
    ref_ms = (time.perf_counter() - t0) / n * 1e3
    saver = FlatSampleSaver(None, opt, capacity=n + 2)
    saver.add_state_dict(sd, 0)
    saver.add_state_dict(sd, 0)
    torch.cuda.synchronize()
    # (a) how long the chain (host thread + compute stream) is held up per sample
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        saver.add_state_dict(sd, i)
    e1.record()
    host_ms = (time.perf_counter() - t0) / n * 1e3
    e1.synchronize()
    stream_ms = e0.elapsed_time(e1) / n
    # (b) until the data is in host memory
    t0 = time.perf_counter()
    saver.flush(final=True)
    drain_ms = (time.perf_counter() - t0) * 1e3
    out[tag] = dict(entries=len(sd), params=fg.n_params, reference_ms_per_sample=round(ref_ms, 3),
                    flat_host_ms_per_sample=round(host_ms, 3), flat_compute_stream_ms_per_sample=round(stream_ms, 3),
                    flat_drain_ms_after_last=round(drain_ms, 3))
    del saver, opt, params, fg
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
