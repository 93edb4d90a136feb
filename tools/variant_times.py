#!/usr/bin/env python
"""Kernel time of the step-kernel variants for one build of libbnnp.so (BNNP_LIB picks it), back to back and
with the L2 evicted before every launch (the production regime).  GPU box only.

    for f in bnn_priors_b200/_lib/tune/*.so; do BNNP_LIB=$PWD/$f python tools/variant_times.py; done
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
K = int(os.environ.get("TUNE_STEPS", "60"))
peak, _ = bench.measured_peak()
flush = bench.L2Flush(dev)
out = {"lib": os.path.basename(os.environ.get("BNNP_LIB", "default"))}
cases = (("sgld", "SGLD", False, lambda o: o.step(calc_metrics=False)),
         ("sgld_metrics", "SGLD", False, lambda o: o.step(calc_metrics=True)),
         ("verlet", "VerletSGLD", False, lambda o: o.step(calc_metrics=False)),
         ("verlet_fused", "VerletSGLD", True, lambda o: o.step(calc_metrics=False)),
         ("verlet_fused_metrics", "VerletSGLD", True, lambda o: o.step(calc_metrics=True)),
         ("hmc", "HMC", False, lambda o: o.step(calc_metrics=False)),
         ("hmc_fused", "HMC", True, lambda o: o.step(calc_metrics=False)))
only = os.environ.get("TUNE_CASES")
for name, smp, fused, call in cases:
    if only and name not in only.split(","):
        continue
    opt, params, fg = bench.make_chain(dev, 0, smp, fused_prior=fused)
    call(opt)
    b2b, same, prod = bench.regime_numbers(fg, fg.n_params, K, dev, False, flush, peak)
    out[name] = {"b2b_us": round(b2b["kernel_us"], 2), "production_us": round(prod["kernel_us"], 2)}
    del opt, params, fg
    torch.cuda.empty_cache()
print(json.dumps(out))
