#!/usr/bin/env python
"""Wall-clock of the reference's OWN training script (experiments/train_bnn.py `main`, unmodified, from the
oracle/_ref snapshot; synthetic data of the data set's shape) on one GPU: stock reference code vs the same
code with `overlay.install(...)`.  GPU box only (needs oracle/_ref: `python oracle/make_ref.py` in the build
container).

    python tools/bench_runner.py [config ...]        # default: the four BASELINE configs
"""
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import runner_harness as H  # noqa: E402
from bnn_priors_b200 import overlay  # noqa: E402

CONFIGS = {
    "cfg2_densenet_verlet_reject_gaussian": dict(inference="VerletSGLDReject", model="classificationdensenet", data="mnist",
                                                 weight_prior="gaussian"),
    "cfg3_convnet_verlet_reject_laplace_T0.1": dict(inference="VerletSGLDReject", model="classificationconvnet", data="mnist",
                                                    weight_prior="laplace", temperature=0.1),
    "cfg4_googleresnet_verlet_reject_studentt": dict(inference="VerletSGLDReject", model="googleresnet", data="cifar10",
                                                     weight_prior="student-t"),
    "cfg5_googleresnet_hmc_reject_gaussian": dict(inference="HMCReject", model="googleresnet", data="cifar10",
                                                  weight_prior="gaussian", momentum=1.0),
}
# 50 minibatches per epoch (HMC: 50 leapfrog steps), 2 cycles x (1 warm-up + 2 sampling epochs)
RUN = dict(n_train=6400, n_test=1024, n_samples=4, cycles=2, warmup=1, burnin=0, metrics_skip=10, reject_samples=True)

out = {}
for name in (sys.argv[1:] or list(CONFIGS)):
    cfg = dict(CONFIGS[name], **RUN)
    rec = {}
    for label, kw in (("reference", None),
                      ("overlay", dict()),
                      ("overlay_all", dict(evaluate=True, fuse_prior=True, sample_sink=True))):
        times = []
        for rep in range(2):                      # the first run pays cuDNN autotuning and module loading
            d = tempfile.mkdtemp()
            if kw is not None:
                overlay.install(**kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                run, _ = H.run_train_bnn(d, **cfg)
            finally:
                if kw is not None:
                    overlay.uninstall()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        rec[label] = {"seconds": round(times[-1], 3), "first_run_seconds": round(times[0], 3),
                      "result": {k: round(v, 4) for k, v in run.result.items()}}
    steps = 2 * 3 * 50
    rec["sampler_steps"] = steps
    rec["speedup_overlay"] = round(rec["reference"]["seconds"] / rec["overlay"]["seconds"], 2)
    rec["speedup_overlay_all"] = round(rec["reference"]["seconds"] / rec["overlay_all"]["seconds"], 2)
    out[name] = rec
    print(json.dumps({name: rec}), flush=True)
with open(os.path.join(ROOT, "gpurun_out", "r02_runner_wallclock.json"), "w") as f:
    json.dump(out, f, indent=1)
