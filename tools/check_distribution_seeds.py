#!/usr/bin/env python
"""The statistical checks of tests/test_cuda_reference_suite.py (the reference's distribution-
preservation tests, testing/test_verlet_sgld.py:58-146) under several in-kernel noise streams:
prints the KS p-values per seed.  A sampler that preserves the target gives p-values spread over
(0, 1); a biased one piles them up at 0.  GPU box only."""
import json
import math
import os
import sys

import numpy as np
import scipy.stats
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_cuda_reference_suite as T  # noqa: E402
from bnn_priors_b200.mcmc import HMC, VerletSGLD  # noqa: E402

out = []
for seed in range(int(os.environ.get("SEEDS", "10"))):
    row = {"seed": seed}
    for name in ("verlet", "hmc"):
        torch.manual_seed(1000 + seed)
        mean, std = 1., 2.
        temperature = 0.75 if name == "verlet" else 1.0
        model = T.gaussian_model(50, 1000, mean, std, temperature)
        if name == "verlet":
            s = VerletSGLD(model.parameters(), lr=3 / 8, num_data=1, momentum=0.9, temperature=temperature, seed=seed)
        else:
            s = HMC(model.parameters(), lr=1 / 4, num_data=1, seed=seed)
        for _, state in s.state.items():
            state['preconditioner'] = (torch.rand(()).item() + 0.2) / 2
        if name == "verlet":
            s.sample_momentum()
        acc, n_rej = T._mh_loop(s, model, 200, 4, hmc=(name == "hmc"))
        parameters, kinetic, config = T._collect(s, 50, 1000)
        sd = std * temperature ** .5
        p_par = scipy.stats.ks_1samp(parameters, lambda x: scipy.stats.norm.cdf(x, loc=mean, scale=sd))[1]
        chi = lambda x: scipy.stats.chi2.cdf(x, df=1000, loc=0., scale=temperature / 1000)  # noqa: E731
        row[name] = dict(acc=round(acc, 3), rejected=n_rej, p_params=round(p_par, 4),
                         p_config=round(scipy.stats.ks_1samp(config, chi)[1], 4),
                         p_kinetic=round(scipy.stats.ks_1samp(kinetic, chi)[1], 4),
                         std_ratio=round(float(parameters.std() / sd), 4),
                         config_mean=round(float(config.mean() / temperature), 4),
                         kinetic_mean=round(float(kinetic.mean() / temperature), 4))
    out.append(row)
    print(json.dumps(row), flush=True)
ps = np.array([r[n][k] for r in out for n in ("verlet", "hmc") for k in ("p_params", "p_config", "p_kinetic")])
print(json.dumps({"n_pvalues": int(ps.size), "below_0.01": int((ps < 0.01).sum()), "below_0.05": int((ps < 0.05).sum()),
                  "median": float(np.median(ps)), "ks_uniform_p": float(scipy.stats.kstest(ps, "uniform")[1])}))
