#!/usr/bin/env python
"""Per-epoch test evaluation: wall time of one `evaluate_model` call on a
10,000-point test set, batch 128, one sample = the live state (what the runners do after
every epoch, inference.py:199-213).

  reference way   the loop of exp_utils.py:266-338 restated: per batch two blocking
                  float64 copies to the CPU, [E, N, C] tensors and the reductions on the host
  device way      bnn_priors_b200.evaluate.evaluate_model (one launch per batch, five
                  doubles read back once)
GPU box only.    python tools/bench_evaluate.py
"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import local_models as LM  # noqa: E402
from bnn_priors_b200.evaluate import evaluate_model  # noqa: E402

dev = torch.device("cuda", 0)


def reference_way(model, loader, samples, labels):
    n = labels.shape[0]
    E = len(next(iter(samples.values())))
    c = int(labels.max()) + 1
    lps = torch.zeros((E, n), dtype=torch.float64)
    acc_data = torch.zeros((E, n, c), dtype=torch.float64)
    for e in range(E):
        with torch.no_grad():
            model.load_state_dict({k: v[e] for k, v in samples.items()})
            i = 0
            for bx, by in loader:
                preds = model(bx.to(dev))
                j = i + len(bx)
                lps[e, i:j] = preds.log_prob(by.to(dev)).detach()
                acc_data[e, i:j] = preds.logits.detach()
                i = j
    lse = lambda t, d: t.logsumexp(d) - math.log(t.size(d))   # noqa: E731
    ens = torch.distributions.Categorical(logits=lse(acc_data, 0))
    last = torch.distributions.Categorical(logits=acc_data[-1])
    return {"lp_ensemble": lse(lps, 0).mean().item(), "lp_last": lps.mean(1)[-1].item(),
            "acc_ensemble": ens.logits.argmax(1).eq(labels).float().mean().item(),
            "acc_last": last.logits.argmax(1).eq(labels).float().mean().item()}


out = {}
for name, din, width in (("densenet_mnist", 784, 50), ("wide_mlp", 784, 2048)):
    torch.manual_seed(0)
    model = LM.TinyClassifier(din, 10, width).to(dev)
    x = torch.rand(10000, din, device=dev)
    y = torch.randint(0, 10, (10000,), device=dev)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=128, shuffle=False)
    live = {k: v.unsqueeze(0) for k, v in model.state_dict().items()}
    labels_cpu = y.cpu()
    res = {}
    for label, fn in (("reference_way", lambda: reference_way(model, loader, live, labels_cpu)),
                      ("device_way", lambda: evaluate_model(model, loader, live, True, True, False))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            r = fn()
        torch.cuda.synchronize()
        res[label] = {"ms_per_call": round((time.perf_counter() - t0) / 5 * 1e3, 2), "result": r}
    a, b = res["reference_way"]["result"], res["device_way"]["result"]
    res["max_abs_diff"] = max(abs(a[k] - b[k]) for k in a)
    out[name] = res
print(json.dumps(out, indent=1))
