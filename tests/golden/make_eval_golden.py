#!/usr/bin/env python
"""Golden vectors for the test-set evaluation: run the UNMODIFIED
reference `evaluate_model` (bnn_priors/exp_utils.py:250-340) on the reference's own
ClassificationDenseNet / DenseNet with a few random "posterior samples", on CPU.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_eval_golden.py

Stores in tests/golden/eval.npz, per case (cls = Categorical, reg = Normal):
    x, y                 the test set (N = 203 points, batch 64 -> ragged last batch)
    sample.<key>         [E, ...] every state_dict entry of every sample
    acc_data, lps        [E, N, C] / [E, N] what the reference accumulates (recomputed
                         here with the same model calls, float32)
    results              the dict the reference returned (as a JSON string)
Build container only (the GPU box has no /root/reference).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("BNNP_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REFERENCE)

from bnn_priors import exp_utils  # noqa: E402
from bnn_priors.models import ClassificationDenseNet, DenseNet  # noqa: E402

N, E, BATCH = 203, 4, 64


def run_case(tag, model, x, y, out):
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BATCH, shuffle=False)
    samples = {}
    for e in range(E):
        model.sample_all_priors()
        for k, v in model.state_dict().items():
            samples.setdefault(k, []).append(v.detach().clone())
    samples = {k: torch.stack(v) for k, v in samples.items()}
    res = exp_utils.evaluate_model(model, loader, samples, likelihood_eval=True, accuracy_eval=True,
                                   calibration_eval=False)
    res1 = exp_utils.evaluate_model(model, loader, {k: v[-1:] for k, v in samples.items()}, likelihood_eval=True,
                                    accuracy_eval=True, calibration_eval=False)
    acc_data, lps = [], []
    with torch.no_grad():
        for e in range(E):
            model.load_state_dict({k: v[e] for k, v in samples.items()})
            a, l = [], []
            for bx, by in loader:
                preds = model(bx)
                if isinstance(preds, torch.distributions.Categorical):
                    a.append(preds.logits)
                    l.append(preds.log_prob(by))
                else:
                    a.append(preds.mean)
                    l.append(preds.log_prob(by).sum(-1))
            acc_data.append(torch.cat(a))
            lps.append(torch.cat(l))
    out[f"{tag}.x"] = x.numpy()
    out[f"{tag}.y"] = y.numpy()
    out[f"{tag}.acc_data"] = torch.stack(acc_data).numpy()
    out[f"{tag}.lps"] = torch.stack(lps).numpy()
    out[f"{tag}.results"] = np.array(json.dumps(res))
    out[f"{tag}.results_last_only"] = np.array(json.dumps(res1))
    for k, v in samples.items():
        out[f"{tag}.sample.{k}"] = v.numpy()
    print(tag, res, res1)


def main():
    torch.manual_seed(7)
    out = {}
    x = torch.rand(N, 20)
    y = torch.randint(0, 7, (N,))
    y[:7] = torch.arange(7)                       # every class present (exp_utils.py:263-264)
    run_case("cls", ClassificationDenseNet(20, 7, 16, depth=3, softmax_temp=1.0), x, y, out)
    xr = torch.randn(N, 6)
    yr = torch.randn(N, 3)
    run_case("reg", DenseNet(6, 3, 8, depth=3, noise_std=0.7), xr, yr, out)
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)


if __name__ == "__main__" and not os.environ.get("CALIBRATION"):
    main()


def calibration_case():
    """Second fixture (tests/golden/eval_calibration.npz): the same classification case evaluated by the
    reference with calibration_eval=True -- ensemble probabilities (recomputed like exp_utils.py:309-324)
    and the ece / ace / rmsce values of the reference's third_party/calibration_error.py."""
    torch.manual_seed(7)
    x = torch.rand(N, 20)
    y = torch.randint(0, 7, (N,))
    y[:7] = torch.arange(7)
    model = ClassificationDenseNet(20, 7, 16, depth=3, softmax_temp=1.0)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=BATCH, shuffle=False)
    samples = {}
    for e in range(E):
        model.sample_all_priors()
        for k, v in model.state_dict().items():
            samples.setdefault(k, []).append(v.detach().clone())
    samples = {k: torch.stack(v) for k, v in samples.items()}
    res = exp_utils.evaluate_model(model, loader, samples, likelihood_eval=True, accuracy_eval=True,
                                   calibration_eval=True)
    acc = []
    with torch.no_grad():
        for e in range(E):
            model.load_state_dict({k: v[e] for k, v in samples.items()})
            acc.append(torch.cat([model(bx).logits for bx, _ in loader]))
    acc = torch.stack(acc).double()
    ens = torch.distributions.Categorical(logits=acc.logsumexp(0) - np.log(E))
    np.savez_compressed(os.path.join(HERE, "eval_calibration.npz"), acc_data=acc.float().numpy(), y=y.numpy(),
                        probs_mean=ens.probs.numpy(), results=np.array(json.dumps(res)))
    print("calibration", res)


if __name__ == "__main__" and os.environ.get("CALIBRATION"):
    calibration_case()
