"""CPU restatement (numpy, float64) of the bookkeeping of the reference's
`evaluate_model` (bnn_priors/exp_utils.py:250-340) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module; the product path (bnn_priors_b200/evaluate.py -> csrc/bnnp_eval.cu) never does.

Pinned: tests/golden/eval.npz holds inputs and results recorded from the unmodified
reference (tests/golden/make_eval_golden.py); tests/test_oracle_golden.py checks this
restatement against them.

Inputs are what the reference accumulates on the CPU while it loops over samples and
test batches (exp_utils.py:266-297):
    acc_data [E, N, C]  preds.logits (Categorical) or preds.mean (Normal), as float64
    lps      [E, N]     preds.log_prob(y) (summed over the last axis for Normal)
"""
from __future__ import annotations

import math

import numpy as np

CATEGORICAL, NORMAL = 0, 1


def _logsumexp(a: np.ndarray, axis: int) -> np.ndarray:
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis) + np.log(np.sum(np.exp(a - m), axis=axis))


def _log_space_mean(a: np.ndarray, axis: int) -> np.ndarray:
    "exp_utils.py:299-300"
    return _logsumexp(a, axis) - math.log(a.shape[axis])


def evaluate(acc_data: np.ndarray, lps: np.ndarray, labels: np.ndarray, kind: int) -> dict:
    """exp_utils.py:302-338 without the calibration metrics.  Returns lp_ensemble,
    lp_last, acc_ensemble, acc_last (+ probs_mean and lp_ensemble_check for
    Categorical predictions)."""
    acc_data = np.asarray(acc_data, dtype=np.float64)
    lps = np.asarray(lps, dtype=np.float64)
    out = {}
    lps_each_model = lps.mean(1)                                   # :304
    out["lp_ensemble"] = float(_log_space_mean(lps, 0).mean())     # :305
    out["lp_last"] = float(lps_each_model[-1])
    if kind == CATEGORICAL:
        labels = np.asarray(labels, dtype=np.int64)
        idx = np.arange(labels.shape[0])
        ens = _log_space_mean(acc_data, 0)                         # :309
        ens = ens - _logsumexp(ens, 1)[:, None]                    # Categorical(logits=...) normalises
        last = acc_data[-1] - _logsumexp(acc_data[-1], 1)[:, None]
        out["lp_ensemble_check"] = float(ens[idx, labels].mean())  # :312-313
        out["acc_ensemble"] = float((np.argmax(ens, 1) == labels).astype(np.float32).mean())   # models/base.py:184-185
        out["acc_last"] = float((np.argmax(last, 1) == labels).astype(np.float32).mean())
        out["probs_mean"] = np.exp(ens)
    else:
        y = np.asarray(labels, dtype=np.float64).reshape(acc_data.shape[1], -1)
        d_ens = acc_data.mean(0) - y                               # :317
        d_last = acc_data[-1] - y                                  # :318
        out["acc_ensemble"] = float((d_ens * d_ens).sum(1).mean()) # models/base.py:155-158
        out["acc_last"] = float((d_last * d_last).sum(1).mean())
    return out
