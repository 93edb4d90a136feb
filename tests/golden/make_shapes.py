#!/usr/bin/env python
"""Segment tables (tensor shapes + prior kind/loc/scale/df per parameter) of the
models the BASELINE configs name, taken from the unmodified reference's
exp_utils.get_model (exp_utils.py:108-232).  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_shapes.py

Writes tests/golden/model_shapes.json (used by bench.py and the size tests; the
GPU box has no /root/reference).
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("BNNP_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REFERENCE)

from bnn_priors import exp_utils, prior as ref_prior  # noqa: E402

KIND = {"Normal": 1, "Laplace": 2, "StudentT": 3}


def table(model):
    prior_of = {id(pm.p): pm for _, pm in ref_prior.named_priors(model)}
    rows = []
    for name, p in model.named_parameters():
        pm = prior_of.get(id(p))
        row = dict(name=name, shape=list(p.shape), kind=0, loc=0.0, scale=1.0, df=3.0)
        if pm is not None:
            row["prior_class"] = type(pm).__name__
            if type(pm).__name__ in KIND:
                row.update(kind=KIND[type(pm).__name__], loc=float(pm.loc), scale=float(pm.scale))
                if hasattr(pm, "df"):
                    row["df"] = float(pm.df)
        rows.append(row)
    return rows


def main():
    torch.manual_seed(0)
    out = {}
    x_mnist, y = torch.rand(16, 784), torch.arange(16) % 10
    x_mnist_img = torch.rand(16, 1, 28, 28)
    x_cifar = torch.randn(16, 3, 32, 32)
    # experiments/train_bnn.py:55-65,103 defaults
    common = dict(width=50, depth=3, weight_loc=0., weight_scale=2.**0.5, bias_loc=0., bias_scale=1.,
                  batchnorm=True, weight_prior_params={}, bias_prior_params={})
    cases = {
        "classificationdensenet_mnist_gaussian": dict(x_train=x_mnist, y_train=y, model="classificationdensenet",
                                                      weight_prior="gaussian", bias_prior="gaussian", **common),
        "classificationconvnet_mnist_laplace": dict(x_train=x_mnist_img, y_train=y, model="classificationconvnet",
                                                    weight_prior="laplace", bias_prior="gaussian", **common),
        "googleresnet_cifar10_studentt": dict(x_train=x_cifar, y_train=y, model="googleresnet",
                                              weight_prior="student-t", bias_prior="gaussian", **common),
        "googleresnet_cifar10_gaussian": dict(x_train=x_cifar, y_train=y, model="googleresnet",
                                              weight_prior="gaussian", bias_prior="gaussian", **common),
        "vwidth_resnet18_w96_cifar10_gaussian": dict(x_train=x_cifar, y_train=y, model="vwidth_resnet18",
                                                     weight_prior="gaussian", bias_prior="gaussian",
                                                     **{**common, "width": 96}),
    }
    for tag, kw in cases.items():
        m = exp_utils.get_model(**kw)
        rows = table(m)
        n = sum(int(torch.Size(r["shape"]).numel()) for r in rows)
        out[tag] = dict(n_params=n, tensors=rows)
        print(tag, len(rows), "tensors", n, "params", {r.get("prior_class", "-") for r in rows})
    with open(os.path.join(HERE, "model_shapes.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
