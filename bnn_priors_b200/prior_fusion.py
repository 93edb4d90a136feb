"""Fuse the elementwise Normal / Laplace / Student-t prior into the sampler kernel.

In the reference the prior reaches the sampler through autograd: every minibatch
the runner evaluates `model.log_prior()` = sum of `Prior.log_prob()`
(models/base.py:25-30, prior/base.py:57-58), forms `potential = loss - log_prior/N`
(models/base.py:72-77) and backpropagates it (inference.py:218), which costs a
separate autograd graph and ~6 small kernels forward + backward per Prior module.

`fuse_prior(model, sampler)` removes that work for the priors the kernel knows in
closed form:

  * the segment table of the sampler gets (kind, loc, scale, df) per tensor and
    the step kernel adds  -(1/N) dlog p/dtheta  to the likelihood gradient
    in-register (and re-applies the runner's +-grad_max clamp to the sum);
  * `model.log_prior` is replaced by a function that returns the sum the kernel
    reduced while it updated the parameters (one extra read-only launch only if
    somebody changed the parameters since).  The value carries a do-nothing
    grad_fn so that the runners' `.backward()` calls keep working
    (inference_reject.py:20-22).

Priors that are not plain Normal / Laplace / StudentT with constant loc / scale / df
(hierarchical, empirical-Bayes, mixtures, correlated, Improper, ...) are left alone:
their log_prob stays in autograd and their gradient arrives in p.grad as before.
`p.grad` of a fused tensor holds the LIKELIHOOD gradient only.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributions as td

from . import _native as N

_KIND_OF_DIST = {td.Normal: N.PRIOR_NORMAL, td.Laplace: N.PRIOR_LAPLACE, td.StudentT: N.PRIOR_STUDENT_T}


def _is_prior_module(m: torch.nn.Module) -> bool:
    return isinstance(getattr(m, "p", None), torch.nn.Parameter) and hasattr(m, "log_prob") \
        and hasattr(m, "kwargs_keys")


def describe_prior(m: torch.nn.Module) -> Optional[tuple]:
    """(kind, loc, scale, df) if `m` is a prior the kernel can evaluate exactly like
    the reference does, else None.  Requirements: the class's `_dist` is
    torch.distributions.{Normal, Laplace, StudentT}; `log_prob` is the base
    `Prior.log_prob` (prior/base.py:57-58), not an override; loc / scale / df are
    one-element buffers (constants), not Parameters or Prior modules."""
    kind = _KIND_OF_DIST.get(getattr(type(m), "_dist", None))
    if kind is None:
        return None
    definer = next((k for k in type(m).__mro__ if "log_prob" in k.__dict__), None)
    if definer is None or definer.__name__ != "Prior":
        return None
    keys = set(getattr(m, "kwargs_keys", ()))
    want = {"loc", "scale"} | ({"df"} if kind == N.PRIOR_STUDENT_T else set())
    if keys != want:
        return None
    vals = {}
    for k in want:
        v = m._buffers.get(k)
        if not isinstance(v, torch.Tensor) or v.numel() != 1 or v.requires_grad:
            return None
        vals[k] = float(v)
    if not vals["scale"] > 0:
        return None
    return kind, vals["loc"], vals["scale"], vals.get("df", 3.0)


class _EngineValue(torch.autograd.Function):
    """A number computed outside autograd, attached to the graph through an anchor
    so that `.backward()` on it (or on anything built from it) is legal and free."""

    @staticmethod
    def forward(ctx, anchor, value):
        return value.clone()

    @staticmethod
    def backward(ctx, grad):
        return None, None


class FusedPrior:
    """Handle returned by `fuse_prior`; `.unfuse()` restores the model."""

    def __init__(self, model, sampler, grad_max: Optional[float]):
        self.model, self.sampler = model, sampler
        self.fused_modules: List[torch.nn.Module] = []
        self.other_modules: List[torch.nn.Module] = []
        self._had_attr = "log_prior" in model.__dict__
        self._old_attr = model.__dict__.get("log_prior")
        where: Dict[int, tuple] = {}
        for gi, fg in enumerate(sampler.flat_groups):
            for i, p in enumerate(fg.params):
                where[id(p)] = (gi, i)
        self.groups = set()
        for _, m in model.named_modules():
            if not _is_prior_module(m):
                continue
            spec = describe_prior(m)
            if spec is None or id(m.p) not in where:
                self.other_modules.append(m)
                continue
            gi, i = where[id(m.p)]
            fg = sampler.flat_groups[gi]
            fg.set_prior(i, *spec)
            self.groups.add(gi)
            self.fused_modules.append(m)
        for gi in self.groups:
            fg = sampler.flat_groups[gi]
            fg.prior_fused = True
            fg.grad_max = None if grad_max is None else float(grad_max)
            fg.invalidate_sums()
        dev = sampler.flat_groups[0].device
        self._anchor = torch.zeros((), device=dev, requires_grad=True)
        if self.fused_modules:
            model.log_prior = self.log_prior          # instance attribute shadows the method

    # the replacement of AbstractModel.log_prior (models/base.py:25-30)
    def log_prior(self) -> torch.Tensor:
        total = None
        for gi in sorted(self.groups):
            fg = self.sampler.flat_groups[gi]
            if not fg.log_prior_fresh():
                fg.sync_views(raise_on_no_grad=False)
                fg.reduce_now(1.0 / self.sampler.param_groups[gi]['num_data'])
            fg.flush_pending()          # the sum lives in the segment state once the last launch's epilogue ran
            v = fg.state_dev[:, N.S_LOG_PRIOR].sum()
            total = v if total is None else total + v
        lp = _EngineValue.apply(self._anchor, total.to(torch.float32))
        for m in self.other_modules:
            lp = lp + m.log_prob()
        return lp

    def unfuse(self) -> None:
        for gi in self.groups:
            fg = self.sampler.flat_groups[gi]
            for i in range(fg.nseg):
                fg.set_prior(i, N.PRIOR_NONE, 0.0, 1.0, 3.0)
            fg.prior_fused = False
            fg.grad_max = None
            fg.invalidate_sums()
        if self._had_attr:
            self.model.log_prior = self._old_attr
        elif "log_prior" in self.model.__dict__:
            del self.model.__dict__["log_prior"]


def fuse_prior(model: torch.nn.Module, sampler, grad_max: Optional[float] = None) -> FusedPrior:
    """Move the supported priors of `model` from autograd into `sampler`'s kernel.
    `grad_max` is the runner's gradient clamp (inference.py:219-220, default 1e6
    in experiments/train_bnn.py:84): the reference clamps likelihood + prior
    gradient together, so the kernel re-applies it to the fused sum."""
    return FusedPrior(model, sampler, grad_max)
