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

Covered: Normal, Laplace, StudentT (the north star's three) and, as the next row of the
scope table, Cauchy, GenNorm, LogNormal, Uniform, Improper and DoubleGamma
with constant hyper-parameters, plus Normal / Laplace / StudentT whose SCALE is a sampled
scalar with a Gamma, Uniform, HalfCauchy or improper hyper-prior (prior/hierarchical.py:
NormalGamma, NormalUniform, Horseshoe, LaplaceGamma, LaplaceUniform, StudentTGamma,
StudentTUniform; prior/empirical_bayes.py: NormalEmpirical, LaplaceEmpirical).  For those the
sampler runs a read-only pre-pass before a step (FlatGroup.hyper_prepass) that yields the
current scale, the log-prior and the hyper-parameter's gradient on the device.
Everything else (sampled df / beta, mixtures, correlated and multivariate priors) is left
alone: their log_prob stays in autograd and their gradient arrives in p.grad as before.
`p.grad` of a fused tensor holds the LIKELIHOOD gradient only.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import math

import torch
import torch.distributions as td

from . import _native as N

_KIND_OF_DIST = {td.Normal: N.PRIOR_NORMAL, td.Laplace: N.PRIOR_LAPLACE, td.StudentT: N.PRIOR_STUDENT_T,
                 td.Cauchy: N.PRIOR_CAUCHY}
# name of the third hyper-parameter (BnnpSegment.prior_df) per kind
_THIRD = {N.PRIOR_STUDENT_T: "df", N.PRIOR_GENNORM: "beta", N.PRIOR_DOUBLE_GAMMA: "concentration"}


def _is_prior_module(m: torch.nn.Module) -> bool:
    return isinstance(getattr(m, "p", None), torch.nn.Parameter) and hasattr(m, "log_prob") \
        and hasattr(m, "kwargs_keys")


def _constant(m, key):
    "value of a one-element, non-trainable buffer `key` of `m`, else None"
    v = m._buffers.get(key)
    if not isinstance(v, torch.Tensor) or v.numel() != 1 or v.requires_grad:
        return None
    return float(v)


def describe_prior(m: torch.nn.Module) -> Optional[tuple]:
    """(kind, loc, scale, third) if `m` is a prior the kernel can evaluate exactly like
    the reference does, else None.  Recognised (reference classes, prior/loc_scale.py and
    prior/transformed.py): Normal, Laplace, StudentT, Cauchy, GenNorm -- classes that use
    the base `Prior.log_prob` (prior/base.py:57-58) with the matching `_dist` -- and
    LogNormal, Improper / PositiveImproper, Uniform, DoubleGamma, which override `log_prob`
    and are recognised by the overriding class.  All hyper-parameters must be one-element
    buffers (constants), not Parameters or Prior modules.  `fuse_prior` additionally checks
    every recognised module numerically against its own `log_prob` + autograd."""
    cls = type(m)
    definer = next((k for k in cls.__mro__ if "log_prob" in k.__dict__), None)
    if definer is None:
        return None
    dist = getattr(cls, "_dist", None)
    keys = set(getattr(m, "kwargs_keys", ()))
    name = definer.__name__
    kind = None
    if name == "Prior":
        kind = _KIND_OF_DIST.get(dist)
        if kind is None and getattr(dist, "__name__", "") == "GeneralizedNormal":
            kind = N.PRIOR_GENNORM
    elif name == "LogNormal" and dist is td.Normal:
        kind = N.PRIOR_LOGNORMAL
    elif name == "Improper":
        return N.PRIOR_IMPROPER, 0.0, 1.0, 3.0
    elif name == "Uniform" and dist is td.Uniform:
        lo, hi = _constant(m, "low"), _constant(m, "high")
        if keys != {"low", "high"} or lo is None or hi is None or not hi > lo:
            return None
        return N.PRIOR_UNIFORM, lo, hi - lo, 3.0
    elif name == "DoubleGamma":
        kind = N.PRIOR_DOUBLE_GAMMA
    if kind is None:
        return None
    third = _THIRD.get(kind)
    if keys != {"loc", "scale"} | ({third} if third else set()):
        return None
    loc, scale = _constant(m, "loc"), _constant(m, "scale")
    t = _constant(m, third) if third else 3.0
    if loc is None or scale is None or t is None or not scale > 0 or not t > 0:
        return None
    return kind, loc, scale, t


_HYPER_TARGETS = {td.Normal: N.PRIOR_NORMAL, td.Laplace: N.PRIOR_LAPLACE, td.StudentT: N.PRIOR_STUDENT_T}


def _definer(cls, name):
    return next((k.__name__ for k in cls.__mro__ if name in k.__dict__), None)


def describe_hyper(h: torch.nn.Module) -> Optional[tuple]:
    """(hyper kind, a, b) if `h` is a scalar scale prior the kernel knows: the reference's
    Gamma / HalfCauchy (softplus-transformed, prior/transformed.py:50-80), Uniform (Gaussian
    CDF-transformed, :12-47) or PositiveImproper (prior/loc_scale.py:100-103); else None."""
    if not _is_prior_module(h) or h.p.numel() != 1:
        return None
    cls, keys = type(h), set(h.kwargs_keys)
    lp_def, fw_def = _definer(cls, "log_prob"), _definer(cls, "forward")
    dist = getattr(cls, "_dist", None)
    if lp_def == "Gamma" and fw_def == "Gamma" and dist is td.Gamma and keys == {"concentration", "rate"}:
        a, b = _constant(h, "concentration"), _constant(h, "rate")
        return (N.PRIOR_HYPER_GAMMA, a, b) if a and b and a > 0 and b > 0 else None
    if lp_def == "Uniform" and fw_def == "Uniform" and dist is td.Uniform and keys == {"low", "high"}:
        lo, hi = _constant(h, "low"), _constant(h, "high")
        return (N.PRIOR_HYPER_UNIFORM, lo, hi - lo) if lo is not None and hi is not None and hi > lo >= 0 else None
    if lp_def == "HalfCauchy" and fw_def == "HalfCauchy" and dist is td.HalfCauchy and keys == {"scale"}:
        a, mult = _constant(h, "scale"), getattr(h, "multiplier", None)
        ok = a is not None and a > 0 and isinstance(mult, (int, float)) and mult > 0
        return (N.PRIOR_HYPER_HALFCAUCHY, a, float(mult)) if ok else None
    if lp_def == "Improper" and fw_def == "PositiveImproper":
        return N.PRIOR_HYPER_IMPROPER, 0.0, 1.0
    return None


def describe_hier_prior(m: torch.nn.Module) -> Optional[tuple]:
    """(kind, loc, df, hyper module, hyper kind, a, b) if `m` is a Normal / Laplace / StudentT
    prior (base `Prior.log_prob`) whose `scale` is a recognised scalar prior module and whose
    other hyper-parameters are constants; else None."""
    cls = type(m)
    if _definer(cls, "log_prob") != "Prior":
        return None
    kind = _HYPER_TARGETS.get(getattr(cls, "_dist", None))
    keys = set(getattr(m, "kwargs_keys", ()))
    if kind is None or keys != {"loc", "scale"} | ({"df"} if kind == N.PRIOR_STUDENT_T else set()):
        return None
    h = m._modules.get("scale")
    loc = _constant(m, "loc")
    df = _constant(m, "df") if kind == N.PRIOR_STUDENT_T else 3.0
    hyper = describe_hyper(h) if h is not None else None
    if hyper is None or loc is None or df is None or not df > 0:
        return None
    return (kind, loc, df, h) + hyper


def hyper_closed_form(hkind: int, u: float, a: float, b: float):
    "(s, ds/du, log density of s, d/du of it) in float64: csrc/bnnp_kernels.cu hyper_scale / hyper_epilogue"
    if hkind == N.PRIOR_HYPER_UNIFORM:
        s = a + b * 0.5 * math.erfc(-u / math.sqrt(2.0))
        ds = b * math.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)
        return s, ds, -math.log(b), 0.0
    sp = u if u > 20.0 else math.log1p(math.exp(u))
    sg = 1.0 if u > 20.0 else 1.0 / (1.0 + math.exp(-u))
    if hkind == N.PRIOR_HYPER_GAMMA:
        return sp, sg, a * math.log(b) + (a - 1) * math.log(sp) - b * sp - math.lgamma(a), ((a - 1) / sp - b) * sg
    if hkind == N.PRIOR_HYPER_HALFCAUCHY:
        s, ds = sp * b, sg * b
        q = s / a
        return s, ds, math.log(2 / math.pi) - math.log(a) - math.log1p(q * q), -(2 * q / a) / (1 + q * q) * ds
    if hkind == N.PRIOR_HYPER_IMPROPER:
        return sp, sg, 0.0, 0.0
    raise ValueError(hkind)


def matches_hier_module(m: torch.nn.Module, spec: tuple, rtol: float = 1e-4) -> bool:
    """Do the kernel's closed forms reproduce log_prob() + scale.log_prob() of `m` and the
    autograd gradients w.r.t. the weights and the hyper-parameter at the current values?"""
    kind, loc, df, h, hkind, a, b = spec
    with torch.enable_grad():
        lp = m.log_prob() + h.log_prob()
        g_p, g_u = torch.autograd.grad(lp, [m.p, h.p], allow_unused=True)
    g_u = 0.0 if g_u is None else float(g_u)
    s, ds, lph, dlph = hyper_closed_form(hkind, float(h.p.detach()), a, b)
    if abs(float(h().detach()) - s) > rtol * s:
        return False
    want_lp, want_g = closed_form(kind, m.p, loc, s, df)
    d = m.p.detach().double() - loc
    n = d.numel()
    if kind == N.PRIOR_NORMAL:
        dl_ds = float((d * d).sum()) / s ** 3 - n / s
    elif kind == N.PRIOR_LAPLACE:
        dl_ds = float(d.abs().sum()) / s ** 2 - n / s
    else:
        dl_ds = (df + 1) * float((d * d / (df * s * s + d * d)).sum()) / s - n / s
    want_u = dl_ds * ds + dlph
    lp = float(lp.detach())
    if not (math.isfinite(lp) and math.isfinite(g_u) and bool(torch.isfinite(g_p).all())):
        return False
    if abs(lp - (float(want_lp) + lph)) > rtol * max(1.0, abs(lp)):
        return False
    if abs(g_u - want_u) > rtol * max(1.0, abs(g_u)):
        return False
    scale = float(g_p.double().abs().mean()) + 1e-30
    return bool(((g_p.double() - want_g).abs() <= rtol * (g_p.double().abs() + scale)).all())


def closed_form(kind: int, p: torch.Tensor, loc: float, scale: float, third: float):
    """(sum of log density, d log density / d p) in float64 -- the formulas the kernel
    implements (csrc/bnnp_kernels.cu: prior_grad_term, log_prior_term, log_prior_const),
    used to check a recognised module against its own log_prob before it is fused."""
    p = p.detach().double()
    d = p - loc
    z = d / scale
    n = p.numel()
    if kind == N.PRIOR_NORMAL or kind == N.PRIOR_LOGNORMAL:
        lp = (-0.5 * z * z).sum() - n * (math.log(scale) + 0.5 * math.log(2 * math.pi))
        g = -d / scale ** 2
        if kind == N.PRIOR_LOGNORMAL:
            lp, g = lp - p.sum(), g - 1.0
    elif kind == N.PRIOR_LAPLACE:
        lp = -z.abs().sum() - n * math.log(2 * scale)
        g = -torch.sign(d) / scale
    elif kind in (N.PRIOR_STUDENT_T, N.PRIOR_CAUCHY):
        df = 1.0 if kind == N.PRIOR_CAUCHY else third
        lp = (-0.5 * (df + 1) * torch.log1p(z * z / df)).sum() - n * (
            math.log(scale) + 0.5 * math.log(df) + 0.5 * math.log(math.pi)
            + math.lgamma(0.5 * df) - math.lgamma(0.5 * (df + 1)))
        g = -(df + 1) * d / (df * scale ** 2 + d * d)
    elif kind == N.PRIOR_GENNORM:
        lp = -(z.abs() ** third).sum() + n * (-math.log(2 * scale) - math.lgamma(1 / third) + math.log(third))
        g = -third * z.abs() ** (third - 1) * torch.sign(d) / scale
    elif kind == N.PRIOR_UNIFORM:
        lp, g = torch.tensor(-n * math.log(scale), dtype=torch.float64, device=p.device), torch.zeros_like(p)
    elif kind == N.PRIOR_IMPROPER:
        lp, g = torch.zeros((), dtype=torch.float64, device=p.device), torch.zeros_like(p)
    elif kind == N.PRIOR_DOUBLE_GAMMA:
        lp = ((third - 1) * d.abs().log() - z.abs()).sum() + n * (
            -third * math.log(scale) - math.lgamma(third) - math.log(2))
        g = (third - 1) / d - torch.sign(d) / scale
    else:
        raise ValueError(kind)
    return lp, g


def matches_module(m: torch.nn.Module, spec: tuple, rtol: float = 1e-4) -> bool:
    """Does the kernel's closed form reproduce `m.log_prob()` and its autograd gradient at
    the current parameter value?  (Guards the class-name based recognition.)"""
    with torch.enable_grad():
        lp = m.log_prob()
        if isinstance(lp, torch.Tensor) and lp.requires_grad:
            (g,) = torch.autograd.grad(lp, m.p, allow_unused=True)
        else:
            g = None
    g = torch.zeros_like(m.p) if g is None else g
    want_lp, want_g = closed_form(spec[0], m.p, *spec[1:])
    lp = float(lp.detach()) if isinstance(lp, torch.Tensor) else float(lp)
    if not (math.isfinite(lp) and bool(torch.isfinite(g).all())):
        return False
    if abs(lp - float(want_lp)) > rtol * max(1.0, abs(lp)):
        return False
    scale = float(g.double().abs().mean()) + 1e-30
    return bool(((g.double() - want_g).abs() <= rtol * (g.double().abs() + scale)).all())


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

    def __init__(self, model, sampler, grad_max: Optional[float], verify: bool = True, clamp: str = "always"):
        if clamp not in ("always", "armed"):
            raise ValueError("clamp must be 'always' or 'armed'")
        self.model, self.sampler = model, sampler
        self.clamp_mode = clamp
        self.fused_modules: List[torch.nn.Module] = []
        self.other_modules: List[torch.nn.Module] = []
        self._had_attr = "log_prior" in model.__dict__
        self._old_attr = model.__dict__.get("log_prior")
        where: Dict[int, tuple] = {}
        for gi, fg in enumerate(sampler.flat_groups):
            for i, p in enumerate(fg.params):
                where[id(p)] = (gi, i)
        self.groups = set()
        hyper_done = set()           # scale priors fused together with their parent (visited after it)
        for _, m in model.named_modules():
            if not _is_prior_module(m) or id(m) in hyper_done:
                continue
            hier = describe_hier_prior(m)
            if hier is not None and id(m.p) in where and id(hier[3].p) in where \
                    and where[id(m.p)][0] == where[id(hier[3].p)][0] and (not verify or matches_hier_module(m, hier)):
                kind, loc, df, h, hkind, a, b = hier
                gi, i = where[id(m.p)]
                fg = sampler.flat_groups[gi]
                fg.set_prior(i, kind, loc, float(h().detach()), df)
                fg.set_hyper_link(i, where[id(h.p)][1], hkind, a, b)
                self.groups.add(gi)
                self.fused_modules += [m, h]
                hyper_done.add(id(h))
                continue
            spec = describe_prior(m)
            if spec is not None and verify and not matches_module(m, spec):
                spec = None                   # same name, different density: leave it to autograd
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
            fg.clamp_armed = clamp == "always"
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
            inv_n = 1.0 / self.sampler.param_groups[gi]['num_data']
            if fg.has_hyper:
                if not (fg.hyper_fresh() and fg.log_prior_fresh()):
                    fg.sync_views(raise_on_no_grad=False)
                    fg.hyper_prepass(inv_n)       # scales, log-prior and hyper gradients in one read of P
            elif not fg.log_prior_fresh():
                fg.sync_views(raise_on_no_grad=False)
                fg.reduce_log_prior(inv_n)      # reads P only: legal while p.grad is None
            fg.flush_pending()          # the sum lives in the segment state once the last launch's epilogue ran
            v = fg.state_dev[:, N.S_LOG_PRIOR].sum()
            total = v if total is None else total + v
        lp = _EngineValue.apply(self._anchor, total.to(torch.float32))
        for m in self.other_modules:
            lp = lp + m.log_prob()
        return lp

    def arm_clamp(self, on: bool) -> None:
        """clamp="armed": whether the +-grad_max clamp applies to the gradient now in p.grad.  The
        overlay arms it for minibatch gradients (the runner clamps those, inference.py:219-220) and
        disarms it for the full-data gradient (inference_reject.py:18-33 does not clamp)."""
        if self.clamp_mode != "armed":
            return
        for gi in self.groups:
            fg = self.sampler.flat_groups[gi]
            if fg.clamp_armed != bool(on):
                fg.clamp_armed = bool(on)
                fg.invalidate_sums()

    def unfuse(self) -> None:
        for gi in self.groups:
            fg = self.sampler.flat_groups[gi]
            for i in range(fg.nseg):
                fg.set_prior(i, N.PRIOR_NONE, 0.0, 1.0, 3.0)
            fg.clear_hyper_links()
            fg.prior_fused = False
            fg.grad_max = None
            fg.clamp_armed = True
            fg.invalidate_sums()
        if self._had_attr:
            self.model.log_prior = self._old_attr
        elif "log_prior" in self.model.__dict__:
            del self.model.__dict__["log_prior"]


def fuse_prior(model: torch.nn.Module, sampler, grad_max: Optional[float] = None, verify: bool = True,
               clamp: str = "always") -> FusedPrior:
    """Move the supported priors of `model` from autograd into `sampler`'s kernel.
    `grad_max` is the runner's gradient clamp (inference.py:219-220, default 1e6 in
    inference.py:13): the reference clamps likelihood + prior gradient together, so the kernel
    applies it to the fused sum.  With an unchanged runner that still clamps p.grad (now the
    likelihood part alone) first, the two differ when |likelihood gradient| > grad_max; the
    overlay (`overlay.install(fuse_prior=True)`) removes the runner-side clamp and uses
    clamp="armed" -- the kernel clamps only gradients the runner would have clamped -- which is
    the reference's arithmetic exactly."""
    return FusedPrior(model, sampler, grad_max, verify, clamp)
