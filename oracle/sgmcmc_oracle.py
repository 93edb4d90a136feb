"""CPU oracle for the SG-MCMC sampler hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm in the reference's
``bnn_priors/mcmc/{sgld,verlet_sgld,hmc}.py`` and of the three elementwise
priors in ``bnn_priors/prior/loc_scale.py``.  It exists so that the CUDA path
in ``bnn_priors_b200`` has something independent to be checked against.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``bnn_priors_b200/`` imports it, and the product path has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` replays the call
traces in ``tests/golden/*.npz`` -- recorded from the unmodified reference
samplers (and from the reference's ``VerletSGLDRunnerReject`` driving them) by
``tests/golden/make_golden.py`` -- through this oracle, and
``tests/test_oracle_live_reference.py`` does the same against the reference
imported live whenever ``/root/reference`` is present.

Layout.  A chain is a list of *segments* (one per parameter tensor, in
``param_groups`` order) over flat fp32 arrays, plus one hyper-parameter group.
All elementwise arithmetic is fp32 (like the reference's tensors); every scalar
that the reference holds as a Python float is a Python float here.

Arithmetic that comes from a third party: ``torch.distributions.{Normal,
Laplace,StudentT}.log_prob`` (torch 2.11.0; the reference pins torch>=1.5,<1.6
in setup.py:14) -- restated in ``prior_log_prob`` from their published
closed forms; call sites prior/base.py:57-58, prior/loc_scale.py:34-35,66-67,
74-77.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

F32 = np.float32

PRIOR_NONE, PRIOR_NORMAL, PRIOR_LAPLACE, PRIOR_STUDENT_T = 0, 1, 2, 3
# the other elementwise priors with constant hyper-parameters
PRIOR_CAUCHY, PRIOR_GENNORM, PRIOR_LOGNORMAL, PRIOR_UNIFORM, PRIOR_IMPROPER, PRIOR_DOUBLE_GAMMA = 4, 5, 6, 7, 8, 9
# SURVEY 8f/N4, second half: scalar hyper-parameters with their own prior (prior/hierarchical.py,
# prior/empirical_bayes.py).  A hyper segment has one element u; its kind says how the scale of the
# linked weight segment follows from u and which density the scale has.
PRIOR_HYPER_GAMMA, PRIOR_HYPER_UNIFORM, PRIOR_HYPER_HALFCAUCHY, PRIOR_HYPER_IMPROPER = 10, 11, 12, 13
HYPER_KINDS = (PRIOR_HYPER_GAMMA, PRIOR_HYPER_UNIFORM, PRIOR_HYPER_HALFCAUCHY, PRIOR_HYPER_IMPROPER)
PHASE_INITIAL, PHASE_MID, PHASE_FINAL = 0, 1, 2


# --------------------------------------------------------------------------
# Counter-based noise: Philox4x32-10 + Box-Muller.  The reference draws
# torch.randn_like(p) (sgld.py:67-69,142; verlet_sgld.py:163); the production
# kernel cannot reproduce ATen's stream, so it uses the generator specified
# here and the oracle is the specification it is checked against.
# --------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32 with 10 rounds (Salmon et al., SC'11).  Counters are uint32
    arrays of equal shape, the key two Python ints.  Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1,
                          hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def philox_key(seed: int, stream: int):
    """64-bit Philox key from a user seed and a per-sampler stream id."""
    k = splitmix64((seed + stream * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    return k & 0xFFFFFFFF, k >> 32


def _box_muller(x, y):
    """Two uint32 arrays -> two fp32 standard normals; every step is a single
    correctly rounded fp32 operation so that the device can mirror it."""
    two_m32 = F32(2.0 ** -32)
    u = x.astype(F32) * two_m32 + F32(2.0 ** -33)       # (0, 1]
    t = y.astype(F32) * two_m32 + F32(-0.5)              # [-.5, .5]
    theta = t * F32(6.283185307179586)
    r = np.sqrt(F32(-2.0) * np.log(u))
    return r * np.cos(theta), r * np.sin(theta)


def philox_normal(key, call: int, quad_index):
    """fp32 normals for the flat element quads `quad_index` (flat offset // 4):
    returns an array [len(quad_index), 4].  `call` is the 64-bit launch counter."""
    q = np.asarray(quad_index, dtype=np.uint64)
    c0 = (q & _MASK32)
    c1 = (q >> np.uint64(32))
    c2 = np.full(q.shape, call & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(q.shape, (call >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(c0, c1, c2, c3, key[0], key[1])
    z0, z1 = _box_muller(x0, x1)
    z2, z3 = _box_muller(x2, x3)
    return np.stack([z0, z1, z2, z3], axis=-1).astype(F32)


def philox_normal_segment(key, call: int, flat_offset: int, numel: int):
    """Noise for one segment that starts at `flat_offset` (a multiple of 4) in
    the flat layout."""
    assert flat_offset % 4 == 0
    nq = (numel + 3) // 4
    z = philox_normal(key, call, flat_offset // 4 + np.arange(nq, dtype=np.uint64))
    return z.reshape(-1)[:numel].copy()


# --------------------------------------------------------------------------
# Priors (prior/loc_scale.py:34-35 Normal, :66-67 Laplace, :74-77 StudentT;
# summed per tensor in prior/base.py:57-58).
# --------------------------------------------------------------------------
def prior_log_prob(kind: int, p, loc: float, scale: float, df: float = 3.0) -> float:
    """Sum over the tensor of log density, accumulated in fp64 of fp32 terms.
    `df` is the third hyper-parameter of the kind: StudentT df, GenNorm beta,
    DoubleGamma concentration.  For UNIFORM, loc = low and scale = high - low."""
    if kind in (PRIOR_NONE, PRIOR_IMPROPER):        # Improper.log_prob == 0.0 (prior/loc_scale.py:94-97)
        return 0.0
    p = np.asarray(p, dtype=F32)
    loc, scale, df = F32(loc), F32(scale), F32(df)
    if kind == PRIOR_NORMAL:
        z = (p - loc) / scale
        lp = F32(-0.5) * z * z - np.log(scale) - F32(0.5 * math.log(2 * math.pi))
    elif kind == PRIOR_LAPLACE:
        lp = -np.abs(p - loc) / scale - np.log(F32(2.0) * scale)
    elif kind == PRIOR_STUDENT_T:
        z = (p - loc) / scale
        norm = (np.log(scale) + F32(0.5) * np.log(df) + F32(0.5 * math.log(math.pi))
                + F32(math.lgamma(0.5 * float(df)) - math.lgamma(0.5 * (float(df) + 1.0))))
        lp = F32(-0.5) * (df + F32(1.0)) * np.log1p(z * z / df) - norm
    elif kind == PRIOR_CAUCHY:                      # td.Cauchy.log_prob (prior/loc_scale.py:70-71)
        z = (p - loc) / scale
        lp = -F32(math.log(math.pi)) - np.log(scale) - np.log1p(z * z)
    elif kind == PRIOR_GENNORM:                     # prior/distributions.py:75-79, beta = df
        lp = (-np.log(F32(2.0) * scale) - F32(math.lgamma(1.0 / float(df))) + np.log(df)
              - np.power(np.abs(p - loc) / scale, df))
    elif kind == PRIOR_LOGNORMAL:                   # Normal.log_prob(p) - p  (prior/loc_scale.py:86-92)
        z = (p - loc) / scale
        lp = F32(-0.5) * z * z - np.log(scale) - F32(0.5 * math.log(2 * math.pi)) - p
    elif kind == PRIOR_UNIFORM:                     # -log(high - low) per element (prior/transformed.py:32-45)
        lp = np.full(p.shape, -np.log(scale), dtype=F32)
    elif kind == PRIOR_DOUBLE_GAMMA:                # Gamma(c, 1/s).log_prob(|p - loc|) - log 2
        a = np.abs(p - loc)                         # (prior/transformed.py:83-96, distributions.py:97-110)
        c = df
        rate = F32(1.0) / scale
        lp = (c * np.log(rate) + (c - F32(1.0)) * np.log(a) - rate * a
              - F32(math.lgamma(float(c))) - F32(math.log(2.0)))
    else:
        raise ValueError(kind)
    return float(np.sum(lp.astype(np.float64)))


def prior_grad_log_prob(kind: int, p, loc: float, scale: float, df: float = 3.0):
    """d log density / d p, fp32 (what autograd returns for Prior.log_prob)."""
    p = np.asarray(p, dtype=F32)
    if kind in (PRIOR_NONE, PRIOR_UNIFORM, PRIOR_IMPROPER):
        return np.zeros_like(p)
    loc, scale, df = F32(loc), F32(scale), F32(df)
    d = p - loc
    if kind == PRIOR_NORMAL:
        return -d / (scale * scale)
    if kind == PRIOR_LAPLACE:
        return -np.sign(d) / scale
    if kind == PRIOR_STUDENT_T:
        return -(df + F32(1.0)) * d / (df * scale * scale + d * d)
    if kind == PRIOR_CAUCHY:
        return F32(-2.0) * d / (scale * scale + d * d)
    if kind == PRIOR_GENNORM:
        return -df * np.power(np.abs(d) / scale, df - F32(1.0)) * np.sign(d) / scale
    if kind == PRIOR_LOGNORMAL:
        return -d / (scale * scale) - F32(1.0)
    if kind == PRIOR_DOUBLE_GAMMA:
        return (df - F32(1.0)) / d - np.sign(d) / scale
    raise ValueError(kind)


# --------------------------------------------------------------------------
# Hierarchical priors: scale = f(u), u a sampled scalar with a hyper-prior.
#   Gamma(concentration=a, rate=b) on s = softplus(u)      prior/transformed.py:50-63, hierarchical.py:17-22
#   Uniform(low=a, high=a+b)       on s = a + b Phi(u)     prior/transformed.py:12-47, hierarchical.py:25-30
#   HalfCauchy(scale=a) on s = softplus(u) * b             prior/transformed.py:66-80, hierarchical.py:85-90
#   PositiveImproper: s = softplus(u), log density 0       prior/loc_scale.py:100-103, empirical_bayes.py:24-29
# Scalars are float64 here (the reference's are fp32 0-dim tensors).
# --------------------------------------------------------------------------
def _softplus(u: float) -> float:
    return u if u > 20.0 else math.log1p(math.exp(u))      # torch.nn.functional.softplus, threshold 20


def _sigmoid(u: float) -> float:
    return 1.0 / (1.0 + math.exp(-u))


def hyper_scale(kind: int, u: float, a: float, b: float):
    "(s, ds/du) of a hyper segment"
    u = float(u)
    if kind in (PRIOR_HYPER_GAMMA, PRIOR_HYPER_IMPROPER):
        return _softplus(u), (1.0 if u > 20.0 else _sigmoid(u))
    if kind == PRIOR_HYPER_HALFCAUCHY:
        return _softplus(u) * b, (1.0 if u > 20.0 else _sigmoid(u)) * b
    if kind == PRIOR_HYPER_UNIFORM:
        cdf = 0.5 * math.erfc(-u / math.sqrt(2.0))
        pdf = math.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)
        return a + b * cdf, b * pdf
    raise ValueError(kind)


def hyper_log_prob(kind: int, u: float, a: float, b: float):
    "(log density of the scale, d/du of it) -- what scale_prior.log_prob() and autograd give"
    s, ds = hyper_scale(kind, u, a, b)
    if kind == PRIOR_HYPER_GAMMA:          # td.Gamma(a, b).log_prob(s)
        return a * math.log(b) + (a - 1.0) * math.log(s) - b * s - math.lgamma(a), ((a - 1.0) / s - b) * ds
    if kind == PRIOR_HYPER_UNIFORM:        # -log(high - low), prior/transformed.py:32-45
        return -math.log(b), 0.0
    if kind == PRIOR_HYPER_HALFCAUCHY:     # td.HalfCauchy(a).log_prob(s)
        r = s / a
        return math.log(2.0 / math.pi) - math.log(a) - math.log1p(r * r), -(2.0 * r / a) / (1.0 + r * r) * ds
    if kind == PRIOR_HYPER_IMPROPER:
        return 0.0, 0.0
    raise ValueError(kind)


def prior_dlog_prob_dscale(kind: int, p, loc: float, scale: float, df: float = 3.0) -> float:
    "sum_i d log p(w_i | loc, s) / ds for the kinds a hyper-parameter may drive"
    d = np.asarray(p, dtype=np.float64) - float(loc)
    s, n = float(scale), d.size
    if kind == PRIOR_NORMAL:
        return float(np.sum(d * d)) / s ** 3 - n / s
    if kind == PRIOR_LAPLACE:
        return float(np.sum(np.abs(d))) / s ** 2 - n / s
    if kind == PRIOR_STUDENT_T:
        return (df + 1.0) * float(np.sum(d * d / (df * s * s + d * d))) / s - n / s
    raise ValueError(f"prior kind {kind} cannot have a sampled scale")


# --------------------------------------------------------------------------
# Chain state
# --------------------------------------------------------------------------
@dataclass
class Segment:
    p: np.ndarray                       # parameter, fp32 flat
    g: np.ndarray                       # p.grad as the sampler sees it, fp32 flat
    m: Optional[np.ndarray] = None      # state['momentum_buffer']
    square_avg: Optional[np.ndarray] = None
    preconditioner: float = 1.0         # state['preconditioner'] (M^-1/2)
    est_temperature: float = math.nan
    est_config_temp: float = math.nan
    delta_energy: float = 0.0
    prev_new_momentum_delta: float = 0.0
    prev_p: Optional[np.ndarray] = None
    prev_g: Optional[np.ndarray] = None
    prev_m: Optional[np.ndarray] = None
    # fused prior description (PRIOR_NONE => the prior gradient is already in g)
    prior_kind: int = PRIOR_NONE
    prior_loc: float = 0.0
    prior_scale: float = 1.0
    prior_df: float = 3.0
    # hierarchical priors: a weight segment names its hyper segment, a hyper segment (one element,
    # prior_kind in HYPER_KINDS, hyper-parameters a = prior_loc, b = prior_scale) its weight segment
    link: int = -1


@dataclass
class Group:
    lr: float
    num_data: float
    momentum: float = 0.0
    temperature: float = 1.0
    rmsprop_alpha: float = 0.99
    rmsprop_eps: float = 1e-8
    derived: dict = field(default_factory=dict)


class Chain:
    """One Markov chain: segments + one hyper-parameter group.

    `dot_dtype` float32 restates the reference (`dot`, sgld.py:9-11, is an fp32
    inner product turned into a Python float); float64 is the tighter yardstick
    used when checking the CUDA reductions, which accumulate in fp64."""

    def __init__(self, params: Sequence[np.ndarray], group: Group, dot_dtype=np.float32):
        self.segs: List[Segment] = [
            Segment(p=np.array(p, dtype=F32).reshape(-1),
                    g=np.zeros(int(np.size(p)), dtype=F32)) for p in params]
        self.group = group
        self.dot_dtype = dot_dtype
        update_preconditioner(self)      # sgld.py:44 (constructor)

    def dot(self, a, b) -> float:
        if self.dot_dtype == np.float32:
            return float(np.dot(a, b))
        return float(np.dot(a.astype(np.float64), b.astype(np.float64)))


NoiseFn = Callable[[int, int], np.ndarray]   # (segment index, numel) -> fp32 N(0,1)


# --------------------------------------------------------------------------
# Shared pieces
# --------------------------------------------------------------------------
def update_preconditioner(chain: Chain) -> None:
    """sgld.py:156-179: s_t = mean(square_avg_t) + eps; M_t = (s_t/min s)^(-1/4).
    Creates square_avg = ones on first use."""
    eps = chain.group.rmsprop_eps
    s = []
    for seg in chain.segs:
        if seg.square_avg is None:
            seg.square_avg = np.ones_like(seg.p)
        s.append(float(np.mean(seg.square_avg, dtype=F32)) + eps)
    lo = min(s) if s else math.inf
    for seg, v in zip(chain.segs, s):
        seg.preconditioner = (v / lo) ** (-1 / 4)


def sample_momentum(chain: Chain, noise: NoiseFn, keep: float = 0.0) -> None:
    """sgld.py:57-69."""
    assert 0.0 <= keep <= 1.0
    if keep == 1.0:
        return
    std = math.sqrt(chain.group.temperature * (1 - keep))
    for i, seg in enumerate(chain.segs):
        eps = np.asarray(noise(i, seg.p.size), dtype=F32)
        if keep == 0.0:
            seg.m = eps * F32(std)
        else:
            seg.m = seg.m * F32(math.sqrt(keep)) + F32(std) * eps


def fuse_prior_into_grad(chain: Chain, grad_max: Optional[float] = None) -> None:
    """What the fused kernel does in-register: g <- clamp(g_lik - dlogp/dp / N).
    Restates models/base.py:72-77 (potential = loss - log_prior/N), the backward
    at inference.py:218 and the clamp at inference.py:219-220."""
    n = F32(chain.group.num_data)
    refresh_hyper_scales(chain)
    for seg in chain.segs:
        if seg.prior_kind == PRIOR_NONE:
            continue
        if seg.prior_kind in HYPER_KINDS:
            # d/du [ sum_i log p(w_i | s(u)) + log p_hyper(s(u)) ] through autograd in the reference
            w = chain.segs[seg.link]
            _, ds = hyper_scale(seg.prior_kind, seg.p[0], seg.prior_loc, seg.prior_scale)
            _, dlp = hyper_log_prob(seg.prior_kind, seg.p[0], seg.prior_loc, seg.prior_scale)
            dl = np.array([prior_dlog_prob_dscale(w.prior_kind, w.p, w.prior_loc, w.prior_scale, w.prior_df) * ds
                           + dlp], dtype=F32)
        else:
            dl = prior_grad_log_prob(seg.prior_kind, seg.p, seg.prior_loc,
                                     seg.prior_scale, seg.prior_df)
        seg.g = (seg.g - dl / n).astype(F32)
        if grad_max is not None:
            np.clip(seg.g, -F32(grad_max), F32(grad_max), out=seg.g)


def refresh_hyper_scales(chain: Chain) -> None:
    "prior_scale of every weight segment that has a hyper segment, from the current u"
    for seg in chain.segs:
        if seg.prior_kind in HYPER_KINDS:
            s, _ = hyper_scale(seg.prior_kind, seg.p[0], seg.prior_loc, seg.prior_scale)
            chain.segs[seg.link].prior_scale = float(F32(s))


def link_hyper(chain: Chain, weight: int, hyper: int, kind: int, a: float, b: float) -> None:
    "declare segment `hyper` (one element) the scale hyper-parameter of segment `weight`"
    h, w = chain.segs[hyper], chain.segs[weight]
    assert h.p.size == 1 and kind in HYPER_KINDS
    assert w.prior_kind in (PRIOR_NORMAL, PRIOR_LAPLACE, PRIOR_STUDENT_T)
    h.prior_kind, h.prior_loc, h.prior_scale, h.link = kind, float(a), float(b), weight
    w.link = hyper
    refresh_hyper_scales(chain)


def log_prior(chain: Chain) -> float:
    """models/base.py:25-30 restricted to the fused kinds (hyper segments contribute the
    log density of their scale)."""
    refresh_hyper_scales(chain)
    total = 0.0
    for s in chain.segs:
        if s.prior_kind in HYPER_KINDS:
            total += hyper_log_prob(s.prior_kind, s.p[0], s.prior_loc, s.prior_scale)[0]
        else:
            total += prior_log_prob(s.prior_kind, s.p, s.prior_loc, s.prior_scale, s.prior_df)
    return total


def _rmsprop(seg: Segment, alpha: float) -> None:
    # sgld.py:153-154 / verlet_sgld.py:196-197 / hmc.py:78-79
    a = F32(alpha)
    seg.square_avg = seg.square_avg * a + F32(1 - alpha) * (seg.g * seg.g)


def _save_state(seg: Segment, has_momentum: bool) -> None:
    # verlet_sgld.py:72-83
    seg.prev_p = seg.p.copy()
    seg.prev_g = seg.g.copy()
    if has_momentum:
        seg.prev_m = seg.m.copy()


# --------------------------------------------------------------------------
# SGLD (sgld.py:114-154)
# --------------------------------------------------------------------------
def sgld_step(chain: Chain, noise: Optional[NoiseFn], calc_metrics: bool = True,
              is_final: bool = False) -> None:
    g = chain.group
    hn = math.sqrt(g.lr * g.num_data)
    h = math.sqrt(g.lr / g.num_data)
    noise_std = math.sqrt(2 * (1 - g.momentum) * g.temperature)
    g.derived.update(hn=hn, h=h, noise_std=noise_std)
    for i, seg in enumerate(chain.segs):
        M = seg.preconditioner
        d = seg.p.size
        if g.momentum > 0:
            if seg.m is None:
                raise RuntimeError("No 'momentum_buffer' stored in state. "
                                   "Perhaps you forgot to call `sample_momentum`?")
            mom = seg.m
            if calc_metrics:
                seg.est_temperature = chain.dot(mom, mom) / d
            if not is_final:
                mom = mom * F32(g.momentum) + F32(-hn * M) * seg.g
                seg.m = mom
        else:
            mom = seg.g * F32(-hn * M) if not is_final else None
            if calc_metrics:
                if mom is None:       # the reference hits an unbound local here
                    raise UnboundLocalError("momentum")
                seg.est_temperature = chain.dot(mom, mom) / d
        if not is_final and g.temperature > 0:
            eps = np.asarray(noise(i, d), dtype=F32)
            mom += F32(noise_std) * eps
        if calc_metrics:
            seg.est_config_temp = chain.dot(seg.p, seg.g) * (g.num_data / d)
        if not is_final:
            seg.p = seg.p + F32(h * M) * mom
            _rmsprop(seg, g.rmsprop_alpha)


# --------------------------------------------------------------------------
# VerletSGLD / GGMC (verlet_sgld.py:86-197)
# --------------------------------------------------------------------------
def verlet_group_constants(g: Group, phase: int) -> dict:
    """verlet_sgld.py:138-146 (intermediate), :96-101 (initial), :129-134 (final)."""
    a = g.momentum
    c = dict(b2h2=g.lr / g.num_data, bh=math.sqrt(g.lr / g.num_data),
             bhn=math.sqrt(g.lr * g.num_data))
    if phase == PHASE_MID:
        c.update(mom_decay=a, grad_v=1 + a, noise_std=math.sqrt((1 - a ** 2) * g.temperature))
    elif phase == PHASE_INITIAL:
        c.update(mom_decay=math.sqrt(a), grad_v=1.0, noise_std=math.sqrt((1 - a) * g.temperature))
    else:
        c.update(mom_decay=math.sqrt(a), grad_v=math.sqrt(a),
                 noise_std=math.sqrt((1 - a) * g.temperature))
    return c


def verlet_point_energy(chain: Chain, seg: Segment) -> float:
    # verlet_sgld.py:44-47
    g = chain.group
    curv = seg.preconditioner ** 2 * g.num_data ** 2 * g.derived["b2h2"] / 8
    return curv * chain.dot(seg.g, seg.g)


def verlet_step(chain: Chain, noise: NoiseFn, phase: int = PHASE_MID,
                save_state: bool = False, calc_metrics: bool = True) -> None:
    g = chain.group
    c = verlet_group_constants(g, phase)
    g.derived.update(c)
    for i, seg in enumerate(chain.segs):
        if seg.m is None:
            raise RuntimeError("No 'momentum_buffer' stored in state. "
                               "Perhaps you forgot to call `sample_momentum`?")
        if save_state:
            _save_state(seg, g.momentum > 0)
        M = seg.preconditioner
        d = seg.p.size
        old = seg.m
        new = np.asarray(noise(i, d), dtype=F32) * F32(c["noise_std"])   # drawn even if std == 0
        new = new + F32(-.5 * c["grad_v"] * c["bhn"] * M) * seg.g
        if c["mom_decay"] > 0:
            new = new + F32(c["mom_decay"]) * old
        c_gm = -.5 * c["bhn"] * M
        if phase == PHASE_INITIAL:
            seg.delta_energy = -verlet_point_energy(chain, seg)
        else:
            seg.delta_energy += seg.prev_new_momentum_delta
            seg.delta_energy += c_gm * chain.dot(seg.g, old)
        seg.prev_new_momentum_delta = c_gm * chain.dot(seg.g, new)
        if calc_metrics:
            which = new if phase == PHASE_FINAL else old
            seg.est_temperature = chain.dot(which, which) / d
            seg.est_config_temp = chain.dot(seg.p, seg.g) * (g.num_data / d)
        seg.m = new
        if phase != PHASE_FINAL:
            seg.p = seg.p + F32(c["bh"] * M) * new
            _rmsprop(seg, g.rmsprop_alpha)


def verlet_delta_energy(chain: Chain, prev_potential: float, potential: float) -> float:
    # verlet_sgld.py:27-42
    tot = 0.0
    for seg in chain.segs:
        tot += seg.delta_energy + verlet_point_energy(chain, seg)
    return tot + (float(potential) - prev_potential) * chain.group.num_data


def maybe_reject(chain: Chain, delta_energy: float, u: Optional[float]):
    """verlet_sgld.py:49-70.  `u` is the uniform the reference takes from
    torch.rand(()) on the host generator (one draw iff temperature != 0)."""
    T = chain.group.temperature
    if T == 0.0:
        return False, 0.0
    log_accept = -delta_energy / T
    reject = math.log(u) > log_accept
    if reject:
        for seg in chain.segs:
            seg.p = seg.prev_p.copy()
            seg.g = seg.prev_g.copy()
            if seg.prev_m is not None and seg.m is not None:
                seg.m = seg.prev_m.copy()
    return reject, log_accept


# --------------------------------------------------------------------------
# HMC (hmc.py:25-79)
# --------------------------------------------------------------------------
def hmc_step(chain: Chain, phase: int = PHASE_MID, save_state: bool = False,
             calc_metrics: bool = True) -> None:
    g = chain.group
    assert g.momentum == 1.0 and g.temperature == 1.0          # hmc.py:39
    c = verlet_group_constants(g, phase)
    g.derived.update(c)
    for seg in chain.segs:
        if seg.m is None:
            raise RuntimeError("No 'momentum_buffer' stored in state. "
                               "Perhaps you forgot to call `sample_momentum`?")
        if save_state:
            _save_state(seg, True)
        M = seg.preconditioner
        d = seg.p.size
        if phase == PHASE_INITIAL:
            mm = chain.dot(seg.m, seg.m)
            seg.delta_energy = -.5 * mm
            if calc_metrics:
                seg.est_temperature = mm / d
        if calc_metrics:
            if phase == PHASE_MID:
                seg.est_temperature = chain.dot(seg.m, seg.m) / d
            seg.est_config_temp = chain.dot(seg.p, seg.g) * (g.num_data / d)
        seg.m = seg.m + F32(-.5 * c["grad_v"] * c["bhn"] * M) * seg.g
        if phase == PHASE_FINAL:
            if calc_metrics:
                seg.est_temperature = chain.dot(seg.m, seg.m) / d
        else:
            seg.p = seg.p + F32(c["bh"] * M) * seg.m
            _rmsprop(seg, g.rmsprop_alpha)


def hmc_delta_energy(chain: Chain, prev_potential: float, potential: float) -> float:
    # verlet_sgld.py:27-42 with hmc.py:32-33 as the point energy
    tot = 0.0
    for seg in chain.segs:
        tot += seg.delta_energy + .5 * chain.dot(seg.m, seg.m)
    return tot + (float(potential) - prev_potential) * chain.group.num_data
