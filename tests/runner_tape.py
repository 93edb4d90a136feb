"""Record what a runner does to a sampler, and replay it on another sampler.

The parity claim of this repo is about the SAMPLER: same inputs (hyper-parameters, p.grad,
the N(0,1) tensors, the Metropolis uniform, the potentials handed to delta_energy) -> same
outputs (parameters, momentum, per-tensor scalars, delta energy, accept / reject).  To test it
under the reference's own runners (bnn_priors/inference.py, inference_reject.py) on the real
BASELINE models, the runner is run twice:

  run A   reference runner + reference eager sampler, wrapped by `recording_class`: every public
          sampler call is logged on a `Tape` with everything that went in and came out;
  run B   the same runner + the sampler under test (after `overlay.install()`: the B200 kernel),
          wrapped by `replaying_class`: before every call the inputs of the matching call of
          run A are put in place (so the two samplers see identical gradients and noise no
          matter how the minibatches were shuffled), after it the outputs are compared.

All of the runner's own code -- zero_grad / backward / clamp, the lr scheduler, temperature
switches, update_preconditioner, state reads in store_metrics, state_dict / load_state_dict in
the evaluation, the sample saver -- runs unchanged in both.  Test infrastructure only.
"""
from __future__ import annotations

import contextlib
import math
from typing import List, Optional

import torch

STEP_OPS = ("initial_step", "step", "final_step")
OPS = STEP_OPS + ("sample_momentum", "delta_energy", "maybe_reject", "update_preconditioner")
SCALAR_KEYS = ("preconditioner", "est_temperature", "est_config_temp", "delta_energy", "prev_new_momentum_delta")
GROUP_KEYS = ("lr", "num_data", "momentum", "temperature", "rmsprop_alpha", "rmsprop_eps")


class Tape:
    def __init__(self):
        self.events: List[dict] = []
        self.cursor = 0

    def next(self) -> dict:
        if self.cursor >= len(self.events):
            raise AssertionError("the replayed run makes more sampler calls than the recorded one")
        ev = self.events[self.cursor]
        self.cursor += 1
        return ev

    def summary(self) -> dict:
        out = {}
        for ev in self.events:
            out[ev["op"]] = out.get(ev["op"], 0) + 1
        return out


@contextlib.contextmanager
def _rng_hooks(randn_like=None, rand=None):
    real_randn_like, real_rand = torch.randn_like, torch.rand
    if randn_like is not None:
        torch.randn_like = lambda t, *a, **k: randn_like(real_randn_like, t, *a, **k)
    if rand is not None:
        torch.rand = lambda *a, **k: rand(real_rand, *a, **k)
    try:
        yield
    finally:
        torch.randn_like, torch.rand = real_randn_like, real_rand


def _params(opt):
    return [p for g in opt.param_groups for p in g["params"]]


def _snapshot(opt, params):
    st = opt.state
    m = []
    for p in params:
        b = st[p].get("momentum_buffer") if p in st else None
        m.append(None if b is None else b.detach().clone())
    scal = {}
    for k in SCALAR_KEYS:
        vals = []
        for p in params:
            v = st[p].get(k) if p in st else None
            vals.append(None if v is None else float(v))
        scal[k] = vals
    scal["square_avg_mean"] = [
        float(st[p]["square_avg"].double().mean()) if (p in st and "square_avg" in st[p]) else None for p in params]
    return dict(p=[p.detach().clone() for p in params], m=m, scalars=scal)


def _to_float(a):
    return float(a.detach()) if isinstance(a, torch.Tensor) else float(a)


def recording_class(base, tape: Tape, prior_grad_fn=None):
    """Subclass of a (reference) sampler class that logs every public call on `tape`.
    `prior_grad_fn() -> list of tensors`: the prior's share of p.grad at the current parameters
    (recorded with every gradient, for replays with the prior fused into the kernel)."""

    class Recording(base):
        _tape_depth = 0

        def _tape_call(self, op, args, kwargs):
            if self._tape_depth > 0:          # a public method calling another one
                return getattr(base, op)(self, *args, **kwargs)
            params = _params(self)
            g0 = self.param_groups[0]
            ev = dict(op=op, kwargs=dict(kwargs), group={k: float(g0[k]) for k in GROUP_KEYS if k in g0})
            if op in STEP_OPS:
                ev["grads"] = [None if p.grad is None else p.grad.detach().clone() for p in params]
                # sum |p_i g_i| per tensor: the size of the terms dot(p, g) adds up (est_config_temp, sgld.py:146)
                with torch.no_grad():
                    ev["pg_abs"] = torch.stack([(p.detach() * p.grad).abs().sum().double() if p.grad is not None
                                                else torch.zeros((), dtype=torch.float64, device=p.device)
                                                for p in params]).tolist()
                if prior_grad_fn is not None:
                    ev["prior_grads"] = [g.detach().clone() for g in prior_grad_fn()]
            if op in ("delta_energy", "maybe_reject"):
                ev["args"] = [_to_float(a) for a in args]
            noise, unif = [], []

            def randn_like(real, t, *a, **k):
                z = real(t, *a, **k)
                noise.append(z.detach().clone())
                return z

            def rand(real, *a, **k):
                u = real(*a, **k)
                unif.append(float(u))
                return u
            self._tape_depth += 1
            try:
                with _rng_hooks(randn_like, rand):
                    out = getattr(base, op)(self, *args, **kwargs)
            finally:
                self._tape_depth -= 1
            ev["noise"] = noise
            if unif:
                assert len(unif) == 1
                ev["u"] = unif[0]
            if op == "delta_energy":
                ev["out"] = float(out)
            elif op == "maybe_reject":
                ev["out"] = (bool(out[0]), float(out[1]))
            ev["after"] = _snapshot(self, params)
            tape.events.append(ev)
            return out

    for _op in OPS:
        if hasattr(base, _op):
            setattr(Recording, _op, (lambda op: lambda self, *a, **k: self._tape_call(op, a, k))(_op))
    Recording.__name__ = base.__name__
    return Recording


class Report:
    def __init__(self):
        self.p_err = self.m_err = 0.0
        self.scalar_err = {}
        self.de_abs_err = self.de_rel_err = 0.0
        self.de_scale = 0.0
        self.de_term_err = 0.0       # |delta energy error| / (sum of the magnitudes of its terms)
        self.decisions = self.decisions_equal = self.rejections = 0
        self.min_margin = math.inf
        self.log_accept_err = 0.0
        self.n_events = 0
        self.ops = {}
        self.kink_guards = 0         # elements put back on the reference's side of a prior's kink (fused replays)

    def bump(self, key, v):
        self.scalar_err[key] = max(self.scalar_err.get(key, 0.0), v)

    def as_dict(self):
        return {k: v for k, v in self.__dict__.items()}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    "max |a - b| relative to the RMS of the reference tensor b (the scale of the tensor)"
    a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1).to(a.device)
    if a.numel() == 0:
        return 0.0
    scale = max(float(b.pow(2).mean().sqrt()), 1e-30)
    return float((a - b).abs().max()) / scale


def replaying_class(base, tape: Tape, report: Report, fused_prior: bool = False):
    """Subclass of the sampler class under test whose public calls take their inputs from
    `tape` and are compared with what the recorded sampler produced."""

    class Replaying(base):
        _tape_depth = 0
        _last_de = None
        _last_pg_abs = None

        def _tape_call(self, op, args, kwargs):
            if self._tape_depth > 0:
                return getattr(base, op)(self, *args, **kwargs)
            ev = tape.next()
            assert ev["op"] == op, f"call #{tape.cursor - 1}: recorded {ev['op']}, replayed run calls {op}"
            assert dict(kwargs) == ev["kwargs"], (op, kwargs, ev["kwargs"])
            params = _params(self)
            g0 = self.param_groups[0]
            for k, want in ev["group"].items():
                have = float(g0[k])
                assert abs(have - want) <= 1e-12 * max(1.0, abs(want)), f"{op}: group[{k}] {have} != {want}"
            report.n_events += 1
            report.ops[op] = report.ops.get(op, 0) + 1
            if "pg_abs" in ev and ev["kwargs"].get("calc_metrics", True):
                self._last_pg_abs = ev["pg_abs"]     # the step that (re)computes est_config_temp
            if op in STEP_OPS and fused_prior and tape.cursor >= 2:
                # A fused replay gets the likelihood gradient and computes the prior's term from ITS parameters.
                # A Laplace prior's term is sign(p - loc) / (N scale) (prior/loc_scale.py:66-67): discontinuous at
                # the kink, so an element that lies within the comparison tolerance of the kink on the other side
                # than the reference's (both are right to 1e-7) would get the opposite term and leave the
                # reference trajectory for good -- a property of the prior, not of the sampler.  Such elements
                # (loc = 0 in every prior `get_prior` builds; expected: none to a few per run) are put on the
                # reference's value before the step; the move is far below TOL and counted in the report.
                for p, q in zip(params, tape.events[tape.cursor - 2]["after"]["p"]):
                    q = q.to(p.device)
                    flip = (torch.sign(p.detach()) != torch.sign(q)) & ((p.detach() - q).abs() <= 1e-5 * q.abs().max())
                    n_flip = int(flip.sum())
                    if n_flip:
                        with torch.no_grad():
                            p[flip] = q[flip]
                        report.kink_guards += n_flip
            if op in STEP_OPS:
                for i, (p, g) in enumerate(zip(params, ev["grads"])):
                    if g is None:
                        p.grad = None
                        continue
                    g = g.to(p.device)
                    if fused_prior:
                        g = (g.double() - ev["prior_grads"][i].to(p.device).double()).float()
                    p.grad = g.clone()          # a tensor of its own, like the one autograd hands over
            noise = [z.to(params[0].device) for z in ev["noise"]]
            if op == "delta_energy":
                args = tuple(ev["args"])
            elif op == "maybe_reject":
                # decide on this sampler's OWN delta energy if the runner passes it on (it does)
                if self._last_de is not None and abs(ev["args"][0] - self._last_de[0]) <= 1e-9 * max(1.0, abs(self._last_de[0])):
                    args = (self._last_de[1],)
                else:
                    args = tuple(ev["args"])
            use_hook = not hasattr(self, "set_replay_noise")
            if noise and not use_hook:
                if len(noise) != len(params):
                    raise AssertionError(f"{op}: {len(noise)} noise tensors for {len(params)} parameters")
                self.set_replay_noise(noise)
            it = iter(noise)

            def randn_like(real, t, *a, **k):
                return next(it).clone()

            def rand(real, *a, **k):
                return torch.tensor(ev["u"], dtype=torch.float32) if "u" in ev else real(*a, **k)
            self._tape_depth += 1
            try:
                with _rng_hooks(randn_like if use_hook else None, rand):
                    out = getattr(base, op)(self, *args, **kwargs)
            finally:
                self._tape_depth -= 1
            self._compare(op, ev, out, params)
            return out

        def _compare(self, op, ev, out, params):
            if op == "delta_energy":
                self._last_de = (ev["out"], float(out))
                report.de_abs_err = max(report.de_abs_err, abs(float(out) - ev["out"]))
                report.de_rel_err = max(report.de_rel_err, abs(float(out) - ev["out"]) / max(abs(ev["out"]), 1e-12))
                if math.isfinite(ev["out"]):
                    report.de_scale = max(report.de_scale, abs(ev["out"]))
                    # delta energy = sum_t (running sum_t + point energy_t) + N (U - U_prev): judged
                    # against the size of the terms it adds up (they cancel)
                    terms = [abs(x) for x in ev["after"]["scalars"]["delta_energy"] if x is not None and math.isfinite(x)]
                    nd = float(self.param_groups[0]["num_data"])
                    scale = sum(terms) + abs((ev["args"][1] - ev["args"][0]) * nd) + 1.0
                    report.de_term_err = max(report.de_term_err, abs(float(out) - ev["out"]) / scale)
            elif op == "maybe_reject":
                report.decisions += 1
                report.decisions_equal += int(bool(out[0]) == ev["out"][0])
                report.rejections += int(ev["out"][0])
                report.log_accept_err = max(report.log_accept_err, abs(float(out[1]) - ev["out"][1]))
                if "u" in ev:
                    report.min_margin = min(report.min_margin, abs(math.log(ev["u"]) - ev["out"][1]))
            want = ev["after"]
            st = self.state
            for i, p in enumerate(params):
                report.p_err = max(report.p_err, rel_err(p, want["p"][i]))
                wm = want["m"][i]
                if wm is not None:
                    hm = st[p].get("momentum_buffer")
                    assert hm is not None, f"{op}: no momentum_buffer for tensor {i}"
                    report.m_err = max(report.m_err, rel_err(hm, wm))
            for key in SCALAR_KEYS + ("square_avg_mean",):
                wants = want["scalars"][key]
                for i, p in enumerate(params):
                    b = wants[i]
                    if b is None or not math.isfinite(b):
                        continue
                    if key == "square_avg_mean":
                        a = float(st[p]["square_avg"].double().mean())
                    else:
                        a = st[p].get(key)
                        if a is None:
                            raise AssertionError(f"{op}: state[{key}] missing for tensor {i}")
                        a = float(a)
                    if key == "est_config_temp" and self._last_pg_abs is not None:
                        # dot(p, g) N / d: a sum of signed terms, judged against sum |p_i g_i| N / d
                        nd = float(self.param_groups[0]["num_data"])
                        scale = max(abs(b), self._last_pg_abs[i] * nd / p.numel(), 1e-30)
                    elif key in ("delta_energy", "prev_new_momentum_delta"):
                        # running sums of signed terms: judged against the size of the terms
                        mags = [abs(x) for x in wants if x is not None and math.isfinite(x)]
                        scale = max(max(mags, default=0.0), 1e-3)
                    else:
                        scale = max(abs(b), 1e-30)
                    report.bump(key, abs(a - b) / scale)

    for _op in OPS:
        if hasattr(base, _op):
            setattr(Replaying, _op, (lambda op: lambda self, *a, **k: self._tape_call(op, a, k))(_op))
    Replaying.__name__ = base.__name__
    return Replaying


@contextlib.contextmanager
def bound_sampler_classes(ref_mcmc, wrap):
    """Re-bind SGLD / VerletSGLD / HMC on the (reference) `bnn_priors.mcmc` module object to
    `wrap(current class)` for the duration of a run -- the runners look the classes up there at
    run time (inference.py:89-94, inference_reject.py:12-16,183-198)."""
    names = ("SGLD", "VerletSGLD", "HMC")
    saved = {n: getattr(ref_mcmc, n) for n in names}
    try:
        for n in names:
            setattr(ref_mcmc, n, wrap(saved[n]))
        yield
    finally:
        for n, c in saved.items():
            setattr(ref_mcmc, n, c)
