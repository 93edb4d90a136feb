"""Replay a golden call trace (tests/golden/*.npz, written by
tests/golden/make_golden.py from the unmodified reference) through an engine and
measure how far the engine's outputs are from the reference's.

Engines: `OracleEngine` (oracle/sgmcmc_oracle.py, numpy) and `CudaEngine`
(the product samplers in bnn_priors_b200.mcmc, via the C-ABI library).
Test infrastructure only.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_OPS = ("initial_step", "step", "final_step")
PHASE = {"initial_step": 0, "step": 1, "final_step": 2}


class Trace:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.arrays = {k: z[k] for k in z.files if k != "meta"}
        self.events = self.meta["events"]
        self.shapes = [tuple(s) for s in self.meta["shapes"]]
        self.sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self.sampler = self.meta["sampler"]
        self.ctor = self.meta["ctor"]
        self.priors = self.meta.get("priors")

    def arr(self, idx) -> np.ndarray:
        return self.arrays[f"a{idx}"]

    def split(self, flat) -> List[np.ndarray]:
        out, o = [], 0
        for n in self.sizes:
            out.append(np.asarray(flat[o:o + n]))
            o += n
        return out


@dataclass
class Report:
    """Worst deviations seen during a replay (relative to the tensor's scale)."""
    p_err: float = 0.0
    m_err: float = 0.0
    scalar_err: Dict[str, float] = field(default_factory=dict)
    de_abs_err: float = 0.0
    de_rel_err: float = 0.0
    decisions: int = 0
    decisions_equal: int = 0
    min_margin: float = math.inf       # min |log u - log_accept| over decisions
    log_accept_err: float = 0.0
    n_events: int = 0

    def bump(self, key, v):
        self.scalar_err[key] = max(self.scalar_err.get(key, 0.0), v)


def _rel(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = max(float(np.sqrt(np.mean(b * b))), 1e-30)
    return float(np.max(np.abs(a - b)) / scale)


def replay(trace: Trace, engine, fused_prior: bool = False, check_every: int = 1) -> Report:
    rep = Report()
    last_de = None
    for i, ev in enumerate(trace.events):
        op = ev["op"]
        if op == "set_preconditioner":
            engine.set_preconditioners(ev["values"])
            continue
        rep.n_events += 1
        engine.set_group(ev["group"])
        have = engine.preconditioners()
        for a, b in zip(have, ev["precond"]):
            rep.bump("preconditioner_in", abs(a - b) / abs(b))
        if op in STEP_OPS and ev["grad"] != "same":
            g = trace.arr(ev["grad"]).astype(np.float32)
            if fused_prior:
                g = (g.astype(np.float64) - trace.arr(ev["grad_prior"]).astype(np.float64)).astype(np.float32)
            engine.set_grad(trace.split(g))
        noise = trace.split(trace.arr(ev["noise"])) if "noise" in ev else None
        args = list(ev.get("args", []))
        if op == "maybe_reject" and last_de is not None and args[0] == last_de[0]:
            args[0] = last_de[1]      # decide on the engine's OWN delta energy
        out = engine.call(op, ev.get("kwargs", {}), args, noise, ev.get("u"))
        if op == "delta_energy":
            last_de = (ev["out"], out)
            rep.de_abs_err = max(rep.de_abs_err, abs(out - ev["out"]))
            rep.de_rel_err = max(rep.de_rel_err, abs(out - ev["out"]) / max(abs(ev["out"]), 1e-12))
        elif op == "maybe_reject":
            rep.decisions += 1
            rep.decisions_equal += int(bool(out[0]) == bool(ev["out"][0]))
            rep.log_accept_err = max(rep.log_accept_err, abs(out[1] - ev["out"][1]))
            if "u" in ev:
                rep.min_margin = min(rep.min_margin, abs(math.log(ev["u"]) - ev["out"][1]))
        if "p" in ev and (i % check_every == 0 or i == len(trace.events) - 1):
            rep.p_err = max(rep.p_err, _rel(engine.p_flat(), trace.arr(ev["p"])))
            if trace.ctor.get("momentum", 1.0) > 0:
                rep.m_err = max(rep.m_err, _rel(engine.m_flat(), trace.arr(ev["m"])))
        if "scalars" in ev:
            got = engine.scalars()
            for key, want in ev["scalars"].items():
                if key not in got:
                    continue
                for a, b in zip(got[key], want):
                    if b is None or a is None or (isinstance(b, float) and math.isnan(b)):
                        continue
                    if key in ("delta_energy", "prev_new_momentum_delta"):
                        # sums of signed terms: judge against the magnitude of the run
                        rep.bump(key, abs(a - b) / max(abs(b), 1.0))
                    else:
                        rep.bump(key, abs(a - b) / max(abs(b), 1e-12))
    return rep


# ---------------------------------------------------------------------------
class OracleEngine:
    """Adapter around oracle/sgmcmc_oracle.py."""

    def __init__(self, trace: Trace, fused_prior: bool = False, dot_dtype=np.float32):
        from oracle import sgmcmc_oracle as O
        self.O = O
        c = trace.ctor
        self.kind = trace.sampler
        if self.kind == "HMC":
            group = O.Group(lr=c["lr"], num_data=c["num_data"], momentum=1.0, temperature=1.0)
        else:
            group = O.Group(lr=c["lr"], num_data=c["num_data"], momentum=c.get("momentum", 0.0),
                            temperature=c.get("temperature", 1.0))
        self.chain = O.Chain(trace.split(trace.arr(trace.meta["p0"])), group, dot_dtype=dot_dtype)
        self.fused = fused_prior
        self.needs_fuse = False
        if fused_prior:
            for seg, spec in zip(self.chain.segs, trace.priors):
                seg.prior_kind, seg.prior_loc = spec["kind"], spec["loc"]
                seg.prior_scale, seg.prior_df = spec["scale"], spec["df"]

    def set_group(self, g):
        self.chain.group.lr = g["lr"]
        self.chain.group.temperature = g["temperature"]

    def set_preconditioners(self, values):
        for seg, v in zip(self.chain.segs, values):
            seg.preconditioner = float(v)

    def preconditioners(self):
        return [seg.preconditioner for seg in self.chain.segs]

    def set_grad(self, grads):
        for seg, g in zip(self.chain.segs, grads):
            seg.g = np.array(g, dtype=np.float32)
        self.needs_fuse = self.fused

    def call(self, op, kwargs, args, noise, u):
        O, ch = self.O, self.chain

        def nz(i, n):
            assert noise[i].size == n
            return noise[i]
        if op == "sample_momentum":
            return O.sample_momentum(ch, nz, keep=kwargs.get("keep", 0.0))
        if op == "update_preconditioner":
            return O.update_preconditioner(ch)
        if op in STEP_OPS:
            if self.needs_fuse:
                O.fuse_prior_into_grad(ch)
                self.needs_fuse = False
            cm = kwargs.get("calc_metrics", True)
            if self.kind == "SGLD":
                return O.sgld_step(ch, nz, calc_metrics=cm, is_final=(op == "final_step"))
            if self.kind == "VerletSGLD":
                return O.verlet_step(ch, nz, phase=PHASE[op], calc_metrics=cm,
                                     save_state=kwargs.get("save_state", op == "initial_step"))
            return O.hmc_step(ch, phase=PHASE[op], calc_metrics=cm,
                              save_state=kwargs.get("save_state", op == "initial_step"))
        if op == "delta_energy":
            fn = O.hmc_delta_energy if self.kind == "HMC" else O.verlet_delta_energy
            return fn(ch, args[0], args[1])
        if op == "maybe_reject":
            return O.maybe_reject(ch, args[0], u)
        raise ValueError(op)

    def p_flat(self):
        return np.concatenate([s.p for s in self.chain.segs])

    def m_flat(self):
        return np.concatenate([s.m if s.m is not None else np.zeros_like(s.p) for s in self.chain.segs])

    def scalars(self):
        segs = self.chain.segs
        return dict(
            preconditioner=[s.preconditioner for s in segs],
            est_temperature=[s.est_temperature for s in segs],
            est_config_temp=[s.est_config_temp for s in segs],
            delta_energy=[s.delta_energy for s in segs],
            prev_new_momentum_delta=[s.prev_new_momentum_delta for s in segs],
            square_avg_mean=[None if s.square_avg is None else float(np.mean(s.square_avg, dtype=np.float64))
                             for s in segs])


# ---------------------------------------------------------------------------
class CudaEngine:
    """Adapter around the product samplers (bnn_priors_b200.mcmc) on cuda:0: the
    same calls the trace recorded, made through the reference-shaped Python API,
    every one of them ending in `bnnp_launch` of libbnnp.so."""

    def __init__(self, trace: Trace, fused_prior: bool = False, foreign_grads: bool = False):
        import torch
        from bnn_priors_b200 import mcmc
        self.torch = torch
        dev = torch.device("cuda", 0)
        p0 = trace.split(trace.arr(trace.meta["p0"]))
        self.params = [torch.nn.Parameter(torch.tensor(np.asarray(a, dtype=np.float32).reshape(s), device=dev))
                       for a, s in zip(p0, trace.shapes)]
        c = dict(trace.ctor)
        self.kind = trace.sampler
        self.opt = getattr(mcmc, self.kind)(self.params, **c, seed=1234)
        self.foreign_grads = foreign_grads
        self.fused = fused_prior
        if fused_prior:
            (fg,) = self.opt.flat_groups
            for i, spec in enumerate(trace.priors):
                fg.set_prior(i, spec["kind"], spec["loc"], spec["scale"], spec["df"])
            fg.prior_fused = True

    def set_group(self, g):
        for pg in self.opt.param_groups:
            pg["lr"] = g["lr"]
            pg["temperature"] = g["temperature"]

    def set_preconditioners(self, values):
        for p, v in zip(self.params, values):
            self.opt.state[p]["preconditioner"] = float(v)

    def preconditioners(self):
        return [self.opt.state[p]["preconditioner"] for p in self.params]

    def set_grad(self, grads):
        torch = self.torch
        for p, g in zip(self.params, grads):
            t = torch.tensor(np.asarray(g, dtype=np.float32).reshape(tuple(p.shape)), device=p.device)
            if p.grad is None or self.foreign_grads:
                p.grad = t            # a tensor the sampler has never seen: must be adopted
            else:
                p.grad.copy_(t)       # in place, like autograd accumulation into the flat view

    def call(self, op, kwargs, args, noise, u):
        torch, opt = self.torch, self.opt
        if noise is not None:
            opt.set_replay_noise([torch.tensor(np.asarray(n, dtype=np.float32)) for n in noise])
        if op == "maybe_reject":
            real_rand = torch.rand
            if u is not None:
                torch.rand = lambda *a, **k: torch.tensor(u, dtype=torch.float32)
            try:
                return opt.maybe_reject(args[0])
            finally:
                torch.rand = real_rand
        if op == "delta_energy":
            return opt.delta_energy(args[0], args[1])
        return getattr(opt, op)(**kwargs)

    def _cat(self, tensors):
        return np.concatenate([t.detach().reshape(-1).cpu().numpy() for t in tensors])

    def p_flat(self):
        return self._cat(self.params)

    def m_flat(self):
        out = []
        for p in self.params:
            m = self.opt.state[p].get("momentum_buffer")
            out.append(m if m is not None else self.torch.zeros_like(p))
        return self._cat(out)

    def scalars(self):
        from bnn_priors_b200 import _native as N
        st = self.opt.state
        keys = ("preconditioner", "est_temperature", "est_config_temp", "delta_energy",
                "prev_new_momentum_delta")
        out = {k: [st[p].get(k) for p in self.params] for k in keys}
        (fg,) = self.opt.flat_groups
        out["square_avg_mean"] = [float(v) for v in fg.fetch()[:, N.S_SQ_MEAN]]
        return out
