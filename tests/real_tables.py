"""Golden traces at the REAL segment tables of the BASELINE configs (tests/golden/model_shapes.json:
classificationdensenet 42,310 / classificationconvnet-laplace 47,560 / googleresnet student-t and
gaussian 272,474 / vwidth_resnet18-96 25,124,842 parameters), compact enough to commit.

Whole-array traces at these sizes would be hundreds of MB, so a compact trace stores

  * INPUTS as seeds: initial parameters, the likelihood part of every gradient and every N(0,1)
    tensor are regenerated from `torch.Generator` seeds (CPU generator, fixed call order); the
    Metropolis uniforms and the potentials are stored as numbers;
  * OUTPUTS as fingerprints: per event and tensor a strided sample of 16 elements of the parameters
    and of the momentum plus their first two moments in fp64, all per-tensor scalars, every delta
    energy and every accept / reject decision.

The prior's share of p.grad is not an input: the generator lets the reference differentiate its
own prior (torch.distributions + autograd, like prior/base.py:57-58), the replay either does the
same at the engine's own parameters (sampler only) or leaves it to the engine (prior fused into
the kernel / closed forms of the oracle).

tests/golden/make_real_golden.py records them from the UNMODIFIED reference samplers;
`replay_compact` feeds the same inputs to an engine of tests/replay.py (numpy oracle on the CPU,
CUDA samplers on the GPU) and returns the worst deviations.  Test infrastructure only.
"""
from __future__ import annotations

import json
import math
import os
from typing import List

import numpy as np
import torch

import replay as R

GOLDEN_DIR = R.GOLDEN_DIR
NSAMPLE = 16
GRAD_STD = 3e-3


def load_table(tag: str):
    with open(os.path.join(GOLDEN_DIR, "model_shapes.json")) as f:
        return json.load(f)[tag]["tensors"]


# name -> (table tag, sampler, constructor, script)
def _verlet_script(n_mid=3, cycles=2):
    ev = [dict(op="sample_momentum", kwargs={})]
    g = 0
    for c in range(cycles):
        ev.append(dict(op="initial_step", kwargs=dict(save_state=True, calc_metrics=(c == 0)), grad=g, lr_scale=1.0 - 0.2 * c))
        for i in range(n_mid):
            g += 1
            ev.append(dict(op="step", kwargs=dict(calc_metrics=(i == n_mid - 1)), grad=g))
        g += 1
        ev.append(dict(op="final_step", kwargs=dict(calc_metrics=True), grad=g))
        ev.append(dict(op="delta_energy", u0=2.3 + 0.01 * c, target=(3.0 if c else -0.5)))
        ev.append(dict(op="maybe_reject", u=(0.5 if c else 0.9)))
        ev.append(dict(op="update_preconditioner", kwargs={}))
    return ev


def _hmc_script(n_mid, cycles=2):
    ev = []
    g = 0
    for c in range(cycles):
        ev.append(dict(op="sample_momentum", kwargs={}))
        ev.append(dict(op="initial_step", kwargs=dict(save_state=True, calc_metrics=(c == 0)), grad=g))
        for i in range(n_mid):
            g += 1
            ev.append(dict(op="step", kwargs=dict(calc_metrics=(i % 25 == 24)), grad=g, skip_fp=(i % 10 != 9)))
        g += 1
        ev.append(dict(op="final_step", kwargs=dict(calc_metrics=True), grad=g))
        ev.append(dict(op="delta_energy", u0=2.3, target=(3.0 if c else -0.5)))
        ev.append(dict(op="maybe_reject", u=(0.5 if c else 0.9)))
        ev.append(dict(op="update_preconditioner", kwargs={}))
    return ev


def _sgld_script(n):
    ev = [dict(op="sample_momentum", kwargs={})]
    for i in range(n):
        ev.append(dict(op="step", kwargs=dict(calc_metrics=(i % 2 == 0)), grad=i, lr_scale=1.0 - 0.1 * i))
        if i == n - 2:
            ev.append(dict(op="update_preconditioner", kwargs={}))
    return ev


CASES = {
    # BASELINE.json config 2
    "real_densenet_verlet_gaussian": ("classificationdensenet_mnist_gaussian", "VerletSGLD",
                                      dict(lr=5e-4, num_data=60000.0, momentum=0.994, temperature=1.0), _verlet_script()),
    # config 3: laplace weights, temperature 0.1
    "real_convnet_verlet_laplace_T0.1": ("classificationconvnet_mnist_laplace", "VerletSGLD",
                                         dict(lr=5e-4, num_data=60000.0, momentum=0.994, temperature=0.1), _verlet_script()),
    # config 4: 65 tensors, 42 without a prior, StudentT on the last dense layer only
    "real_googleresnet_verlet_studentt": ("googleresnet_cifar10_studentt", "VerletSGLD",
                                          dict(lr=5e-4, num_data=50000.0, momentum=0.994, temperature=1.0), _verlet_script()),
    # config 5: HMC, 50 leapfrog steps
    "real_googleresnet_hmc50_gaussian": ("googleresnet_cifar10_gaussian", "HMC",
                                         dict(lr=2e-5, num_data=50000.0), _hmc_script(49)),
    # config 1 / the metric's workload
    "real_densenet_sgld_gaussian": ("classificationdensenet_mnist_gaussian", "SGLD",
                                    dict(lr=5e-4, num_data=60000.0, momentum=0.994, temperature=1.0), _sgld_script(6)),
    "real_resnet18w96_sgld_gaussian": ("vwidth_resnet18_w96_cifar10_gaussian", "SGLD",
                                       dict(lr=5e-4, num_data=50000.0, momentum=0.994, temperature=1.0), _sgld_script(3)),
    "real_resnet18w96_verlet_gaussian": ("vwidth_resnet18_w96_cifar10_gaussian", "VerletSGLD",
                                         dict(lr=5e-4, num_data=50000.0, momentum=0.994, temperature=1.0),
                                         _verlet_script(n_mid=1, cycles=1)),
}


class Inputs:
    "the seeded inputs of one case, regenerated on demand (CPU generator; fixed order of draws)"

    def __init__(self, name: str):
        self.name = name
        self.tag, self.sampler, self.ctor, self.script = CASES[name]
        self.table = load_table(self.tag)
        self.shapes = [tuple(t["shape"]) for t in self.table]
        self.sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self.base = sum(ord(c) for c in name) * 1000
        self.priors = [dict(kind=t["kind"], loc=t["loc"], scale=t["scale"], df=t["df"]) for t in self.table]

    def _randn(self, seed: int) -> List[torch.Tensor]:
        g = torch.Generator().manual_seed(self.base + seed)
        return [torch.randn(s, generator=g) for s in self.shapes]

    def p0(self) -> List[torch.Tensor]:
        return [z * (t["scale"] if t["kind"] else 0.5) + (0.0 if t["kind"] else 1.0)
                for z, t in zip(self._randn(1), self.table)]

    def lik_grad(self, k: int) -> List[torch.Tensor]:
        return [z * GRAD_STD for z in self._randn(100 + k)]

    def noise(self, event_index: int) -> List[torch.Tensor]:
        return self._randn(10_000 + event_index)

    def prior_grad(self, params: List[torch.Tensor], num_data: float) -> List[torch.Tensor]:
        """d/dp [-log p(p) / N] the way the reference gets it: torch.distributions + autograd
        (prior/base.py:57-58, prior/loc_scale.py:34-77, models/base.py:72-77)."""
        import torch.distributions as td
        out = []
        for p, t in zip(params, self.table):
            if t["kind"] == 0:
                out.append(torch.zeros_like(p))
                continue
            q = p.detach().clone().requires_grad_(True)
            loc = torch.tensor(t["loc"], device=p.device)
            scale = torch.tensor(t["scale"], device=p.device)
            if t["kind"] == 1:
                d = td.Normal(loc, scale)
            elif t["kind"] == 2:
                d = td.Laplace(loc, scale)
            elif t["kind"] == 3:
                d = td.StudentT(torch.tensor(t["df"], device=p.device), loc, scale)
            else:
                raise ValueError(t["kind"])
            (g,) = torch.autograd.grad(d.log_prob(q).sum() / -num_data, q)
            out.append(g)
        return out


def fingerprint(tensors: List[torch.Tensor]):
    """per tensor: NSAMPLE strided elements (fp32) + sum and sum of squares (fp64)"""
    samples, moments = [], []
    for t in tensors:
        f = t.detach().reshape(-1)
        idx = torch.linspace(0, f.numel() - 1, NSAMPLE, device=f.device).round().long()
        samples.append(f[idx].float().cpu().numpy())
        d = f.double()
        moments.append([float(d.sum()), float((d * d).sum())])
    return np.stack(samples), np.asarray(moments)


class CompactTrace:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.arrays = {k: z[k] for k in z.files if k != "meta"}
        self.events = self.meta["events"]
        self.inputs = Inputs(name)


class Report:
    def __init__(self):
        self.sample_err = 0.0        # strided samples of p / m, relative to the tensor's RMS
        self.moment_err = 0.0        # sum / sum of squares, relative to sqrt(n) RMS resp. n RMS^2
        self.scalar_err = {}
        self.de_term_err = 0.0
        self.decisions = self.decisions_equal = self.rejections = 0
        self.n_events = 0

    def bump(self, k, v):
        self.scalar_err[k] = max(self.scalar_err.get(k, 0.0), v)


def _compare_fp(rep: Report, got, want_s, want_m, sizes):
    gs, gm = got
    for i, n in enumerate(sizes):
        rms = math.sqrt(max(want_m[i][1], 1e-300) / n)
        rep.sample_err = max(rep.sample_err, float(np.max(np.abs(gs[i].astype(np.float64) - want_s[i]))) / max(rms, 1e-30))
        rep.moment_err = max(rep.moment_err, abs(gm[i][0] - want_m[i][0]) / max(math.sqrt(n) * rms, 1e-30),
                             abs(gm[i][1] - want_m[i][1]) / max(want_m[i][1], 1e-30))


class _TraceShim:
    "what the engine constructors of tests/replay.py read from a Trace"

    def __init__(self, inp: Inputs):
        self.sampler, self.ctor, self.shapes, self.sizes, self.priors = inp.sampler, dict(inp.ctor), inp.shapes, inp.sizes, inp.priors
        self.meta = {"p0": "p0"}
        self._p0 = np.concatenate([t.reshape(-1).numpy() for t in inp.p0()])

    def arr(self, idx):
        assert idx == "p0"
        return self._p0

    def split(self, flat):
        out, o = [], 0
        for n in self.sizes:
            out.append(np.asarray(flat[o:o + n]))
            o += n
        return out


def make_engine(name: str, kind: str, fused_prior: bool):
    inp = Inputs(name)
    shim = _TraceShim(inp)
    if kind == "oracle":
        return inp, R.OracleEngine(shim, fused_prior=fused_prior, dot_dtype=np.float64)
    return inp, R.CudaEngine(shim, fused_prior=fused_prior, foreign_grads=True)


def engine_params(engine) -> List[torch.Tensor]:
    if hasattr(engine, "params"):
        return [p.detach() for p in engine.params]
    return [torch.from_numpy(s.p.reshape(shape)) for s, shape in zip(engine.chain.segs, engine._shapes)]


def replay_compact(name: str, kind: str, fused_prior: bool) -> Report:
    trace = CompactTrace(name)
    inp, engine = make_engine(name, kind, fused_prior)
    engine._shapes = inp.shapes
    rep = Report()
    lr0 = inp.ctor["lr"]
    temperature = inp.ctor.get("temperature", 1.0)
    num_data = inp.ctor["num_data"]
    last_de = None
    metrics_pg_abs = None            # sum |p g| of the step that last computed est_config_temp
    for k, ev in enumerate(trace.events):
        op = ev["op"]
        rep.n_events += 1
        if "pg_abs" in ev and ev.get("kwargs", {}).get("calc_metrics", True):
            metrics_pg_abs = ev["pg_abs"]
        engine.set_group(dict(lr=lr0 * ev.get("lr_scale", 1.0), temperature=temperature))
        if op in R.STEP_OPS and ev.get("new_grad", True):
            g = inp.lik_grad(ev["grad"])
            if not fused_prior:
                params = engine_params(engine)
                pg = inp.prior_grad(params, num_data)
                g = [a.to(b.device) + b for a, b in zip(g, pg)]
            engine.set_grad([t.detach().cpu().numpy().reshape(-1) for t in g])
        noise = [t.numpy().reshape(-1) for t in inp.noise(k)] if ev.get("noise") else None
        args = []
        if op == "delta_energy":
            args = list(ev["potentials"])
        elif op == "maybe_reject":
            args = [last_de[1] if last_de is not None else ev["arg"]]
        out = engine.call(op, ev.get("kwargs", {}), args, noise, ev.get("u"))
        if op == "delta_energy":
            last_de = (ev["out"], out)
            terms = [abs(x) for x in ev["scalars"]["delta_energy"] if x is not None and math.isfinite(x)]
            scale = sum(terms) + abs((args[1] - args[0]) * num_data) + 1.0
            rep.de_term_err = max(rep.de_term_err, abs(out - ev["out"]) / scale)
        elif op == "maybe_reject":
            rep.decisions += 1
            rep.decisions_equal += int(bool(out[0]) == bool(ev["out"][0]))
            rep.rejections += int(bool(ev["out"][0]))
        if "fp" in ev:
            i = ev["fp"]
            params = engine_params(engine)
            _compare_fp(rep, fingerprint(params), trace.arrays[f"ps{i}"], trace.arrays[f"pm{i}"], inp.sizes)
            if inp.ctor.get("momentum", 1.0) > 0:
                if hasattr(engine, "params"):
                    ms = [engine.opt.state[p]["momentum_buffer"] for p in engine.params]
                else:
                    ms = [torch.from_numpy(s.m.reshape(shape)) for s, shape in zip(engine.chain.segs, inp.shapes)]
                _compare_fp(rep, fingerprint(ms), trace.arrays[f"ms{i}"], trace.arrays[f"mm{i}"], inp.sizes)
        if "scalars" not in ev:
            continue
        got = engine.scalars()
        for key, want in ev["scalars"].items():
            if key not in got:
                continue
            mags = [abs(x) for x in want if x is not None and math.isfinite(x)]
            for i, (a, b) in enumerate(zip(got[key], want)):
                if b is None or a is None or not math.isfinite(b):
                    continue
                if key in ("delta_energy", "prev_new_momentum_delta"):
                    scale = max(max(mags, default=0.0), 1e-3)
                elif key == "est_config_temp" and metrics_pg_abs is not None:
                    scale = max(abs(b), metrics_pg_abs[i] * num_data / inp.sizes[i], 1e-30)
                else:
                    scale = max(abs(b), 1e-30)
                rep.bump(key, abs(a - b) / scale)
    return rep
