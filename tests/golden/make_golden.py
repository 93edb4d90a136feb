#!/usr/bin/env python
"""Generate the golden call traces under tests/golden/ by running the UNMODIFIED
reference samplers (imported read-only from /root/reference) on CPU in fp32.

Run from the repo root, in the build container only (the GPU box has no
/root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

A trace is the list of calls made on one reference sampler object.  For every
call we store what went in (hyper-parameters at call time, p.grad of every
tensor, the N(0,1) tensors the call drew through torch.randn_like, the uniform
maybe_reject drew through torch.rand) and what came out (return value, every
parameter and momentum_buffer afterwards, the per-tensor scalars the reference
keeps in optimizer.state).  tests/replay.py feeds the same inputs to the oracle
and to the CUDA samplers and compares the outputs.

Nothing here is on the product path.
"""
import contextlib
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("BNNP_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REFERENCE)

from bnn_priors import mcmc as ref_mcmc  # noqa: E402


class Recorder:
    """Logs every public call made on a reference sampler."""

    STEP_OPS = ("initial_step", "step", "final_step")

    def __init__(self, opt, kind, ctor, extra_meta=None):
        self.opt = opt
        self.kind = kind
        self.arrays = []
        self.events = []
        self.params = [p for g in opt.param_groups for p in g["params"]]
        self.meta = dict(sampler=kind, ctor=ctor,
                         shapes=[list(p.shape) for p in self.params])
        self.meta.update(extra_meta or {})
        self._last_grad = None
        self.grad_prior_fn = None    # optional: () -> flat fp32 prior part of p.grad
        self.record_initial()

    # -- array pool ---------------------------------------------------------
    def _put(self, arr):
        self.arrays.append(np.ascontiguousarray(arr, dtype=np.float32))
        return len(self.arrays) - 1

    def _flat(self, tensors):
        return np.concatenate([t.detach().reshape(-1).numpy().astype(np.float32) for t in tensors]) \
            if tensors else np.zeros(0, np.float32)

    def _mom(self):
        out = []
        for p in self.params:
            m = self.opt.state[p].get("momentum_buffer")
            out.append(m if m is not None else torch.zeros_like(p))
        return out

    def record_initial(self):
        self.meta["p0"] = self._put(self._flat(self.params))

    def _scalars(self):
        keys = ("preconditioner", "est_temperature", "est_config_temp",
                "delta_energy", "prev_new_momentum_delta")
        out = {}
        for k in keys:
            vals = [self.opt.state[p].get(k, None) for p in self.params]
            out[k] = [None if v is None else float(v) for v in vals]
        out["square_avg_mean"] = [
            float(self.opt.state[p]["square_avg"].double().mean()) if "square_avg" in self.opt.state[p] else None
            for p in self.params]
        return out

    def _group(self):
        g = self.opt.param_groups[0]
        return {k: float(g[k]) for k in ("lr", "num_data", "momentum", "temperature",
                                         "rmsprop_alpha", "rmsprop_eps")}

    @contextlib.contextmanager
    def _capture_rng(self, noise_out, unif_out):
        real_randn_like, real_rand = torch.randn_like, torch.rand

        def randn_like(t, *a, **k):
            z = real_randn_like(t, *a, **k)
            noise_out.append(z.detach().clone())
            return z

        def rand(*a, **k):
            u = real_rand(*a, **k)
            unif_out.append(float(u))
            return u
        torch.randn_like, torch.rand = randn_like, rand
        try:
            yield
        finally:
            torch.randn_like, torch.rand = real_randn_like, real_rand

    # -- the recorded API -----------------------------------------------------
    def call(self, op, *args, **kwargs):
        ev = dict(op=op, kwargs={k: v for k, v in kwargs.items()}, group=self._group(),
                  precond=[float(self.opt.state[p].get("preconditioner", 1.0)) for p in self.params])
        if op in self.STEP_OPS:
            grads = self._flat([p.grad for p in self.params])
            if self._last_grad is None or not np.array_equal(grads, self._last_grad):
                ev["grad"] = self._put(grads)
                self._last_grad = grads
                if self.grad_prior_fn is not None:
                    ev["grad_prior"] = self._put(self.grad_prior_fn())
            else:
                ev["grad"] = "same"
        if op == "delta_energy":
            ev["args"] = [float(a.detach()) if isinstance(a, torch.Tensor) else float(a) for a in args]
        if op == "maybe_reject":
            ev["args"] = [float(args[0])]
        noise, unif = [], []
        with self._capture_rng(noise, unif):
            out = getattr(self.opt, op)(*args, **kwargs)
        if noise:
            ev["noise"] = self._put(self._flat(noise))
        if unif:
            assert len(unif) == 1
            ev["u"] = unif[0]
        if op == "delta_energy":
            ev["out"] = float(out)
        elif op == "maybe_reject":
            ev["out"] = [bool(out[0]), float(out[1])]
            # after a reject p.grad is the restored one
            self._last_grad = self._flat([p.grad for p in self.params])
        if op not in ("delta_energy", "update_preconditioner"):
            ev["p"] = self._put(self._flat(self.params))
            ev["m"] = self._put(self._flat(self._mom()))
        ev["scalars"] = self._scalars()
        self.events.append(ev)
        return out

    def set_group(self, **kv):
        for g in self.opt.param_groups:
            g.update(kv)

    def set_preconditioners(self, values):
        for p, v in zip(self.params, values):
            self.opt.state[p]["preconditioner"] = float(v)
        self.events.append(dict(op="set_preconditioner", values=[float(v) for v in values]))

    def save(self, name):
        self.meta["events"] = self.events
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, meta=np.frombuffer(json.dumps(self.meta).encode(), dtype=np.uint8),
                            **{f"a{i}": a for i, a in enumerate(self.arrays)})
        print(f"{name}: {len(self.events)} events, {sum(a.size for a in self.arrays)} floats, "
              f"{os.path.getsize(path) / 1024:.0f} KiB")


def make_params(shapes, scale=1.0):
    return [torch.nn.Parameter(torch.randn(*s) * scale) for s in shapes]


def make_closure(params, mean=0.3, std=0.7, quartic=0.05, num_data=1.0):
    """Anharmonic potential per data point; returns the potential like the
    reference's test closures do."""
    def closure():
        for p in params:
            p.grad = None
        u = sum((0.5 * ((p - mean) / std) ** 2).sum() + quartic * ((p - mean) ** 4).sum() for p in params)
        u = u / num_data
        u.backward()
        return u
    return closure


# -- scenarios ------------------------------------------------------------------
def golden_sgld():
    torch.manual_seed(1001)
    shapes = [(257,), (8, 5), (5,), (3, 2, 2, 2)]
    params = make_params(shapes)
    ctor = dict(lr=1 / 256, num_data=7.0, momentum=0.9, temperature=0.75)
    rec = Recorder(ref_mcmc.SGLD(params, **ctor), "SGLD", ctor)
    closure = make_closure(params, num_data=7.0)
    rec.set_preconditioners([torch.rand(()).item() + 0.2 for _ in params])
    rec.call("sample_momentum")
    for i in range(12):
        rec.set_group(lr=(1 / 256) * (0.5 * (math.cos(math.pi * i / 20) + 1)))
        if i in (5, 6):
            rec.set_group(temperature=0.0)         # descent: no noise is drawn
        else:
            rec.set_group(temperature=0.75)
        closure()
        rec.call("step", calc_metrics=(i % 3 == 0))
    rec.call("update_preconditioner")
    rec.call("sample_momentum", keep=0.5)
    for i in range(4):
        closure()
        rec.call("initial_step" if i == 0 else "step", calc_metrics=True)
    closure()
    rec.call("final_step", calc_metrics=True)
    rec.save("sgld_trace")


def golden_sgld_nomomentum():
    torch.manual_seed(1002)
    shapes = [(130,), (4, 4)]
    params = make_params(shapes)
    ctor = dict(lr=1 / 128, num_data=3.0, momentum=0.0, temperature=1.0)
    rec = Recorder(ref_mcmc.SGLD(params, **ctor), "SGLD", ctor)
    closure = make_closure(params, num_data=3.0)
    rec.call("sample_momentum")
    for i in range(6):
        if i == 4:
            rec.set_group(temperature=0.0)
        closure()
        rec.call("step", calc_metrics=True)
    rec.call("update_preconditioner")
    closure()
    rec.call("step", calc_metrics=False)
    rec.save("sgld_nomomentum_trace")


def _mh_cycles(rec, closure, n_cycles, mh_freq, hmc, lr0):
    """The call pattern of testing/test_verlet_sgld.py:90-118 / test_hmc.py:80-104."""
    prev_u = None
    n_rej = 0
    for step in range(n_cycles * mh_freq + 1):
        if step % mh_freq == 0:
            if step != 0:
                u = closure().item()
                rec.call("final_step", calc_metrics=True)
                de = rec.call("delta_energy", prev_u, u)
                rej, _ = rec.call("maybe_reject", de)
                n_rej += int(rej)
                if step == n_cycles * mh_freq:
                    break
            if hmc:
                rec.call("sample_momentum")
            rec.set_group(lr=lr0 * (1 + 0.1 * (step // mh_freq % 3)))
            prev_u = closure().item()
            rec.call("initial_step", save_state=True, calc_metrics=(step % 8 == 0))
        else:
            closure()
            rec.call("step", calc_metrics=(step % 3 == 0))
    return n_rej


def golden_verlet():
    torch.manual_seed(1003)
    shapes = [(300,), (16, 12), (12,), (7,)]
    params = make_params(shapes)
    ctor = dict(lr=1 / 6, num_data=1.0, momentum=0.9, temperature=0.75)
    rec = Recorder(ref_mcmc.VerletSGLD(params, **ctor), "VerletSGLD", ctor)
    closure = make_closure(params)
    rec.set_preconditioners([(torch.rand(()).item() + 0.2) / 2 for _ in params])
    rec.call("sample_momentum")
    n_rej = _mh_cycles(rec, closure, n_cycles=12, mh_freq=4, hmc=False, lr0=1 / 6)
    rec.call("update_preconditioner")
    # a descent stretch (temperature 0: noise is still drawn, maybe_reject never rejects)
    rec.set_group(temperature=0.0)
    u0 = closure().item()
    rec.call("initial_step", save_state=True, calc_metrics=True)
    closure()
    rec.call("step", calc_metrics=True)
    u1 = closure().item()
    rec.call("final_step", calc_metrics=True)
    de = rec.call("delta_energy", u0, u1)
    rec.call("maybe_reject", de)
    print("  verlet rejects:", n_rej)
    assert 2 <= n_rej <= 10, "want both decisions represented"
    rec.save("verlet_trace")


def golden_hmc():
    torch.manual_seed(1004)
    shapes = [(200,), (10, 9), (9,)]
    params = make_params(shapes)
    ctor = dict(lr=1.0, num_data=1.0)
    rec = Recorder(ref_mcmc.HMC(params, **ctor), "HMC", ctor)
    closure = make_closure(params)
    rec.set_preconditioners([(torch.rand(()).item() + 0.2) / 1.4 for _ in params])
    n_rej = _mh_cycles(rec, closure, n_cycles=10, mh_freq=5, hmc=True, lr0=1.0)
    print("  hmc rejects:", n_rej)
    assert 2 <= n_rej <= 8
    rec.save("hmc_trace")


def _golden_runner_case(tag, runner_cls, sampler_name, prior_w, wparams, runner_kw, out_name):
    """One reference runner driving a recording sampler on a reference ClassificationDenseNet,
    synthetic data.  Pins the call ORDER the runner produces and the prior's contribution to
    p.grad."""
    from bnn_priors import prior as ref_prior
    from bnn_priors.models import ClassificationDenseNet

    torch.manual_seed(1005)
    n, din, width, dout = 96, 20, 16, 4
    x = torch.rand(n, din)
    y = torch.randint(0, dout, (n,))
    model = ClassificationDenseNet(din, dout, width, depth=3, prior_w=prior_w,
                                   weight_prior_params=wparams)
    ds = torch.utils.data.TensorDataset(x, y)
    dl = torch.utils.data.DataLoader(ds, batch_size=32, shuffle=True)
    dl_test = torch.utils.data.DataLoader(ds, batch_size=96)

    class NullMetrics:
        def add_scalar(self, *a, **k): pass
        def flush(self, *a, **k): pass

    holder = {}
    prior_of = {id(pm.p): pm for _, pm in ref_prior.named_priors(model)}
    KIND = {"Normal": 1, "Laplace": 2, "StudentT": 3}

    Base = getattr(ref_mcmc, sampler_name)

    class Recording(Base):
        """Routes the runner's calls through the Recorder."""
        def __init__(self, params, **kw):
            super().__init__(params, **kw)
            ps = [p for g in self.param_groups for p in g["params"]]
            specs = []
            for p in ps:
                pm = prior_of[id(p)]
                specs.append(dict(kind=KIND[type(pm).__name__], loc=float(pm.loc),
                                  scale=float(pm.scale),
                                  df=float(getattr(pm, "df", torch.tensor(3.0)))))
            rec = Recorder(self, sampler_name, {k: float(v) for k, v in kw.items()},
                           extra_meta=dict(priors=specs, runner=runner_cls.__name__))

            def grad_prior():
                with torch.enable_grad():
                    gs = torch.autograd.grad(-model.log_prior() / kw["num_data"], ps)
                return np.concatenate([g.reshape(-1).numpy() for g in gs]).astype(np.float32)
            rec.grad_prior_fn = grad_prior
            holder["rec"] = rec
            self._in_call = False

        def _wrap(name):
            def f(self, *a, **k):
                if getattr(self, "_in_call", True):
                    return getattr(Base, name)(self, *a, **k)
                self._in_call = True
                try:
                    return holder["rec"].call(name, *a, **k)
                finally:
                    self._in_call = False
            return f
        for _n in ("sample_momentum", "initial_step", "step", "final_step",
                   "delta_energy", "maybe_reject", "update_preconditioner"):
            if hasattr(Base, _n):
                locals()[_n] = _wrap(_n)

    class DetGenerator(torch.Generator):
        def seed(self):          # inference_reject.py:72 would use OS entropy
            self.manual_seed(4242)
            return 4242

    saved_cls, saved_gen = getattr(ref_mcmc, sampler_name), torch.Generator
    setattr(ref_mcmc, sampler_name, Recording)
    torch.Generator = DetGenerator
    try:
        runner = runner_cls(model=model, dataloader=dl, dataloader_test=dl_test,
                            metrics_saver=NullMetrics(), model_saver=None, **runner_kw)
        # the constructor's update_preconditioner happened before recording
        runner.run(progressbar=False)
    finally:
        setattr(ref_mcmc, sampler_name, saved_cls)
        torch.Generator = saved_gen
    rec = holder["rec"]
    ops = [e["op"] for e in rec.events]
    print("  runner", tag, "ops:", {o: ops.count(o) for o in sorted(set(ops))},
          "rejects:", sum(1 for e in rec.events if e["op"] == "maybe_reject" and e["out"][0]))
    rec.save(out_name)


def golden_runner():
    """The reference's own VerletSGLDRunnerReject (inference_reject.py:11-176) with Normal /
    Laplace / StudentT priors (the analogs of BASELINE configs 2-4)."""
    from bnn_priors import prior as ref_prior
    from bnn_priors import inference_reject
    kw = dict(epochs_per_cycle=4, warmup_epochs=1, sample_epochs=2, learning_rate=0.02, skip=1, metrics_skip=2,
              temperature=1.0, momentum=0.9, cycles=2, precond_update=2, reject_samples=True)
    for tag, prior_w, wparams in (("normal", ref_prior.Normal, {}),
                                  ("laplace", ref_prior.Laplace, {}),
                                  ("studentt", ref_prior.StudentT, {"df": 3.0})):
        _golden_runner_case(tag, inference_reject.VerletSGLDRunnerReject, "VerletSGLD", prior_w, wparams, kw,
                            f"runner_verlet_{tag}_trace")


def golden_runner_more():
    """HMCRunnerReject (inference_reject.py:182-189; BASELINE config 5: leapfrog trajectories with
    an M-H test per sampling epoch) and the plain SGLDRunner (inference.py:9-294; config 1)."""
    from bnn_priors import prior as ref_prior
    from bnn_priors import inference, inference_reject
    _golden_runner_case("hmc_normal", inference_reject.HMCRunnerReject, "HMC", ref_prior.Normal, {},
                        dict(epochs_per_cycle=3, warmup_epochs=1, sample_epochs=2, learning_rate=0.3, skip=1,
                             metrics_skip=2, temperature=1.0, momentum=1.0, cycles=2, precond_update=2,
                             reject_samples=True),
                        "runner_hmc_normal_trace")
    _golden_runner_case("sgld_normal", inference.SGLDRunner, "SGLD", ref_prior.Normal, {},
                        dict(epochs_per_cycle=4, warmup_epochs=1, sample_epochs=2, learning_rate=0.02, skip=1,
                             metrics_skip=2, temperature=1.0, momentum=0.9, cycles=2, precond_update=2),
                        "runner_sgld_normal_trace")


def golden_priors():
    """log_prob and its gradient for the three fused priors, straight from the
    reference Prior classes (prior/loc_scale.py) + autograd."""
    from bnn_priors import prior as ref_prior
    torch.manual_seed(1006)
    out = {}
    cases = [("normal", ref_prior.Normal, dict(loc=0.0, scale=0.0505)),
             ("normal_shift", ref_prior.Normal, dict(loc=1.0, scale=2.0)),
             ("laplace", ref_prior.Laplace, dict(loc=0.0, scale=0.3)),
             ("laplace_shift", ref_prior.Laplace, dict(loc=-0.5, scale=1.7)),
             ("studentt", ref_prior.StudentT, dict(loc=0.0, scale=0.2, df=3.0)),
             ("studentt_df7", ref_prior.StudentT, dict(loc=0.25, scale=1.5, df=7.0)),
             # the other elementwise priors with constant hyper-parameters
             ("cauchy", ref_prior.Cauchy, dict(loc=0.0, scale=0.05)),
             ("cauchy_shift", ref_prior.Cauchy, dict(loc=-1.0, scale=2.5)),
             ("gennorm", ref_prior.GenNorm, dict(loc=0.0, scale=0.4, beta=0.5)),
             ("gennorm_b15", ref_prior.GenNorm, dict(loc=0.5, scale=1.2, beta=1.5)),
             ("lognormal", ref_prior.LogNormal, dict(loc=-1.0, scale=0.2)),
             ("uniform", ref_prior.Uniform, dict(low=-2.0, high=3.0)),
             ("improper", ref_prior.Improper, dict(loc=0.0, scale=1.0)),
             ("doublegamma", ref_prior.DoubleGamma, dict(loc=0.0, scale=0.3, concentration=1.7)),
             ("doublegamma_c05", ref_prior.DoubleGamma, dict(loc=0.2, scale=2.0, concentration=0.5))]
    KIND = {"Normal": 1, "Laplace": 2, "StudentT": 3, "Cauchy": 4, "GenNorm": 5, "LogNormal": 6, "Uniform": 7,
            "Improper": 8, "DoubleGamma": 9}
    meta = []
    for name, cls, kw in cases:
        pm = cls(torch.Size([513]), **kw)
        loc = kw.get("loc", 0.0)
        scale = kw.get("scale", 1.0)
        smooth_at_loc = cls.__name__ in ("Normal", "Laplace", "StudentT", "Cauchy", "LogNormal", "Uniform", "Improper")
        with torch.no_grad():
            pm.p.copy_(torch.randn(513) * 3 * scale + loc)
            if smooth_at_loc:
                pm.p[0] = loc                       # the kink of the Laplace density
            pm.p[1] = loc + 50 * scale              # far tail
        lp = pm.log_prob()
        if isinstance(lp, float) or not lp.requires_grad:   # Improper: the float 0.0; Uniform: a constant
            lp_val, g = float(lp), torch.zeros(513)
        else:
            lp_val = float(lp)
            (g,) = torch.autograd.grad(lp, pm.p, allow_unused=True)
            g = torch.zeros(513) if g is None else g
        out[name + "_p"] = pm.p.detach().numpy().copy()
        out[name + "_grad"] = g.numpy().copy()
        row = dict(name=name, kind=KIND[cls.__name__], log_prob=lp_val, loc=loc, scale=scale)
        if "df" in kw:
            row["df"] = kw["df"]
        if "beta" in kw:
            row["df"] = kw["beta"]                  # third hyper-parameter slot
        if "concentration" in kw:
            row["df"] = kw["concentration"]
        if cls.__name__ == "Uniform":
            row.update(loc=kw["low"], scale=kw["high"] - kw["low"])
        meta.append(row)
    np.savez_compressed(os.path.join(HERE, "priors.npz"),
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **out)
    print("priors:", [m["name"] for m in meta])


def golden_hier_priors():
    """SURVEY 8f N4, second half: priors whose scale is itself a sampled scalar with its
    own prior (prior/hierarchical.py, prior/empirical_bayes.py).  For each class: the
    weights p, the unconstrained hyper-parameter u (= scale_prior.p), the total
    log-density  Prior.log_prob() + scale_prior.log_prob()  -- both are summed by
    AbstractModel.log_prior (models/base.py:25-30, prior/base.py:79-81) -- and its
    autograd gradients w.r.t. p and u, from the unmodified reference classes."""
    from bnn_priors import prior as ref_prior
    torch.manual_seed(2024)
    H_GAMMA, H_UNIFORM, H_HALFCAUCHY, H_IMPROPER = 10, 11, 12, 13
    cases = [
        # name, class, ctor kwargs, weight kind, (hyper kind, a, b), u values tried
        ("normal_gamma", ref_prior.NormalGamma, dict(loc=0.0, scale=0.8, rate=1.5), 1, (H_GAMMA, 0.8, 1.5)),
        ("normal_uniform", ref_prior.NormalUniform, dict(loc=0.3, scale=0.6), 1, (H_UNIFORM, 0.0, 1.2)),
        ("horseshoe", ref_prior.Horseshoe, dict(loc=0.0, scale=0.5, hyperscale=2.0), 1, (H_HALFCAUCHY, 2.0, 0.5)),
        ("laplace_gamma", ref_prior.LaplaceGamma, dict(loc=-0.2, scale=1.3, rate=0.7), 2, (H_GAMMA, 1.3, 0.7)),
        ("laplace_uniform", ref_prior.LaplaceUniform, dict(loc=0.0, scale=0.4), 2, (H_UNIFORM, 0.0, 0.8)),
        ("studentt_gamma", ref_prior.StudentTGamma, dict(loc=0.0, scale=0.9, rate=1.0, df=2), 3, (H_GAMMA, 0.9, 1.0)),
        ("studentt_uniform", ref_prior.StudentTUniform, dict(loc=0.1, scale=0.7, df=5), 3, (H_UNIFORM, 0.0, 1.4)),
        ("normal_empirical", ref_prior.NormalEmpirical, dict(loc=0.0, scale=0.3), 1, (H_IMPROPER, 0.0, 1.0)),
        ("laplace_empirical", ref_prior.LaplaceEmpirical, dict(loc=0.5, scale=1.1), 2, (H_IMPROPER, 0.0, 1.0)),
    ]
    out, meta = {}, []
    for name, cls, kw, wkind, (hkind, ha, hb) in cases:
        for j, du in enumerate((0.0, 0.9, -1.3)):
            pm = cls(torch.Size([513]), **kw)
            sp = pm.scale                                   # the hyper-prior module (a Prior with shape [])
            assert isinstance(sp, ref_prior.Prior) and sp.p.numel() == 1
            with torch.no_grad():
                sp.p.add_(du)
                s_now = float(sp())
                pm.p.copy_(torch.randn(513) * 2 * s_now + kw["loc"])
                pm.p[1] = kw["loc"] + 30 * s_now
            lp = pm.log_prob() + sp.log_prob()
            g_p, g_u = torch.autograd.grad(lp, [pm.p, sp.p])
            tag = f"{name}_{j}"
            out[tag + "_p"] = pm.p.detach().numpy().copy()
            out[tag + "_grad_p"] = g_p.numpy().copy()
            meta.append(dict(name=tag, kind=wkind, loc=kw["loc"], df=float(kw.get("df", 3.0)), hyper_kind=hkind,
                             hyper_a=ha, hyper_b=hb, u=float(sp.p), scale=s_now, log_prob=float(lp),
                             log_prob_weights=float(pm.log_prob()), log_prob_hyper=float(sp.log_prob()),
                             grad_u=float(g_u)))
    np.savez_compressed(os.path.join(HERE, "hier_priors.npz"),
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **out)
    print("hier priors:", [(m["name"], round(m["scale"], 4), round(m["grad_u"], 3)) for m in meta])


if __name__ == "__main__":
    torch.set_num_threads(1)
    if os.environ.get("ONLY"): globals()[os.environ["ONLY"]](); sys.exit(0)
    golden_sgld()
    golden_sgld_nomomentum()
    golden_verlet()
    golden_hmc()
    golden_runner()
    golden_runner_more()
    golden_priors()
    golden_hier_priors()
