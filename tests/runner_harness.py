"""Run one of the reference's runner classes on a BASELINE config with synthetic data, with the
sampler classes wrapped by tests/runner_tape.py.  Shared by the CPU self-check of the harness
and the GPU parity tests.  Test infrastructure only."""
from __future__ import annotations

import contextlib
import importlib

import torch

import refenv
import runner_tape as RT

@contextlib.contextmanager
def deterministic_generators(base_seed: int = 20240):
    """The Reject runners seed the minibatch order from OS entropy (`generator.seed()`,
    inference_reject.py:68-72), so every run sees another trajectory.  For reproducible tests every
    `torch.Generator()` created while this context is active answers `seed()` with a counter-based
    value instead (test-side; the reference's code is untouched)."""
    real = torch.Generator
    counter = {"n": 0}

    class CountingSeedGenerator(real):
        def seed(self):
            counter["n"] += 1
            s = base_seed + counter["n"]
            self.manual_seed(s)
            return s
    torch.Generator = CountingSeedGenerator
    try:
        yield
    finally:
        torch.Generator = real


RUNNERS = {
    "SGLD": ("bnn_priors.inference", "SGLDRunner"),
    "VerletSGLD": ("bnn_priors.inference", "VerletSGLDRunner"),
    "OurHMC": ("bnn_priors.inference", "HMCRunner"),
    "VerletSGLDReject": ("bnn_priors.inference_reject", "VerletSGLDRunnerReject"),
    "HMCReject": ("bnn_priors.inference_reject", "HMCRunnerReject"),
    "SGLDReject": ("bnn_priors.inference_reject", "SGLDRunnerReject"),
}


def make_runner(inference: str, config: str, device, *, n_train=512, n_test=256, batch_size=128, cycles=2,
                descent=1, warmup=1, sample_epochs=2, lr=5e-4, temperature=1.0, momentum=0.994,
                metrics_skip=10, reject_samples=True, model_saver=None, metrics_saver=None, seed=0,
                runner_kwargs=None):
    refenv.exp_utils()
    mod, name = RUNNERS[inference]
    cls = getattr(importlib.import_module(mod), name)
    data = refenv.synthetic_dataset(refenv.CONFIGS[config]["data"], device, n_train, n_test, seed=seed)
    model = refenv.build_model(config, data, seed=seed)
    dl = torch.utils.data.DataLoader(data.norm.train, batch_size=batch_size, shuffle=True, drop_last=False)
    dl_test = torch.utils.data.DataLoader(data.norm.test, batch_size=batch_size, shuffle=False, drop_last=False)
    if "HMC" in inference:
        descent, momentum, temperature = 0, 1.0, 1.0
    if inference in ("SGLD", "SGLDReject"):
        reject_samples = False
    runner = cls(model=model, dataloader=dl, dataloader_test=dl_test,
                 epochs_per_cycle=descent + warmup + sample_epochs, warmup_epochs=warmup,
                 sample_epochs=sample_epochs, learning_rate=lr, skip=1, metrics_skip=metrics_skip,
                 sampling_decay="cosine", cycles=cycles, temperature=temperature, momentum=momentum,
                 precond_update=1, metrics_saver=metrics_saver or refenv.FakeMetrics(), model_saver=model_saver,
                 reject_samples=reject_samples, **(runner_kwargs or {}))
    return runner, model, data


def prior_grad_fn_for(model, num_data_ref):
    """() -> the prior's share of p.grad, d/dp [ -log_prior / N ], at the model's current parameters
    (models/base.py:72-77 as the runner backpropagates it)."""
    params = [p for _, p in model.named_parameters()]

    def fn():
        with torch.enable_grad():
            lp = model.log_prior() / -num_data_ref()
            gs = torch.autograd.grad(lp, params, allow_unused=True)
        return [torch.zeros_like(p) if g is None else g for p, g in zip(params, gs)]
    return fn


def record_run(inference, config, device, seed=0, with_prior_grads=False, **kw):
    """run A: the reference runner with the reference's eager sampler; returns (tape, runner)."""
    refenv.setup()
    ref_mcmc = importlib.import_module("bnn_priors.mcmc")
    tape = RT.Tape()
    torch.manual_seed(seed)
    runner, model, _ = make_runner(inference, config, device, seed=seed, **kw)
    pg = prior_grad_fn_for(model, lambda: runner.eff_num_data) if with_prior_grads else None
    with RT.bound_sampler_classes(ref_mcmc, lambda c: RT.recording_class(c, tape, pg)), deterministic_generators():
        runner.run(progressbar=False)
    return tape, runner


def replay_run(inference, config, device, tape, seed=0, fused_prior=False, before_run=None, **kw):
    """run B: the same runner with whatever classes `bnn_priors.mcmc` currently names (the reference's
    again for the harness self-check, the B200 ones after overlay.install()); returns (report, runner)."""
    refenv.setup()
    ref_mcmc = importlib.import_module("bnn_priors.mcmc")
    report = RT.Report()
    tape.cursor = 0
    torch.manual_seed(seed)
    runner, model, _ = make_runner(inference, config, device, seed=seed, **kw)
    if before_run is not None:
        before_run(runner, model)
    with RT.bound_sampler_classes(ref_mcmc, lambda c: RT.replaying_class(c, tape, report, fused_prior)), \
            deterministic_generators():
        runner.run(progressbar=False)
    assert tape.cursor == len(tape.events), "the replayed run made fewer sampler calls than the recorded one"
    return report, runner


def run_train_bnn(log_dir, device="try_cuda", n_train=512, n_test=256, **config_updates):
    """experiments/train_bnn.py's `main` (train_bnn.py:155-259), unmodified, through the sacred / h5py
    test shims, on synthetic data of the named data set's shape.  Returns (run, run directory)."""
    mod = refenv.load_train_bnn()
    eu = refenv.exp_utils()
    real_get_data = eu.get_data

    def get_data(data, device):
        kind = "cifar10" if data.startswith("cifar10") else "mnist"
        return refenv.synthetic_dataset(kind, device, n_train, n_test)

    cfg = dict(data="mnist", model="classificationdensenet", inference="VerletSGLDReject", width=50, depth=3,
               n_samples=4, cycles=2, warmup=1, burnin=0, skip=1, metrics_skip=2, skip_first=1, batch_size=128,
               reject_samples=True, save_samples=True, progressbar=False, log_dir=str(log_dir), device=device)
    cfg.update(config_updates)
    mod.ex.observers.clear()
    eu.get_data = get_data
    real_dp = torch.nn.DataParallel

    class OneDeviceDataParallel(real_dp):
        "one process per GPU (SURVEY 8e): the wrapper exp_utils.py:229 adds stays on the model's device"

        def __init__(self, module, device_ids=None, output_device=None, dim=0):
            p = next(module.parameters())
            super().__init__(module, device_ids=[p.device.index or 0], output_device=output_device, dim=dim)
    eu.t.nn.DataParallel = OneDeviceDataParallel
    try:
        with contextlib.redirect_stdout(None), deterministic_generators():
            run = mod.ex.run(config_updates=cfg)
    finally:
        eu.get_data = real_get_data
        eu.t.nn.DataParallel = real_dp
    return run, run.observers[0].dir
