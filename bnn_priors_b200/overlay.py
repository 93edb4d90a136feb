"""Make the reference's own code use the B200 samplers.

The reference's runners reach the samplers through the module attribute
`bnn_priors.mcmc` (`from . import mcmc` in inference.py:4 and inference_reject.py:4,
then `mcmc.SGLD(...)`, `mcmc.VerletSGLD(...)`, `mcmc.HMC(...)` at run time:
inference.py:89-94,298-302,368-374; inference_reject.py:12-16,183-198).  `install()`
re-binds those three names -- on the already imported module object and in
`sys.modules` -- so `experiments/train_bnn.py`, the runner classes and the reference's
tests drive the kernels in this package without any change to the reference.

`install(evaluate=True)` additionally re-binds `evaluate_model` (exp_utils.py:250-340;
imported by name into inference.py:6 and inference_reject.py:6, reached as
`exp_utils.evaluate_model` by experiments/train_bnn.py:148-149 and eval_bnn.py:62) to the
device-side evaluation of `bnn_priors_b200.evaluate`.

`install(sample_sink=True)` re-binds `exp_utils.HDF5ModelSaver` (the sample file writer train_bnn.py
constructs, :201-203) to `sample_sink.FlatSampleSaver`: one device-to-device snapshot + one asynchronous
device-to-host copy per sample instead of one blocking copy per tensor, same file.

`install(fuse_prior=True)` also takes the prior out of autograd under an UNCHANGED runner:
every runner class's `_make_optimizer` (inference.py:89,298,368; inference_reject.py:12,183,192)
is wrapped so that the sampler it returns gets `fuse_prior(self.model, sampler,
grad_max=self.grad_max)` (prior_fusion.py), and `_model_potential_and_grad`
(inference.py:215-223) is re-stated without its per-tensor `p.grad.clamp_` loop: the kernel
clamps likelihood + prior gradient together, exactly what the reference clamps, and only for
gradients that came through this method (the full-data pass of inference_reject.py:18-33 does
not clamp in the reference, and does not here).

    import bnn_priors_b200.overlay as overlay
    overlay.install()            # before or after `import bnn_priors`
    ...                          # run experiments/train_bnn.py's main, or a Runner
    overlay.uninstall()
"""
from __future__ import annotations

import importlib
import sys

_NAMES = ("SGLD", "VerletSGLD", "HMC")
_saved = {}


_EVAL_MODULES = ("exp_utils", "inference", "inference_reject")


def _install_evaluate(pkg: str) -> None:
    from . import evaluate as _ev
    from .evaluate import evaluate_model as fast_eval
    eu = sys.modules.get(f"{pkg}.exp_utils")
    if eu is not None and all(hasattr(eu, n) for n in ("ece", "ace", "rmsce")):
        _ev.set_calibration_metrics(eu.ece, eu.ace, eu.rmsce)      # exp_utils.py:13
    saved = _saved.setdefault("evaluate", {})
    for sub in _EVAL_MODULES:
        name = f"{pkg}.{sub}"
        try:
            m = sys.modules.get(name) or importlib.import_module(name)
        except ImportError:
            continue                       # e.g. exp_utils needs h5py / sacred; nothing to patch then
        if hasattr(m, "evaluate_model"):
            saved.setdefault(sub, m.evaluate_model)
            m.evaluate_model = fast_eval


_RUNNERS = (("inference", ("SGLDRunner", "VerletSGLDRunner", "HMCRunner")),
            ("inference_reject", ("VerletSGLDRunnerReject", "HMCRunnerReject", "SGLDRunnerReject")))


def _install_fuse_prior(pkg: str) -> None:
    import torch
    from .prior_fusion import fuse_prior
    saved = _saved.setdefault("runner_methods", [])

    def wrap_make_optimizer(orig):
        def _make_optimizer(self, params):
            opt = orig(self, params)
            if hasattr(opt, "flat_groups"):
                # clamp only gradients that came through _model_potential_and_grad (see below)
                self._bnnp_fused_prior = fuse_prior(self.model, opt, grad_max=self.grad_max, clamp="armed")
            return opt
        return _make_optimizer

    def _model_potential_and_grad(self, x, y):
        "inference.py:215-223; the +-grad_max clamp of :219-220 is applied by the kernel to likelihood + prior"
        fused = getattr(self, "_bnnp_fused_prior", None)
        if fused is None or not fused.groups:
            return _saved["potential_and_grad"](self, x, y)
        self.optimizer.zero_grad()
        loss, log_prior, potential, accs_batch, _ = self.model.split_potential_and_acc(x, y, self.eff_num_data)
        potential.backward()
        fused.arm_clamp(True)
        if torch.isnan(potential).item():
            raise ValueError("Potential is NaN")
        return loss, log_prior, potential, accs_batch.mean()

    def wrap_exact(orig):
        def _exact_model_potential_and_grad(self, dataloader):
            fused = getattr(self, "_bnnp_fused_prior", None)
            if fused is not None:
                fused.arm_clamp(False)          # inference_reject.py:18-33 does not clamp
            return orig(self, dataloader)
        return _exact_model_potential_and_grad

    for sub, classes in _RUNNERS:
        name = f"{pkg}.{sub}"
        try:
            m = sys.modules.get(name) or importlib.import_module(name)
        except ImportError:
            continue
        for cname in classes:
            cls = getattr(m, cname, None)
            if cls is None:
                continue
            if "_make_optimizer" in cls.__dict__:
                saved.append((cls, "_make_optimizer", cls.__dict__["_make_optimizer"]))
                cls._make_optimizer = wrap_make_optimizer(cls.__dict__["_make_optimizer"])
            if "_model_potential_and_grad" in cls.__dict__:
                _saved.setdefault("potential_and_grad", cls.__dict__["_model_potential_and_grad"])
                saved.append((cls, "_model_potential_and_grad", cls.__dict__["_model_potential_and_grad"]))
                cls._model_potential_and_grad = _model_potential_and_grad
            if "_exact_model_potential_and_grad" in cls.__dict__:
                saved.append((cls, "_exact_model_potential_and_grad", cls.__dict__["_exact_model_potential_and_grad"]))
                cls._exact_model_potential_and_grad = wrap_exact(cls.__dict__["_exact_model_potential_and_grad"])


def _install_sample_sink(pkg: str) -> None:
    """exp_utils.HDF5ModelSaver (exp_utils.py:409-487; constructed by experiments/train_bnn.py:201-203 as
    `exp_utils.HDF5ModelSaver(path, "w")`) -> sample_sink.FlatSampleSaver, same constructor arguments.
    HDF5Metrics, a subclass defined in the same module, keeps the original base class."""
    from .sample_sink import FlatSampleSaver
    name = f"{pkg}.exp_utils"
    try:
        m = sys.modules.get(name) or importlib.import_module(name)
    except ImportError:
        return
    _saved["model_saver"] = m.HDF5ModelSaver
    m.HDF5ModelSaver = FlatSampleSaver


def install(reference_package: str = "bnn_priors", evaluate: bool = False, fuse_prior: bool = False,
            sample_sink: bool = False) -> None:
    from . import mcmc as fast
    ref = importlib.import_module(reference_package + ".mcmc")
    if "classes" not in _saved:
        _saved["classes"] = {n: getattr(ref, n) for n in _NAMES}
        _saved["package"] = reference_package
    for n in _NAMES:
        setattr(ref, n, getattr(fast, n))
    # submodules: `from bnn_priors.mcmc.sgld import SGLD` style imports
    for sub, names in (("sgld", ("SGLD",)), ("verlet_sgld", ("VerletSGLD",)), ("hmc", ("HMC",))):
        m = sys.modules.get(f"{reference_package}.mcmc.{sub}")
        if m is not None:
            # (a repeated install() must not record the classes it put there itself as the originals)
            _saved.setdefault("sub", {}).setdefault(sub, {n: getattr(m, n) for n in names})
            for n in names:
                setattr(m, n, getattr(fast, n))
    if evaluate:
        _install_evaluate(reference_package)
    if fuse_prior and "runner_methods" not in _saved:
        _install_fuse_prior(reference_package)
    if sample_sink and "model_saver" not in _saved:
        _install_sample_sink(reference_package)


def uninstall() -> None:
    if "classes" not in _saved:
        return
    pkg = _saved["package"]
    ref = importlib.import_module(pkg + ".mcmc")
    for n, c in _saved["classes"].items():
        setattr(ref, n, c)
    for sub, names in _saved.get("sub", {}).items():
        m = sys.modules.get(f"{pkg}.mcmc.{sub}")
        for n, c in names.items():
            setattr(m, n, c)
    if "model_saver" in _saved:
        m = sys.modules.get(f"{pkg}.exp_utils")
        if m is not None:
            m.HDF5ModelSaver = _saved["model_saver"]
    for cls, name, fn in reversed(_saved.get("runner_methods", [])):
        setattr(cls, name, fn)
    for sub, fn in _saved.get("evaluate", {}).items():
        m = sys.modules.get(f"{pkg}.{sub}")
        if m is not None:
            m.evaluate_model = fn
    _saved.clear()
