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
    from .evaluate import evaluate_model as fast_eval
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


def install(reference_package: str = "bnn_priors", evaluate: bool = False) -> None:
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
            _saved.setdefault("sub", {})[sub] = {n: getattr(m, n) for n in names}
            for n in names:
                setattr(m, n, getattr(fast, n))
    if evaluate:
        _install_evaluate(reference_package)


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
    for sub, fn in _saved.get("evaluate", {}).items():
        m = sys.modules.get(f"{pkg}.{sub}")
        if m is not None:
            m.evaluate_model = fn
    _saved.clear()
