"""The reference package for tests: importable from `oracle/_ref` (the snapshot
`oracle/make_ref.py` makes in the build container and that travels to the GPU box) or, in the
build container, from /root/reference -- with the test shims for the packages this image
lacks (gpytorch, h5py, sacred, pyro; tests/golden/_shims).  Test infrastructure only.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIMS = os.path.join(HERE, "golden", "_shims")
SNAPSHOT = os.path.join(ROOT, "oracle", "_ref")
CHECKOUT = os.environ.get("BNNP_REFERENCE", "/root/reference")


def reference_root() -> str:
    "directory that holds the reference's `bnn_priors/` package, or '' if there is none"
    for cand in (SNAPSHOT, CHECKOUT):
        if os.path.exists(os.path.join(cand, "bnn_priors", "mcmc", "sgld.py")):
            return cand
    return ""


def available() -> bool:
    return bool(reference_root())


def setup():
    """Put the shims and the reference on sys.path (idempotent) and return the imported
    `bnn_priors` package."""
    root = reference_root()
    if not root:
        raise RuntimeError("no reference available: run `python oracle/make_ref.py` in the build container")
    sys.dont_write_bytecode = True
    for p in (root, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    pkg = importlib.import_module("bnn_priors")
    importlib.import_module("bnn_priors.mcmc")
    return pkg


def exp_utils():
    setup()
    eu = importlib.import_module("bnn_priors.exp_utils")
    numpy2_metrics_compat(eu)
    return eu


def numpy2_metrics_compat(eu) -> None:
    """exp_utils.HDF5Metrics fills its caches with `arr[:] = np.nan` also for integer
    metrics and relies on numpy 1.18 casting NaN to -2**63 (setup.py pins numpy<1.19 for this,
    exp_utils.py:467 "int64 stores NaN as -2**63").  numpy 2 raises instead, so the two helper
    methods are re-stated here with the cast made explicit.  Environment adaptation of the
    metrics writer only; the sampler path is untouched."""
    if getattr(eu.HDF5Metrics, "_bnnp_numpy2", False):
        return

    def nan_of(dtype):
        return np.iinfo(dtype).min if np.issubdtype(dtype, np.integer) else np.nan

    def _scrub_cache(self):
        for v in self._cache.values():
            v[:] = nan_of(v.dtype)

    def _append(self, name, value, dtype):
        try:
            arr = self._cache[name]
        except KeyError:
            arr = self._cache[name] = np.empty(self.chunk_size, dtype=dtype)
            arr[:] = nan_of(arr.dtype)
        if np.issubdtype(arr.dtype, np.integer) and isinstance(value, float) and np.isnan(value):
            value = nan_of(arr.dtype)
        arr[self._chunk_i] = value

    eu.HDF5Metrics._scrub_cache = _scrub_cache
    eu.HDF5Metrics._append = _append
    eu.HDF5Metrics._bnnp_numpy2 = True


def load_train_bnn():
    """experiments/train_bnn.py as a module (its sacred Experiment `ex` is built at import;
    nothing runs because the module is not __main__)."""
    setup()
    exp_utils()
    path = os.path.join(reference_root(), "experiments", "train_bnn.py")
    spec = importlib.util.spec_from_file_location("bnnp_ref_train_bnn", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bnnp_ref_train_bnn"] = mod
    spec.loader.exec_module(mod)
    return mod


class FakeMetrics:
    "the 5-line metrics_saver of SURVEY 8c: records add_scalar calls"

    def __init__(self):
        self.rows = {}

    def add_scalar(self, name, value, step, dtype=None):
        self.rows.setdefault(name, []).append((step, value))

    def flush(self, every_s=0):
        pass


def synthetic_dataset(kind: str, device, n_train: int, n_test: int, seed: int = 0):
    """SURVEY 8d synthetic inputs with the reference's dataset interface (data/base.py:12-27):
    `synthetic-MNIST` X ~ U[0,1) [n, 784], `synthetic-CIFAR10` X ~ N(0,1) [n, 3, 32, 32];
    y int64 over 0..9 with every class present.  Returns an object with `.norm` / `.unnorm`."""
    import torch
    setup()
    from bnn_priors.data import Dataset
    g = torch.Generator().manual_seed(seed)
    n = n_train + n_test
    if kind == "mnist":
        X = torch.rand(n, 784, generator=g)
    elif kind == "cifar10":
        X = torch.randn(n, 3, 32, 32, generator=g)
    else:
        raise ValueError(kind)
    y = torch.randint(0, 10, (n,), generator=g)
    y[:10] = torch.arange(10)
    y[n_train:n_train + 10] = torch.arange(10)
    idx_train, idx_test = torch.arange(n_train), torch.arange(n_train, n)

    class _Data:
        pass
    d = _Data()
    d.norm = Dataset(X, y, idx_train, idx_test, device=device)
    d.unnorm = d.norm
    d.num_train_set = n_train
    return d


# BASELINE.json configs 2-5 as `exp_utils.get_model` keyword sets (experiments/train_bnn.py:38-123)
MODEL_DEFAULTS = dict(width=50, depth=3, weight_loc=0., weight_scale=2. ** 0.5, bias_prior="gaussian", bias_loc=0.,
                      bias_scale=1., batchnorm=True, weight_prior_params={}, bias_prior_params={})
CONFIGS = {
    "densenet_gaussian": dict(data="mnist", model="classificationdensenet", weight_prior="gaussian"),
    "convnet_laplace": dict(data="mnist", model="classificationconvnet", weight_prior="laplace"),
    "googleresnet_studentt": dict(data="cifar10", model="googleresnet", weight_prior="student-t"),
    "googleresnet_gaussian": dict(data="cifar10", model="googleresnet", weight_prior="gaussian"),
}


def build_model(config: str, data, seed: int = 0):
    """The reference's own model for a BASELINE config, He-initialised like train_bnn.py:173-174."""
    import torch
    eu = exp_utils()
    cfg = dict(MODEL_DEFAULTS, **{k: v for k, v in CONFIGS[config].items() if k != "data"})
    torch.manual_seed(seed)
    model = eu.get_model(x_train=data.norm.train_X, y_train=data.norm.train_y, **cfg)
    eu.he_initialize(model)
    net = getattr(model, "net", None)
    if isinstance(net, torch.nn.DataParallel):
        # one process per GPU: keep the wrapper exp_utils.py:229 adds, but on this
        # device only, so that it calls the module directly instead of replicating it
        dev = data.norm.train_X.device
        net.device_ids = [dev.index if dev.index is not None else 0]
        net.output_device = net.device_ids[0]
    return model
