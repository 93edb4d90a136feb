"""experiments/train_bnn.py -- the reference's training script, unmodified (oracle/_ref) -- run end to
end on the GPU with the B200 samplers plugged in by `overlay.install()`: `main` builds the reference's
model, its runner constructs `mcmc.VerletSGLD` / `mcmc.HMC` / `mcmc.SGLD` (now this repo's classes), runs
cycles with Metropolis tests, writes samples and metrics through the reference's HDF5 savers (h5py test
shim), re-loads the samples and evaluates them.  The north star's "plugs into experiments/train_bnn.py
unchanged" (synthetic data of the data set's shape; sacred / h5py / pyro are test shims)."""
import math
import os

import pytest
import torch

import refenv

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refenv.available(), reason="no oracle/_ref snapshot (python oracle/make_ref.py)")]


def _spy(monkeypatch):
    "remember every sampler the run constructs"
    from bnn_priors_b200 import mcmc as fast
    made = []
    real_init = fast.SGLD.__init__

    def init(self, *a, **k):
        real_init(self, *a, **k)
        made.append(self)
    monkeypatch.setattr(fast.SGLD, "__init__", init)
    return made


@pytest.mark.parametrize("inference,model,data,prior,extra", [
    ("VerletSGLDReject", "classificationdensenet", "mnist", "gaussian", {}),
    ("VerletSGLDReject", "classificationconvnet", "mnist", "laplace", dict(temperature=0.1)),
    ("VerletSGLDReject", "googleresnet", "cifar10", "student-t", {}),
    ("HMCReject", "googleresnet", "cifar10", "gaussian", dict(momentum=1.0)),
    ("SGLD", "classificationdensenet", "mnist", "gaussian", dict(reject_samples=False)),
    ("VerletSGLD", "classificationdensenet", "mnist", "gaussian", {}),
])
@pytest.mark.parametrize("fuse", [False, True])
def test_train_bnn_main_with_the_overlay(tmp_path, monkeypatch, inference, model, data, prior, extra, fuse):
    """fuse=False: samplers + evaluation re-bound; fuse=True: everything the overlay offers -- the prior fused
    into the kernel and the sample file written by FlatSampleSaver (in the reference's HDF5 layout, through
    the h5py stand-in) and read back by the reference's own load_samples inside main."""
    import runner_harness as H
    from bnn_priors_b200 import mcmc as fast, overlay
    made = _spy(monkeypatch)
    overlay.install(evaluate=True, fuse_prior=fuse, sample_sink=fuse)
    try:
        run, rundir = H.run_train_bnn(tmp_path, inference=inference, model=model, data=data, weight_prior=prior,
                                      n_train=512, n_test=256, **extra)
    finally:
        overlay.uninstall()
    assert len(made) == 1
    opt = made[0]
    want = {"VerletSGLDReject": fast.VerletSGLD, "VerletSGLD": fast.VerletSGLD, "HMCReject": fast.HMC, "SGLD": fast.SGLD}
    assert type(opt) is want[inference]
    (fg,) = opt.flat_groups
    assert fg.launches > 20 and fg.device.type == "cuda"
    assert fg.prior_fused == fuse
    assert set(run.result) == {"lp_ensemble", "lp_last", "acc_ensemble", "acc_last"}
    assert all(math.isfinite(v) for v in run.result.values()), run.result
    eu = refenv.exp_utils()
    samples = eu.load_samples(os.path.join(rundir, "samples.pt"))
    assert samples["steps"].shape == (4,)
    assert sum(int(v[0].numel()) for k, v in samples.items() if v.dtype == torch.float32) >= fg.n_params
    for k, v in samples.items():
        if v.dtype.is_floating_point:
            assert torch.isfinite(v).all(), k
    import h5py
    with h5py.File(os.path.join(rundir, "metrics.h5"), "r") as f:
        assert "est_temperature/all" in f and "lr" in f
        if inference.endswith("Reject") and extra.get("reject_samples", True):
            rejected = f["acceptance/rejected"][:]
            assert (rejected != -2 ** 63).sum() == 5
