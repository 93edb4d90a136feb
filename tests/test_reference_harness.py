"""CPU checks of the test infrastructure around the reference (build container and GPU box):
the record / replay harness reports zero deviation when the reference replays itself, the
reference's train_bnn.py runs unmodified through the sacred / h5py shims, and the oracle/_ref
snapshot is byte-identical to the checkout it was made from."""
import hashlib
import json
import os
import warnings

import pytest
import torch

import refenv

pytestmark = pytest.mark.skipif(not refenv.available(), reason="no reference (oracle/_ref or /root/reference)")


@pytest.mark.parametrize("inference,config,kw", [
    ("VerletSGLDReject", "densenet_gaussian", dict(lr=2e-2)),
    ("HMCReject", "densenet_gaussian", dict(lr=2e-3)),
    ("SGLD", "convnet_laplace", dict(temperature=0.1)),
])
def test_reference_replays_itself_exactly(inference, config, kw):
    import runner_harness as H
    warnings.filterwarnings("ignore", message="Detected call of `lr_scheduler.step")
    dev = torch.device("cpu")
    tape, ra = H.record_run(inference, config, dev, n_train=256, n_test=128, **kw)
    report, rb = H.replay_run(inference, config, dev, tape, n_train=256, n_test=128, **kw)
    assert report.n_events == len(tape.events) > 20
    assert report.p_err == 0.0 and report.m_err == 0.0 and report.de_abs_err == 0.0
    assert all(v == 0.0 for v in report.scalar_err.values()), report.scalar_err
    assert report.decisions_equal == report.decisions
    if "Reject" in inference and inference != "SGLDReject":
        assert report.decisions == 4
    sa, sb = ra.get_samples(), rb.get_samples()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)


def test_train_bnn_main_runs_unmodified_through_the_shims(tmp_path):
    import runner_harness as H
    run, rundir = H.run_train_bnn(tmp_path, device="cpu", n_train=256, n_test=128)
    assert set(run.result) == {"lp_ensemble", "lp_last", "acc_ensemble", "acc_last"}
    assert all(map(lambda v: v == v, run.result.values()))
    eu = refenv.exp_utils()
    samples = eu.load_samples(os.path.join(rundir, "samples.pt"))
    assert samples["steps"].shape == (4,) and samples["net.module.0.weight_prior.p"].shape == (4, 50, 784)
    import h5py
    with h5py.File(os.path.join(rundir, "metrics.h5"), "r") as f:
        rejected = f["acceptance/rejected"][:]
        assert ((rejected == 0) | (rejected == 1) | (rejected == -2 ** 63)).all()
        assert (rejected != -2 ** 63).sum() == 5        # the first sample + 4 Metropolis tests
    with open(os.path.join(rundir, "run.json")) as f:
        assert sorted(json.load(f)["artifacts"]) == ["metrics.h5", "samples.pt"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/bnn_priors"), reason="no reference checkout here")
def test_snapshot_is_the_unmodified_reference():
    import sys
    sys.path.insert(0, os.path.join(refenv.ROOT, "oracle"))
    import make_ref
    dest = make_ref.make()
    with open(os.path.join(dest, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    assert "bnn_priors/mcmc/sgld.py" in files and "experiments/train_bnn.py" in files
    for rel, sha in files.items():
        for root in (dest, "/root/reference"):
            with open(os.path.join(root, rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == sha, (root, rel)
