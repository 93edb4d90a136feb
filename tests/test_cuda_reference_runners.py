"""The reference's OWN runners (bnn_priors/inference.py, inference_reject.py), unmodified, on the
reference's OWN models at the BASELINE configs, driving the B200 sampler on the GPU.

Each case runs the runner twice on cuda:0 (tests/runner_tape.py): once with the reference's eager
sampler (recorded), once after `overlay.install()` with the kernel (replayed with the recorded
gradients, N(0,1) tensors, Metropolis uniforms and potentials) and asserts the north star's parity
definition: parameter / momentum trajectories within 1e-5 (relative to the tensor's RMS), the
per-tensor scalars the runner logs within 1e-5, delta energies within 1e-5 of the size of their
terms, IDENTICAL accept / reject decisions, and identical stored samples.

The reference comes from oracle/_ref (oracle/make_ref.py; made by __graft_entry__.build() in the
build container, travels with the tree).  BASELINE.json configs: (2) densenet / VerletSGLDReject /
gaussian, (3) convnet / VerletSGLDReject / laplace / T=0.1, (4) googleresnet / VerletSGLDReject /
student-t, (5) HMC with 50 leapfrog steps / googleresnet / gaussian, (1) densenet / SGLD.
"""
import importlib
import json
import os
import warnings

import pytest
import torch

import refenv

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refenv.available(), reason="no oracle/_ref snapshot (python oracle/make_ref.py)")]

TOL = 1e-5

# (inference, config, runner settings)
CASES = {
    "cfg2_densenet_verlet_reject_gaussian": ("VerletSGLDReject", "densenet_gaussian", dict(n_train=1024, lr=2e-2)),
    "cfg3_convnet_verlet_reject_laplace_T0.1": ("VerletSGLDReject", "convnet_laplace",
                                                dict(n_train=1024, lr=5e-3, temperature=0.1)),
    "cfg4_googleresnet_verlet_reject_studentt": ("VerletSGLDReject", "googleresnet_studentt", dict(n_train=1024, lr=5e-3)),
    # 50 leapfrog steps = 50 minibatches between initial_step and final_step: 6400 points at batch 128
    "cfg5_googleresnet_hmc50_gaussian": ("HMCReject", "googleresnet_gaussian",
                                         dict(n_train=6400, lr=2e-3, cycles=1, sample_epochs=3)),
    "cfg1_densenet_sgld_gaussian": ("SGLD", "densenet_gaussian", dict(n_train=1024, lr=5e-4)),
    "densenet_verlet_noreject": ("VerletSGLD", "densenet_gaussian", dict(n_train=1024, lr=1e-2)),
    "googleresnet_sgld_reject_runner": ("SGLDReject", "googleresnet_studentt", dict(n_train=512, lr=5e-4)),
    # inference.py:368-374 (HMCRunner: the non-"Reject" HMC loop, momentum re-sampled at every sampling epoch)
    "convnet_hmc_runner_laplace": ("OurHMC", "convnet_laplace", dict(n_train=1024, lr=2e-3)),
}


def _check(report, tape, runner_a, runner_b, name):
    d = report.as_dict()
    print(f"\n[{name}] calls={tape.summary()} report={d}")
    if os.environ.get("BNNP_REPORT_FILE"):           # tools/gpu_session.sh keeps the numbers for profiles/
        with open(os.environ["BNNP_REPORT_FILE"], "a") as f:
            f.write(json.dumps(dict(case=name, calls=tape.summary(), **{k: (v if v != float("inf") else None)
                                                                        for k, v in d.items()})) + "\n")
    assert report.n_events == len(tape.events)
    assert report.p_err <= TOL, d
    assert report.m_err <= TOL, d
    for k in ("preconditioner", "est_temperature", "est_config_temp", "square_avg_mean"):
        if k in report.scalar_err:
            assert report.scalar_err[k] <= TOL, (k, d)
    # the per-tensor running sums of the energy bookkeeping (verlet_sgld.py:170-176) add one signed
    # c * dot(g, m) term per step, each rounded in fp32 by the reference's `dot` (fp64 here), and are
    # judged against the largest of them at the event, not against the (unknown) sum of the magnitudes
    # of everything that went in: a few 1e-6 per step.  The quantity they exist for, the total delta
    # energy, is held to 1e-5 of the size of its terms just below.
    for k in ("delta_energy", "prev_new_momentum_delta"):
        if k in report.scalar_err:
            assert report.scalar_err[k] <= 5 * TOL, (k, d)
    assert report.de_term_err <= TOL, d
    assert report.decisions_equal == report.decisions, d
    # what the runner stored as samples (inference.py:189-194): the parameters.  (The BatchNorm running
    # statistics in the same state_dict follow the minibatch order, which the Reject runners draw from OS
    # entropy, inference_reject.py:68-72: not a function of the sampler.)
    sa, sb = runner_a.get_samples(), runner_b.get_samples()
    assert sa.keys() == sb.keys()
    import runner_tape as RT
    for k in runner_a.param_names:
        assert sa[k].shape == sb[k].shape and sa[k].shape[0] >= 2, k
        assert RT.rel_err(sb[k], sa[k]) <= TOL, k


def _run_case(name, fused=False, before_run=None):
    import runner_harness as H
    from bnn_priors_b200 import mcmc as fast, overlay
    warnings.filterwarnings("ignore", message="Detected call of `lr_scheduler.step")
    inference, config, kw = CASES[name]
    dev = torch.device("cuda:0")
    tape, runner_a = H.record_run(inference, config, dev, with_prior_grads=fused, **kw)
    ref_mcmc = importlib.import_module("bnn_priors.mcmc")
    overlay.install(fuse_prior=fused)
    try:
        assert ref_mcmc.VerletSGLD is fast.VerletSGLD
        report, runner_b = H.replay_run(inference, config, dev, tape, fused_prior=fused, before_run=before_run, **kw)
        opt = runner_b.optimizer
        assert isinstance(opt, fast.SGLD) and opt.flat_groups[0].launches > 0     # the kernel did the work
        if fused:
            assert opt.flat_groups[0].prior_fused
    finally:
        overlay.uninstall()
    _check(report, tape, runner_a, runner_b, name + ("+fused_prior" if fused else ""))
    return report, tape


@pytest.mark.parametrize("name", list(CASES))
def test_reference_runner_drives_the_kernel_on_the_reference_trajectory(name):
    report, tape = _run_case(name)
    if "reject" in name and "sgld" not in name:
        assert report.decisions >= 4


@pytest.mark.parametrize("name", ["cfg2_densenet_verlet_reject_gaussian", "cfg3_convnet_verlet_reject_laplace_T0.1",
                                  "cfg4_googleresnet_verlet_reject_studentt", "cfg5_googleresnet_hmc50_gaussian",
                                  "cfg1_densenet_sgld_gaussian"])
def test_unchanged_runner_with_the_prior_fused_into_the_kernel(name, monkeypatch):
    """overlay.install(fuse_prior=True): the runner is unchanged, the prior never goes through autograd
    (Prior.log_prob is not entered once after the sampler exists), and the trajectory is still the
    reference's."""
    refenv.setup()
    base = importlib.import_module("bnn_priors.prior.base")
    calls = {"n": 0, "armed": False}
    real = base.Prior.log_prob

    def counting(self):
        calls["n"] += int(calls["armed"])
        return real(self)

    def before_run(runner, model):
        make = runner._make_optimizer

        def make_and_arm(params):
            # fuse_prior evaluates every module's log_prob once, inside _make_optimizer, to verify the
            # closed forms against it; everything after that must stay out of autograd
            opt = make(params)
            calls["armed"] = True
            return opt
        runner._make_optimizer = make_and_arm
    try:
        _run_case(name, fused=True, before_run=lambda r, m: (monkeypatch.setattr(base.Prior, "log_prob", counting),
                                                              before_run(r, m)))
    finally:
        calls["armed"] = False
    assert calls["n"] == 0, f"Prior.log_prob entered {calls['n']} times under the fused overlay"
