"""The oracle (oracle/sgmcmc_oracle.py) against the golden call traces recorded
from the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import json
import math
import os

import numpy as np
import pytest

from oracle import sgmcmc_oracle as O
from replay import GOLDEN_DIR, OracleEngine, Trace, replay

TRACES = ["sgld_trace", "sgld_nomomentum_trace", "verlet_trace", "hmc_trace",
          "runner_verlet_normal_trace", "runner_verlet_laplace_trace",
          "runner_verlet_studentt_trace",
          # the reference's HMCRunnerReject and SGLDRunner (BASELINE configs 5 and 1)
          "runner_hmc_normal_trace", "runner_sgld_normal_trace"]

# fp32 elementwise work replayed op for op: a few ulp of drift over <=100 calls
TRAJ_TOL = 1e-5
SCALAR_TOL = 1e-5
DE_ABS_TOL = 1e-4


def _check(rep):
    assert rep.n_events > 0
    assert rep.p_err < TRAJ_TOL and rep.m_err < TRAJ_TOL, rep
    assert rep.decisions_equal == rep.decisions, rep
    assert rep.de_abs_err < DE_ABS_TOL, rep
    assert rep.log_accept_err < 10 * DE_ABS_TOL, rep
    for k, v in rep.scalar_err.items():
        assert v < SCALAR_TOL, (k, v)


@pytest.mark.parametrize("name", TRACES)
@pytest.mark.parametrize("dot_dtype", [np.float32, np.float64])
def test_oracle_reproduces_reference_trace(name, dot_dtype):
    t = Trace(name)
    _check(replay(t, OracleEngine(t, dot_dtype=dot_dtype)))


@pytest.mark.parametrize("name", [n for n in TRACES if n.startswith("runner")])
def test_oracle_fused_prior_reproduces_reference_trace(name):
    """Likelihood-only gradient in, closed-form prior gradient added by the
    oracle: must land on the reference trajectory, whose p.grad came from
    autograd through Prior.log_prob (prior/base.py:57-58)."""
    t = Trace(name)
    _check(replay(t, OracleEngine(t, fused_prior=True), fused_prior=True))


def test_runner_trace_has_the_call_order_of_the_reference_runner():
    """inference_reject.py:57-59 then, per sampling epoch, :120-127,:156."""
    t = Trace("runner_verlet_studentt_trace")
    ops = [e["op"] for e in t.events]
    assert ops[:2] == ["sample_momentum", "initial_step"]
    for i, op in enumerate(ops):
        if op == "maybe_reject":
            assert ops[i - 2:i] == ["final_step", "delta_energy"]
            assert ops[i + 1] == "initial_step"
    assert any(e["op"] == "maybe_reject" and e["out"][0] for e in t.events)
    assert any(e["op"] == "maybe_reject" and not e["out"][0] for e in t.events)


def test_priors_against_reference_log_prob_and_autograd():
    z = np.load(os.path.join(GOLDEN_DIR, "priors.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    assert {m["kind"] for m in meta} == {1, 2, 3, 4, 5, 6, 7, 8, 9}
    for m in meta:
        p, g = z[m["name"] + "_p"], z[m["name"] + "_grad"]
        lp = O.prior_log_prob(m["kind"], p, m["loc"], m["scale"], m.get("df", 3.0))
        assert math.isclose(lp, m["log_prob"], rel_tol=2e-6), (m["name"], lp, m["log_prob"])
        got = O.prior_grad_log_prob(m["kind"], p, m["loc"], m["scale"], m.get("df", 3.0))
        np.testing.assert_allclose(got, g, rtol=3e-6, atol=1e-30, err_msg=m["name"])


def test_host_closed_forms_match_reference_log_prob_and_autograd():
    """bnn_priors_b200.prior_fusion.closed_form (the check fuse_prior runs before it fuses a
    module) against the same golden vectors."""
    import torch
    from bnn_priors_b200.prior_fusion import closed_form
    z = np.load(os.path.join(GOLDEN_DIR, "priors.npz"))
    for m in json.loads(bytes(z["meta"]).decode()):
        p, g = torch.tensor(z[m["name"] + "_p"]), z[m["name"] + "_grad"]
        lp, grad = closed_form(m["kind"], p, m["loc"], m["scale"], m.get("df", 3.0))
        assert math.isclose(float(lp), m["log_prob"], rel_tol=5e-6, abs_tol=1e-6), m["name"]
        np.testing.assert_allclose(grad.numpy(), g, rtol=1e-5, atol=1e-7 * max(1.0, float(np.abs(g).max())), err_msg=m["name"])


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    def run(c, k):
        r = O.philox4x32_10(*[np.array([x], dtype=np.uint32) for x in c], k[0], k[1])
        return [int(x[0]) for x in r]
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normal_is_standard_normal():
    import scipy.stats
    z = O.philox_normal(O.philox_key(7, 3), 11, np.arange(50000)).reshape(-1).astype(np.float64)
    assert np.isfinite(z).all()
    assert scipy.stats.kstest(z, "norm").pvalue > 0.01
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    # different launch counters / keys decorrelate
    z2 = O.philox_normal(O.philox_key(7, 3), 12, np.arange(50000)).reshape(-1)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 0.01


# ---- test-set evaluation: oracle/eval_oracle.py against the reference's results
@pytest.mark.parametrize("tag,kind", [("cls", 0), ("reg", 1)])
def test_eval_oracle_reproduces_reference_results(tag, kind):
    from oracle import eval_oracle as EO
    z = np.load(os.path.join(GOLDEN_DIR, "eval.npz"))
    want = json.loads(str(z[f"{tag}.results"]))
    got = EO.evaluate(z[f"{tag}.acc_data"], z[f"{tag}.lps"], z[f"{tag}.y"], kind)
    for k in ("lp_ensemble", "lp_last", "acc_ensemble", "acc_last"):
        # the reference takes the accuracy mean in float32 (models/base.py:185): 1e-7
        assert math.isclose(got[k], want[k], rel_tol=1e-7, abs_tol=1e-7), (k, got[k], want[k])
    if kind == 0:
        assert math.isclose(got["lp_ensemble_check"], got["lp_ensemble"], rel_tol=1e-6)
        assert np.allclose(got["probs_mean"].sum(1), 1.0)
    # a one-sample ensemble (the per-epoch call of the runners, inference.py:199-213)
    want1 = json.loads(str(z[f"{tag}.results_last_only"]))
    got1 = EO.evaluate(z[f"{tag}.acc_data"][-1:], z[f"{tag}.lps"][-1:], z[f"{tag}.y"], kind)
    for k in ("lp_ensemble", "lp_last", "acc_ensemble", "acc_last"):
        assert math.isclose(got1[k], want1[k], rel_tol=1e-7, abs_tol=1e-7), (k, got1[k], want1[k])


# ---- hierarchical priors (second half) against the reference's autograd
def _hier_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "hier_priors.npz"))
    return z, json.loads(bytes(z["meta"]).decode())


@pytest.mark.parametrize("case", _hier_cases()[1], ids=[m["name"] for m in _hier_cases()[1]])
def test_oracle_hierarchical_prior_matches_reference_autograd(case):
    z, _ = _hier_cases()
    p, g_ref = z[case["name"] + "_p"], z[case["name"] + "_grad_p"]
    s, _ = O.hyper_scale(case["hyper_kind"], case["u"], case["hyper_a"], case["hyper_b"])
    assert math.isclose(s, case["scale"], rel_tol=2e-6)
    ch = O.Chain([p, np.array([case["u"]], np.float32)], O.Group(lr=1.0, num_data=1.0))
    ch.segs[0].prior_kind, ch.segs[0].prior_loc, ch.segs[0].prior_df = case["kind"], case["loc"], case["df"]
    O.link_hyper(ch, 0, 1, case["hyper_kind"], case["hyper_a"], case["hyper_b"])
    assert math.isclose(O.log_prior(ch), case["log_prob"], rel_tol=5e-6, abs_tol=1e-4)
    O.fuse_prior_into_grad(ch)            # zero likelihood gradient, N = 1: g = -dlogp
    got_p, got_u = -ch.segs[0].g, -float(ch.segs[1].g[0])
    assert np.allclose(got_p, g_ref, rtol=2e-5, atol=1e-6 * np.abs(g_ref).max())
    assert math.isclose(got_u, case["grad_u"], rel_tol=2e-5, abs_tol=1e-4), (got_u, case["grad_u"])


def test_torch_op_restatement_agrees_with_the_numpy_oracle():
    """oracle/sgmcmc_torch.py (the CPU arm of bench.py) against oracle/sgmcmc_oracle.py on
    seeded inputs: same noise, same gradients, a few steps."""
    import torch
    from oracle import sgmcmc_torch as OT
    rng = np.random.default_rng(5)
    shapes = [(257,), (8, 5), (5,), (33, 3)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    hp = dict(lr=1e-2, num_data=50.0, momentum=0.9, temperature=0.7)
    ch = O.Chain(p0, O.Group(**hp))
    tc = OT.TorchSGLDChain([torch.tensor(a) for a in p0], **hp)
    z = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    O.sample_momentum(ch, lambda i, n: z[i].reshape(-1))
    tc.sample_momentum(noise=[torch.tensor(a) for a in z])
    for it in range(6):
        gs = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
        z = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        for seg, g, tg in zip(ch.segs, gs, tc.g):
            seg.g = g.reshape(-1).copy()
            tg.copy_(torch.tensor(g))
        O.sgld_step(ch, lambda i, n: z[i].reshape(-1), calc_metrics=True)
        tc.step(calc_metrics=True, noise=[torch.tensor(a) for a in z])
        for i, seg in enumerate(ch.segs):
            assert np.allclose(tc.p[i].numpy().reshape(-1), seg.p, rtol=1e-5, atol=1e-6)   # fp32, fused vs separate roundings
            assert np.allclose(tc.m[i].numpy().reshape(-1), seg.m, rtol=1e-5, atol=1e-6)
            assert math.isclose(tc.est_temperature[i], seg.est_temperature, rel_tol=1e-5)
            assert math.isclose(tc.est_config_temp[i], seg.est_config_temp, rel_tol=1e-4, abs_tol=1e-6)


def test_torch_op_verlet_restatement_agrees_with_the_numpy_oracle():
    import torch
    from oracle import sgmcmc_torch as OT
    rng = np.random.default_rng(6)
    shapes = [(130,), (7, 6), (3,)]
    p0 = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    hp = dict(lr=2e-2, num_data=20.0, momentum=0.8, temperature=1.0)
    ch = O.Chain(p0, O.Group(**hp), dot_dtype=np.float64)
    tc = OT.TorchVerletChain([torch.tensor(a) for a in p0], **hp)
    z = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    O.sample_momentum(ch, lambda i, n: z[i].reshape(-1))
    tc.sample_momentum(noise=[torch.tensor(a) for a in z])
    for it in range(5):
        gs = [rng.standard_normal(s).astype(np.float32) * 0.1 for s in shapes]
        z = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        for seg, g, tg in zip(ch.segs, gs, tc.g):
            seg.g = g.reshape(-1).copy()
            tg.copy_(torch.tensor(g))
        O.verlet_step(ch, lambda i, n: z[i].reshape(-1), phase=O.PHASE_MID, calc_metrics=True)
        tc.step(calc_metrics=True, noise=[torch.tensor(a) for a in z])
        for i, seg in enumerate(ch.segs):
            assert np.allclose(tc.p[i].numpy().reshape(-1), seg.p, rtol=1e-5, atol=1e-6)
            assert np.allclose(tc.m[i].numpy().reshape(-1), seg.m, rtol=1e-5, atol=1e-6)
            assert math.isclose(tc.delta_energy[i], seg.delta_energy, rel_tol=1e-4, abs_tol=1e-5)
            assert math.isclose(tc.est_temperature[i], seg.est_temperature, rel_tol=1e-5)


def test_eval_oracle_ensemble_probabilities_and_calibration_inputs():
    """tests/golden/eval_calibration.npz: the reference's evaluate_model with calibration_eval=True.  The
    ensemble probabilities the evaluation hands to the calibration metrics (exp_utils.py:324-327) must be
    the reference's; with the reference checkout present, its own ece / ace / rmsce on OUR probabilities
    must give the recorded values."""
    from oracle import eval_oracle as EO
    z = np.load(os.path.join(GOLDEN_DIR, "eval_calibration.npz"))
    want = json.loads(str(z["results"]))
    acc = z["acc_data"]
    labels = z["y"]
    lps = np.take_along_axis(acc, labels[None, :, None].repeat(acc.shape[0], 0), 2)[..., 0]
    got = EO.evaluate(acc, lps, labels, EO.CATEGORICAL)
    assert np.allclose(got["probs_mean"], z["probs_mean"], rtol=1e-12, atol=1e-15)
    assert math.isclose(got["lp_ensemble"], want["lp_ensemble"], rel_tol=1e-7)
    if os.path.isdir("/root/reference/bnn_priors"):
        import sys
        sys.path.insert(0, "/root/reference")
        try:
            from bnn_priors.third_party.calibration_error import ace, ece, rmsce
        finally:
            sys.path.remove("/root/reference")
        assert math.isclose(float(ece(labels, got["probs_mean"]).mean()), want["ece"], rel_tol=1e-9)
        assert math.isclose(float(ace(labels, got["probs_mean"]).mean()), want["ace"], rel_tol=1e-9)
        assert math.isclose(float(rmsce(labels, got["probs_mean"]).mean()), want["rmsce"], rel_tol=1e-9)
