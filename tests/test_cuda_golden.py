"""The CUDA samplers (bnn_priors_b200.mcmc -> libbnnp.so) against the golden call
traces recorded from the unmodified reference (tests/golden/make_golden.py).
Every call goes through the reference-shaped Python API and ends in the C ABI."""
import numpy as np
import pytest

from replay import CudaEngine, OracleEngine, Trace, replay

pytestmark = pytest.mark.gpu

TRACES = ["sgld_trace", "sgld_nomomentum_trace", "verlet_trace", "hmc_trace",
          "runner_verlet_normal_trace", "runner_verlet_laplace_trace",
          "runner_verlet_studentt_trace",
          # the reference's HMCRunnerReject and SGLDRunner (BASELINE configs 5 and 1)
          "runner_hmc_normal_trace", "runner_sgld_normal_trace"]

# north star: trajectories within 1e-5 relative fp32, accept decisions identical
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
        assert v < SCALAR_TOL, (k, v, rep)


@pytest.mark.parametrize("name", TRACES)
@pytest.mark.parametrize("foreign_grads", [False, True])
def test_cuda_reproduces_reference_trace(name, foreign_grads):
    t = Trace(name)
    rep = replay(t, CudaEngine(t, foreign_grads=foreign_grads))
    _check(rep)
    if t.sampler != "SGLD":
        assert rep.decisions > 0


@pytest.mark.parametrize("name", [n for n in TRACES if n.startswith("runner")])
def test_cuda_fused_prior_reproduces_reference_trace(name):
    """Likelihood-only gradients in; the kernel adds the closed-form prior
    gradient in-register and must land on the reference trajectory, whose p.grad
    came from autograd through Prior.log_prob (prior/base.py:57-58)."""
    t = Trace(name)
    _check(replay(t, CudaEngine(t, fused_prior=True), fused_prior=True))


@pytest.mark.parametrize("name", TRACES)
def test_cuda_matches_oracle_scalars_tightly(name):
    """Same trace through oracle (fp64 dots) and CUDA: the per-tensor scalars the
    kernel reduces in fp64 agree to ~1e-6 of the run's magnitude."""
    t = Trace(name)
    a = replay(t, OracleEngine(t, dot_dtype=np.float64))
    b = replay(t, CudaEngine(t))
    assert abs(a.de_abs_err - b.de_abs_err) < DE_ABS_TOL
    assert b.decisions_equal == b.decisions == a.decisions
