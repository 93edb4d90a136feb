"""The numpy oracle against the compact goldens recorded from the unmodified reference at the REAL
segment tables of the BASELINE configs (tests/real_tables.py): densenet / convnet-laplace-T0.1 /
googleresnet-studentt (65 tensors, 42 without prior) / googleresnet HMC with 50 leapfrog steps, and
the 25M-parameter table of the metric.  CPU."""
import pytest

import real_tables as RT

TOL = 1e-5
SMALL = [n for n in RT.CASES if "resnet18w96" not in n]


def _check(rep, name):
    assert rep.sample_err <= TOL and rep.moment_err <= TOL, (name, rep.__dict__)
    for k, v in rep.scalar_err.items():
        assert v <= TOL, (name, k, rep.__dict__)
    assert rep.de_term_err <= TOL, (name, rep.__dict__)
    assert rep.decisions_equal == rep.decisions, (name, rep.__dict__)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", SMALL)
def test_oracle_on_the_real_tables(name, fused):
    rep = RT.replay_compact(name, "oracle", fused_prior=fused)
    _check(rep, name)
    if "verlet" in name or "hmc" in name:
        assert rep.decisions == 2 and rep.rejections == 1      # one acceptance, one rejection


def test_oracle_on_the_25m_parameter_table():
    rep = RT.replay_compact("real_resnet18w96_sgld_gaussian", "oracle", fused_prior=True)
    _check(rep, "real_resnet18w96_sgld_gaussian")
