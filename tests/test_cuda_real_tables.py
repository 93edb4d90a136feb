"""The CUDA samplers against the compact goldens recorded from the unmodified reference at the REAL
segment tables of the BASELINE configs (tests/real_tables.py), sampler only (the prior's gradient
from torch.distributions + autograd, like the reference) and with the prior fused into the kernel.
Needs neither the reference nor oracle/_ref: inputs are regenerated from seeds."""
import json
import os

import pytest

import real_tables as RT
from test_oracle_real_tables import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", list(RT.CASES))
def test_cuda_on_the_real_tables(name, fused):
    rep = RT.replay_compact(name, "cuda", fused_prior=fused)
    print(name, fused, rep.__dict__)
    if os.environ.get("BNNP_REPORT_FILE"):
        with open(os.environ["BNNP_REPORT_FILE"], "a") as f:
            f.write(json.dumps(dict(case=name + ("+fused_prior" if fused else ""), golden="real_tables",
                                    **{k: (float(v) if isinstance(v, float) or hasattr(v, "item") else v)
                                       for k, v in rep.__dict__.items()})) + "\n")
    _check(rep, name)
    if "verlet" in name or "hmc" in name:
        assert rep.decisions >= 1
