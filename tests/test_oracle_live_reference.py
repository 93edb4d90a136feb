"""The oracle against the reference imported LIVE (build container only; the GPU box has
no /root/reference): the golden-trace generator (tests/golden/make_golden.py) is run again
on the unmodified reference samplers, (a) the traces it produces must be the committed
fixtures (the fixtures are reproducible from the reference, nobody edited them by hand) and
(b) the oracle must replay them.  CPU only."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "bnn_priors")),
                                reason="no reference checkout here")


@pytest.fixture(scope="module")
def generator(tmp_path_factory):
    spec = importlib.util.spec_from_file_location("make_golden_live", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    try:
        spec.loader.exec_module(mod)            # imports bnn_priors.mcmc from /root/reference
        mod.HERE = str(tmp_path_factory.mktemp("live_golden"))
        yield mod
    finally:
        sys.path[:] = saved


@pytest.mark.parametrize("fn,name", [("golden_sgld", "sgld_trace"), ("golden_sgld_nomomentum", "sgld_nomomentum_trace"),
                                     ("golden_verlet", "verlet_trace"), ("golden_hmc", "hmc_trace")])
def test_live_reference_trace_equals_fixture_and_oracle_replays_it(generator, fn, name, monkeypatch):
    import torch
    import replay as R
    threads = torch.get_num_threads()
    torch.set_num_threads(1)                    # like make_golden.py's main: same reduction order
    try:
        getattr(generator, fn)()
    finally:
        torch.set_num_threads(threads)
    live = np.load(os.path.join(generator.HERE, name + ".npz"))
    fixture = np.load(os.path.join(R.GOLDEN_DIR, name + ".npz"))
    assert sorted(live.files) == sorted(fixture.files)
    for k in live.files:
        if k == "meta":
            assert bytes(live[k]) == bytes(fixture[k]), "events / scalars recorded from the reference changed"
        else:
            assert np.array_equal(live[k], fixture[k]), k
    monkeypatch.setattr(R, "GOLDEN_DIR", generator.HERE)
    t = R.Trace(name)
    rep = R.replay(t, R.OracleEngine(t, dot_dtype=np.float32))
    assert rep.n_events > 0 and rep.p_err < 1e-5 and rep.m_err < 1e-5
    assert rep.decisions_equal == rep.decisions
