"""examples/chains_resnet_reject.py (BASELINE config 4 / 5 on synthetic data) runs end to end:
ResNet-20, fused Student-t prior, M-H decisions, device-side evaluation, sample ring."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("sampler", ["VerletSGLD", "HMC"])
def test_example_runs_a_short_cycle(sampler):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "chains_resnet_reject.py"), "--sampler", sampler,
                        "--cycles", "1", "--epochs", "2", "--n-train", "512"], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    log = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert log["sampler"] == sampler and log["params"] == 272282
    assert len(log["decisions"]) == 2 and len(log["test"]) == 2
    assert log["gathered"] == [[1, 2, log["gathered"][0][2]]]
    assert all(0.0 <= t["acc_last"] <= 1.0 for t in log["test"])
