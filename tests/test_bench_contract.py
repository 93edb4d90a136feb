"""bench.py's output contract on a machine without a GPU: the reference arm prints ONE JSON
line with the keys the driver reads; the GPU arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "param-updates/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("SGLD param-updates/sec") and d["config"]["workload"] and d["config"]["n_params"] == 25124842
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["cpu_arms"]) >= {"numpy_port_one_subchain_per_thread", "torch_ops_like_the_reference"}
    assert d["value"] == max(a["value"] for a in d["cpu_arms"].values())
    assert d["steps"] == 1 and d["warmup"] == 1 or d["cpu_baseline"]["kind"] == "reference"
    # both arms of the bench describe the workload with the same keys (the driver compares them)
    assert set(d["config"]) == {"workload", "n_params", "tensors", "sampler", "calc_metrics", "lr", "num_data",
                                "momentum", "temperature", "chains", "parallelism", "l2"}
    assert d["config"]["tensors"] == 54
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bnn_priors", "mcmc", "sgld.py")):
        # the unmodified reference classes were timed too
        assert d["cpu_arms"]["reference_bnn_priors_mcmc_SGLD"]["kind"] == "reference"
        assert {"SGLD.step(calc_metrics=False)", "VerletSGLD.step(calc_metrics=True)", "HMC.step(calc_metrics=False)",
                "VerletSGLD.delta_energy", "VerletSGLD.maybe_reject(rejecting)"} <= set(d["reference_calls"])


def test_other_ranks_of_the_reference_arm_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CPU implementation" in (r.stderr + r.stdout)
