"""Multi-GPU: independent chains, one per GPU, with the single cycle-end NCCL all-gather
(SURVEY 8e).  Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_cuda_chains_nccl.py -m gpu`);
skipped on a one-GPU box.  The worker asserts the parity definition: the gathered block of
chain r == an independent single-process run with chain r's seed, bit for bit."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


def test_one_chain_per_gpu_gathers_what_independent_runs_store():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 8)
    r = _torchrun(world, os.path.join(HERE, "nccl_chains_worker.py"))
    assert r.returncode == 0, r.stdout[-4000:]
    for rank in range(world):
        assert f"rank {rank}/{world} ok" in r.stdout, r.stdout[-4000:]


def test_single_gpu_chain_worker_runs_without_a_process_group():
    "world_size 1: the same worker, no collective (the gather is the identity)"
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(HERE, "nccl_chains_worker.py")], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "rank 0/1 ok" in r.stdout
