"""Multi-chain host logic on CPU: two processes over gloo, each with its own chain
seed and sample ring; the cycle-end all-gather must return, on every rank, the
concatenation of what the ranks stored (SURVEY 8e parity definition)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest
import torch

from bnn_priors_b200 import chains as CH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    from bnn_priors_b200 import chains as CH
    rank, world, device = CH.init_chains(backend="gloo")
    assert world == 2 and device.type == "cpu"
    seed = CH.chain_seed(100, rank)
    g = torch.Generator().manual_seed(seed)
    total, extra = 96, 5
    ring = CH.SampleRing(3, total + extra, device)
    mine = []
    for s in range(3):
        flat = torch.randn(total, generator=g)
        bn = torch.randn(extra, generator=g)
        ring.push(flat, extras=[bn], step=10 * s + rank, rejected=(s == 1 and rank == 1))
        mine.append(torch.cat([flat, bn]))
    out, meta = ring.gather()
    assert out.shape == (2, 3, total + extra) and meta.shape == (2, 3, 2)
    # what each rank must have stored, recomputed from the seeds alone
    for r in range(2):
        g2 = torch.Generator().manual_seed(CH.chain_seed(100, r))
        for s in range(3):
            want = torch.cat([torch.randn(total, generator=g2), torch.randn(extra, generator=g2)])
            assert torch.equal(out[r, s], want), (rank, r, s)
            assert meta[r, s, 0].item() == 10 * s + r
            assert meta[r, s, 1].item() == int(s == 1 and r == 1)
    try:
        ring.push(torch.zeros(total))
        raise SystemExit("ring overflow not detected")
    except IndexError:
        pass
    print("rank", rank, "ok")
""") % ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_chains_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_single_process_ring_and_unflatten():
    ring = CH.SampleRing(2, 10, torch.device("cpu"))
    flat = torch.arange(10, dtype=torch.float32)
    ring.push(flat, step=7)
    out, meta = ring.gather()
    assert out.shape == (1, 2, 10) and torch.equal(out[0, 0], flat) and meta[0, 0, 0] == 7
    d = CH.unflatten_sample(out[0, 0], [0, 6, 9], [(2, 3), (3,), ()], ["w", "b", "s"])
    assert d["w"].shape == (2, 3) and d["b"].tolist() == [6., 7., 8.] and d["s"].item() == 9.
    assert CH.chain_seed(5, 3) == 8
