"""Worker of tests/test_cuda_chains_nccl.py -- run under torchrun, one rank per GPU.

Every rank runs its own VerletSGLD-with-rejection chain (googleresnet segment table,
student-t prior fused, in-kernel Philox noise, seed = base + rank) for a few cycles,
stores one sample per cycle in its SampleRing and takes part in the ONE cycle-end
all-gather (NCCL).  Parity definition of SURVEY 8e: block r of the gathered tensor
must equal what an independent single-process run with chain r's seed stores -- so
every rank re-runs its neighbour's chain locally and compares bit for bit.
"""
import json
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from bnn_priors_b200 import chains as CH  # noqa: E402
from bnn_priors_b200 import mcmc  # noqa: E402

TAG = os.environ.get("CHAINS_TAG", "googleresnet_cifar10_studentt")
TENSORS = json.load(open(os.path.join(HERE, "golden", "model_shapes.json")))[TAG]["tensors"]
CYCLES, STEPS = 3, 4
HP = dict(lr=2e-3, num_data=500.0, momentum=0.9, temperature=1.0)


def run_chain(seed: int, device) -> tuple:
    """One chain: returns (ring rows [CYCLES, total], meta [CYCLES, 2])."""
    torch.manual_seed(seed)                       # the CPU generator maybe_reject draws from
    g = torch.Generator(device=device).manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(tuple(t["shape"]), device=device, generator=g)
                                 * (t["scale"] if t["kind"] else 1.0)) for t in TENSORS]
    opt = mcmc.VerletSGLD(params, **HP, seed=seed)
    (fg,) = opt.flat_groups
    for i, t in enumerate(TENSORS):
        fg.set_prior(i, t["kind"], t["loc"], t["scale"], t["df"])
    fg.prior_fused = True
    for p, v in zip(params, fg.g_views):
        p.grad = v

    pad = torch.ones(fg.total, dtype=torch.bool, device=device)
    for o, n in zip(fg.off, fg.numel):
        pad[o:o + n] = False

    def potential_and_grad():
        # a quadratic "likelihood" with a known gradient: U = sum_i .5 (p_i - .1)^2 / N
        fg.G.copy_((fg.P - 0.1) / HP["num_data"])
        fg.G[pad] = 0.0                          # the padding between tensors stays zero
        return float((0.5 * (fg.unpack(fg.P) - 0.1).double().pow(2)).sum() / HP["num_data"])

    ring = CH.SampleRing(CYCLES, fg.total, device)
    opt.sample_momentum()
    step = 0
    for c in range(CYCLES):
        u0 = potential_and_grad()
        opt.initial_step(save_state=True, calc_metrics=False)
        for _ in range(STEPS):
            potential_and_grad()
            opt.step(calc_metrics=False)
            step += 1
        u1 = potential_and_grad()
        opt.final_step(calc_metrics=False)
        de = opt.delta_energy(u0, u1)
        rejected, _ = opt.maybe_reject(de)
        assert math.isfinite(de)
        ring.push(fg.P, step=step, rejected=rejected)
        opt.sample_momentum(keep=0.5)
    return ring, fg


def main():
    rank, world, device = CH.init_chains()
    assert device.type == "cuda", "this worker needs GPUs"
    base = 1234
    ring, fg = run_chain(CH.chain_seed(base, rank), device)
    out, meta = ring.gather()
    assert out.shape == (world, CYCLES, fg.total), out.shape
    assert torch.equal(out[rank], ring.rows) and torch.equal(meta[rank], ring.meta)
    # the neighbour's chain, recomputed here from its seed alone
    other = (rank + 1) % world
    ring2, _ = run_chain(CH.chain_seed(base, other), device)
    assert torch.equal(out[other], ring2.rows), "gathered block differs from an independent run of that chain"
    assert torch.equal(meta[other], ring2.meta), (meta[other].tolist(), ring2.meta.tolist())
    if world > 1:
        assert not torch.equal(out[0], out[1]), "chains with different seeds must differ"
    torch.cuda.synchronize(device)
    print(f"rank {rank}/{world} ok: steps/rejected = {meta[rank].tolist()}", flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
