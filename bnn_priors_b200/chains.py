"""Independent Markov chains, one per GPU.

The reference replicates chains as separate OS processes started by a shell loop
(experiments/run_experiment.sh:15) and has no communication at all.  Here every
rank (one process per GPU, `torch.distributed`) runs its own chain with its own
seed; nothing is exchanged during sampling.  At the end of a cycle the ranks do
exactly ONE all-gather of that cycle's posterior samples, sent straight from a
contiguous ring of flat parameter snapshots -- the same flat layout the sampler
kernel works on, so there is no packing pass.

Host logic only (torch.distributed does the transport: NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def pin_host_cores(local_rank: int, local_world: int) -> Optional[List[int]]:
    """Give this rank its own slice of the host cores its GPU is closest to (NVML's CPU affinity of
    the device = the NUMA node the GPU hangs off), so that the launch threads of N chains neither
    migrate nor share a core, and first-touch places this rank's pinned staging buffers on that
    node.  Returns the cores chosen, or None if nothing was changed (no NVML, one rank, or
    BNNP_PIN_CORES=0)."""
    if local_world <= 1 or os.environ.get("BNNP_PIN_CORES", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = sorted(os.sched_getaffinity(0))
    near = allowed
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
        gpu_cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        if gpu_cpus & set(allowed):
            near = sorted(gpu_cpus & set(allowed))
    except Exception:
        pass
    per = len(near) // local_world
    if per < 1:
        return None
    # ranks that share an affinity set split it evenly (on this pod: one NUMA node, 32 cores, 8 GPUs)
    mine = near[local_rank * per:(local_rank + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return mine


def init_chains(backend: Optional[str] = None) -> tuple:
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT (torchrun), pin this process to its GPU and to its own slice of
    the host cores next to that GPU.  Returns (rank, world_size, device).  A single process needs
    no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
        pin_host_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if use_cuda and (backend or "nccl") == "nccl":
            kw["device_id"] = device
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    return rank, world, device


def chain_seed(base_seed: int, rank: int) -> int:
    "Seed of chain `rank` (SURVEY 8d: torch.manual_seed(s) with s = base + rank)."
    return int(base_seed) + int(rank)


class SampleRing:
    """The samples one chain keeps during a cycle: `capacity` rows of `width` fp32
    (flat parameter array, optionally followed by the model's fp32 buffers such as
    BatchNorm statistics), plus the step index and the rejected flag of each row
    (inference_reject.py:127-144 stores a sample every sampling epoch, rejected or
    not, together with its step)."""

    def __init__(self, capacity: int, width: int, device, dtype=torch.float32):
        self.capacity, self.width = int(capacity), int(width)
        self.rows = torch.zeros(self.capacity, self.width, dtype=dtype, device=device)
        self.meta = torch.zeros(self.capacity, 2, dtype=torch.int64, device=device)   # step, rejected
        self.count = 0

    def reset(self) -> None:
        self.count = 0

    @torch.no_grad()
    def push(self, flat: torch.Tensor, extras: Sequence[torch.Tensor] = (), step: int = 0,
             rejected: bool = False) -> int:
        """Snapshot `flat` (the chain's flat P array: one device-to-device copy) and
        the extra tensors into the next row."""
        if self.count >= self.capacity:
            raise IndexError("SampleRing is full; gather() and reset() at the end of the cycle")
        row = self.rows[self.count]
        n = flat.numel()
        row[:n].copy_(flat.reshape(-1), non_blocking=True)
        for t in extras:
            k = t.numel()
            row[n:n + k].copy_(t.reshape(-1).to(row.dtype), non_blocking=True)
            n += k
        if n > self.width:
            raise ValueError("sample wider than the ring")
        self.meta[self.count, 0] = int(step)
        self.meta[self.count, 1] = int(bool(rejected))
        self.count += 1
        return self.count - 1

    @torch.no_grad()
    def gather(self, group=None):
        """The cycle-end collective.  Returns (samples [world, capacity, width],
        meta [world, capacity, 2]) -- rank r's block equals what an independent
        single-process run with chain r's seed would have saved."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self.rows.unsqueeze(0), self.meta.unsqueeze(0)
        world = dist.get_world_size(group)
        out = torch.empty(world, self.capacity, self.width, dtype=self.rows.dtype, device=self.rows.device)
        meta = torch.empty(world, self.capacity, 2, dtype=torch.int64, device=self.rows.device)
        dist.all_gather_into_tensor(out.view(world * self.capacity, self.width), self.rows,
                                    group=group)                 # the one data-path collective
        dist.all_gather_into_tensor(meta.view(world * self.capacity, 2), self.meta, group=group)   # 16 B per sample
        return out, meta


def unflatten_sample(row: torch.Tensor, offsets: Sequence[int], shapes: Sequence[Sequence[int]],
                     names: Sequence[str]) -> dict:
    """One ring row -> {parameter name: tensor}, the state_dict layout the
    reference's `_save_sample` stores (inference.py:189-197)."""
    out = {}
    for name, off, shape in zip(names, offsets, shapes):
        n = 1
        for s in shape:
            n *= int(s)
        out[name] = row[off:off + n].view(*shape) if len(shape) else row[off:off + 1].view(())
    return out
