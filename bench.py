#!/usr/bin/env python
"""bench.py -- SG-MCMC sampler throughput on B200 (BASELINE.json metric:
"SGLD param-updates/sec (25M-param net) ...; % HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one sampler transition (SGLD.step, calc_metrics=False) over the flat
parameter / gradient / momentum arrays of the 25,124,842-parameter
`vwidth_resnet18 width=96` segment table (tests/golden/model_shapes.json, taken
from the reference's get_model), synthetic gradients.  The working set
(3 x 100 MB) is larger than the 126 MB L2, so consecutive steps stream from HBM.

One JSON line on stdout (rank 0):
  value        param-updates/s, all ranks, inputs resident in HBM, through the
               sampler's public API (`opt.step`)
  roofline     12 B/param algorithmic bytes / kernel time against MEASURED_PEAKS.json, in BOTH
               regimes: `frac` = back-to-back C-ABI launches (launch k+1 starts on what launch k
               left in L2), `production` = the same launch with the L2 evicted in between, which
               is what a training loop sees (a forward/backward pass sits between two steps)
  samplers     the same two numbers for VerletSGLD.step and HMC.step, all ranks
  e2e          same metric with the step's gradient coming from pinned HOST memory
               and the diagnostics read back to the host every step
  cpu_baseline the numpy oracle port of the reference sampler on the host cores
`--impl reference` times the reference's CPU sampler on the same workload: the
unmodified `bnn_priors.mcmc` classes from oracle/_ref (when the snapshot is there)
next to the two ports under oracle/; `value` is the fastest of them.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SGLD param-updates/sec (25M-param net)"
UNIT = "param-updates/s"
WORKLOAD = "vwidth_resnet18_w96_cifar10_gaussian"
HP = dict(lr=5e-4, num_data=50000.0, momentum=0.994, temperature=1.0)
ALG_BYTES_PER_PARAM = 12          # p, grad, momentum as fp32 (BASELINE.json north_star)


def workload_config(world: int, n_params: int, tensors: int, total_floats: int) -> dict:
    "what is being computed -- identical keys and values in both arms (`--impl reference` and the GPU arm)"
    return {"workload": WORKLOAD, "n_params": n_params, "tensors": tensors, "sampler": "SGLD",
            "calc_metrics": False, **HP, "chains": world,
            "parallelism": f"{world} independent chains, no per-step collective",
            "l2": "inputs (3 x %.0f MB) larger than the 126 MB L2, no flush inside the timed loop" % (4 * total_floats / 1e6)}


def padded_total(tensors) -> int:
    "floats of one flat array in the chain's layout (segments start on 32-float lines)"
    import numpy as np
    return int(sum((int(np.prod(t["shape"]) if t["shape"] else 1) + 31) // 32 * 32 for t in tensors))


def load_tensors(tag=WORKLOAD):
    with open(os.path.join(ROOT, "tests", "golden", "model_shapes.json")) as f:
        return json.load(f)[tag]["tensors"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """(steady-state, cold-cache) DRAM bytes per launch of the step kernel from the committed ncu
    captures (profiles/ncu_step_kernel.json), or (None, None)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_step_kernel.json")) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), d.get("dram_bytes_per_launch_cold")
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------
# clocks: sampled with NVML during the timed regions
# ---------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {getattr(nv, n): n[len("nvmlClocksEventReason"):] for n in dir(nv)
                 if n.startswith("nvmlClocksEventReason") and isinstance(getattr(nv, n), int)}
        if not names:
            names = {getattr(nv, n): n[len("nvmlClocksThrottleReason"):] for n in dir(nv)
                     if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bit and (r & bit) and name not in ("None", "All"):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference sampler (oracle/sgmcmc_oracle.py),
# the 25M-parameter chain cut into one sub-chain per host thread.
# ---------------------------------------------------------------------------------
def build_cpu_chains(threads: int, seed: int = 0):
    import numpy as np
    from oracle import sgmcmc_oracle as O
    rng = np.random.default_rng(seed)
    tensors = load_tensors()
    pieces = [[] for _ in range(threads)]
    load = [0] * threads
    for t in tensors:
        n = int(np.prod(t["shape"])) if t["shape"] else 1
        # cut big tensors so that every thread gets a similar number of elements
        k = max(1, min(threads, n // 65536))
        for part in np.array_split(np.arange(n), k):
            j = load.index(min(load))
            pieces[j].append(part.size)
            load[j] += part.size
    chains = []
    for sizes in pieces:
        if not sizes:
            continue
        ps = [rng.standard_normal(s).astype(np.float32) * np.float32(0.05) for s in sizes]
        ch = O.Chain(ps, O.Group(**HP))
        for seg in ch.segs:
            seg.g = rng.standard_normal(seg.p.size).astype(np.float32) * np.float32(1e-3)
        chains.append(ch)
    n_params = sum(load)
    return O, chains, n_params


def time_cpu(steps: int, warmup: int, threads: int):
    """Seconds per step of the oracle's sgld_step over the whole 25M chain."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    O, chains, n_params = build_cpu_chains(threads)
    rngs = [np.random.default_rng(100 + i) for i in range(len(chains))]
    noises = [(lambda i, d, r=r: r.standard_normal(d, dtype=np.float32)) for r in rngs]
    for ch, nz in zip(chains, noises):
        O.sample_momentum(ch, nz)

    def one(j):
        O.sgld_step(chains[j], noises[j], calc_metrics=False)

    with ThreadPoolExecutor(max_workers=len(chains)) as ex:
        for _ in range(warmup):
            list(ex.map(one, range(len(chains))))
        t0 = time.perf_counter()
        for _ in range(steps):
            list(ex.map(one, range(len(chains))))
        dt = time.perf_counter() - t0
    return dt / steps, n_params, len(chains)


def time_cpu_torch(steps: int, warmup: int, threads: int):
    """Seconds per step of the reference's own op sequence (oracle/sgmcmc_torch.py: one tensor at
    a time, in-place ATen CPU kernels with intra-op threading, torch.randn_like for the noise)."""
    import torch
    from oracle import sgmcmc_torch as OT
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    tensors = load_tensors()
    params = [torch.randn(tuple(t["shape"]), generator=g) * (t["scale"] if t["kind"] else 1.0) for t in tensors]
    ch = OT.TorchSGLDChain(params, **HP)
    for t in ch.g:
        t.normal_(0.0, 1e-3, generator=g)
    ch.sample_momentum()
    for _ in range(warmup):
        ch.step(calc_metrics=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        ch.step(calc_metrics=False)
    dt = time.perf_counter() - t0
    return dt / steps, sum(int(p.numel()) for p in params)


def reference_package():
    """The unmodified reference (`bnn_priors.mcmc`) from the oracle/_ref snapshot, or None.  The snapshot
    is made by oracle/make_ref.py in the build container and travels with the tree; nothing outside
    `--impl reference` and tests/ imports it."""
    root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(root, "bnn_priors", "mcmc", "sgld.py")):
        return None
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    return importlib.import_module("bnn_priors.mcmc")          # needs torch + numpy only (mcmc/sgld.py:1-6)


def time_reference_package(ref, threads: int, steps: int, warmup: int):
    """the reference's own classes, CPU tensors of the workload's shapes, all host
    threads for ATen.  Returns {call: {ms_per_step, value, steps}}."""
    import torch
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    tensors = load_tensors()

    def fresh():
        ps = [torch.nn.Parameter(torch.randn(tuple(t["shape"]), generator=g) * (t["scale"] if t["kind"] else 1.0))
              for t in tensors]
        for p in ps:
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
        return ps
    n = sum(int(p.numel()) for p in fresh())
    out = {}

    def timed(name, fn, k, w):
        for _ in range(w):
            fn()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        sec = (time.perf_counter() - t0) / k
        out[name] = {"ms_per_step": sec * 1e3, "value": n / sec, "steps": k, "warmup": w, "kind": "reference"}

    ps = fresh()
    opt = ref.SGLD(ps, **HP)
    opt.sample_momentum()
    timed("SGLD.step(calc_metrics=False)", lambda: opt.step(calc_metrics=False), steps, warmup)
    few = max(2, min(steps, 5))
    timed("SGLD.step(calc_metrics=True)", lambda: opt.step(calc_metrics=True), few, 1)
    del opt
    opt = ref.VerletSGLD(ps, **HP)
    opt.sample_momentum()
    timed("VerletSGLD.initial_step(save_state=True)", lambda: opt.initial_step(save_state=True, calc_metrics=False), few, 1)
    timed("VerletSGLD.step(calc_metrics=False)", lambda: opt.step(calc_metrics=False), few, 1)
    timed("VerletSGLD.step(calc_metrics=True)", lambda: opt.step(calc_metrics=True), few, 1)
    timed("VerletSGLD.final_step(calc_metrics=True)", lambda: opt.final_step(calc_metrics=True), few, 1)
    timed("VerletSGLD.delta_energy", lambda: opt.delta_energy(1.0, 1.1), few, 1)
    for gr in opt.param_groups:
        gr["temperature"] = 1e-30                 # the Metropolis test always rejects: the restore is timed
    timed("VerletSGLD.maybe_reject(rejecting)", lambda: opt.maybe_reject(1e30), few, 1)
    del opt
    opt = ref.HMC(ps, lr=HP["lr"], num_data=HP["num_data"], raise_on_nan=False)
    opt.sample_momentum()
    opt.initial_step(save_state=False, calc_metrics=False)
    timed("HMC.step(calc_metrics=False)", lambda: opt.step(calc_metrics=False), few, 1)
    timed("HMC.step(calc_metrics=True)", lambda: opt.step(calc_metrics=True), few, 1)
    return out, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # (a) the numpy port, the 25M chain cut into one sub-chain per host thread (parallel noise):
    #     exactly --steps / --warmup
    sec_np, n_params, used = time_cpu(steps, max(1, warmup), threads)
    arms = {"numpy_port_one_subchain_per_thread": {"ms_per_step": sec_np * 1e3, "value": n_params / sec_np, "steps": steps,
                                                   "warmup": max(1, warmup), "kind": "port", "cores": used}}
    # (b) the reference's own torch op sequence restated (oracle/sgmcmc_torch.py), all host threads for
    #     ATen's intra-op parallelism; bounded: its serial randn_like makes a step take ~0.3 s
    t_steps = max(2, min(steps, 12))
    sec_t, _ = time_cpu_torch(t_steps, 1, threads)
    arms["torch_ops_like_the_reference"] = {"ms_per_step": sec_t * 1e3, "value": n_params / sec_t, "steps": t_steps,
                                            "warmup": 1, "kind": "port", "cores": threads}
    # (c) the UNMODIFIED reference classes (oracle/_ref); bounded like (b)
    ref = reference_package()
    ref_calls = None
    if ref is not None:
        ref_calls, _ = time_reference_package(ref, threads, t_steps, 1)
        main = ref_calls["SGLD.step(calc_metrics=False)"]
        arms["reference_bnn_priors_mcmc_SGLD"] = dict(main, cores=threads)
    # headline of this arm: the FASTEST CPU arm (the strongest baseline, so the ratio is conservative)
    best = min(arms, key=lambda k: arms[k]["ms_per_step"])
    sec = arms[best]["ms_per_step"] * 1e-3
    value = n_params / sec
    tensors = load_tensors()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": arms[best]["steps"], "warmup": arms[best]["warmup"], "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, n_params, len(tensors), padded_total(tensors)),
        "impl_notes": "CPU arms of the reference sampler on the host cores (one chain, however many GPUs the other "
                      "arm uses); value = the fastest arm (" + best + ")",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arms[best]["cores"], "kind": arms[best]["kind"],
                         "sample": f"{arms[best]['steps']} full SGLD steps over all {n_params} parameters ({best})"},
        "cpu_arms": arms,
        "reference_calls": ref_calls,      # BASELINE.md section 4 table, from the unmodified reference (None: no snapshot)
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------
def make_chain(device, seed, sampler="SGLD", tag=WORKLOAD, fused_prior=False, **extra):
    import torch
    from bnn_priors_b200 import mcmc
    tensors = load_tensors(tag)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    params = []
    for t in tensors:
        shape = tuple(t["shape"])
        scale = t["scale"] if t["kind"] else 1.0
        params.append(torch.nn.Parameter(torch.randn(shape, device=device, generator=g) * scale))
    hp = dict(HP)
    if sampler == "HMC":
        hp = dict(lr=HP["lr"], num_data=HP["num_data"], raise_on_nan=False)
    opt = getattr(mcmc, sampler)(params, **hp, seed=seed, **extra)
    (fg,) = opt.flat_groups
    if fused_prior:
        for i, t in enumerate(tensors):
            fg.set_prior(i, t["kind"], t["loc"], t["scale"], t["df"])
        fg.prior_fused = True
    for p, v in zip(params, fg.g_views):
        p.grad = v
        v.normal_(0.0, 1e-3, generator=g)      # segment by segment: the padding stays zero
    opt.sample_momentum()
    return opt, params, fg


LEAD_IN = 2      # untimed calls between the barrier and the first timed one (see timed_gpu)


def timed_gpu(fn, steps, device, dist_on):
    """K calls of fn bracketed by barrier + synchronize, CUDA events on the current stream; returns
    milliseconds (max over ranks).  After the barrier a rank has waited for the slowest one with an idle
    GPU, so LEAD_IN more untimed calls run first and the start event is recorded behind them, in stream
    order: the K timed calls start on a busy, clocked-up GPU with the host already ahead, as they do in
    the middle of a run -- with only 20 timed steps the ramp-up of an idle GPU was 3-9 % of the window."""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    for _ in range(LEAD_IN):
        fn()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    return max_over_ranks(ms, device, dist_on)


def max_over_ranks(ms, device, dist_on):
    if not dist_on:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(t.item())


class L2Flush:
    """Evicts the 126 MB L2 between two timed launches: a 512 MB buffer overwritten by a memset
    (what a forward / backward pass does to the L2 between two sampler steps of a training loop)."""

    def __init__(self, device, mbytes=512):
        import torch
        self.buf = torch.empty(mbytes * 1024 * 1024, dtype=torch.uint8, device=device)

    def __call__(self):
        self.buf.zero_()


def timed_flushed(fn, steps, device, dist_on, flush, before=None):
    """Mean milliseconds of ONE call of fn with the L2 evicted before each call: a CUDA event pair
    around every call (the flush and `before` -- host-side preparation such as zero_grad() -- are
    outside the pairs), max over ranks."""
    import torch
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize(device)
    for e0, e1 in pairs:
        if before is not None:
            before()
        flush()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize(device)
    ms = sum(e0.elapsed_time(e1) for e0, e1 in pairs) / steps
    return max_over_ranks(ms, device, dist_on)


def regime_numbers(fg, n, K, device, dist_on, flush, peak):
    """kernel time of the chain's LAST launch re-issued through the C ABI, in three regimes"""
    for _ in range(3):
        fg.relaunch()
    back_to_back = timed_gpu(fg.relaunch, K, device, dist_on) / K
    fg.serpentine = False            # every launch in the same direction: nothing useful left in L2
    for _ in range(3):
        fg.relaunch()
    same_dir = timed_gpu(fg.relaunch, K, device, dist_on) / K
    fg.serpentine = True
    Kp = max(5, min(K, 50))
    production = timed_flushed(fg.relaunch, Kp, device, dist_on, flush)
    alg = ALG_BYTES_PER_PARAM * n

    def f(ms):
        return {"kernel_us": ms * 1e3, "achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak,
                "touched_GBs": 20 * n / (ms * 1e-3) / 1e9, "touched_frac": 20 * n / (ms * 1e-3) / 1e9 / peak}
    return f(back_to_back), f(same_dir), f(production)


def run_gpu(args):
    import torch
    from bnn_priors_b200 import chains as CH
    from bnn_priors_b200 import _native as N
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sampler path has no CPU implementation "
                         "(use --impl reference for the CPU arm)")
    rank, world, device = CH.init_chains()
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
    K, W = args.steps, max(3, args.warmup)
    peak, peak_src = measured_peak()
    flush = L2Flush(device)

    opt, params, fg = make_chain(device, CH.chain_seed(0, rank))
    n, nseg, total = fg.n_params, fg.nseg, fg.total
    step = lambda: opt.step(calc_metrics=False)   # noqa: E731
    sampler = ClockSampler(device.index)

    for _ in range(W):
        step()
    torch.cuda.synchronize(device)
    sampler.start()
    l0 = fg.launches
    ms_api = timed_gpu(step, K, device, dist_on)
    launches = fg.launches - l0 - LEAD_IN        # kernels between the two events (the lead-in steps run before the first)
    # host cost of one opt.step (enqueue only), for the record
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    host_us = (time.perf_counter() - t0) / K * 1e6
    torch.cuda.synchronize(device)

    # ---- kernel time: back-to-back launches / same direction / L2 evicted in between (production)
    r_b2b, r_same, r_prod = regime_numbers(fg, n, K, device, dist_on, flush, peak)

    # ---- the reference runner's loop: zero_grad() drops p.grad, backward() hands every gradient over
    #      in a tensor of its own, the L2 is gone; the step reads those tensors in place (no copy)
    grad_bufs = [torch.randn_like(p) * 1e-3 for p in params]

    def foreign_grads():
        opt.zero_grad()
        for p, gb in zip(params, grad_bufs):
            p.grad = gb.view_as(gb)          # a new tensor object at the address the allocator re-uses
    Kp = max(5, min(K, 50))
    for _ in range(3):
        foreign_grads()
        step()
    c0, tw0 = fg.copies, fg.table_writes
    ms_foreign = timed_flushed(step, Kp, device, dist_on, flush, before=foreign_grads)
    foreign = {"us_per_step": ms_foreign * 1e3, "frac": ALG_BYTES_PER_PARAM * n / (ms_foreign * 1e-3) / 1e9 / peak,
               "gradient_copies": fg.copies - c0, "pointer_table_writes": fg.table_writes - tw0,
               "what": "opt.step(calc_metrics=False) right after zero_grad() + gradients in tensors of their own, "
                       "L2 evicted by a 512 MB memset before every step; CUDA events around the step only"}
    for p, v in zip(params, fg.g_views):
        p.grad = v
    del grad_bufs

    # ---- end to end: gradient from pinned host memory in, diagnostics out, every step
    E = max(3, min(K, 30))
    # the step's input (the gradient, in the chain's flat layout) lives in pinned host memory
    host_g = torch.zeros(fg.total, dtype=torch.float32).pin_memory()
    for o, k in zip(fg.off, fg.numel):
        host_g[o:o + k].normal_(0, 1e-3)
    p0 = params[0]

    def e2e_step():
        fg.G.copy_(host_g, non_blocking=True)         # H2D, 4 bytes per parameter
        opt.step(calc_metrics=True)
        return opt.state[p0]["est_temperature"]      # D2H of the segment-state array + sync

    for _ in range(2):
        e2e_step()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(E):
        e2e_step()
    torch.cuda.synchronize(device)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # the bare copy alone, same buffers: what the host side of this box can deliver to N GPUs at once
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(E):
        fg.G.copy_(host_g, non_blocking=True)
    torch.cuda.synchronize(device)
    h2d_ms = (time.perf_counter() - t0) * 1e3 / E
    e2e_ms = max_over_ranks(e2e_ms, device, dist_on)
    h2d_ms = max_over_ranks(h2d_ms, device, dist_on)
    sampler.stop()
    h2d = 4 * fg.total
    d2h = fg.nseg * N.STATE_STRIDE * 8
    del host_g

    # ---- cycle-end all-gather of the samples (one per cycle, outside the step loop)
    gather_ms = None
    if dist_on:
        ring = CH.SampleRing(1, fg.total, device)
        ring.push(fg.P, step=0)
        ring.gather()
        torch.cuda.synchronize(device)
        gather_ms = timed_gpu(lambda: ring.gather(), 3, device, True) / 3
        del ring

    # ---- the other samplers of the path, ALL ranks (max over ranks), same chain size
    samplers = {}
    Ks = max(5, min(K, 100))
    for name, smp in (("VerletSGLD.step", "VerletSGLD"), ("HMC.step", "HMC")):
        del opt, params, fg
        torch.cuda.empty_cache()
        opt, params, fg = make_chain(device, CH.chain_seed(0, rank), smp)
        sstep = lambda: opt.step(calc_metrics=False)   # noqa: E731
        for _ in range(W):
            sstep()
        ms = timed_gpu(sstep, Ks, device, dist_on)
        b2b, _, prod = regime_numbers(fg, n, Ks, device, dist_on, flush, peak)
        samplers[name] = {"value": world * n * Ks / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / Ks, "steps": Ks,
                          "kernel_us": b2b["kernel_us"], "frac": b2b["frac"],
                          "production_kernel_us": prod["kernel_us"], "production_frac": prod["frac"]}

    # ---- other transitions of the path (kernel time, rank 0), for context
    extra = {"SGLD.step after zero_grad()+foreign grads, L2 flushed": foreign}
    if rank == 0 and not args.no_extra:
        for name, smp, fused, call in (
                ("VerletSGLD.step+fused_prior", "VerletSGLD", True, lambda o: o.step(calc_metrics=False)),
                ("VerletSGLD.step(calc_metrics=True)", "VerletSGLD", False, lambda o: o.step(calc_metrics=True)),
                ("VerletSGLD.initial_step(save_state)", "VerletSGLD", False,
                 lambda o: o.initial_step(save_state=True, calc_metrics=False)),
                ("HMC.initial_step(save_state)", "HMC", False,
                 lambda o: o.initial_step(save_state=True, calc_metrics=False))):
            del opt, params, fg
            torch.cuda.empty_cache()
            opt, params, fg = make_chain(device, 0, smp, fused_prior=fused)
            call(opt)
            b2b, _, prod = regime_numbers(fg, n, min(K, 50), device, False, flush, peak)
            extra[name] = {"us_per_step": b2b["kernel_us"], "production_us_per_step": prod["kernel_us"],
                           "param_updates_per_s": n / (b2b["kernel_us"] * 1e-6), "alg_GBs": b2b["achieved"]}

        # hierarchical priors: every prior-carrying weight tensor gets a sampled scale
        # (NormalGamma); a step = ONE launch (BNNP_F_HYPER_POST, its epilogue rides on the next step:
        # BNNP_F_HYPER_CHAIN), timed through the API
        try:
            from bnn_priors_b200 import mcmc
            del opt, params, fg
            torch.cuda.empty_cache()
            gen = torch.Generator(device=device).manual_seed(0)
            params, links = [], []
            for t in load_tensors():
                params.append(torch.nn.Parameter(torch.randn(tuple(t["shape"]), device=device, generator=gen)
                                                 * (t["scale"] if t["kind"] else 1.0)))
                if t["kind"] and len(t["shape"]) > 1:
                    links.append((len(params) - 1, len(params), t))
                    params.append(torch.nn.Parameter(torch.tensor(0.1, device=device)))
            opt = mcmc.VerletSGLD(params, **HP, seed=0)
            (fg,) = opt.flat_groups
            for w, h, t in links:
                fg.set_prior(w, N.PRIOR_NORMAL, 0.0, t["scale"], 3.0)
                fg.set_hyper_link(w, h, N.PRIOR_HYPER_GAMMA, 1.0, 1.0)
            fg.prior_fused = True
            for p, v in zip(params, fg.g_views):
                p.grad = v
                v.normal_(0.0, 1e-3, generator=gen)
            opt.sample_momentum()
            hstep = lambda: opt.step(calc_metrics=False)   # noqa: E731
            for _ in range(5):
                hstep()
            ms = timed_gpu(hstep, min(K, 50), device, False) / min(K, 50)
            extra[f"VerletSGLD.step+{len(links)}_sampled_scales(one launch, chained epilogue)"] = {
                "us_per_step": ms * 1e3, "param_updates_per_s": n / (ms * 1e-3), "alg_GBs": ALG_BYTES_PER_PARAM * n / (ms * 1e-3) / 1e9}
        except Exception as e:      # context only: never lose the headline because of it
            extra["VerletSGLD.step+sampled_scales"] = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N == 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sec, n_cpu, used = time_cpu(steps=8, warmup=1, threads=1)
        cpu = {"value": n_cpu / sec, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"8 full SGLD steps over all {n_cpu} parameters, oracle/sgmcmc_oracle.py (numpy, 1 thread)"}

    if rank != 0:
        return
    traffic, traffic_cold = ncu_traffic()
    value = world * n * K / (ms_api * 1e-3)
    ms_kernel = r_b2b["kernel_us"] * 1e-3
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_api / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, n, nseg, total),
        "impl_notes": {"noise": "in-kernel Philox4x32-10 + Box-Muller",
                       "timing": f"W warm-up steps; barrier + synchronize; {LEAD_IN} more untimed steps; start event; K timed steps; "
                                 "stop event; synchronize; max over ranks",
                       "l2": "consecutive launches walk the chain in opposite directions, so each starts on the lines "
                             "the previous one left in L2 (roofline.frac); roofline.production evicts the L2 in between",
                       "host_us_per_step": host_us},
        "roofline": {"bound": "hbm", "achieved": r_b2b["achieved"], "peak": peak, "unit": "GB/s",
                     "frac": r_b2b["frac"], "traffic": traffic,
                     "traffic_source": "ncu capture committed as profiles/ncu_step_kernel.json (not measured in this run)",
                     "peak_source": peak_src, "kernel_us": r_b2b["kernel_us"],
                     "alg_bytes_per_launch": ALG_BYTES_PER_PARAM * n,
                     # what a training loop sees: the L2 evicted between two steps (512 MB memset), one CUDA
                     # event pair per launch; all 20 B/param (read p, g, m; write p, m) cross the HBM pins
                     "production": dict(r_prod, traffic=traffic_cold, touched_bytes_per_launch=20 * n,
                                        how="L2 evicted by a 512 MB memset before every launch; event pair per launch"),
                     # every launch in the same direction, back to back (no useful L2 carry-over either)
                     "same_direction": dict(r_same, traffic=traffic_cold, touched_bytes_per_launch=20 * n),
                     # alternating directions (the default): part of the 20 B/param is served by L2
                     "dram_GBs": (traffic / (ms_kernel * 1e-3) / 1e9) if traffic else None,
                     "dram_frac": (traffic / (ms_kernel * 1e-3) / 1e9 / peak) if traffic else None},
        "samplers": samplers,
        "cpu_baseline": cpu,
        "e2e": {"value": world * n * E / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": E, "ms_per_step": e2e_ms / E,
                "bare_h2d_ms": h2d_ms, "bare_h2d_GBs_per_gpu": h2d / (h2d_ms * 1e-3) / 1e9,
                "note": "per step: the gradient (4 B/param) from pinned host memory in, the per-tensor diagnostics "
                        "(not the parameters: the chain stays resident) out; bare_h2d_* = the same copy alone, all "
                        "ranks at once -- the host-side ceiling of this box"},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "extra": extra,
    }
    if gather_ms is not None:
        line["cycle_gather_ms"] = gather_ms
    emit(line)


class QuietStdout:
    """Rank 0 must print exactly ONE line on stdout.  Libraries write there too (NCCL prints its
    version banner to stdout when the first communicator comes up), so file descriptor 1 is pointed
    at stderr for the whole run and the JSON line goes to the saved descriptor at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str) -> None:
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


OUT = None


def emit(obj) -> None:
    line = json.dumps(obj)
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global OUT
    OUT = QuietStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other transitions")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()                  # rank 0 may still be timing the context cases
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
