"""The reference's SGLD transition restated with the torch ops the reference itself issues
(one tensor at a time, in-place ATen kernels, `.item()` for every dot product) -- TEST AND
BENCH INFRASTRUCTURE ONLY, like everything under oracle/.

Why it exists next to the numpy oracle: `bench.py --impl reference` has to time "the
reference's own CPU sampler" on the GPU box, where /root/reference does not exist.  The
numpy port (sgmcmc_oracle.py) is the parity yardstick; this file keeps the reference's
*cost structure* -- ATen's intra-op threading, its serial CPU `randn_like`, ~7 kernels and
2 syncs per tensor (bnn_priors/mcmc/sgld.py:119-154, looped by :88-112) -- so the CPU arm
is not flattered or penalised by numpy's way of doing the same arithmetic.  It runs on any
device; on a CUDA device it is the "eager GPU" baseline of SURVEY 8d.

tests/test_oracle_golden.py checks it against the numpy oracle on the golden SGLD traces.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch


class TorchSGLDChain:
    """params / grads / momenta as separate tensors, like `optimizer.state` in the reference."""

    def __init__(self, params: List[torch.Tensor], lr: float, num_data: float, momentum: float = 0.0,
                 temperature: float = 1.0, rmsprop_alpha: float = 0.99):
        self.p = [t.detach().clone() for t in params]
        self.g = [torch.zeros_like(t) for t in self.p]
        self.m: List[Optional[torch.Tensor]] = [None] * len(self.p)
        self.sq = [torch.ones_like(t) for t in self.p]            # sgld.py:170
        self.precond = [1.0] * len(self.p)
        self.lr, self.num_data, self.a, self.T, self.alpha = lr, num_data, momentum, temperature, rmsprop_alpha
        self.est_temperature = [math.nan] * len(self.p)
        self.est_config_temp = [math.nan] * len(self.p)

    @torch.no_grad()
    def sample_momentum(self, noise=None):
        "sgld.py:57-69 with keep = 0"
        std = math.sqrt(self.T)
        for i, t in enumerate(self.p):
            z = torch.randn_like(t) if noise is None else noise[i]
            self.m[i] = z * std

    @torch.no_grad()
    def step(self, calc_metrics: bool = False, noise=None):
        "sgld.py:114-154 for every tensor of the group (momentum > 0)"
        hn = math.sqrt(self.lr * self.num_data)
        h = math.sqrt(self.lr / self.num_data)
        noise_std = math.sqrt(2 * (1 - self.a) * self.T)
        for i, (p, g, m, sq) in enumerate(zip(self.p, self.g, self.m, self.sq)):
            M = self.precond[i]
            d = p.numel()
            if calc_metrics:
                self.est_temperature[i] = (m.view(-1) @ m.view(-1)).item() / d
            m.mul_(self.a).add_(g, alpha=-hn * M)
            if self.T > 0:
                z = torch.randn_like(m) if noise is None else noise[i]
                m.add_(z, alpha=noise_std)
            if calc_metrics:
                self.est_config_temp[i] = (p.view(-1) @ g.view(-1)).item() * (self.num_data / d)
            p.add_(m, alpha=h * M)
            sq.mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)


class TorchVerletChain(TorchSGLDChain):
    """The intermediate GGMC transition (bnn_priors/mcmc/verlet_sgld.py:138-197) with the
    reference's ops: a fresh momentum tensor per step, two `.item()` dot products per tensor for the
    delta-energy bookkeeping and two more for the diagnostics."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.delta_energy = [0.0] * len(self.p)
        self.prev_new_momentum_delta = [0.0] * len(self.p)

    @torch.no_grad()
    def step(self, calc_metrics: bool = False, noise=None):
        a = self.a
        bh = math.sqrt(self.lr / self.num_data)
        bhn = math.sqrt(self.lr * self.num_data)
        mom_decay, grad_v, noise_std = a, 1 + a, math.sqrt((1 - a * a) * self.T)
        for i, (p, g, m, sq) in enumerate(zip(self.p, self.g, self.m, self.sq)):
            M = self.precond[i]
            d = p.numel()
            z = torch.randn_like(p) if noise is None else noise[i]
            new_m = z.mul_(noise_std) if noise is None else z * noise_std
            new_m.add_(g, alpha=-.5 * grad_v * bhn * M).add_(m, alpha=mom_decay)
            c = -.5 * bhn * M
            self.delta_energy[i] += self.prev_new_momentum_delta[i] + c * (g.view(-1) @ m.view(-1)).item()
            self.prev_new_momentum_delta[i] = c * (g.view(-1) @ new_m.view(-1)).item()
            if calc_metrics:
                self.est_temperature[i] = (m.view(-1) @ m.view(-1)).item() / d
                self.est_config_temp[i] = (p.view(-1) @ g.view(-1)).item() * (self.num_data / d)
            self.m[i] = new_m
            p.add_(new_m, alpha=bh * M)
            sq.mul_(self.alpha).addcmul_(g, g, value=1 - self.alpha)
