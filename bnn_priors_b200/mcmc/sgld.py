"""SGLD with momentum (SGHMC), preconditioning and temperature diagnostics.

Mirror of the reference's `bnn_priors/mcmc/sgld.py` (class `SGLD`, :14-179): same
constructor, methods, `param_groups` / `state` keys and errors -- but a step is ONE
launch of the sm_100a kernel in csrc/bnnp_kernels.cu over the group's flat arrays
instead of ~7 eager ops and 2-4 `.item()` syncs per tensor.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, Optional, Sequence, Union

import torch

from .. import _native as N
from ._flat import FlatGroup


def dot(a, b):
    "return (a*b).sum().item()   (reference: mcmc/sgld.py:9-11; kept for API parity)"
    return (a.view(-1) @ b.view(-1)).item()


class SGLD(torch.optim.Optimizer):
    """Stochastic-gradient Langevin dynamics with momentum (SGHMC), a per-tensor RMSProp
    preconditioner and the kinetic / configurational temperature diagnostics of Wenzel et al. 2020.

    Constructor arguments: the reference's, in the reference's order (mcmc/sgld.py:31-34) --
        params            parameters, or dicts that define parameter groups
        lr                step size (rescaled by num_data inside: h = sqrt(lr / N), hn = sqrt(lr N))
        num_data          N, the size of the training set the minibatch gradients are averages over
        momentum          a in [0, 1): 0 gives plain SGLD (no momentum buffer is kept)
        temperature       T of the tempered posterior; 0 turns the sampler into SGD with momentum
        rmsprop_alpha     decay of the running mean of the squared gradients
        rmsprop_eps       regulariser added to that mean when the preconditioner is formed
        raise_on_no_grad  a parameter without gradient is an error (else it is skipped)
        raise_on_nan      a non-finite gradient raises ValueError

    Engine-only extras (keyword-only, not in the reference): `seed` / `chain` pick
    the Philox stream of the in-kernel noise (default: torch.initial_seed(), 0).
    """
    _OP = N.OP_SGLD

    def __init__(self, params: Sequence[Union[torch.nn.Parameter, Dict]], lr: float,
                 num_data: int, momentum: float = 0, temperature: float = 1.,
                 rmsprop_alpha: float = 0.99, rmsprop_eps: float = 1e-8,
                 raise_on_no_grad: bool = True, raise_on_nan: bool = False,
                 *, seed: Optional[int] = None, chain: int = 0):
        assert lr >= 0 and num_data >= 0 and momentum >= 0 and temperature >= 0
        defaults = dict(lr=lr, num_data=num_data, momentum=momentum,
                        rmsprop_alpha=rmsprop_alpha, rmsprop_eps=rmsprop_eps,
                        temperature=temperature)
        super(SGLD, self).__init__(params, defaults)
        self.raise_on_no_grad = raise_on_no_grad
        self.raise_on_nan = raise_on_nan
        if seed is None:
            seed = torch.initial_seed()
        self._flat = [FlatGroup(g['params'], seed, (chain << 16) + gi)
                      for gi, g in enumerate(self.param_groups)]
        for fg in self._flat:
            for p, s in zip(fg.params, fg.seg_states):
                self.state[p] = s
        self.update_preconditioner()
        self._step_count = 0  # keep the `torch.optim.scheduler` happy

    def add_param_group(self, param_group):
        """Groups are laid out in HBM when the sampler is constructed (the reference's runners pass
        all parameters to the constructor, inference.py:86-94); adding one later is refused rather
        than silently ignored."""
        if getattr(self, "_flat", None) is not None:
            raise NotImplementedError("bnn_priors_b200 samplers lay their parameter groups out at construction; "
                                      "pass every group to the constructor")
        super().add_param_group(param_group)

    # ------------------------------------------------------------------ engine access
    @property
    def flat_groups(self):
        "The FlatGroup (flat P/G/M arrays, segment table) of every param group."
        return self._flat

    def set_replay_noise(self, tensors) -> None:
        """Parity-test hook: the N(0,1) values the NEXT noise-consuming call uses
        instead of the in-kernel Philox stream -- one tensor per parameter, in
        `param_groups` order (what the reference drew through torch.randn_like)."""
        tensors = list(tensors)
        k = 0
        for fg in self._flat:
            fg.replay = fg.pack(tensors[k:k + fg.nseg])
            k += fg.nseg

    def _consume_replay(self, fg):
        fg.replay = None

    def _preconditioner_default(self, state, p) -> float:
        try:
            return state['preconditioner']
        except KeyError:
            v = state['preconditioner'] = 1.
            return v

    def zero_grad(self, set_to_none: bool = True):
        """inference.py:216 calls this every minibatch.  Like torch >= 2's default it drops the
        gradients (`p.grad = None`): backward() then hands every gradient over in a fresh tensor
        without an accumulation kernel, and the next sampler call copies all of them into the flat
        G array with one multi-tensor copy (FlatGroup.sync_views).  `set_to_none=False` zeroes G
        with one memset and binds p.grad to its views, so that autograd accumulates in place."""
        for fg in self._flat:
            if set_to_none:
                for p in fg.params:
                    p.grad = None
            else:
                fg.G.zero_()
                fg.bind_grad_views()

    def delta_energy(self, a, b) -> float:
        return math.inf

    # ------------------------------------------------------------------ sample_momentum
    @torch.no_grad()
    def sample_momentum(self, keep=0.0):
        "Refresh the momentum of every tensor: m <- sqrt(keep) m + sqrt(T (1 - keep)) eps  (mcmc/sgld.py:57-69)"
        assert 0 <= keep and keep <= 1.
        if keep == 1.:
            return
        for group, fg in zip(self.param_groups, self._flat):
            std = math.sqrt(group['temperature'] * (1 - keep))
            if keep == 0.0:
                fg.ensure_momentum_storage()
                # m = eps * std
                flags, cm = N.F_WRITE_M | N.F_NOISE_FIRST, 0.0
            else:
                if fg.M is None or any(s.raw_get('momentum_buffer') is None for s in fg.seg_states):
                    raise KeyError('momentum_buffer')
                # m.mul_(sqrt(keep)).add_(eps, alpha=std)
                flags, cm = N.F_READ_M | N.F_WRITE_M, math.sqrt(keep)
            noise = fg.take_noise_mode(True)
            fg.launch(N.OP_SAMPLE_MOMENTUM, N.PHASE_MID, flags, noise, cm=cm, cn=std)
            self._consume_replay(fg)
            fg.publish_momentum()
            fg._mm_version = fg.M._version

    # ------------------------------------------------------------------ steps
    @torch.no_grad()
    def step(self, closure: Optional[Callable[..., torch.Tensor]] = None,
             calc_metrics=True, save_state=False):
        assert save_state is False
        return self._step_internal(self._update_group_fn, self._step_fn,
                                   closure, calc_metrics=calc_metrics)
    initial_step = step

    @torch.no_grad()
    def final_step(self, closure: Optional[Callable[..., torch.Tensor]] = None,
                   calc_metrics=True, save_state=False):
        assert save_state is False
        return self._step_internal(self._update_group_fn, self._step_fn,
                                   closure, calc_metrics=calc_metrics,
                                   is_final=True)

    def _step_internal(self, update_group_fn, step_fn, closure, **step_fn_kwargs):
        """mcmc/sgld.py:88-112 with the per-tensor loop replaced by one launch per
        param group (`step_fn(group, flat_group, missing, **kw)`)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group, fg in zip(self.param_groups, self._flat):
            update_group_fn(group)
            missing = fg.sync_views(self.raise_on_no_grad)
            chunks = None
            if missing:
                chunks = fg.chunks_without(missing)
                if chunks is None:
                    continue
            step_fn(group, fg, chunks, **step_fn_kwargs)
            if self.raise_on_nan:
                self._raise_if_nonfinite(fg, missing)
        return loss

    def _raise_if_nonfinite(self, fg, missing=()):
        st = fg.fetch()
        for i, p in enumerate(fg.params):
            if i not in missing and st[i, N.S_NONFINITE] != 0.0:
                raise ValueError(
                    f"Gradient of shape {p.shape} is not finite: {p.grad}")

    def _update_group_fn(self, g):
        g['hn'] = math.sqrt(g['lr'] * g['num_data'])
        g['h'] = math.sqrt(g['lr'] / g['num_data'])
        g['noise_std'] = math.sqrt(2 * (1 - g['momentum']) * g['temperature'])

    def _prior_flags(self, fg, group):
        "flags / 1/N of the fused prior gradient (prior_fusion.py), if enabled"
        if not fg.prior_fused:
            return 0, 0.0
        f = N.F_PRIOR_GRAD
        if fg.clamp_active:
            f |= N.F_CLAMP_GRAD
        if fg.has_hyper and not fg.hyper_fresh():
            # hierarchical priors: current scales and -(1/N) dlog p/du come from a read-only
            # pre-pass (free if model.log_prior() already ran it for these parameters)
            fg.hyper_prepass(1.0 / group['num_data'])
        return f, 1.0 / group['num_data']

    def _step_fn(self, group, fg: FlatGroup, chunks, calc_metrics=True, is_final=False):
        """One SGLD transition of a whole group (mcmc/sgld.py:119-154); a final step only produces
        the diagnostics and leaves parameters and momentum alone."""
        a = group['momentum']
        pf, inv_n = self._prior_flags(fg, group)
        flags = N.F_READ_P | N.F_READ_G | pf
        if calc_metrics:
            flags |= N.F_CALC_METRICS
        if a > 0:
            fg.check_momentum()
            flags |= N.F_READ_M
        else:
            if is_final and calc_metrics:
                # the reference reads an unbound local here (sgld.py:132-137)
                raise UnboundLocalError("cannot access local variable 'momentum' where it is not "
                                        "associated with a value")
            flags |= N.F_MM_PRE_NOISE
        if is_final:
            if not (calc_metrics or self.raise_on_nan):
                return
            noise = N.NOISE_NONE
            # no writes; m' = 1*m so that the sums describe the stored momentum
            fg.launch(self._OP, N.PHASE_FINAL, flags, noise, cm=1.0 if a > 0 else 0.0,
                      inv_num_data=inv_n, chunks=chunks)
        else:
            flags |= N.F_WRITE_P | N.F_UPDATE_SQ
            if a > 0:
                flags |= N.F_WRITE_M
            flags |= fg.step_prior_flags(pf, chunks)
            noise = fg.take_noise_mode(group['temperature'] > 0)
            fg.launch(self._OP, N.PHASE_MID, flags, noise,
                      cm=a, cg=-group['hn'], cn=group['noise_std'], cp=group['h'],
                      inv_num_data=inv_n, rms_alpha=group['rmsprop_alpha'], chunks=chunks)
            if noise != N.NOISE_NONE:
                self._consume_replay(fg)
        if calc_metrics:
            fg.have_metrics = True
            fg.metrics_num_data = group['num_data']
        fg.note_step_sums(flags, self._OP)
        if flags & N.F_HYPER_POST:
            fg.after_hyper_post()

    # ------------------------------------------------------------------ preconditioner
    @torch.no_grad()
    def update_preconditioner(self):
        """Recompute every tensor's `state['preconditioner']` from the running mean of its squared
        gradients: M_t = ((mean_t + eps) / min_t' (mean_t' + eps)) ** (-1/4)  (mcmc/sgld.py:156-179).
        The engine keeps mean(square_avg) per tensor on the device -- square_avg is never consumed
        in any other way -- so this is one small D2H copy."""
        precond = OrderedDict()
        min_s = math.inf

        for group, fg in zip(self.param_groups, self._flat):
            eps = group['rmsprop_eps']
            st = fg.fetch()
            for i, p in enumerate(fg.params):
                precond[p] = float(st[i, N.S_SQ_MEAN]) + eps
                min_s = min(min_s, precond[p])

        for p, new_M in precond.items():
            # the sampler uses M^(-1/2) with M = sqrt(mean / min mean): one exponent of -1/4
            self.state[p]['preconditioner'] = (new_M / min_s)**(-1 / 4)
