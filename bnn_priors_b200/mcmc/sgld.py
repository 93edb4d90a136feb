"""SGLD with momentum (SGHMC), preconditioning and temperature diagnostics.

Mirror of the reference's `bnn_priors/mcmc/sgld.py` (class `SGLD`, :14-179): same
constructor, methods, `param_groups` / `state` keys and errors -- but a step is ONE
launch of the sm_100a kernel in csrc/bnnp_kernels.cu over the group's flat arrays
instead of ~7 eager ops and 2-4 `.item()` syncs per tensor.
"""
from __future__ import annotations

import math
import sys
import weakref
from collections import OrderedDict
from typing import Callable, Dict, Optional, Sequence, Union

import torch

from .. import _native as N
from ._flat import FlatGroup


_OPT_MOD = sys.modules[torch.optim.Optimizer.__module__]      # torch.optim.optimizer (global step hooks)
_PROFILER = torch.autograd.profiler


_LIVE = weakref.WeakSet()


def live_samplers():
    "the samplers of this process that are still alive (sample_sink.FlatSampleSaver finds its chain here)"
    return list(_LIVE)


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
    the Philox stream of the in-kernel noise (default: torch.initial_seed(), 0);
    `capturable=True` keeps the launch state (Philox counter, walking direction, pending
    bookkeeping, coefficients) in device memory, so that steps can be recorded in a CUDA graph
    together with the forward / backward pass and replayed.
    """
    _OP = N.OP_SGLD

    def __init__(self, params: Sequence[Union[torch.nn.Parameter, Dict]], lr: float,
                 num_data: int, momentum: float = 0, temperature: float = 1.,
                 rmsprop_alpha: float = 0.99, rmsprop_eps: float = 1e-8,
                 raise_on_no_grad: bool = True, raise_on_nan: bool = False,
                 *, seed: Optional[int] = None, chain: int = 0, capturable: bool = False):
        assert lr >= 0 and num_data >= 0 and momentum >= 0 and temperature >= 0
        defaults = dict(lr=lr, num_data=num_data, momentum=momentum,
                        rmsprop_alpha=rmsprop_alpha, rmsprop_eps=rmsprop_eps,
                        temperature=temperature)
        super(SGLD, self).__init__(params, defaults)
        self.raise_on_no_grad = raise_on_no_grad
        self.raise_on_nan = raise_on_nan
        if seed is None:
            seed = torch.initial_seed()
        self.capturable = bool(capturable)
        self._flat = [FlatGroup(g['params'], seed, (chain << 16) + gi, capturable=self.capturable)
                      for gi, g in enumerate(self.param_groups)]
        for fg in self._flat:
            for p, s in zip(fg.params, fg.seg_states):
                self.state[p] = s
        self.update_preconditioner()
        self._step_count = 0  # keep the `torch.optim.scheduler` happy
        _LIVE.add(self)

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
        without an accumulation kernel, and the next sampler call reads those tensors where they
        lie (FlatGroup.sync_views: per-segment gradient pointers, no copy).  `set_to_none=False`
        zeroes G with one memset and binds p.grad to its views, so that autograd accumulates in
        place."""
        for fg in self._flat:
            if set_to_none:
                fg.drop_grads()
            else:
                fg.G.zero_()
                fg.bind_grad_views()

    # ------------------------------------------------------------------ torch.optim plumbing
    def _patch_step_function(self) -> None:
        """torch.optim.Optimizer wraps `step` in a profiler range plus pre / post hook dispatch,
        which costs more host time than this sampler's whole step.  `step` below does the same
        dispatch itself, but only when a hook is registered or a profiler is running."""
        self._zero_grad_profile_name = f"Optimizer.zero_grad#{self.__class__.__name__}.zero_grad"

    def _hooks_active(self) -> bool:
        return bool(self._optimizer_step_pre_hooks or self._optimizer_step_post_hooks
                    or _OPT_MOD._global_optimizer_pre_hooks or _OPT_MOD._global_optimizer_post_hooks
                    or _PROFILER._is_profiler_enabled)

    def _hooked(self, name, *args, **kwargs):
        "the slow path: the method `name` under torch's profile_hook_step wrapper"
        fn = torch.optim.Optimizer.profile_hook_step(getattr(type(self), "_" + name + "_impl"))
        return fn(self, *args, **kwargs)

    def state_dict(self):
        """torch's state_dict with every per-parameter state as a PLAIN dict (a snapshot of the lazy,
        device-backed one) plus `square_avg` (the engine keeps only its mean; written out as the
        constant tensor with that mean, which is all `update_preconditioner` reads: sgld.py:170-173)."""
        from ._flat import SegState
        sd = super().state_dict()
        plain = {}
        for idx, st in sd['state'].items():
            if isinstance(st, SegState):
                d = dict(st.items())
                d['square_avg'] = st._fg.square_avg_tensor(st._i)
                plain[idx] = d
            else:
                plain[idx] = st
        sd['state'] = plain
        return sd

    def load_state_dict(self, state_dict):
        """torch replaces `self.state[p]` by plain dicts; route the loaded values back into the flat
        arrays / the device-side segment state (momentum_buffer, preconditioner, square_avg,
        delta_energy, prev_*) and restore the lazy state dicts."""
        super().load_state_dict(state_dict)
        for fg in self._flat:
            loaded_momentum = False
            for p, s in zip(fg.params, fg.seg_states):
                loaded = self.state.get(p)
                if loaded is s:
                    continue
                self.state[p] = s
                for k, v in (loaded or {}).items():
                    if k in ('prev_parameter', 'prev_grad', 'prev_momentum_buffer'):
                        fg.ensure_prev_storage(with_momentum=(k == 'prev_momentum_buffer'))
                        with torch.no_grad():
                            s.raw_get(k).copy_(v)
                    elif k in ('est_temperature', 'est_config_temp'):
                        continue              # diagnostics of the last step: recomputed by the next one
                    else:
                        s[k] = v
                        loaded_momentum |= k == 'momentum_buffer'
            if loaded_momentum:
                fg.publish_momentum()
            fg.invalidate_sums()

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
    def _step_impl(self, closure: Optional[Callable[..., torch.Tensor]] = None,
                   calc_metrics=True, save_state=False):
        assert save_state is False
        return self._step_internal(self._update_group_fn, self._step_fn,
                                   closure, calc_metrics=calc_metrics)

    def step(self, closure: Optional[Callable[..., torch.Tensor]] = None,
             calc_metrics=True, save_state=False):
        if self._hooks_active():
            return self._hooked("step", closure, calc_metrics=calc_metrics, save_state=save_state)
        return self._step_impl(closure, calc_metrics, save_state)
    step.hooked = True          # torch.optim.Optimizer: do not wrap again (see _patch_step_function)
    initial_step = step

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
        grad_was_enabled = torch.is_grad_enabled()
        torch._C._set_grad_enabled(False)
        try:
            for group, fg in zip(self.param_groups, self._flat):
                update_group_fn(group)
                missing = fg.sync_views(self.raise_on_no_grad)
                chunks = None
                if missing:
                    chunks = fg.chunks_without(missing)
                    if chunks is None:
                        continue
                if self.raise_on_nan:
                    self._raise_if_nonfinite(fg, chunks, missing)
                step_fn(group, fg, chunks, **step_fn_kwargs)
        finally:
            torch._C._set_grad_enabled(grad_was_enabled)
        return loss

    def _raise_if_nonfinite(self, fg, chunks, missing=()):
        """sgld.py:102-104: a non-finite gradient raises BEFORE anything is updated.  One read-only
        launch over the gradients (4 B/param) + the read-back of the per-tensor flags; only with
        `raise_on_nan=True` (HMC's default)."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("raise_on_nan=True reads a flag back from the device at every step and cannot be "
                               "recorded in a CUDA graph: construct the sampler with raise_on_nan=False")
        fg.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_G, N.NOISE_NONE, cm=1.0, chunks=chunks)
        fg._gg_sig = None
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

    def _coefs(self, group, fg: FlatGroup, phase: int):
        """(cm, cg, cn, cp, 1/N, c_gm_base, curv_base, rms_alpha) of a transition, from the derived
        entries `_update_group_fn` left in the group (include/bnnp.h BnnpCoef)"""
        a = group['momentum']
        inv_n = 1.0 / group['num_data'] if fg.prior_fused else 0.0
        if phase == N.PHASE_FINAL:
            # no writes; m' = 1*m so that the sums describe the stored momentum
            return (1.0 if a > 0 else 0.0, 0.0, 0.0, 0.0, inv_n, 0.0, 0.0, 0.0)
        return (a, -group['hn'], group['noise_std'], group['h'], inv_n, 0.0, 0.0, group['rmsprop_alpha'])

    def _update_group_for(self, group, phase: int) -> None:
        self._update_group_fn(group)

    def sync_hyperparameters(self) -> None:
        """capturable=True: write the coefficients that follow from the CURRENT `param_groups`
        (lr, temperature, momentum, num_data -- what a scheduler or a runner changes between steps)
        into the device control block, so that the next replay of a captured step uses them.
        Eager calls do this themselves; only graph replays need it."""
        if not self.capturable:
            return
        for group, fg in zip(self.param_groups, self._flat):
            for phase in (N.PHASE_INITIAL, N.PHASE_MID, N.PHASE_FINAL):
                self._update_group_for(group, phase)
                fg.poke_coef(phase, self._coefs(group, fg, phase))

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
            if not calc_metrics:
                return
            noise = N.NOISE_NONE
            fg.launch_coef(self._OP, N.PHASE_FINAL, flags, noise, self._coefs(group, fg, N.PHASE_FINAL), chunks)
        else:
            flags |= N.F_WRITE_P | N.F_UPDATE_SQ
            if a > 0:
                flags |= N.F_WRITE_M
            flags |= fg.step_prior_flags(pf, chunks)
            noise = fg.take_noise_mode(group['temperature'] > 0)
            fg.launch_coef(self._OP, N.PHASE_MID, flags, noise, self._coefs(group, fg, N.PHASE_MID), chunks)
            if noise != N.NOISE_NONE:
                self._consume_replay(fg)
        if calc_metrics:
            fg.have_metrics = True
            fg.metrics_num_data = group['num_data']
        fg.note_step_sums(flags, self._OP, capture_grads=False)     # SGLD has no point energy to keep fresh
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
