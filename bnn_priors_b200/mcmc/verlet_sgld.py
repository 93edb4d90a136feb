"""GGMC / OBABO ("VerletSGLD") with Metropolis-Hastings correction.

Mirror of the reference's `bnn_priors/mcmc/verlet_sgld.py` (class `VerletSGLD`,
:8-197).  The initial / intermediate / final transitions, the snapshot for
rejection, the per-tensor delta-energy bookkeeping and the temperature diagnostics
are all done by one kernel launch per transition (csrc/bnnp_kernels.cu); the
accept/reject decision stays on the host, with the uniform drawn from the CPU
generator exactly as the reference does (verlet_sgld.py:61).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch

from .. import _native as N
from ._flat import FlatGroup
from .sgld import SGLD, dot  # noqa: F401  (dot re-exported like the reference)


class VerletSGLD(SGLD):
    """Gradient-guided Monte Carlo: underdamped Langevin dynamics split as OBABO, with the
    energy bookkeeping that makes a Metropolis-Hastings correction possible.  Same arguments
    as `SGLD` (the reference's signature, mcmc/sgld.py:31-34)."""
    _OP = N.OP_VERLET

    # ------------------------------------------------------------------ energies
    def _group_of(self, p):
        for group, fg in zip(self.param_groups, self._flat):
            for i, q in enumerate(fg.params):
                if q is p:
                    return group, fg, i
        raise KeyError("parameter is not managed by this sampler")

    def _refresh_sums(self, group, fg: FlatGroup):
        """Make SUM_GG (and SUM_MM) in the segment state describe the current
        p.grad / momentum: free right after a step, one read-only launch if
        somebody changed them since."""
        fg.sync_views(raise_on_no_grad=True)     # the gradient pointers follow a re-bound p.grad
        if not fg.sums_fresh(need_mm=self._OP == N.OP_HMC):
            fg.reduce_now(1.0 / group['num_data'] if fg.prior_fused else 0.0)

    def delta_energy(self, prev_potential: float, potential: float) -> float:
        """Energy difference between the last `initial_step` and now (verlet_sgld.py:27-42):
        per-tensor running sums from the device + point energies + N (U - U_prev)."""
        num_data = self.param_groups[0]['num_data']
        assert all(g['num_data'] == num_data for g in self.param_groups), "unclear which `num_data` to use"
        delta_energy = 0.
        for group, fg in zip(self.param_groups, self._flat):
            self._refresh_sums(group, fg)
            if not fg.have_delta:
                raise KeyError('delta_energy')
            st = fg.fetch()
            for i, p in enumerate(fg.params):
                point_energy = self._point_energy_i(group, fg, i)
                delta_energy += float(st[i, N.S_DELTA_ENERGY]) + point_energy

        if isinstance(potential, torch.Tensor):
            potential = potential.item()
        delta_energy += (potential - prev_potential) * num_data
        return delta_energy

    def _point_energy_i(self, group, fg: FlatGroup, i: int) -> float:
        M_rsqrt = float(fg.table["precond"][i])
        curv = M_rsqrt**2 * group['num_data']**2 * group['b^2h^2'] / 8
        return curv * float(fg.fetch()[i, N.S_SUM_GG])

    def _point_energy(self, group, p, state):
        "verlet_sgld.py:44-47; dot(p.grad, p.grad) comes from the kernel's reduction"
        _, fg, i = self._group_of(p)
        self._refresh_sums(group, fg)
        return self._point_energy_i(group, fg, i)

    @torch.no_grad()
    def maybe_reject(self, delta_energy: float) -> (bool, float):
        """Metropolis-Hastings test (verlet_sgld.py:49-70).  The uniform comes from torch's CPU
        generator with the same single draw as the reference, so host RNG streams stay aligned;
        a rejection restores P, G, M from the snapshot with one device-to-device kernel."""
        temperature = self.param_groups[0]['temperature']
        assert all(g['temperature'] == temperature for g in self.param_groups), "unclear which `temperature` to use"
        if temperature == 0.0:
            return False, 0.  # descent phase: never reject
        log_accept_prob = -delta_energy / temperature
        reject = (math.log(torch.rand(()).item()) > log_accept_prob)
        if reject:
            for fg in self._flat:
                fg.sync_views(raise_on_no_grad=False)
                fg.rollback()
        return reject, log_accept_prob

    # ------------------------------------------------------------------ transitions
    # OBABO splitting: an `initial_step` (theta(n), m(n) -> theta(n+1), u(n+1)), any number of
    # `step`s (theta(n), u(n) -> theta(n+1), u(n+1)) and a `final_step` (theta(n), u(n) ->
    # theta(n), m(n)), where u = sqrt(a) m + noise is the half-updated momentum.  The three
    # differ only in the scalars below (verlet_sgld.py:96-101, :129-134, :138-146).
    @staticmethod
    def _phase_scalars(a, temperature, phase):
        "(mom_decay, grad_v, noise_std) of a transition for momentum parameter a"
        if phase == N.PHASE_MID:
            return a, 1 + a, math.sqrt((1 - a**2) * temperature)
        root_a = math.sqrt(a)
        half_noise = math.sqrt((1 - a) * temperature)
        return root_a, (1. if phase == N.PHASE_INITIAL else root_a), half_noise

    def _update_group_fn(self, g, *, phase=N.PHASE_MID):
        "derived entries of a param group, from its CURRENT lr / num_data / momentum / temperature"
        g['b^2h^2'] = g['lr'] / g['num_data']
        g['bh'] = math.sqrt(g['b^2h^2'])
        g['bhn'] = math.sqrt(g['lr'] * g['num_data'])
        g['mom_decay'], g['grad_v'], g['noise_std'] = self._phase_scalars(
            g['momentum'], g['temperature'], phase)

    def _update_group_for(self, group, phase: int) -> None:
        self._update_group_fn(group, phase=phase)

    def _coefs(self, group, fg: FlatGroup, phase: int):
        bhn = group['bhn']
        inv_n = 1.0 / group['num_data'] if fg.prior_fused else 0.0
        return (group['mom_decay'], -.5 * group['grad_v'] * bhn, group['noise_std'], group['bh'], inv_n,
                -.5 * bhn, group['num_data']**2 * group['b^2h^2'] / 8, group['rmsprop_alpha'])

    def _transition(self, phase, closure, **step_kwargs):
        if phase != N.PHASE_MID:
            # keep a `torch.optim.lr_scheduler` happy (verlet_sgld.py:95,128)
            self._step_count = getattr(self, '_step_count', 0) + 1
        return self._step_internal(lambda g: self._update_group_fn(g, phase=phase), self._step_fn, closure,
                                   is_initial=(phase == N.PHASE_INITIAL), is_final=(phase == N.PHASE_FINAL),
                                   **step_kwargs)

    def initial_step(self, closure: Optional[Callable[..., torch.Tensor]] = None,
                     save_state=True, calc_metrics=True):
        "First transition after (re)sampling / accepting: optionally snapshots the state for a rejection."
        return self._transition(N.PHASE_INITIAL, closure, save_state=save_state, calc_metrics=calc_metrics)

    def _step_impl(self, closure: Optional[Callable[..., torch.Tensor]] = None, calc_metrics=True):
        return self._transition(N.PHASE_MID, closure, calc_metrics=calc_metrics)

    def step(self, closure: Optional[Callable[..., torch.Tensor]] = None,
             calc_metrics=True):
        "An intermediate transition."
        if self._hooks_active():
            return self._hooked("step", closure, calc_metrics=calc_metrics)
        return self._transition(N.PHASE_MID, closure, calc_metrics=calc_metrics)
    step.hooked = True

    def final_step(self, closure: Optional[Callable[..., torch.Tensor]] = None,
                   calc_metrics=True):
        "Last transition before the M-H test: completes the momentum, leaves the parameters where they are."
        return self._transition(N.PHASE_FINAL, closure, calc_metrics=calc_metrics)

    def _phase(self, is_initial, is_final):
        return N.PHASE_INITIAL if is_initial else (N.PHASE_FINAL if is_final else N.PHASE_MID)

    def _step_fn(self, group, fg: FlatGroup, chunks, is_initial=False, is_final=False,
                 save_state=False, calc_metrics=True):
        """One GGMC transition of a whole group (mcmc/verlet_sgld.py:149-197):
        m' = noise_std*eps + grad_lr*g + mom_decay*m ;  p += bh*M*m' (unless final)."""
        fg.check_momentum()
        pf, inv_n = self._prior_flags(fg, group)
        flags = N.F_READ_P | N.F_READ_G | N.F_READ_M | N.F_WRITE_M | N.F_NOISE_FIRST | pf
        if calc_metrics:
            flags |= N.F_CALC_METRICS
        if save_state:
            fg.ensure_prev_storage(with_momentum=group['momentum'] > 0)
            flags |= N.F_SAVE_STATE
        if not is_final:
            flags |= N.F_WRITE_P | N.F_UPDATE_SQ
            flags |= fg.step_prior_flags(pf, chunks)
        # the reference draws randn_like(p) even when noise_std == 0 (:163); a replayed
        # trace carries that draw, the Philox stream simply skips it
        noise = fg.take_noise_mode(True)
        if noise == N.NOISE_PHILOX and group['noise_std'] == 0.0:
            noise = N.NOISE_NONE
        phase = self._phase(is_initial, is_final)
        fg.launch_coef(self._OP, phase, flags, noise, self._coefs(group, fg, phase), chunks)
        self._consume_replay(fg)
        if calc_metrics:
            fg.have_metrics = True
            fg.metrics_num_data = group['num_data']
        fg.have_prev_new = True
        if is_initial:
            fg.have_delta = True
        # delta_energy() / _point_energy() follow initial / final steps and steps that log metrics
        # (inference.py:321-358, inference_reject.py:96-123): remember which gradient SUM_GG describes
        fg.note_step_sums(flags, self._OP, capture_grads=(is_initial or is_final or calc_metrics))
        if flags & N.F_HYPER_POST:
            fg.after_hyper_post()
