"""HMC with Verlet (leapfrog) integration.

Mirror of the reference's `bnn_priors/mcmc/hmc.py` (class `HMC`, :10-79): really
`VerletSGLD` with momentum=1 and temperature=1, no noise inside a trajectory, and
the kinetic energy 1/2 m.m as the point energy of the M-H acceptance probability.
The user calls `sample_momentum` to refresh the momentum between trajectories.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

import torch

from .. import _native as N
from ._flat import FlatGroup
from .sgld import dot  # noqa: F401
from .verlet_sgld import VerletSGLD


class HMC(VerletSGLD):
    """Hamiltonian Monte Carlo with the leapfrog integrator: GGMC at momentum 1 and temperature 1.

    Constructor arguments: the reference's (mcmc/hmc.py:25-27) -- params, lr, num_data as for
    `SGLD`; raise_on_no_grad; raise_on_nan (default True here: a trajectory with a non-finite
    gradient is useless).
    """
    _OP = N.OP_HMC

    def __init__(self, params: Sequence[Union[torch.nn.Parameter, Dict]],
                 lr: float, num_data: int,
                 raise_on_no_grad: bool = True, raise_on_nan: bool = True,
                 *, seed: Optional[int] = None, chain: int = 0, capturable: bool = False):
        super().__init__(params, lr, num_data, 1., 1.,
                         raise_on_no_grad=raise_on_no_grad,
                         raise_on_nan=raise_on_nan, seed=seed, chain=chain, capturable=capturable)

    def _point_energy_i(self, group, fg: FlatGroup, i: int) -> float:
        "hmc.py:32-33: .5 * dot(momentum, momentum) of the momentum now stored"
        return .5 * float(fg.fetch()[i, N.S_SUM_MM])

    def _update_group_fn(self, g, *, phase=N.PHASE_MID):
        # whatever a runner or scheduler wrote into the group, HMC is a = 1, T = 1 (hmc.py:35-39)
        super()._update_group_fn(g, phase=phase)
        assert g['momentum'] == 1. and g['temperature'] == 1.

    def _coefs(self, group, fg: FlatGroup, phase: int):
        inv_n = 1.0 / group['num_data'] if fg.prior_fused else 0.0
        return (1.0, -.5 * group['grad_v'] * group['bhn'], 0.0, group['bh'], inv_n, 0.0, 0.0, group['rmsprop_alpha'])

    def _step_fn(self, group, fg: FlatGroup, chunks, is_initial=False, is_final=False,
                 save_state=False, calc_metrics=True):
        """One leapfrog transition of a whole group (mcmc/hmc.py:41-79):
        m += grad_lr*g  (grad_v = 1 / 2 / 1 for initial / intermediate / final);
        p += bh*M*m unless final.  No noise."""
        fg.check_momentum()
        pf, inv_n = self._prior_flags(fg, group)
        flags = N.F_READ_P | N.F_READ_G | N.F_READ_M | N.F_WRITE_M | pf
        if calc_metrics:
            flags |= N.F_CALC_METRICS
        if is_initial or is_final:
            flags |= N.F_ALL_SUMS        # the kinetic energy .5 m.m enters delta_energy (:50-53, :32-33)
        if save_state:
            fg.ensure_prev_storage(with_momentum=True)
            flags |= N.F_SAVE_STATE
        if not is_final:
            flags |= N.F_WRITE_P | N.F_UPDATE_SQ
            flags |= fg.step_prior_flags(pf, chunks)
        phase = self._phase(is_initial, is_final)
        fg.launch_coef(self._OP, phase, flags, N.NOISE_NONE, self._coefs(group, fg, phase), chunks)
        if calc_metrics:
            fg.have_metrics = True
            fg.metrics_num_data = group['num_data']
        if is_initial:
            fg.have_delta = True
        fg.note_step_sums(flags, self._OP, capture_grads=(is_initial or is_final or calc_metrics))
        if flags & N.F_HYPER_POST:
            fg.after_hyper_post()
