"""Flat HBM storage of one Markov chain's parameter group and the launch plumbing
around `bnnp_launch` (include/bnnp.h).

The reference keeps one tensor per parameter for p, p.grad, momentum_buffer,
square_avg, prev_* and loops over them in Python (mcmc/sgld.py:94-105).  Here a
group owns flat fp32 arrays for the parameters, the momenta and (once asked for)
the snapshots; the model's parameters and state['momentum_buffer'] are re-pointed
to views of them, so the whole group is updated by one kernel launch and the
reference's callers (runners, lr schedulers, load_state_dict) keep seeing ordinary
tensors.  Gradients are read where autograd leaves them: a small device table holds
one pointer per tensor (the tensor backward() produced, or the tensor's slice of the
flat G array) and is rewritten only when an address changes.

torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import operator
import os
import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _native as N

# BNNP_NVTX=1: an NVTX range around every launch (names the op / phase / flags in an nsys or ncu timeline)
_NVTX = os.environ.get("BNNP_NVTX", "0") == "1"

_GRAD = operator.attrgetter("grad")
_VERSION = operator.attrgetter("_version")
_DATA_PTR = torch.Tensor.data_ptr
_IS = operator.is_
_NO_MISSING: List[int] = []

# keys of optimizer.state[p] whose values live on the device between launches
LAZY_SCALARS = ("est_temperature", "est_config_temp", "delta_energy", "prev_new_momentum_delta")


_NO_PENDING = (0, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0)


class SegState(dict):
    """optimizer.state[p].  A real dict (torch.optim.Optimizer expects one) whose
    device-resident entries are filled in from the segment-state array the first
    time somebody looks, so a step costs no host synchronisation unless a scalar
    is actually read (SURVEY 8b "Threading / sync")."""

    __slots__ = ("_fg", "_i")

    def __init__(self, fg: "FlatGroup", i: int):
        super().__init__()
        self._fg = fg
        self._i = i

    # -- reads
    def __getitem__(self, k):
        if k in LAZY_SCALARS:
            self._fg.materialize()
        elif k == "square_avg":
            return self._fg.square_avg_tensor(self._i)
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        try:
            return self[k]
        except KeyError:
            return default

    def __contains__(self, k):
        if k in LAZY_SCALARS:
            self._fg.materialize()
        elif k == "square_avg":
            return True
        return dict.__contains__(self, k)

    def _all(self):
        self._fg.materialize()
        return self

    def keys(self):
        return dict.keys(self._all())

    def items(self):
        return dict.items(self._all())

    def values(self):
        return dict.values(self._all())

    def __iter__(self):
        return dict.__iter__(self._all())

    def __len__(self):
        return dict.__len__(self._all())

    def __repr__(self):
        return dict.__repr__(self._all())

    # -- writes
    def __setitem__(self, k, v):
        fg, i = self._fg, self._i
        if k == "preconditioner":
            v = float(v)
            fg.set_preconditioner(i, v)
        elif k == "momentum_buffer":
            v = fg.set_momentum(i, v)
        elif k == "delta_energy":
            fg.poke(i, N.S_DELTA_ENERGY, float(v))
            fg.have_delta = True
        elif k == "prev_new_momentum_delta":
            fg.poke(i, N.S_PREV_NEW_MOM, float(v))
            fg.have_prev_new = True
        elif k == "square_avg":
            fg.poke(i, N.S_SQ_MEAN, float(v.double().mean()))
            return
        dict.__setitem__(self, k, v)

    def __delitem__(self, k):
        if k == "momentum_buffer":
            self._fg._mom_published = False
        dict.__delitem__(self, k)

    def pop(self, k, *default):
        if k == "momentum_buffer":
            self._fg._mom_published = False
        return dict.pop(self, k, *default)

    def raw_get(self, k, default=None):
        return dict.get(self, k, default)

    def raw_set(self, k, v):
        dict.__setitem__(self, k, v)


class FlatGroup:
    """One param group of one chain in HBM."""

    def __init__(self, params: Sequence[torch.nn.Parameter], seed: int, stream_id: int, capturable: bool = False):
        self.params: List[torch.nn.Parameter] = list(params)
        self.capturable = bool(capturable)
        if not self.params:
            raise ValueError("empty parameter group")
        dev = self.params[0].device
        for p in self.params:
            if p.device.type != "cuda":
                raise RuntimeError(
                    "bnn_priors_b200 samplers run on CUDA tensors only (got a parameter on "
                    f"{p.device}); there is no CPU implementation of this path")
            if p.device != dev:
                raise RuntimeError("all parameters of a group must live on one device")
            if p.dtype != torch.float32:
                raise RuntimeError(f"fp32 parameters only (got {p.dtype}); the kernels are fp32")
            if p.numel() == 0:
                raise RuntimeError("zero-sized parameters are not supported")
        self.device = dev
        self._dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.lib = N.lib()
        self.numel = [int(p.numel()) for p in self.params]
        self.nseg = len(self.params)
        off, first, nch, total, chunks = N.plan_layout(self.numel)
        self.off = [int(o) for o in off]
        self.total = total
        self.nchunks = int(chunks.size)
        self.n_params = int(sum(self.numel))

        # segment table (host copy + device copy)
        self.table = np.zeros(self.nseg, dtype=N.SEGMENT_DTYPE)
        self.table["off"], self.table["numel"] = off, self.numel
        self.table["precond"] = 1.0
        self.table["prior_scale"], self.table["prior_df"] = 1.0, 3.0
        self.table["first_chunk"], self.table["num_chunks"] = first, nch
        self.table["link"] = -1
        self.table_dev = torch.empty(self.table.nbytes, dtype=torch.uint8, device=dev)
        self._table_dirty = True
        self.chunk_seg_host = np.ascontiguousarray(chunks["seg"])
        self.chunks_dev = torch.from_numpy(np.ascontiguousarray(chunks).view(np.uint8).copy()).to(dev)   # BnnpChunk[nchunks]

        # flat arrays
        self.P = torch.zeros(total, dtype=torch.float32, device=dev)
        self.G = torch.zeros(total, dtype=torch.float32, device=dev)
        self.M: Optional[torch.Tensor] = None
        self.prev_p = self.prev_g = self.prev_m = None
        self._ptrs = {"P": self.P.data_ptr(), "G": self.G.data_ptr()}     # device addresses of the flat arrays
        self.p_views = [self.P[o:o + n].view(p.shape) for o, n, p in zip(self.off, self.numel, self.params)]
        self.g_views = [self.G[o:o + n].view(p.shape) for o, n, p in zip(self.off, self.numel, self.params)]
        self.m_views: Optional[List[torch.Tensor]] = None
        self._p_ptrs = [v.data_ptr() for v in self.p_views]
        self._gv_ptrs = [v.data_ptr() for v in self.g_views]
        # per-segment gradient pointers (BnnpLaunch.seg_grad): a segment's gradient is read where it
        # lies -- its slice of G, or the tensor autograd handed over after zero_grad() -- and this
        # small device table is rewritten (bnnp_poke) only when one of the addresses changed
        self.seg_grad_dev = torch.zeros(self.nseg, dtype=torch.int64, device=dev)
        self._g_ptrs: List[int] = []                  # what the device table holds
        self._held_grads: Optional[list] = None       # the p.grad tensors of the last sync (dropped by zero_grad)
        self._gg_sig = None                           # (grad tensors, their versions) SUM_GG describes

        # per-segment scalars (BNNP_S_*) on the device, mirrored on demand
        self.state_dev = torch.zeros(self.nseg, N.STATE_STRIDE, dtype=torch.float64, device=dev)
        self.state_dev[:, N.S_SQ_MEAN] = 1.0          # square_avg = ones (sgld.py:170)
        self.state_host = torch.zeros(self.nseg, N.STATE_STRIDE, dtype=torch.float64).pin_memory()
        self.state_host[:, N.S_SQ_MEAN] = 1.0
        self.state_np = self.state_host.numpy()
        # per-chunk partial records of the last two launches (ping-pong) and their launch stamps
        self.partials = torch.zeros(2 * self.nchunks * N.NRED, dtype=torch.float64, device=dev)
        self.stamps = torch.zeros(2 * self.nchunks, dtype=torch.int64, device=dev)
        self._parity = 0
        # capturable mode: call / parity / pending / coefficients live in a device control block
        # (include/bnnp.h BnnpControl) and every launch is followed by bnnp_advance
        self.ctl_dev = torch.zeros(C.sizeof(N.BnnpControl), dtype=torch.uint8, device=dev) if self.capturable else None
        self._slot_cache = [None] * N.COEF_SLOTS      # coefficients each control-block slot holds
        self._pending = None     # epilogue parameters of the last launch, not yet applied to state_dev
        self._epoch = 0          # bumped by every launch / poke
        self._host_epoch = 0     # epoch state_host corresponds to
        self._mat_epoch = 0      # epoch the SegState dicts correspond to
        self.have_metrics = self.have_delta = self.have_prev_new = False
        self.metrics_num_data = 1.0

        # noise
        self.key = N.philox_key(seed, stream_id)
        self.call = 0
        self.replay: Optional[torch.Tensor] = None

        # fused prior (prior_fusion.py)
        self.prior_fused = False
        self.grad_max: Optional[float] = None
        self.clamp_armed = True       # prior_fusion.FusedPrior.arm_clamp
        self._lp_valid = False
        self._lp_pversion = None
        # hierarchical priors: {hyper segment: weight segment}; see hyper_prepass
        self.hyper_links = {}
        self._hyper_valid = False
        self._hyper_pversion = None
        # validity of SUM_GG / SUM_MM in the segment state
        self._mm_version = None

        self._mom_published = False
        self.serpentine = os.environ.get("BNNP_SERPENTINE", "1") != "0"
        self.seg_states = [SegState(self, i) for i in range(self.nseg)]
        self._g_seen: List[Optional[tuple]] = [None] * self.nseg     # per tensor: (weakref to the foreign grad tensor, its version) last copied into G, or "zero"
        self.args = N.BnnpLaunch()
        self._argbuf = (C.c_char * C.sizeof(N.BnnpLaunch)).from_buffer(self.args)
        self._static_dirty = True        # pointer fields of self.args need a refresh
        self._var_ptrs = None            # (prev_m, replay_noise, chunk_ids) now in self.args
        self.launches = 0
        self.copies = 0                  # gradients that had to be copied into G (not readable in place)
        self.table_writes = 0            # rewrites of the device table of gradient pointers

        self.adopt_parameters()
        hm = N.host_module()
        # optional C++ helper for the per-step scan below (csrc/bnnp_host.cpp); None: the Python scan
        self._scanner = hm.GradScanner(self.params, self._p_ptrs) if hm is not None else None
        self._write_grad_table(list(self._gv_ptrs))

    # ------------------------------------------------------------------ views
    @torch.no_grad()
    def adopt_parameters(self) -> None:
        """Move every parameter's storage into the flat P array (in place for the
        model: `p.data` becomes a view)."""
        for i, (p, v) in enumerate(zip(self.params, self.p_views)):
            v.copy_(p.data)
            p.data = v
            if p.grad is not None and p.grad is not self.g_views[i]:
                self.g_views[i].copy_(p.grad)
                p.grad = self.g_views[i]

    def sync_views(self, raise_on_no_grad: bool) -> List[int]:
        """Make the launch arguments describe what the model holds now.  Gradients are read IN PLACE:
        after `zero_grad()` autograd stores every gradient in a tensor of its own, and the kernel
        reads it there through the per-segment pointer table (BnnpLaunch.seg_grad) -- the table is
        rewritten only when an address changed (in a steady training loop the caching allocator
        hands back the same blocks, so: never).  A parameter whose storage was swapped
        (`Prior.sample()`) is re-adopted.  Returns the indices of parameters that have no gradient
        (sgld.py:96-101)."""
        sc = self._scanner
        if sc is not None:
            # the same scan over the ATen objects: 0 = every gradient lies where the device table says and
            # every parameter is still its flat view (the steady state of a training loop)
            rc = sc.scan()
            if rc == 0:
                self._held_grads = None
                return _NO_MISSING
        grads = list(map(_GRAD, self.params))
        try:
            ptrs = list(map(_DATA_PTR, grads))
        except TypeError:                      # a parameter without gradient
            ptrs = None
        missing = _NO_MISSING
        if ptrs != self._g_ptrs:
            missing = self._sync_grads_slow(grads, raise_on_no_grad)
        self._held_grads = grads
        if list(map(_DATA_PTR, self.params)) != self._p_ptrs:
            self._readopt_parameters()
        return missing

    @torch.no_grad()
    def _sync_grads_slow(self, grads, raise_on_no_grad: bool) -> List[int]:
        missing: List[int] = []
        new_ptrs = list(self._g_ptrs)
        dst: List[torch.Tensor] = []
        src: List[torch.Tensor] = []
        for i, g in enumerate(grads):
            gv = self.g_views[i]
            if g is gv:
                new_ptrs[i] = self._gv_ptrs[i]
            elif g is None:
                if i in self.hyper_links:
                    # a fused hyper-parameter reaches the loss only through the prior, which is no
                    # longer in autograd: its likelihood gradient is zero.  p.grad becomes the (zeroed)
                    # flat view, so that runner code that touches every p.grad (the clamp loop of
                    # inference.py:219-220) keeps working
                    if self._g_seen[i] != "zero":
                        gv.zero_()
                        self._g_seen[i] = "zero"
                    self.params[i].grad = gv
                    grads[i] = gv
                    new_ptrs[i] = self._gv_ptrs[i]
                elif raise_on_no_grad:
                    raise RuntimeError(f"No gradient for parameter with shape {self.params[i].shape}")
                else:
                    missing.append(i)
            elif (g.dtype is torch.float32 and g.device == self.device and g.layout is torch.strided
                  and g.is_contiguous() and g.numel() == self.numel[i] and (g.data_ptr() & 15) == 0):
                new_ptrs[i] = g.data_ptr()          # read where it lies
                self._g_seen[i] = None
            else:
                # not readable in place (other dtype / device / layout, unaligned): copied into G, once
                # per tensor and version
                seen = self._g_seen[i]
                ver = g._version
                if seen is None or seen == "zero" or seen[0]() is not g or seen[1] != ver:
                    dst.append(gv)
                    src.append(g)
                    self._g_seen[i] = (weakref.ref(g), ver)
                new_ptrs[i] = self._gv_ptrs[i]
        if dst:
            self.copies += len(dst)
            try:
                torch._foreach_copy_(dst, src)      # one multi-tensor launch instead of one copy per tensor
            except (RuntimeError, TypeError):
                for d, g in zip(dst, src):
                    d.copy_(g)
        if new_ptrs != self._g_ptrs:
            self._write_grad_table(new_ptrs)
        return missing

    def _write_grad_table(self, ptrs: List[int]) -> None:
        "device table of gradient addresses <- ptrs (stream-ordered, through kernel parameters: bnnp_poke)"
        arr = np.asarray(ptrs, dtype=np.int64)
        base = self.seg_grad_dev.data_ptr()
        with torch.cuda.device(self.device):
            stream = self._stream()
            for lo in range(0, self.nseg, 480):
                part = np.ascontiguousarray(arr[lo:lo + 480])
                N.check(self.lib.bnnp_poke(base + 8 * lo, part.ctypes.data, part.nbytes, stream), "bnnp_poke")
        self._g_ptrs = list(ptrs)
        if self._scanner is not None:
            self._scanner.set_table(self._g_ptrs)
        self.launches += (self.nseg + 479) // 480
        self.table_writes += 1

    @torch.no_grad()
    def _readopt_parameters(self) -> None:
        for i, (p, pv) in enumerate(zip(self.params, self.p_views)):
            if p.data_ptr() != self._p_ptrs[i]:
                pv.copy_(p.data)
                p.data = pv
                self._lp_valid = self._hyper_valid = False

    def bind_grad_views(self) -> None:
        "p.grad <- its view of the flat G array, for every parameter"
        for p, v in zip(self.params, self.g_views):
            if p.grad is not v:
                p.grad = v
        self._g_seen = [None] * self.nseg
        self._held_grads = None
        if self._scanner is not None:
            self._scanner.drop()

    def drop_grads(self) -> None:
        "zero_grad(set_to_none=True): p.grad = None (a fused hyper-parameter keeps its zero view)"
        keep = sorted(self.hyper_links) if (self.prior_fused and self.hyper_links) else []
        if self._scanner is not None and all(self.params[i].grad is self.g_views[i] and self._g_seen[i] == "zero" for i in keep):
            self._scanner.drop_grads(keep)
        else:
            for p in self.params:
                p.grad = None
            for i in keep:
                if self._g_seen[i] != "zero":
                    self.g_views[i].zero_()
                    self._g_seen[i] = "zero"
                self.params[i].grad = self.g_views[i]
            if self._scanner is not None:
                self._scanner.drop()
        self._held_grads = None
        self._gg_sig = None

    def ensure_momentum_storage(self) -> None:
        if self.M is None:
            self.M = torch.zeros(self.total, dtype=torch.float32, device=self.device)
            self._ptrs["M"] = self.M.data_ptr()
            self._static_dirty = True
            self.m_views = [self.M[o:o + n].view(p.shape) for o, n, p in zip(self.off, self.numel, self.params)]

    def set_momentum(self, i: int, value: torch.Tensor) -> torch.Tensor:
        """state['momentum_buffer'] = tensor: the values are copied into the flat M."""
        self.ensure_momentum_storage()
        v = self.m_views[i]
        if value is not v:
            with torch.no_grad():
                v.copy_(value)
        self._mm_version = None
        return v

    def check_momentum(self) -> None:
        """sgld.py:107-111: stepping without sample_momentum is an error."""
        if self._mom_published:
            return
        if self.M is None or any(s.raw_get("momentum_buffer") is None for s in self.seg_states):
            raise RuntimeError("No 'momentum_buffer' stored in state. "
                               "Perhaps you forgot to call `sample_momentum`?")
        self._mom_published = True

    def publish_momentum(self) -> None:
        for s, v in zip(self.seg_states, self.m_views):
            if s.raw_get("momentum_buffer") is not v:
                s.raw_set("momentum_buffer", v)
        self._mom_published = True

    def ensure_prev_storage(self, with_momentum: bool) -> None:
        if self.prev_p is None:
            self.prev_p = torch.zeros_like(self.P)
            self.prev_g = torch.zeros_like(self.P)
            self._ptrs["prev_p"], self._ptrs["prev_g"] = self.prev_p.data_ptr(), self.prev_g.data_ptr()
            self._static_dirty = True
            for s, o, n, p in zip(self.seg_states, self.off, self.numel, self.params):
                s.raw_set("prev_parameter", self.prev_p[o:o + n].view(p.shape))
                s.raw_set("prev_grad", self.prev_g[o:o + n].view(p.shape))
        if with_momentum and self.prev_m is None:
            self.prev_m = torch.zeros_like(self.P)
            self._ptrs["prev_m"] = self.prev_m.data_ptr()
            self._static_dirty = True
            for s, o, n, p in zip(self.seg_states, self.off, self.numel, self.params):
                s.raw_set("prev_momentum_buffer", self.prev_m[o:o + n].view(p.shape))

    def pack(self, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
        """Per-tensor values -> one flat array in this group's layout (replay noise)."""
        flat = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        assert len(tensors) == self.nseg
        for t, o, n in zip(tensors, self.off, self.numel):
            flat[o:o + n].copy_(torch.as_tensor(t, dtype=torch.float32).reshape(-1))
        return flat

    def unpack(self, flat: torch.Tensor) -> torch.Tensor:
        """Flat layout -> the concatenation of the tensors without padding."""
        return torch.cat([flat[o:o + n] for o, n in zip(self.off, self.numel)])

    # ------------------------------------------------------------ segment table
    def set_preconditioner(self, i: int, v: float) -> None:
        if self.table["precond"][i] != v:
            self.table["precond"][i] = v
            self._table_dirty = True

    def set_prior(self, i: int, kind: int, loc: float, scale: float, df: float) -> None:
        t = self.table
        t["prior_kind"][i], t["prior_loc"][i], t["prior_scale"][i], t["prior_df"][i] = kind, loc, scale, df
        self._table_dirty = True
        self._lp_valid = False

    def set_hyper_link(self, weight: int, hyper: int, kind: int, a: float, b: float) -> None:
        """Segment `hyper` (one element u) holds the scale of segment `weight`
        (include/bnnp.h: BNNP_PRIOR_HYPER_*; prior/hierarchical.py, prior/empirical_bayes.py)."""
        t = self.table
        if self.numel[hyper] != 1:
            raise ValueError("a hyper-parameter segment has exactly one element")
        if int(t["prior_kind"][weight]) not in (N.PRIOR_NORMAL, N.PRIOR_LAPLACE, N.PRIOR_STUDENT_T):
            raise ValueError("only Normal / Laplace / StudentT segments can have a sampled scale")
        t["prior_kind"][hyper], t["prior_loc"][hyper], t["prior_scale"][hyper] = kind, a, b
        t["link"][hyper], t["link"][weight] = weight, hyper
        self.hyper_links[hyper] = weight
        self._table_dirty = True
        self._lp_valid = self._hyper_valid = False

    def clear_hyper_links(self) -> None:
        for h, w in self.hyper_links.items():
            self.table["link"][h] = self.table["link"][w] = -1
        self.hyper_links = {}
        self._table_dirty = True
        self._lp_valid = self._hyper_valid = False

    @property
    def clamp_active(self) -> bool:
        "the fused step clamps g + prior gradient to +-grad_max (inference.py:219-220)"
        return self.grad_max is not None and self.clamp_armed

    @property
    def has_hyper(self) -> bool:
        return bool(self.hyper_links)

    @property
    def hyper_post_ok(self) -> bool:
        """All sampled scales belong to Normal / Laplace segments: a step can refresh the scales,
        hyper gradients and log-priors for the parameters it leaves behind (BNNP_F_HYPER_POST),
        so the next step needs no pre-pass."""
        k = self.table["prior_kind"]
        return bool(self.hyper_links) and all(int(k[w]) in (N.PRIOR_NORMAL, N.PRIOR_LAPLACE)
                                              for w in self.hyper_links.values())

    def step_prior_flags(self, prior_flags: int, chunks) -> int:
        "LOG_PRIOR (and HYPER_POST) for a step launch that writes P, given the flags of _prior_flags"
        if not prior_flags:
            return 0
        if not self.hyper_links:
            return N.F_LOG_PRIOR
        if chunks is None and self.hyper_post_ok:
            return N.F_LOG_PRIOR | N.F_HYPER_POST
        return 0

    def after_hyper_post(self) -> None:
        """After a BNNP_F_HYPER_POST step, scales, hyper gradients and log-priors describe the
        parameters the step left in P -- once its epilogue has been applied.  It stays pending: the next
        launch carries it (BNNP_F_HYPER_CHAIN, see `launch`) or whoever needs the numbers flushes it.
        (Capturable mode finalises at once: the host cannot tell what a replayed graph left pending.)"""
        if self.capturable:
            self.flush_pending()
        self._hyper_valid = self._lp_valid = True
        self._hyper_pversion = self._lp_pversion = self._p_version()

    def hyper_fresh(self) -> bool:
        return self._hyper_valid and not self._table_dirty and self._hyper_pversion == self._p_version()

    def hyper_prepass(self, inv_num_data: float) -> None:
        """The read-only pre-pass of the hierarchical priors (BNNP_F_HYPER) followed by its
        epilogue: afterwards the segment table holds the current scales s(u), the state holds
        every segment's log-prior and, per hyper segment, -(1/N) dlog p/du."""
        self.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_P | N.F_LOG_PRIOR | N.F_HYPER, N.NOISE_NONE,
                    cm=1.0, inv_num_data=inv_num_data)
        self.flush_pending()
        self._hyper_valid = self._lp_valid = True
        self._hyper_pversion = self._lp_pversion = self._p_version()

    def _upload_table(self) -> None:
        # rare (preconditioner / prior changes): a plain blocking copy of a few KB.  The pending
        # epilogue still needs the OLD table (it reads the preconditioner), so it runs first.
        self.flush_pending()
        self.table_dev.copy_(torch.from_numpy(self.table.view(np.uint8).copy()))
        self._table_dirty = False
        self._hyper_valid = False       # the device copy of the linked scales is the host's stale one again

    # ------------------------------------------------------------ scalars
    def poke(self, i: int, col: int, v: float) -> None:
        self.flush_pending()
        self.state_dev[i, col] = v
        self._epoch += 1

    def fetch(self) -> np.ndarray:
        """Host mirror of the segment-state array (one D2H copy + one stream sync,
        only if a launch happened since the last fetch)."""
        if self._host_epoch != self._epoch or self.capturable:     # (graph replays launch behind the host's back)
            self.flush_pending()
            self.state_host.copy_(self.state_dev, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self._host_epoch = self._epoch
        return self.state_np

    def materialize(self) -> None:
        if self._mat_epoch == self._epoch and not self.capturable:
            return
        st = self.fetch()
        nd = self.metrics_num_data
        for i, s in enumerate(self.seg_states):
            d = self.numel[i]
            if self.have_metrics:
                s.raw_set("est_temperature", float(st[i, N.S_EST_MM]) / d)
                s.raw_set("est_config_temp", float(st[i, N.S_EST_PG]) * (nd / d))
            if self.have_delta:
                s.raw_set("delta_energy", float(st[i, N.S_DELTA_ENERGY]))
            if self.have_prev_new:
                s.raw_set("prev_new_momentum_delta", float(st[i, N.S_PREV_NEW_MOM]))
        self._mat_epoch = self._epoch

    def square_avg_tensor(self, i: int) -> torch.Tensor:
        """state['square_avg'] is only ever consumed through its mean
        (sgld.py:154,173); the engine keeps that mean.  A reader gets a constant
        tensor with the right mean."""
        return torch.full_like(self.params[i], float(self.fetch()[i, N.S_SQ_MEAN]))

    # ------------------------------------------------------------ launches
    def _stream(self) -> int:
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    def _refresh_static(self) -> None:
        "the pointer / size fields of the launch block that only change when storage is (re)allocated"
        a, ptr = self.args, self._ptrs
        a.P, a.G, a.M = ptr["P"], ptr["G"], ptr.get("M")
        a.prev_p, a.prev_g = ptr.get("prev_p"), ptr.get("prev_g")
        a.segs, a.chunks = self.table_dev.data_ptr(), self.chunks_dev.data_ptr()
        a.seg_grad = self.seg_grad_dev.data_ptr()
        a.ctl = self.ctl_dev.data_ptr() if self.capturable else None
        a.seg_state, a.partials, a.stamps = self.state_dev.data_ptr(), self.partials.data_ptr(), self.stamps.data_ptr()
        a.nseg = self.nseg
        self._var_ptrs = None
        self._static_dirty = False

    def _call_lib(self, fn, what: str) -> None:
        if torch.cuda.current_device() != self._dev_index:
            with torch.cuda.device(self.device):
                rc = fn(C.byref(self.args), self._stream())
        else:
            rc = fn(C.byref(self.args), self._stream())
        if rc != 0:
            N.check(rc, what)

    def _issue(self, nchunks, op, phase, noise, flags, cm, cg, cn, cp, inv_n, c_gm_base, curv_base, rms_alpha) -> None:
        """bnnp_launch with this chain's deferred-epilogue protocol: the launch carries the
        previous launch's epilogue (applied on the device by the first-chunk CTA of every
        segment) and leaves its own pending.  The part of the argument block that changes from
        launch to launch is written with one struct.pack_into."""
        parity, call = self._parity, self.call
        if self.serpentine and parity:
            # alternate the chunk order from launch to launch: each launch starts on the lines the
            # previous one left in L2 (include/bnnp.h: BNNP_F_REVERSE)
            flags |= N.F_REVERSE
        else:
            flags &= ~N.F_REVERSE
        slot = 0
        if self.capturable:
            slot = phase if op <= N.OP_HMC else N.SLOT_OTHER
            self.poke_coef(slot, (cm, cg, cn, cp, inv_n, c_gm_base, curv_base, rms_alpha))
            pend = _NO_PENDING
        else:
            pend = self._pending or _NO_PENDING
        gmax = self.grad_max if self.grad_max is not None else 0.0
        N.DYN_STRUCT.pack_into(self._argbuf, N.DYN_OFFSET,
                               nchunks, self.nchunks, parity, op, phase, noise, flags, slot, 0,
                               self.key[0], self.key[1], call,
                               cm, cg, cn, cp, inv_n, gmax, c_gm_base, curv_base, rms_alpha, *pend)
        if _NVTX:
            torch.cuda.nvtx.range_push(f"bnnp_launch op={op} phase={phase} flags={flags:#x}")
        self._call_lib(self.lib.bnnp_launch, "bnnp_launch")
        if self.capturable:
            self._call_lib(self.lib.bnnp_advance, "bnnp_advance")
            self._pending = True
            self.launches += 1
        else:
            self._pending = (1, op, phase, flags, parity, 0, call, c_gm_base, curv_base, rms_alpha, inv_n)
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        self._last_issue = (nchunks, op, phase, noise, flags, cm, cg, cn, cp, inv_n, c_gm_base, curv_base, rms_alpha)
        self._parity = parity ^ 1
        self.call = call + 1
        self._epoch += 1
        self.launches += 1

    def flush_pending(self) -> None:
        """Apply the last launch's per-segment bookkeeping to state_dev now (bnnp_finalize).
        Needed before the host reads state_dev, before the segment table changes and before a
        launch that skips segments; a plain sequence of steps never calls it."""
        if self._pending is None and not self.capturable:
            return
        if self._static_dirty:
            self._refresh_static()
        if self.capturable:
            # the pending epilogue is in the control block; after graph replays the host does not even
            # know whether there is one: the finalize kernel looks, bnnp_clear_pending marks it done
            self._call_lib(self.lib.bnnp_finalize, "bnnp_finalize")
            with torch.cuda.device(self.device):
                N.check(self.lib.bnnp_clear_pending(self.ctl_dev.data_ptr(), self._stream()), "bnnp_clear_pending")
            self.launches += 2
            self._pending = None
            return
        N.PENDING_STRUCT.pack_into(self._argbuf, N.BnnpLaunch.pending.offset, *self._pending)
        self._call_lib(self.lib.bnnp_finalize, "bnnp_finalize")
        self._pending = None
        self.launches += 1

    def launch(self, op: int, phase: int, flags: int, noise: int, cm=0.0, cg=0.0, cn=0.0, cp=0.0,
               inv_num_data=0.0, c_gm_base=0.0, curv_base=0.0, rms_alpha=0.0,
               chunks=None) -> None:
        if chunks is not None:
            self.flush_pending()      # every segment's pending epilogue needs a CTA; a partial launch has none for some
        pend = self._pending
        if pend is not None and pend is not True and (pend[3] & N.F_HYPER_POST):
            # the epilogue of a step with sampled scales rewrites the segment table: it may ride on a launch
            # over all chunks that derives what it needs from the pending records itself, or is applied first
            if chunks is None and not (flags & N.F_HYPER) and not self._table_dirty:
                flags |= N.F_HYPER_CHAIN
            else:
                self.flush_pending()
        if self._table_dirty:
            self._upload_table()
        if self._static_dirty:
            self._refresh_static()
        a = self.args
        chunk_ids, nchunks = chunks if chunks is not None else (None, self.nchunks)
        replay = None
        if noise == N.NOISE_REPLAY:
            if self.replay is None:
                raise RuntimeError("replay noise requested but none was provided")
            replay = self.replay.data_ptr()
        var = (self._ptrs.get("prev_m") if (flags & N.F_READ_M) else None, replay,
               chunk_ids.data_ptr() if chunk_ids is not None else None)
        if var != self._var_ptrs:
            a.prev_m, a.replay_noise, a.chunk_ids = var
            self._var_ptrs = var
        self._chunk_ids_keepalive = chunk_ids
        self._issue(nchunks, op, phase, noise, flags, cm, cg, cn, cp, inv_num_data, c_gm_base, curv_base, rms_alpha)

    def launch_coef(self, op: int, phase: int, flags: int, noise: int, coef, chunks=None) -> None:
        "launch() with the coefficients as the 8-tuple of a BnnpCoef (the samplers' `_coefs`)"
        cm, cg, cn, cp, inv_n, c_gm_base, curv_base, rms_alpha = coef
        self.launch(op, phase, flags, noise, cm, cg, cn, cp, inv_n, c_gm_base, curv_base, rms_alpha, chunks)

    def poke_coef(self, slot: int, coef) -> None:
        "capturable mode: coefficient slot `slot` of the device control block <- coef (if it differs)"
        if self._slot_cache[slot] == tuple(coef):
            return
        if slot != N.SLOT_OTHER and torch.cuda.is_current_stream_capturing():
            # a write recorded in a graph would be repeated by every replay and undo later updates
            raise RuntimeError("the sampler's hyper-parameters (lr / temperature / momentum / num_data) changed since "
                               "its last launch: call `sync_hyperparameters()` before capturing a CUDA graph")
        raw = N.COEF_STRUCT.pack(*coef)
        dst = self.ctl_dev.data_ptr() + N.BnnpControl.coef.offset + slot * N.COEF_STRUCT.size
        with torch.cuda.device(self.device):
            N.check(self.lib.bnnp_poke(dst, raw, len(raw), self._stream()), "bnnp_poke")
        self._slot_cache[slot] = tuple(coef)
        self.launches += 1

    def relaunch(self) -> None:
        """Launch again with the arguments of the previous launch (same coefficients, next Philox
        counter): the C-ABI-level hot loop that bench.py times for the device-resident number."""
        self._issue(*self._last_issue)

    def chunks_without(self, missing: Sequence[int]):
        """(device list of chunk indices, count) that skips the segments in `missing`
        (raise_on_no_grad=False, sgld.py:96-101); None if nothing is left."""
        keep = np.ones(self.nseg, dtype=bool)
        keep[list(missing)] = False
        ids = np.ascontiguousarray(np.nonzero(keep[self.chunk_seg_host])[0].astype(np.int32))
        if ids.size == 0:
            return None
        return torch.from_numpy(ids).to(self.device), int(ids.size)

    def take_noise_mode(self, needs_noise: bool) -> int:
        if not needs_noise:
            return N.NOISE_NONE
        return N.NOISE_REPLAY if self.replay is not None else N.NOISE_PHILOX

    # ------------------------------------------------------------ reductions on demand
    def _p_version(self) -> int:
        if self._scanner is not None:
            return self._scanner.params_version()
        return sum(p._version for p in self.params)

    def note_step_sums(self, flags: int, op: int, capture_grads: bool = True) -> None:
        """Called after a step launch: SUM_GG in the segment state now describes the gradient the
        step used (remembered as the p.grad tensors and their versions, if `capture_grads`); SUM_MM
        describes M as stored if the launch reduced every sum; LOG_PRIOR describes P as stored if
        it carried BNNP_F_LOG_PRIOR."""
        held = self._held_grads
        if capture_grads and self._scanner is not None:
            self._gg_sig = "scanner" if self._scanner.capture() else None
        elif capture_grads and held is not None:
            try:
                self._gg_sig = (held, list(map(_VERSION, held)))
            except AttributeError:             # a parameter without gradient was skipped
                self._gg_sig = None
        else:
            self._gg_sig = None
        all_sums = bool(flags & (N.F_CALC_METRICS | N.F_ALL_SUMS)) or op in (N.OP_SAMPLE_MOMENTUM, N.OP_REDUCE)
        self._mm_version = self.M._version if (self.M is not None and all_sums) else None
        if flags & N.F_LOG_PRIOR:
            self._lp_valid = True
            self._lp_pversion = self._p_version()
        elif flags & N.F_WRITE_P:
            self._lp_valid = False
        if flags & N.F_WRITE_P:
            self._hyper_valid = False

    def invalidate_sums(self) -> None:
        self._gg_sig = self._mm_version = None
        if self._scanner is not None:
            self._scanner.drop()
        self._lp_valid = self._hyper_valid = False

    def reduce_now(self, inv_num_data: float) -> None:
        """dot(g,g), dot(m,m) and the log-prior of the CURRENT arrays (no writes).  The caller has
        just run sync_views(raise_on_no_grad=True): every gradient pointer is current."""
        flags = N.F_READ_G
        if self.M is not None:
            flags |= N.F_READ_M
        if self.prior_fused:
            flags |= N.F_READ_P | N.F_PRIOR_GRAD
            if self.has_hyper:
                # scales, hyper gradients and the log-prior come from the pre-pass
                if not self.hyper_fresh():
                    self.hyper_prepass(inv_num_data)
            else:
                flags |= N.F_LOG_PRIOR
            if self.clamp_active:
                flags |= N.F_CLAMP_GRAD
        self.launch(N.OP_REDUCE, N.PHASE_MID, flags, N.NOISE_NONE, cm=1.0, inv_num_data=inv_num_data)
        self.note_step_sums(flags, N.OP_REDUCE)
        self._mm_version = self.M._version if self.M is not None else None

    def reduce_log_prior(self, inv_num_data: float) -> None:
        """sum log p(theta) of the parameters now in P, nothing else: reads P only (4 B/param), so it
        is also legal while the gradients are gone (model.log_prior() right after zero_grad(),
        inference_reject.py:19-20)."""
        self.launch(N.OP_REDUCE, N.PHASE_MID, N.F_READ_P | N.F_LOG_PRIOR, N.NOISE_NONE, cm=1.0,
                    inv_num_data=inv_num_data)
        self._lp_valid = True
        self._lp_pversion = self._p_version()

    def sums_fresh(self, need_mm: bool) -> bool:
        sig = self._gg_sig
        if sig is None:
            return False
        if sig == "scanner":
            if not self._scanner.fresh():
                return False
            return not (need_mm and (self.M is None or self._mm_version != self.M._version))
        cur = list(map(_GRAD, self.params))
        if len(cur) != len(sig[0]) or not all(map(_IS, cur, sig[0])) or list(map(_VERSION, cur)) != sig[1]:
            return False
        if need_mm and (self.M is None or self._mm_version != self.M._version):
            return False
        return True

    def log_prior_fresh(self) -> bool:
        return self._lp_valid and self._lp_pversion == self._p_version()

    @torch.no_grad()
    def rollback(self) -> None:
        """verlet_sgld.py:63-69 on the flat arrays."""
        if self.prev_p is None:
            raise KeyError("prev_parameter")
        pm = self.prev_m if (self.prev_m is not None and self.M is not None) else None
        rc = self.lib.bnnp_rollback(self.P.data_ptr(), self.G.data_ptr(),
                                    self.M.data_ptr() if pm is not None else None,
                                    self.prev_p.data_ptr(), self.prev_g.data_ptr(),
                                    pm.data_ptr() if pm is not None else None,
                                    self.total, self._stream())
        N.check(rc, "bnnp_rollback")
        self.launches += 1
        self.invalidate_sums()
        self.bind_grad_views()       # p.grad is the restored gradient again (verlet_sgld.py:66-67)
