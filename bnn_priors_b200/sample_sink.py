"""Posterior-sample sink on the flat layout.

The reference stores a sample by walking `model.state_dict()` and calling
`.cpu().detach().unsqueeze(0).numpy()` on every entry -- one blocking device-to-host
copy per tensor, 175 of them for `googleresnet` -- and then appending to an HDF5 file
(exp_utils.py:426-431, called from inference.py:189-197 at every sampling epoch).

`FlatSampleSaver` offers the same `model_saver` interface (`add_state_dict`, `flush`,
`load_samples`, context manager; exp_utils.py:409-487) on top of the sampler's flat
parameter array: a sample is ONE device-to-device snapshot of P (plus one packed copy
of the few non-parameter buffers) on the compute stream and ONE asynchronous
device-to-host copy into pinned memory on a side stream, so the chain never waits for
the host.  The result is `{name: [n_samples, *shape], "steps": int64[n], "timestamps":
float64[n]}`.  Where h5py is installed it is written as the reference's HDF5 file --
one dataset per key, shape [n, *shape], chunks (1, *shape), maxshape (None, *shape),
fletcher32, NaN fill value, `libver="latest"` (exp_utils.py:418-421,467-477) -- so
`load_samples` and `experiments/eval_bnn.py` read it unchanged; otherwise (this image
has no h5py) with `torch.save`, which `load_samples` reads through its `torch.load`
fallback (exp_utils.py:539-551).  The HDF5 library itself is not re-implemented.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch


def write_samples_hdf5(path: str, samples: Dict[str, torch.Tensor], h5py_module) -> None:
    """The file HDF5ModelSaver leaves behind (exp_utils.py:409-477), written in one go."""
    import numpy as np
    with h5py_module.File(path, "w", libver="latest") as f:
        for k, v in samples.items():
            a = v.detach().cpu().numpy()
            if a.dtype not in (np.float32, np.float64, np.int64):
                raise TypeError(f"{k}: float32, float64 and int64 only (exp_utils.py:467-469), got {a.dtype}")
            shape = tuple(a.shape[1:])
            d = f.create_dataset(k, dtype=a.dtype, shape=(0,) + shape, chunks=(1,) + shape,
                                 maxshape=(None,) + shape, fletcher32=True, fillvalue=np.nan)
            d.resize(a.shape[0], axis=0)
            d[0:a.shape[0]] = a
        f.flush()


def _h5py_or_none():
    try:
        import h5py
        return h5py if hasattr(h5py, "File") else None
    except ImportError:
        return None


class FlatSampleSaver:
    def __init__(self, path: Optional[str], sampler, capacity: int):
        """`path`: file written by `flush(final=True)` / `__exit__` (None: keep in RAM only);
        `sampler`: a bnn_priors_b200 sampler that owns the model's parameters;
        `capacity`: number of samples the run will store (`n_samples`, train_bnn.py:236)."""
        self.path, self.sampler, self.capacity = path, sampler, int(capacity)
        self.groups = sampler.flat_groups
        self.device = self.groups[0].device
        self._ptr = {}                       # data_ptr of a parameter view -> (group index, segment index)
        for gi, fg in enumerate(self.groups):
            for i, v in enumerate(fg.p_views):
                self._ptr[v.data_ptr()] = (gi, i)
        self._layout = None                  # decided at the first add_state_dict
        self.count = 0
        self.steps = torch.zeros(self.capacity, dtype=torch.int64)
        self.timestamps = torch.zeros(self.capacity, dtype=torch.float64)
        self._side = torch.cuda.Stream(device=self.device)
        self._events: List[torch.cuda.Event] = []

    # -- context manager like HDF5ModelSaver (exp_utils.py:418-424)
    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.flush(final=True)

    # ------------------------------------------------------------------
    def _plan(self, state_dict: Dict[str, torch.Tensor]) -> None:
        params, fbufs, ibufs = [], [], []
        for k, v in state_dict.items():
            where = self._ptr.get(v.data_ptr()) if v.is_cuda and v.numel() > 0 else None
            if where is not None and v.dtype == torch.float32:
                params.append((k, where, tuple(v.shape)))
            elif v.dtype == torch.int64:
                ibufs.append((k, tuple(v.shape), v.numel()))
            elif v.dtype in (torch.float32, torch.float64):
                fbufs.append((k, tuple(v.shape), v.numel(), v.dtype))
            else:
                raise TypeError(f"{k}: the sample files hold float32, float64 and int64 only "
                                f"(exp_utils.py:467-469), got {v.dtype}")
        self._layout = (params, fbufs, ibufs)
        widths = [fg.total for fg in self.groups]
        nf, ni = sum(n for _, _, n, _ in fbufs), sum(n for _, _, n in ibufs)
        dev = self.device
        # device staging (two slots: the copy of sample i may still be in flight when i+1 arrives)
        self._dev_p = [[torch.empty(w, dtype=torch.float32, device=dev) for w in widths] for _ in range(2)]
        self._dev_f = [torch.empty(max(nf, 1), dtype=torch.float64, device=dev) for _ in range(2)]
        self._dev_i = [torch.empty(max(ni, 1), dtype=torch.int64, device=dev) for _ in range(2)]
        self._slot_free = [torch.cuda.Event(), torch.cuda.Event()]
        # host rings (pinned: the copies are asynchronous)
        self._host_p = [torch.empty(self.capacity, w, dtype=torch.float32).pin_memory() for w in widths]
        self._host_f = torch.empty(self.capacity, max(nf, 1), dtype=torch.float64).pin_memory()
        self._host_i = torch.empty(self.capacity, max(ni, 1), dtype=torch.int64).pin_memory()

    @torch.no_grad()
    def add_state_dict(self, state_dict: Dict[str, torch.Tensor], step: int) -> None:
        """exp_utils.py:426-431.  Parameter entries are taken from the flat array they
        alias; the other entries (BatchNorm statistics, prior hyper-parameter buffers)
        are packed into one float64 and one int64 staging vector."""
        if self._layout is None:
            self._plan(state_dict)
        if self.count >= self.capacity:
            raise IndexError("FlatSampleSaver is full")
        params, fbufs, ibufs = self._layout
        n, slot = self.count, self.count % 2
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._slot_free[slot])           # the slot's previous D2H must be done
        for dst, fg in zip(self._dev_p[slot], self.groups):
            dst.copy_(fg.P, non_blocking=True)           # the snapshot: one D2D copy per param group
        if fbufs:
            torch.cat([state_dict[k].reshape(-1).to(torch.float64) for k, _, _, _ in fbufs], out=self._dev_f[slot])
        if ibufs:
            torch.cat([state_dict[k].reshape(-1) for k, _, _ in ibufs], out=self._dev_i[slot])
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            for hp, dp in zip(self._host_p, self._dev_p[slot]):
                hp[n].copy_(dp, non_blocking=True)
            if fbufs:
                self._host_f[n].copy_(self._dev_f[slot], non_blocking=True)
            if ibufs:
                self._host_i[n].copy_(self._dev_i[slot], non_blocking=True)
            self._slot_free[slot].record(self._side)
            done = torch.cuda.Event()
            done.record(self._side)
        self._events.append(done)
        self.steps[n] = int(step)
        self.timestamps[n] = time.time()
        self.count += 1

    def flush(self, final: bool = False) -> None:
        """exp_utils.py:486-487 flushes the HDF5 file after every sample; here the sample
        is safe once its copy has landed in host memory, and the file is written once."""
        if final:
            for e in self._events:
                e.synchronize()
            self._events.clear()
            if self.path is not None and self.count:
                h5 = _h5py_or_none()
                if h5 is not None and not str(self.path).endswith((".pt", ".pth")):
                    write_samples_hdf5(self.path, self.load_samples(keep_steps=True), h5)
                else:
                    torch.save(self.load_samples(keep_steps=True), self.path)

    def load_samples(self, idx=slice(None), keep_steps: bool = True) -> Dict[str, torch.Tensor]:
        """Same result layout as exp_utils.load_samples (exp_utils.py:539-551)."""
        for e in self._events:
            e.synchronize()
        self._events.clear()
        n = self.count
        out: Dict[str, torch.Tensor] = {}
        if self._layout is None:
            return out
        params, fbufs, ibufs = self._layout
        for k, (gi, i), shape in params:
            fg = self.groups[gi]
            o, m = fg.off[i], fg.numel[i]
            out[k] = self._host_p[gi][:n, o:o + m].reshape((n,) + shape)[idx].clone()
        o = 0
        for k, shape, m, dtype in fbufs:
            out[k] = self._host_f[:n, o:o + m].reshape((n,) + shape).to(dtype)[idx].clone()
            o += m
        o = 0
        for k, shape, m in ibufs:
            out[k] = self._host_i[:n, o:o + m].reshape((n,) + shape)[idx].clone()
            o += m
        if keep_steps:
            out["steps"] = self.steps[:n][idx].clone()
            out["timestamps"] = self.timestamps[:n][idx].clone()
        return out
