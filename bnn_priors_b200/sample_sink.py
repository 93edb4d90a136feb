"""Posterior-sample sink on the flat layout.

The reference stores a sample by walking `model.state_dict()` and calling
`.cpu().detach().unsqueeze(0).numpy()` on every entry -- one blocking device-to-host
copy per tensor, 175 of them for `googleresnet` -- and then appending to an HDF5 file
and flushing it (exp_utils.py:426-431,486-487, called from inference.py:189-197 at
every sampling epoch).

`FlatSampleSaver` offers the same `model_saver` interface (`add_state_dict`, `flush`,
`load_samples`, context manager; exp_utils.py:409-487) on top of the sampler's flat
parameter array: a sample is ONE device-to-device snapshot of P (plus one packed copy
of the few non-parameter buffers) on the compute stream and ONE asynchronous
device-to-host copy into a pinned staging slot on a side stream, so the chain never
waits for the host.  `flush()` -- which the runner calls after every sample, like the
reference's -- appends every sample whose copy has landed to the file and flushes it,
so a killed run loses at most the samples still in flight and host memory holds a few
staging slots, not the whole run.

It can be constructed like the reference's saver, `FlatSampleSaver(path, "w")`, before
any sampler exists (experiments/train_bnn.py:201-203 builds the saver first; the runner
creates the optimizer inside `run()`): the sampler that owns the parameters is found at
the first `add_state_dict`.  `overlay.install(sample_sink=True)` binds it in place of
`exp_utils.HDF5ModelSaver`.

File format.  Where h5py is importable the file is the reference's HDF5 file -- one
dataset per key, shape [n, *shape], chunks (1, *shape), maxshape (None, *shape),
fletcher32, NaN fill value, `libver="latest"` (exp_utils.py:418-421,467-477) -- grown
one row per sample, so `load_samples` and `experiments/eval_bnn.py` read it unchanged.
Without h5py (this image has none; the tests use a stand-in with h5py's API) rows are
appended to `<path>.rows` (raw, crash-safe; `recover_rows` reads it back) and `close()`
consolidates them into a `torch.save` file at `path`, which `exp_utils.load_samples`
reads through its `torch.load` fallback (exp_utils.py:539-551).  The HDF5 library itself
is not re-implemented.
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict, List, Optional

import numpy as np
import torch


def _h5py_or_none():
    try:
        import h5py
        return h5py if hasattr(h5py, "File") else None
    except ImportError:
        return None


def _create_dset(f, name, shape, dtype):
    "exp_utils.py:467-477"
    if dtype not in (np.float32, np.float64, np.int64):
        raise TypeError(f"{name}: float32, float64 and int64 only (exp_utils.py:467-469), got {dtype}")
    return f.create_dataset(name, dtype=dtype, shape=(0,) + tuple(shape), chunks=(1,) + tuple(shape),
                            maxshape=(None,) + tuple(shape), fletcher32=True, fillvalue=np.nan)


def write_samples_hdf5(path: str, samples: Dict[str, torch.Tensor], h5py_module) -> None:
    """The file HDF5ModelSaver leaves behind (exp_utils.py:409-477), written in one go."""
    with h5py_module.File(path, "w", libver="latest") as f:
        for k, v in samples.items():
            a = v.detach().cpu().numpy()
            d = _create_dset(f, k, a.shape[1:], a.dtype)
            d.resize(a.shape[0], axis=0)
            d[0:a.shape[0]] = a
        f.flush()


class _H5Sink:
    "one row per sample into the reference's HDF5 layout, flushed after every append"

    def __init__(self, path, h5):
        self.f = h5.File(path, "w", libver="latest", rdcc_nbytes=0)
        self.n = 0
        self._init = True

    def append(self, row: Dict[str, np.ndarray]) -> None:
        if self._init:
            for k, v in row.items():
                _create_dset(self.f, k, v.shape, v.dtype)
            try:
                self.f.swmr_mode = True          # readable while the run goes on (exp_utils.py:447-451)
            except Exception:
                pass
            self._init = False
        for k, v in row.items():
            d = self.f[k]
            d.resize(self.n + 1, axis=0)
            d[self.n:self.n + 1] = v[None]
        self.n += 1

    def flush(self):
        self.f.flush()

    def close(self):
        self.f.flush()
        self.f.close()


class _RowsSink:
    """No h5py: append raw rows to `<path>.rows` (+ a JSON header describing them); `close()` writes
    the consolidated `torch.save` file at `path`."""

    def __init__(self, path):
        self.path = str(path)
        self.rows_path = self.path + ".rows"
        self.fh = open(self.rows_path, "wb")
        self.header = None
        self.n = 0

    def append(self, row: Dict[str, np.ndarray]) -> None:
        if self.header is None:
            self.header = [(k, list(v.shape), str(v.dtype)) for k, v in row.items()]
            with open(self.rows_path + ".json", "w") as f:
                json.dump(self.header, f)
        for k, _, _ in self.header:
            self.fh.write(np.ascontiguousarray(row[k]).tobytes())
        self.n += 1

    def flush(self):
        self.fh.flush()
        os.fsync(self.fh.fileno())

    def close(self):
        self.fh.close()
        if self.n:
            torch.save(recover_rows(self.path), self.path)
        for p in (self.rows_path, self.rows_path + ".json"):
            if os.path.exists(p):
                os.remove(p)


def recover_rows(path: str) -> Dict[str, torch.Tensor]:
    """The samples of an interrupted run without h5py: `<path>.rows` -> {key: [n, *shape]}
    (complete rows only)."""
    with open(str(path) + ".rows.json") as f:
        header = json.load(f)
    sizes = [int(np.prod(shape, dtype=np.int64)) * np.dtype(dt).itemsize for _, shape, dt in header]
    row_bytes = sum(sizes)
    raw = np.fromfile(str(path) + ".rows", dtype=np.uint8)
    n = raw.size // row_bytes if row_bytes else 0
    raw = raw[:n * row_bytes].reshape(n, row_bytes)
    out, o = {}, 0
    for (k, shape, dt), sz in zip(header, sizes):
        out[k] = torch.from_numpy(np.ascontiguousarray(raw[:, o:o + sz]).view(np.dtype(dt)).reshape([n] + shape).copy())
        o += sz
    return out


class _MemSink:
    def __init__(self):
        self.rows: List[Dict[str, np.ndarray]] = []

    def append(self, row):
        self.rows.append({k: v.copy() for k, v in row.items()})

    def flush(self):
        pass

    def close(self):
        pass

    def samples(self):
        if not self.rows:
            return {}
        return {k: torch.from_numpy(np.stack([r[k] for r in self.rows])) for k in self.rows[0]}


class FlatSampleSaver:
    def __init__(self, path: Optional[str] = None, mode_or_sampler="w", capacity: Optional[int] = None, *,
                 sampler=None, slots: int = 4):
        """`path`: the sample file (None: keep the samples in host RAM only).
        Second argument: the reference's `mode` string ("w"; exp_utils.py:410) -- or, as in round 1,
        the sampler.  `sampler`: the bnn_priors_b200 sampler that owns the model's parameters; when
        omitted it is found at the first `add_state_dict` among the live samplers.  `capacity` is
        accepted for compatibility and ignored (the sink grows).  `slots`: staging slots, i.e. how
        many samples may be in flight to the host at once."""
        if not isinstance(mode_or_sampler, str):
            sampler = mode_or_sampler
        elif mode_or_sampler not in ("w", "w-", "x"):
            raise ValueError("FlatSampleSaver writes a new file: mode must be 'w'")
        self.path = None if path is None else str(path)
        self.sampler = sampler
        self.slots = max(2, int(slots))
        self.groups = None
        self._layout = None                  # decided at the first add_state_dict
        self.count = 0                       # samples accepted
        self.written = 0                     # samples appended to the sink
        self._inflight: List[tuple] = []     # (slot, event, step, timestamp) in order
        self._sink = None
        self._closed = False

    # -- context manager like HDF5ModelSaver (exp_utils.py:418-424)
    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.close()

    # ------------------------------------------------------------------
    def _bind(self, state_dict: Dict[str, torch.Tensor]) -> None:
        ptrs = {v.data_ptr() for v in state_dict.values() if isinstance(v, torch.Tensor) and v.is_cuda and v.numel()}
        if self.sampler is None:
            from .mcmc.sgld import live_samplers
            best, hits = None, 0
            for s in live_samplers():
                h = sum(1 for fg in s.flat_groups for v in fg.p_views if v.data_ptr() in ptrs)
                if h > hits:
                    best, hits = s, h
            if best is None:
                raise RuntimeError("FlatSampleSaver: no bnn_priors_b200 sampler owns the parameters of this "
                                   "state_dict (construct the sampler first, or pass sampler=...)")
            self.sampler = best
        self.groups = self.sampler.flat_groups
        self.device = self.groups[0].device
        self._ptr = {}                       # data_ptr of a parameter view -> (group index, segment index)
        for gi, fg in enumerate(self.groups):
            for i, v in enumerate(fg.p_views):
                self._ptr[v.data_ptr()] = (gi, i)
        self._side = torch.cuda.Stream(device=self.device)

    def _plan(self, state_dict: Dict[str, torch.Tensor]) -> None:
        self._bind(state_dict)
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
        self._key_order = list(state_dict.keys())
        widths = [fg.total for fg in self.groups]
        nf, ni = sum(n for _, _, n, _ in fbufs), sum(n for _, _, n in ibufs)
        dev = self.device
        S = self.slots
        # device staging slots and their pinned host twins (the copies are asynchronous)
        self._dev_p = [[torch.empty(w, dtype=torch.float32, device=dev) for w in widths] for _ in range(S)]
        self._dev_f = [torch.empty(max(nf, 1), dtype=torch.float64, device=dev) for _ in range(S)]
        self._dev_i = [torch.empty(max(ni, 1), dtype=torch.int64, device=dev) for _ in range(S)]
        self._host_p = [[torch.empty(w, dtype=torch.float32).pin_memory() for w in widths] for _ in range(S)]
        self._host_f = [torch.empty(max(nf, 1), dtype=torch.float64).pin_memory() for _ in range(S)]
        self._host_i = [torch.empty(max(ni, 1), dtype=torch.int64).pin_memory() for _ in range(S)]
        if self.path is None:
            self._sink = _MemSink()
        else:
            h5 = _h5py_or_none()
            if h5 is not None and not self.path.endswith(".pth"):
                self._sink = _H5Sink(self.path, h5)
            else:
                self._sink = _RowsSink(self.path)

    @torch.no_grad()
    def add_state_dict(self, state_dict: Dict[str, torch.Tensor], step: int) -> None:
        """exp_utils.py:426-431.  Parameter entries are taken from the flat array they
        alias; the other entries (BatchNorm statistics, prior hyper-parameter buffers)
        are packed into one float64 and one int64 staging vector."""
        if self._closed:
            raise RuntimeError("FlatSampleSaver is closed")
        if self._layout is None:
            self._plan(state_dict)
        while len(self._inflight) >= self.slots:
            self._drain(block=True, limit=1)     # every slot is in flight: wait for the oldest
        params, fbufs, ibufs = self._layout
        busy = {s for s, _, _, _ in self._inflight}
        slot = next(s for s in range(self.slots) if s not in busy)
        main = torch.cuda.current_stream(self.device)
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
            for hp, dp in zip(self._host_p[slot], self._dev_p[slot]):
                hp.copy_(dp, non_blocking=True)
            if fbufs:
                self._host_f[slot].copy_(self._dev_f[slot], non_blocking=True)
            if ibufs:
                self._host_i[slot].copy_(self._dev_i[slot], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._side)
        # the device slot may be overwritten once its D2H is done: the next user of the slot waits for it
        self._inflight.append((slot, done, int(step), time.time()))
        self.count += 1

    def _row(self, slot: int, step: int, stamp: float) -> Dict[str, np.ndarray]:
        params, fbufs, ibufs = self._layout
        vals = {}
        for k, (gi, i), shape in params:
            fg = self.groups[gi]
            o, m = fg.off[i], fg.numel[i]
            vals[k] = self._host_p[slot][gi][o:o + m].numpy().reshape(shape)
        o = 0
        for k, shape, m, dtype in fbufs:
            vals[k] = self._host_f[slot][o:o + m].to(dtype).numpy().reshape(shape)
            o += m
        o = 0
        for k, shape, m in ibufs:
            vals[k] = self._host_i[slot][o:o + m].numpy().reshape(shape)
            o += m
        row = {k: vals[k] for k in self._key_order}       # the state_dict's own order, like the reference's file
        row["steps"] = np.asarray(step, dtype=np.int64)
        row["timestamps"] = np.asarray(stamp, dtype=np.float64)
        return row

    def _drain(self, block: bool, limit: Optional[int] = None) -> int:
        done = 0
        while self._inflight and (limit is None or done < limit):
            slot, ev, step, stamp = self._inflight[0]
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            self._sink.append(self._row(slot, step, stamp))
            self._inflight.pop(0)
            self.written += 1
            done += 1
        return done

    def flush(self, final: bool = False) -> None:
        """exp_utils.py:486-487: the runner calls this after every sample (inference.py:196-197).
        Every sample whose device-to-host copy has landed is appended to the file, and the file is
        flushed; `final=True` waits for the copies still in flight first."""
        if self._layout is None or self._closed:
            return
        if self._drain(block=final) or final:
            self._sink.flush()

    def close(self) -> None:
        "wait for the copies in flight, append them, finish the file (what `__exit__` does)"
        if self._closed:
            return
        if self._layout is not None:
            self._drain(block=True)
            self._sink.flush()
            self._sink.close()
        self._closed = True

    def load_samples(self, idx=slice(None), keep_steps: bool = True) -> Dict[str, torch.Tensor]:
        """Same result layout as exp_utils.load_samples (exp_utils.py:539-551); like
        HDF5ModelSaver.load_samples (:479-484) it also works after the saver was closed."""
        if self._layout is None:
            return {}
        if not self._closed:
            self._drain(block=True)
            self._sink.flush()
        if isinstance(self._sink, _MemSink):
            out = self._sink.samples()
        elif isinstance(self._sink, _H5Sink):
            with _h5py_or_none().File(self.path, "r", swmr=True) as f:
                out = {k: torch.from_numpy(np.asarray(v[:])) for k, v in f.items()}
        elif self._closed:
            out = torch.load(self.path)
        else:
            out = recover_rows(self.path)
        return {k: v[idx] for k, v in out.items() if keep_steps or k not in ("steps", "timestamps")}
