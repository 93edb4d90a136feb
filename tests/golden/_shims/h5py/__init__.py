"""Test-only stand-in for the slice of h5py the reference uses (exp_utils.py:409-551:
HDF5ModelSaver, HDF5Metrics, load_samples).  h5py is not installed in this image.

It keeps the reference's OBSERVABLE file semantics -- named datasets of shape
[n, *shape] that grow along axis 0 (`resize`), slicing reads / writes, `items()`,
`flush()`, re-opening the path read-only from another `File` object while the writer is
still open -- on top of one pickle per file.  It does not write HDF5: what a test
proves with it is that the reference's saver / loader code round-trips through this
repo's sample sink, not the HDF5 byte format.  Never on the product path.
"""
import os
import pickle

import numpy as np

__version__ = "0.0-shim"


class Dataset:
    def __init__(self, file, name, dtype, shape, chunks=None, maxshape=None, fletcher32=False, fillvalue=None):
        self._file = file
        self.name = "/" + name
        self.dtype = np.dtype(dtype)
        self.chunks = chunks
        self.maxshape = maxshape if maxshape is not None else tuple(shape)
        self.fletcher32 = bool(fletcher32)
        self.fillvalue = fillvalue
        self._data = self._filled(tuple(shape))

    def _fill(self):
        fv = self.fillvalue
        if fv is None:
            return 0
        if np.issubdtype(self.dtype, np.integer) and isinstance(fv, float) and np.isnan(fv):
            return np.iinfo(self.dtype).min      # what HDF5 / numpy 1.18 stored for NaN -> int64 (exp_utils.py:467)
        return fv

    def _filled(self, shape):
        return np.full(shape, self._fill(), dtype=self.dtype)

    @property
    def shape(self):
        return self._data.shape

    def __len__(self):
        return self._data.shape[0]

    def resize(self, size, axis=None):
        if axis is None:
            new_shape = tuple(size)
        else:
            new_shape = list(self._data.shape)
            new_shape[axis] = int(size)
            new_shape = tuple(new_shape)
        for n, m in zip(new_shape, self.maxshape):
            if m is not None and n > m:
                raise ValueError("resize beyond maxshape")
        new = self._filled(new_shape)
        common = tuple(slice(0, min(a, b)) for a, b in zip(self._data.shape, new_shape))
        new[common] = self._data[common]
        self._data = new
        self._file._dirty = True

    def __getitem__(self, idx):
        return self._data[idx]

    def __setitem__(self, idx, value):
        if self._file.mode == "r":
            raise OSError("file is open read-only")
        value = np.asarray(value)
        if np.issubdtype(self.dtype, np.integer) and value.dtype.kind == "f":
            value = np.where(np.isnan(value), float(np.iinfo(self.dtype).min), value)
        self._data[idx] = value.astype(self.dtype, copy=False)
        self._file._dirty = True

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)


class File:
    def __init__(self, name, mode="r", libver=None, rdcc_nbytes=None, swmr=False, **kwargs):
        self.filename = str(name)
        self.mode = mode
        self.swmr_mode = bool(swmr)
        self._dsets = {}
        self._dirty = False
        self._open = True
        if mode == "r":
            if not os.path.exists(self.filename):
                raise OSError(f"Unable to open file (file does not exist: {self.filename})")
            try:
                with open(self.filename, "rb") as f:
                    magic = f.read(len(_MAGIC))
                    if magic != _MAGIC:
                        raise OSError("not a file of the h5py test shim")
                    raw = pickle.load(f)
            except (pickle.UnpicklingError, EOFError) as e:
                raise OSError(str(e))
            for k, d in raw.items():
                ds = Dataset(self, k, d["data"].dtype, d["data"].shape, d["chunks"], d["maxshape"],
                             d["fletcher32"], d["fillvalue"])
                ds._data = d["data"]
                self._dsets[k] = ds
        elif mode in ("w", "x", "w-"):
            if mode in ("x", "w-") and os.path.exists(self.filename):
                raise OSError("file exists")
            self._dirty = True
            self.flush()
        elif mode in ("a", "r+"):
            if os.path.exists(self.filename):
                other = File(self.filename, "r")
                self._dsets = other._dsets
                for d in self._dsets.values():
                    d._file = self
        else:
            raise ValueError(f"invalid mode {mode!r}")

    # -- mapping interface
    def create_dataset(self, name, shape=None, dtype=None, data=None, chunks=None, maxshape=None,
                       fletcher32=False, fillvalue=None, **kwargs):
        if self.mode == "r":
            raise OSError("file is open read-only")
        if self.swmr_mode:
            raise ValueError("cannot create a dataset in SWMR mode")
        if name in self._dsets:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if data is not None:
            data = np.asarray(data, dtype=dtype)
            shape, dtype = data.shape, data.dtype
        ds = Dataset(self, name, dtype, shape, chunks, maxshape, fletcher32, fillvalue)
        if data is not None:
            ds._data = data.copy()
        self._dsets[name] = ds
        self._dirty = True
        return ds

    def __getitem__(self, name):
        return self._dsets[name.lstrip("/")]

    def __contains__(self, name):
        return name.lstrip("/") in self._dsets

    def __iter__(self):
        return iter(self._dsets)

    def __len__(self):
        return len(self._dsets)

    def keys(self):
        return self._dsets.keys()

    def values(self):
        return self._dsets.values()

    def items(self):
        return self._dsets.items()

    # -- persistence
    def flush(self):
        if self.mode == "r" or not self._dirty:
            return
        raw = {k: dict(data=d._data, chunks=d.chunks, maxshape=d.maxshape, fletcher32=d.fletcher32,
                       fillvalue=d.fillvalue) for k, d in self._dsets.items()}
        tmp = self.filename + ".tmp"
        with open(tmp, "wb") as f:
            f.write(_MAGIC)
            pickle.dump(raw, f, protocol=pickle.HIGHEST_PROTOCOL)
        os.replace(tmp, self.filename)
        self._dirty = False

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


_MAGIC = b"BNNP-H5SHIM\n"
