"""The HDF5 layout FlatSampleSaver writes when h5py is there (this image has none): a
recording stand-in for the h5py API checks that every dataset is created exactly like
HDF5ModelSaver._create_dset does (exp_utils.py:467-477) and receives every row.  CPU."""
import numpy as np
import torch

from bnn_priors_b200.sample_sink import write_samples_hdf5


class FakeDataset:
    def __init__(self, kw):
        self.kw, self.data = kw, np.zeros(kw["shape"], dtype=kw["dtype"])

    def resize(self, n, axis=0):
        assert axis == 0 and self.kw["maxshape"][0] is None
        self.data = np.resize(self.data, (n,) + tuple(self.kw["shape"][1:]))

    def __setitem__(self, sl, v):
        self.data[sl] = v


class FakeFile(dict):
    opened = []

    def __init__(self, path, mode, **kw):
        super().__init__()
        self.path, self.mode, self.kw, self.flushed = path, mode, kw, False
        FakeFile.opened.append(self)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def create_dataset(self, name, **kw):
        self[name] = FakeDataset(kw)
        return self[name]

    def flush(self):
        self.flushed = True


class FakeH5py:
    File = FakeFile


def test_hdf5_layout_matches_the_reference_saver():
    samples = {"net.0.weight_prior.p": torch.randn(5, 4, 3), "net.1.num_batches_tracked": torch.arange(5),
               "net.1.running_mean": torch.randn(5, 4).double(), "steps": torch.arange(5) * 10,
               "timestamps": torch.rand(5, dtype=torch.float64)}
    write_samples_hdf5("x.h5", samples, FakeH5py)
    f = FakeFile.opened[-1]
    assert f.mode == "w" and f.kw == {"libver": "latest"} and f.flushed
    assert set(f) == set(samples)
    for k, v in samples.items():
        kw = f[k].kw
        shape = tuple(v.shape[1:])
        assert kw["shape"] == (0,) + shape and kw["chunks"] == (1,) + shape and kw["maxshape"] == (None,) + shape
        assert kw["fletcher32"] is True and np.isnan(kw["fillvalue"]) and kw["dtype"] == v.numpy().dtype
        assert np.array_equal(f[k].data, v.numpy())
    import pytest
    with pytest.raises(TypeError, match="float32, float64 and int64"):
        write_samples_hdf5("y.h5", {"a": torch.zeros(2, 2, dtype=torch.int32)}, FakeH5py)
