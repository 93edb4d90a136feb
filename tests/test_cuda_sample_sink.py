"""FlatSampleSaver against what the reference's HDF5ModelSaver.add_state_dict
would have stored: `{k: v.cpu().detach()}` of the state_dict at the moment of the call
(exp_utils.py:426-431), read back through the `torch.load` path of the reference's
`load_samples` (exp_utils.py:548-551)."""
import pytest
import torch

import local_models as LM

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_sink_stores_what_the_reference_saver_would(tmp_path):
    from bnn_priors_b200 import mcmc
    from bnn_priors_b200.sample_sink import FlatSampleSaver
    torch.manual_seed(0)
    model = LM.TinyClassifier(20, 4, 16, extra_bn=True).to(DEV)       # BatchNorm: float buffers + int64 counters
    opt = mcmc.VerletSGLD(list(model.parameters()), lr=1e-2, num_data=96.0, momentum=0.9, temperature=1.0)
    x = torch.rand(96, 20, device=DEV)
    y = torch.randint(0, 4, (96,), device=DEV)

    def move(n):
        for _ in range(n):
            opt.zero_grad()
            (-model.log_likelihood_avg(x, y) - model.log_prior() / 96.0).backward()
            opt.step(calc_metrics=False)

    opt.sample_momentum()
    path = tmp_path / "samples.pth"       # .pth: the torch.save format even where an h5py is importable
    want = []
    # constructed like the reference's saver, BEFORE it knows any sampler (train_bnn.py:201-203); two
    # staging slots for four samples: the sink drains to the file as it goes
    with FlatSampleSaver(str(path), "w", slots=2) as saver:
        for s in range(4):
            move(3)
            sd = model.state_dict()
            want.append({k: v.cpu().detach().clone() for k, v in sd.items()})
            saver.add_state_dict(sd, step=100 + 7 * s)
            saver.flush()
            move(1)                       # the chain moves on at once; the stored sample must not
            if s == 2:
                # a killed run keeps what was flushed: the first samples are already in the file
                from bnn_priors_b200.sample_sink import recover_rows
                torch.cuda.synchronize()
                saver.flush()
                assert recover_rows(str(path))["steps"].tolist() == [100, 107, 114]
        assert saver.sampler is opt           # found through the parameters it owns
        in_ram = saver.load_samples(keep_steps=False)
    on_disk = torch.load(str(path))       # exp_utils.load_samples' fallback for non-HDF5 files
    keys = list(want[0].keys())
    assert set(on_disk) == set(keys) | {"steps", "timestamps"} and set(in_ram) == set(keys)
    assert any(k.endswith("num_batches_tracked") for k in keys) and any(k.endswith("running_var") for k in keys)
    for k in keys:
        stacked = torch.stack([w[k] for w in want])
        for got in (on_disk[k], in_ram[k]):
            assert got.dtype == stacked.dtype and got.shape == stacked.shape, k
            assert torch.equal(got, stacked), k
    assert on_disk["steps"].tolist() == [100, 107, 114, 121] and on_disk["steps"].dtype == torch.int64
    ts = on_disk["timestamps"]
    assert ts.dtype == torch.float64 and bool((ts[1:] >= ts[:-1]).all())
    # the samples differ from each other (the chain moved) and BatchNorm counted its batches
    w0 = [k for k in keys if k.endswith("weight_prior.p")][0]
    assert not torch.equal(on_disk[w0][0], on_disk[w0][1])
    nb = [k for k in keys if k.endswith("num_batches_tracked")][0]
    assert on_disk[nb].tolist() == [3, 7, 11, 15]


def test_sink_grows_and_ram_only():
    from bnn_priors_b200 import mcmc
    from bnn_priors_b200.sample_sink import FlatSampleSaver
    lin = torch.nn.Linear(8, 4).to(DEV)
    opt = mcmc.SGLD(list(lin.parameters()), lr=1e-2, num_data=1.0)
    saver = FlatSampleSaver(None, opt, capacity=1)       # round-1 signature still accepted; capacity is ignored
    assert saver.load_samples() == {}
    want = []
    for i in range(7):                                   # more samples than staging slots
        with torch.no_grad():
            lin.weight.add_(1.0)
        want.append(lin.weight.detach().cpu().clone())
        saver.add_state_dict(lin.state_dict(), i)
    saver.flush(final=True)
    out = saver.load_samples()
    assert torch.equal(out["weight"], torch.stack(want)) and out["steps"].tolist() == list(range(7))
    saver.close()
    with pytest.raises(RuntimeError):
        saver.add_state_dict(lin.state_dict(), 8)


def test_sink_writes_the_reference_hdf5_layout_incrementally_and_the_reference_reads_it(tmp_path):
    """With h5py importable (here: the test stand-in with h5py's API, tests/golden/_shims/h5py) the sink
    grows the reference's HDF5 file row by row and the REFERENCE's own load_samples
    (exp_utils.py:539-551, from oracle/_ref) reads it back."""
    import refenv
    if not refenv.available():
        pytest.skip("no oracle/_ref snapshot")
    eu = refenv.exp_utils()
    from bnn_priors_b200 import mcmc
    from bnn_priors_b200.sample_sink import FlatSampleSaver
    torch.manual_seed(1)
    model = LM.TinyClassifier(20, 4, 16, extra_bn=True).to(DEV)
    opt = mcmc.SGLD(list(model.parameters()), lr=1e-2, num_data=96.0, momentum=0.9)
    path = str(tmp_path / "samples.h5")
    want = []
    with FlatSampleSaver(path, "w") as saver:
        for s in range(3):
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(0.5)
            sd = model.state_dict()
            want.append({k: v.cpu().detach().clone() for k, v in sd.items()})
            saver.add_state_dict(sd, step=10 * s)
            torch.cuda.synchronize()
            saver.flush()
            partial = eu.load_samples(path)              # readable while the run goes on
            assert partial["steps"].tolist() == [10 * i for i in range(s + 1)]
    got = eu.load_samples(path, keep_steps=False)
    assert list(got.keys()) == list(want[0].keys())
    for k in got:
        assert torch.equal(got[k], torch.stack([w[k] for w in want])), k
    import h5py
    with h5py.File(path, "r") as f:
        d = f[list(want[0].keys())[0]]
        assert d.chunks == (1,) + tuple(d.shape[1:]) and d.maxshape[0] is None and d.fletcher32
