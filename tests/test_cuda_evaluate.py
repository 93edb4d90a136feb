"""Test-set evaluation on the device (bnn_priors_b200/evaluate.py ->
csrc/bnnp_eval.cu) against (a) the results recorded from the unmodified reference
`evaluate_model` (tests/golden/eval.npz), (b) the numpy oracle on seeded inputs incl.
ragged batches, one class, -inf log-probs, (c) size-independent properties at the
BASELINE test-set size (10,000 points x 10 classes)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("lp_ensemble", "lp_last", "acc_ensemble", "acc_last")
# float64 bookkeeping on float32 model outputs; the reference takes the accuracy mean in
# float32 (models/base.py:185) -> 1e-7; everything else agrees to 1e-12
TOL = 1e-7


def golden():
    return np.load(os.path.join(HERE, "golden", "eval.npz"))


def feed(acc_data, lps, labels, kind, batch, want_probs=False):
    "drive the C ABI like evaluate_model does, with precomputed model outputs"
    from bnn_priors_b200 import _native as N
    from bnn_priors_b200.evaluate import EvalAccumulators
    E, n, c = acc_data.shape
    acc = EvalAccumulators(n, c, kind, torch.device(DEV))
    a = torch.tensor(acc_data, dtype=torch.float32, device=DEV)
    l = torch.tensor(lps, dtype=torch.float32, device=DEV)
    if kind == N.EVAL_CATEGORICAL:
        lab = torch.tensor(labels, dtype=torch.int64, device=DEV)
        tgt = None
    else:
        lab = None
        tgt = torch.tensor(labels, dtype=torch.float32, device=DEV).reshape(n, -1)
    for e in range(E):
        for i in range(0, n, batch):
            j = min(n, i + batch)
            acc.batch(a[e, i:j], None if kind == N.EVAL_CATEGORICAL else l[e, i:j],
                      lab[i:j] if lab is not None else None, tgt[i:j] if tgt is not None else None, i, e)
    out, probs = acc.finish(lab, tgt, E, want_probs)
    return out, probs


@pytest.mark.parametrize("tag,kind", [("cls", 0), ("reg", 1)])
def test_c_abi_reproduces_reference_results(tag, kind):
    from bnn_priors_b200 import _native as N
    z = golden()
    want = json.loads(str(z[f"{tag}.results"]))
    out, _ = feed(z[f"{tag}.acc_data"], z[f"{tag}.lps"], z[f"{tag}.y"], kind, batch=64)
    got = dict(zip(KEYS, out[:4]))
    for k in KEYS:
        assert math.isclose(got[k], want[k], rel_tol=TOL, abs_tol=TOL), (k, got[k], want[k])
    if kind == 0:
        assert math.isclose(out[N.EV_LP_ENSEMBLE_CHECK], out[N.EV_LP_ENSEMBLE], rel_tol=1e-6)


def make_model(tag, z):
    import local_models as LM
    if tag == "cls":
        m = LM.TinyClassifier(20, 7, 16)
    else:
        m = LM.TinyRegressor(6, 3, 8, noise_std=0.7)
    return m.to(DEV)


@pytest.mark.parametrize("tag", ["cls", "reg"])
def test_evaluate_model_matches_the_reference_on_the_same_samples(tag):
    """The public call: model + DataLoader + stacked state_dict samples, the reference's
    signature (exp_utils.py:250-253).  The local model restates the reference network's
    forward; its state_dict keys are the reference's."""
    from bnn_priors_b200.evaluate import evaluate_model
    z = golden()
    model = make_model(tag, z)
    samples = {k[len(tag) + 8:]: torch.tensor(z[k], device=DEV) for k in z.files if k.startswith(f"{tag}.sample.")}
    assert set(samples) == set(model.state_dict())
    x = torch.tensor(z[f"{tag}.x"], device=DEV)
    y = torch.tensor(z[f"{tag}.y"], device=DEV)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=64, shuffle=False)
    got = evaluate_model(model, loader, samples, likelihood_eval=True, accuracy_eval=True, calibration_eval=False)
    want = json.loads(str(z[f"{tag}.results"]))
    assert set(got) == set(want)
    for k in KEYS:
        # the network forward runs in float32 on the GPU (TF32 off): 2e-6 on the log-likelihoods
        assert math.isclose(got[k], want[k], rel_tol=2e-6, abs_tol=2e-6), (k, got[k], want[k])
    # 4 samples x 4 batches + the finishing pair
    assert evaluate_model.last_launches == 4 * 4 + 2
    # the per-epoch call of the runners: one sample = the live state_dict (no self-copies)
    model.load_state_dict({k: v[-1] for k, v in samples.items()})
    live = {k: v.unsqueeze(0) for k, v in model.state_dict().items()}
    got1 = evaluate_model(model, loader, live, likelihood_eval=True, accuracy_eval=True, calibration_eval=False)
    want1 = json.loads(str(z[f"{tag}.results_last_only"]))
    for k in KEYS:
        assert math.isclose(got1[k], want1[k], rel_tol=2e-6, abs_tol=2e-6), (k, got1[k], want1[k])
    only = evaluate_model(model, loader, live, likelihood_eval=False, accuracy_eval=True, calibration_eval=False)
    assert set(only) == {"acc_ensemble", "acc_last"}


def test_sample_loader_skips_self_copies_and_is_strict():
    from bnn_priors_b200.evaluate import _SampleLoader
    model = make_model("cls", None)
    sd = model.state_dict()
    ld = _SampleLoader(model)
    assert ld.load(sd) == 0                                   # views of the live tensors: nothing to copy
    other = {k: v.clone() + 1 for k, v in sd.items()}
    assert ld.load(other) == len(sd)
    assert all(torch.equal(model.state_dict()[k], other[k]) for k in other)
    bad = dict(other)
    bad.pop(next(iter(bad)))
    with pytest.raises(RuntimeError, match="missing keys"):
        ld.load(bad)
    with pytest.raises(RuntimeError, match="unexpected keys"):
        ld.load(dict(other, extra=torch.zeros(1, device=DEV)))


@pytest.mark.parametrize("n,c,e,batch", [(1, 1, 1, 4), (37, 3, 2, 5), (130, 100, 3, 64), (64, 10, 5, 64), (257, 33, 2, 100)])
def test_categorical_bookkeeping_matches_the_oracle(n, c, e, batch):
    from oracle import eval_oracle as EO
    rng = np.random.default_rng(n * 1000 + c)
    logits = rng.standard_normal((e, n, c)).astype(np.float32) * 3
    logp = (logits - np.log(np.exp(logits.astype(np.float64)).sum(-1, keepdims=True))).astype(np.float32)
    if c > 2:
        logp[0, 0, 1] = -np.inf                               # a class with zero probability in one sample
    labels = rng.integers(0, c, n)
    lps = np.take_along_axis(logp, labels[None, :, None].repeat(e, 0), 2)[..., 0]
    out, probs = feed(logp, lps, labels, 0, batch, want_probs=True)
    want = EO.evaluate(logp, lps, labels, EO.CATEGORICAL)
    for k, v in zip(KEYS, out[:4]):
        tol = TOL if k.startswith("acc") else 1e-12      # the oracle follows the reference's float32 accuracy mean
        assert math.isclose(v, want[k], rel_tol=tol, abs_tol=tol), (k, v, want[k])
    assert math.isclose(out[4], want["lp_ensemble_check"], rel_tol=1e-12, abs_tol=1e-12)
    assert np.allclose(probs.cpu().numpy(), want["probs_mean"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("n,d,e,batch", [(1, 1, 1, 1), (50, 3, 4, 16), (200, 40, 2, 128)])
def test_normal_bookkeeping_matches_the_oracle(n, d, e, batch):
    from oracle import eval_oracle as EO
    rng = np.random.default_rng(n + d)
    mean = rng.standard_normal((e, n, d)).astype(np.float32)
    y = rng.standard_normal((n, d)).astype(np.float32)
    lps = (-0.5 * ((mean - y) / 0.7) ** 2 - math.log(0.7) - 0.5 * math.log(2 * math.pi)).sum(-1).astype(np.float32)
    out, _ = feed(mean, lps, y, 1, batch)
    want = EO.evaluate(mean, lps, y, EO.NORMAL)
    for k, v in zip(KEYS, out[:4]):
        assert math.isclose(v, want[k], rel_tol=1e-12, abs_tol=1e-12), (k, v, want[k])


def test_full_size_properties():
    """BASELINE test-set size (10,000 x 10) with 8 samples: (i) identical samples -> the
    ensemble equals every member, (ii) results do not depend on the batch size, (iii) the
    two ways of computing the ensemble log-likelihood agree (the reference's assert,
    exp_utils.py:312-314), (iv) agreement with a float64 torch evaluation on the device."""
    n, c, e = 10000, 10, 8
    g = torch.Generator(device=DEV).manual_seed(3)
    logp = torch.log_softmax(torch.randn(e, n, c, device=DEV, generator=g) * 2, -1)
    labels = torch.randint(0, c, (n,), device=DEV, generator=g)
    lp_np, lab_np = logp.cpu().numpy(), labels.cpu().numpy()
    lps = np.take_along_axis(lp_np, lab_np[None, :, None].repeat(e, 0), 2)[..., 0]
    a, _ = feed(lp_np, lps, lab_np, 0, 128)
    b, _ = feed(lp_np, lps, lab_np, 0, 1000)
    assert a == b
    assert math.isclose(a[4], a[0], rel_tol=1e-6)      # equal up to the float32 normalisation of each sample's log-probs
    d = logp.double()
    ens = torch.logsumexp(d, 0) - math.log(e)
    ens = ens - torch.logsumexp(ens, 1, keepdim=True)
    idx = torch.arange(n, device=DEV)
    assert math.isclose(a[4], float(ens[idx, labels].mean()), rel_tol=1e-12)
    assert math.isclose(a[0], float((torch.logsumexp(d[:, idx, labels], 0) - math.log(e)).mean()), rel_tol=1e-12)
    assert math.isclose(a[1], float(d[-1][idx, labels].mean()), rel_tol=1e-12)
    assert math.isclose(a[2], float((ens.argmax(1) == labels).double().mean()), rel_tol=1e-12)
    assert math.isclose(a[3], float((d[-1].argmax(1) == labels).double().mean()), rel_tol=1e-12)
    same = np.repeat(lp_np[:1], 3, 0)
    s, _ = feed(same, np.repeat(lps[:1], 3, 0), lab_np, 0, 512)
    assert math.isclose(s[0], s[1], rel_tol=1e-12) and s[2] == s[3]


def test_tensor_dataset_fast_path_yields_the_loaders_batches():
    from bnn_priors_b200.evaluate import _batches
    x, y = torch.arange(50.).reshape(25, 2), torch.arange(25)
    ds = torch.utils.data.TensorDataset(x, y)
    for kw in (dict(batch_size=8), dict(batch_size=8, drop_last=True), dict(batch_size=25), dict(batch_size=40),
               dict(batch_size=8, shuffle=True), dict(batch_size=4, collate_fn=lambda b: (torch.stack([r[0] for r in b]),
                                                                                       torch.stack([r[1] for r in b])))):
        torch.manual_seed(0)
        want = list(torch.utils.data.DataLoader(ds, **kw))
        torch.manual_seed(0)
        got = list(_batches(torch.utils.data.DataLoader(ds, **kw)))
        assert len(got) == len(want)
        for (a, b), (c, d) in zip(got, want):
            assert torch.equal(a, c) and torch.equal(b, d)
    assert [b for b in _batches([(1, 2), (3, 4)])] == [(1, 2), (3, 4)]     # any iterable of batches


def test_errors():
    from bnn_priors_b200.evaluate import evaluate_model
    import local_models as LM
    cpu_model = LM.TinyClassifier(4, 3, 5)
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.zeros(6, 4), torch.arange(6) % 3), batch_size=4)
    with pytest.raises(RuntimeError, match="CUDA models only"):
        evaluate_model(cpu_model, loader, {k: v.unsqueeze(0) for k, v in cpu_model.state_dict().items()}, True, True, False)
    model = LM.TinyRegressor(4, 2, 5).to(DEV)
    rl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.zeros(6, 4), torch.zeros(6, 2)), batch_size=4)
    with pytest.raises(ValueError, match="calibration"):
        evaluate_model(model, rl, {k: v.unsqueeze(0) for k, v in model.state_dict().items()}, True, True, True)

    class NoLabels:
        dataset = object()
    with pytest.raises(ValueError, match="cannot find the labels"):
        evaluate_model(model, NoLabels(), {}, True, True, False)


def test_calibration_metrics_through_the_overlay_match_the_reference():
    """calibration_eval=True: the binning metrics are the reference's third-party code, handed over by
    overlay.install(evaluate=True); same numbers as the reference's evaluate_model on its own model."""
    import refenv
    if not refenv.available():
        pytest.skip("no oracle/_ref snapshot")
    eu = refenv.exp_utils()
    from bnn_priors_b200 import evaluate as EV, overlay
    data = refenv.synthetic_dataset("mnist", torch.device(DEV), 64, 300, seed=3)
    model = refenv.build_model("densenet_gaussian", data, seed=3)
    loader = torch.utils.data.DataLoader(data.norm.test, batch_size=128)
    sd = model.state_dict()
    g = torch.Generator(device=DEV).manual_seed(0)
    samples = {k: torch.stack([v + 0.05 * i * torch.randn(v.shape, device=v.device, generator=g) if v.dtype.is_floating_point else v
                               for i in range(3)]) for k, v in sd.items()}
    ref_eval = eu.evaluate_model
    want = ref_eval(model, loader, samples, likelihood_eval=True, accuracy_eval=True, calibration_eval=True)
    EV._CALIBRATION = None
    with pytest.raises(RuntimeError, match="set_calibration_metrics"):
        EV.evaluate_model(model, loader, samples, True, True, True)
    overlay.install(evaluate=True)
    try:
        assert eu.evaluate_model is EV.evaluate_model
        got = eu.evaluate_model(model, loader, samples, likelihood_eval=True, accuracy_eval=True, calibration_eval=True)
    finally:
        overlay.uninstall()
    assert set(got) == set(want) == {"lp_ensemble", "lp_last", "acc_ensemble", "acc_last", "ece", "ace", "rmsce"}
    for k in want:
        assert got[k] == pytest.approx(want[k], rel=2e-6, abs=1e-7), k
