"""Test-set evaluation of posterior samples with the bookkeeping on the device.

Mirror of the reference's `evaluate_model` (bnn_priors/exp_utils.py:250-340): same
arguments, same result keys (`lp_ensemble`, `lp_last`, `acc_ensemble`, `acc_last`,
`ece`, `ace`, `rmsce`).  The runners call it after every epoch with the current
state as a one-sample ensemble (inference.py:199-213), `experiments/eval_bnn.py`
with all stored samples.

What changes: the reference copies each batch's `log_prob` and logits to the CPU
as float64 (one blocking copy per batch, exp_utils.py:295-297) and reduces
[E, N] / [E, N, C] float64 tensors on the host.  Here one launch per batch folds the
batch into per-point accumulators in HBM (`bnnp_eval_batch`), one launch pair
finishes (`bnnp_eval_finish`), and the host reads five doubles once.  A sample whose
tensors ARE the model's current parameters (the per-epoch call) is not copied onto
itself.  The forward pass is the model's own torch code.

No CPU implementation: a model on the CPU raises.
"""
from __future__ import annotations

import ctypes as C
import math
import warnings
from typing import Dict, Iterable, Tuple

import torch

from . import _native as N


def _n_samples_dict(samples) -> int:
    "exp_utils.py:237-243"
    n_samples = min(len(v) for _, v in samples.items())
    if not all((len(v) == n_samples) for _, v in samples.items()):
        warnings.warn("Not all samples have the same length. Setting n_samples to the minimum.")
    return n_samples


def _labels_of(dataloader_test) -> torch.Tensor:
    "exp_utils.py:254-259"
    ds = dataloader_test.dataset
    if hasattr(ds, "tensors"):
        return ds.tensors[1]
    if hasattr(ds, "targets"):
        return torch.as_tensor(ds.targets)
    raise ValueError("I cannot find the labels in the dataloader.")


def _batches(dataloader_test):
    """The (x, y) batches `for batch in dataloader_test` yields.  The reference's test
    loaders are `DataLoader(TensorDataset(x, y), batch_size, shuffle=False)`
    (experiments/train_bnn.py:244,257) whose default collation fetches and stacks the rows one by one --
    128 indexing kernels per batch when the tensors live on the GPU.  For exactly that
    configuration the same batches are produced as slices; anything else (samplers,
    custom collate functions, other datasets) goes through the loader itself."""
    from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, TensorDataset
    from torch.utils.data.dataloader import default_collate
    dl = dataloader_test
    if (isinstance(dl, DataLoader) and isinstance(dl.dataset, TensorDataset) and len(dl.dataset.tensors) == 2
            and isinstance(dl.batch_sampler, BatchSampler) and isinstance(dl.sampler, SequentialSampler)
            and type(dl.batch_sampler) is BatchSampler and dl.collate_fn is default_collate
            and dl.batch_size is not None):
        x, y = dl.dataset.tensors
        n, bs = len(dl.dataset), dl.batch_size
        stop = (n // bs) * bs if dl.drop_last else n
        for i in range(0, stop, bs):
            j = min(i + bs, stop)
            yield x[i:j], y[i:j]
        return
    yield from dl


class _SampleLoader:
    """`model.load_state_dict(sample)` (exp_utils.py:276, strict) without the copies
    that would write a tensor onto itself: in the per-epoch evaluation the one sample
    is `model.state_dict()`, i.e. views of the live parameters."""

    def __init__(self, model: torch.nn.Module):
        self.dst = model.state_dict(keep_vars=True)

    @torch.no_grad()
    def load(self, sample: Dict[str, torch.Tensor]) -> int:
        missing = [k for k in self.dst if k not in sample]
        unexpected = [k for k in sample if k not in self.dst]
        if missing or unexpected:
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing}, unexpected keys {unexpected}")
        copied = 0
        for k, d in self.dst.items():
            s = sample[k]
            if s.shape != d.shape:
                raise RuntimeError(f"size mismatch for {k}: copying a param with shape {tuple(s.shape)} from "
                                   f"checkpoint, the shape in current model is {tuple(d.shape)}.")
            if (s.device == d.device and s.dtype == d.dtype and s.data_ptr() == d.data_ptr()
                    and s.stride() == d.stride()):
                continue
            d.detach().copy_(s)
            copied += 1
        return copied


class EvalAccumulators:
    """Device arrays of one evaluation (include/bnnp_eval.h: BnnpEvalState)."""

    def __init__(self, N_points: int, C_out: int, kind: int, device):
        self.N, self.Cn, self.kind, self.device = int(N_points), int(C_out), int(kind), device
        f64 = dict(dtype=torch.float64, device=device)
        self.ens = torch.empty(self.N, self.Cn, **f64)
        self.lps_lse = torch.empty(self.N, **f64)
        self.lps_last = torch.empty(self.N, **f64)
        self.acc_last = torch.empty(self.N, **f64)
        self.rows = torch.empty(self.N, N.EVAL_ROW, **f64)
        self.out = torch.zeros(N.EV_OUT, **f64)
        self.out_host = torch.zeros(N.EV_OUT, dtype=torch.float64).pin_memory()
        st = N.BnnpEvalState()
        st.ens, st.lps_lse, st.lps_last = self.ens.data_ptr(), self.lps_lse.data_ptr(), self.lps_last.data_ptr()
        st.acc_last, st.rows = self.acc_last.data_ptr(), self.rows.data_ptr()
        st.N, st.C, st.kind = self.N, self.Cn, self.kind
        self.st = st
        self.lib = N.lib()
        self.launches = 0
        self._dev_index = device.index if device.index is not None else torch.cuda.current_device()

    def _stream(self) -> int:
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    def batch(self, acc_data: torch.Tensor, lps, labels, targets, n0: int, sample_index: int) -> None:
        B = acc_data.shape[0]
        if acc_data.dtype != torch.float32 or acc_data.stride(-1) != 1:
            acc_data = acc_data.float().contiguous()
        if acc_data.dim() == 1:
            acc_data = acc_data.unsqueeze(-1)
        if lps is not None and (lps.dtype != torch.float32 or not lps.is_contiguous()):
            lps = lps.float().contiguous()
        stride_t = 0
        if targets is not None:
            if targets.dtype != torch.float32 or targets.stride(-1) != 1:
                targets = targets.float().contiguous()
            if targets.dim() == 1:
                targets = targets.unsqueeze(-1)
            stride_t = targets.stride(0)
        if labels is not None and (labels.dtype != torch.int64 or not labels.is_contiguous()):
            labels = labels.long().contiguous()
        rc = self.lib.bnnp_eval_batch(
            C.byref(self.st), acc_data.data_ptr(), acc_data.stride(0),
            lps.data_ptr() if lps is not None else None,
            labels.data_ptr() if labels is not None else None,
            targets.data_ptr() if targets is not None else None, stride_t,
            int(n0), int(B), int(sample_index), self._stream())
        N.check_eval(rc, "bnnp_eval_batch")
        self.launches += 1
        self._keep = (acc_data, lps, labels, targets)     # alive until the next launch is enqueued (same stream)

    def finish(self, labels, targets, n_samples: int, want_probs: bool):
        probs = torch.empty(self.N, self.Cn, dtype=torch.float64, device=self.device) if want_probs else None
        rc = self.lib.bnnp_eval_finish(
            C.byref(self.st), labels.data_ptr() if labels is not None else None,
            targets.data_ptr() if targets is not None else None, int(n_samples),
            self.out.data_ptr(), probs.data_ptr() if probs is not None else None, self._stream())
        N.check_eval(rc, "bnnp_eval_finish")
        self.launches += 2
        self.out_host.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()      # the one host sync of the evaluation
        return self.out_host.tolist(), probs


_CALIBRATION = None


def set_calibration_metrics(ece, ace, rmsce) -> None:
    """The three functions `evaluate_model(..., calibration_eval=True)` applies to (labels, ensemble
    probabilities): the reference imports them into exp_utils (exp_utils.py:13) from its third-party
    module; this package neither ships nor imports them."""
    global _CALIBRATION
    _CALIBRATION = (ece, ace, rmsce)


def evaluate_model(model, dataloader_test: Iterable[Tuple[torch.Tensor, torch.Tensor]],
                   samples: Dict[str, torch.Tensor],
                   likelihood_eval: bool, accuracy_eval: bool, calibration_eval: bool):
    """Drop-in for exp_utils.evaluate_model (exp_utils.py:250-340)."""
    labels = _labels_of(dataloader_test)
    n_points, *possibly_D = labels.shape
    E = _n_samples_dict(samples)
    device = next(iter(model.parameters())).device
    if device.type != "cuda":
        raise RuntimeError("bnn_priors_b200.evaluate_model runs on CUDA models only; there is no CPU "
                           "implementation of this path")

    acc = None
    kind = None
    loader = _SampleLoader(model)
    labels_dev = labels.to(device)
    for sample_i in range(E):
        sample = dict((k, v[sample_i]) for k, v in samples.items())          # exp_utils.py:245-247
        with torch.no_grad():
            loader.load(sample)
            i = 0
            for batch_x, batch_y in _batches(dataloader_test):
                batch_x = batch_x.to(device)
                batch_y = batch_y.to(device)
                preds = model(batch_x)
                if isinstance(preds, torch.distributions.Categorical):
                    this_kind = N.EVAL_CATEGORICAL
                    acc_data_batch, lps_batch = preds.logits, None            # log_prob(y) == logits[y]
                    lab, tgt = batch_y, None
                elif isinstance(preds, torch.distributions.Normal):
                    this_kind = N.EVAL_NORMAL
                    acc_data_batch = preds.mean
                    lps_batch = preds.log_prob(batch_y).sum(-1)
                    lab, tgt = None, batch_y
                else:
                    raise ValueError(f"unknown likelihood {type(preds)}")
                if calibration_eval and this_kind != N.EVAL_CATEGORICAL:
                    raise ValueError("Cannot calculate calibration metrics "
                                     f"for predictions of type {type(preds)}")
                if acc is None:
                    kind = this_kind
                    width = acc_data_batch.shape[-1] if acc_data_batch.dim() > 1 else 1
                    if kind == N.EVAL_CATEGORICAL and len(possibly_D) == 0:
                        # the reference sizes acc_data by labels.max()+1 (exp_utils.py:263-264) and
                        # fails on a mismatch with the logits; same check, on the device
                        n_classes = int(labels_dev.max().item()) + 1
                        if n_classes != width:
                            raise RuntimeError(f"The expanded size of the tensor ({n_classes}) must match "
                                               f"the existing size ({width}) at non-singleton dimension 1")
                    acc = EvalAccumulators(n_points, width, kind, device)
                elif this_kind != kind:
                    raise ValueError("the likelihood type changed between batches")
                acc.batch(acc_data_batch, lps_batch, lab, tgt, i, sample_i)
                i += len(batch_x)
            if i != n_points:
                raise RuntimeError(f"the dataloader yielded {i} points, the label tensor has {n_points}")

    if acc is None:
        raise ValueError("no samples or an empty test set")
    if kind == N.EVAL_CATEGORICAL:
        out, probs = acc.finish(labels_dev.long().contiguous(), None, E, calibration_eval)
        # the identity the reference asserts (exp_utils.py:312-314)
        assert math.isclose(out[N.EV_LP_ENSEMBLE_CHECK], out[N.EV_LP_ENSEMBLE], rel_tol=1e-5, abs_tol=1e-8)
    else:
        tgt = labels_dev.float().reshape(n_points, -1).contiguous()
        out, probs = acc.finish(None, tgt, E, False)

    results = {}
    if likelihood_eval:
        results["lp_ensemble"] = out[N.EV_LP_ENSEMBLE]
        results["lp_last"] = out[N.EV_LP_LAST]
    if accuracy_eval:
        results["acc_ensemble"] = out[N.EV_ACC_ENSEMBLE]
        results["acc_last"] = out[N.EV_ACC_LAST]
    if calibration_eval:
        # the binning metrics are the reference's third-party numpy code (third_party/
        # calibration_error.py, a TensorFlow-Probability derivative), out of this path's scope: the
        # caller hands them over (overlay.install(evaluate=True) does, from the reference's exp_utils);
        # they run on the ensemble probabilities this evaluation produced on the device
        if _CALIBRATION is None:
            raise RuntimeError("calibration_eval=True: hand the calibration metrics over first with "
                               "evaluate.set_calibration_metrics(ece, ace, rmsce) "
                               "(overlay.install(evaluate=True) does it with the reference's own)")
        ece, ace, rmsce = _CALIBRATION
        probs_mean = probs.cpu().numpy()
        labels_np = labels.cpu().numpy()
        results["ece"] = float(ece(labels_np, probs_mean).mean())
        results["ace"] = float(ace(labels_np, probs_mean).mean())
        results["rmsce"] = float(rmsce(labels_np, probs_mean).mean())
    evaluate_model.last_launches = acc.launches
    return results
