#!/usr/bin/env python
"""BASELINE config 4 / 5 end to end, on synthetic data: a ResNet-20 with BatchNorm of the size of
`googleresnet` on synthetic CIFAR-10, VerletSGLD (GGMC) or HMC with the Metropolis-Hastings test
per sampling epoch, Student-t prior on the weights fused into the sampler kernel, one chain per
GPU, ONE all-gather of the cycle's samples at cycle end.

The loop follows the reference's VerletSGLDRunnerReject.run (bnn_priors/inference_reject.py:35-176)
-- exact-gradient pass, initial_step, minibatch steps under a cosine lr schedule, exact-gradient
pass, final_step, delta_energy, maybe_reject, evaluate, store the sample -- restated here because the
reference's runner classes cannot travel to the GPU box (with the reference installed, run ITS
runner through bnn_priors_b200.overlay.install(evaluate=True) instead).

    python examples/chains_resnet_reject.py [--sampler VerletSGLD|HMC] [--cycles 2] [--epochs 3] [--n-train 2048]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/chains_resnet_reject.py
"""
import argparse
import json
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bnn_priors_b200 import _native as N, chains as CH, mcmc  # noqa: E402
from bnn_priors_b200.evaluate import evaluate_model  # noqa: E402
from models import ResNet20  # noqa: E402


class Classifier(torch.nn.Module):
    "p(y | x, params): what the reference's ClassificationModel.forward returns (models/base.py:37-40,179-180)"

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return torch.distributions.Categorical(logits=self.net(x))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sampler", default="VerletSGLD", choices=["VerletSGLD", "HMC"])
    ap.add_argument("--cycles", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=3, help="epochs per cycle; the last two store a sample")
    ap.add_argument("--n-train", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--lr", type=float, default=3e-5)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    rank, world, device = CH.init_chains()
    seed = CH.chain_seed(args.seed, rank)
    torch.manual_seed(seed)                                            # also feeds maybe_reject's uniform
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(args.n_train, 3, 32, 32, device=device, generator=g)
    y = torch.randint(0, 10, (args.n_train,), device=device, generator=g)
    x_test = torch.randn(1024, 3, 32, 32, device=device, generator=g)
    y_test = torch.randint(0, 10, (1024,), device=device, generator=g)
    y_test[:10] = torch.arange(10, device=device)
    test_loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x_test, y_test), batch_size=256)

    model = Classifier(ResNet20().to(device))
    params = list(model.parameters())
    n_data = float(args.n_train)
    if args.sampler == "HMC":
        opt = mcmc.HMC(params, lr=args.lr, num_data=n_data, seed=seed, chain=rank)
    else:
        opt = mcmc.VerletSGLD(params, lr=args.lr, num_data=n_data, momentum=0.98, temperature=1.0, seed=seed, chain=rank)
    (fg,) = opt.flat_groups
    for i, (name, p) in enumerate(model.named_parameters()):           # Student-t(df=3) on weights, N(0,1) on the
        if p.dim() > 1:                                                # last layer's bias, none on BatchNorm
            fg.set_prior(i, N.PRIOR_STUDENT_T, 0.0, math.sqrt(2.0 / p[0].numel()), 3.0)
        elif name.endswith("fc.bias"):
            fg.set_prior(i, N.PRIOR_NORMAL, 0.0, 1.0, 3.0)
    fg.prior_fused, fg.grad_max = True, 1e6

    def log_prior():
        if not fg.log_prior_fresh():
            fg.sync_views(False)
            fg.reduce_now(1.0 / n_data)
        fg.flush_pending()
        return float(fg.state_dev[:, N.S_LOG_PRIOR].sum())

    def exact_potential_and_grad():
        "inference_reject.py:18-33: the whole training set, gradients accumulated over the batches"
        opt.zero_grad()
        loss = 0.0
        for i in range(0, args.n_train, args.batch):
            xb, yb = x[i:i + args.batch], y[i:i + args.batch]
            l = -model(xb).log_prob(yb).sum() / n_data
            l.backward()
            loss += float(l)
        return loss - log_prior() / n_data

    def minibatch_potential_and_grad(xb, yb):
        opt.zero_grad()
        l = -model(xb).log_prob(yb).mean()
        l.backward()
        return float(l)

    steps_per_epoch = args.n_train // args.batch
    total_steps = args.cycles * args.epochs * (steps_per_epoch + 1)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 0.5 * (1 + math.cos(math.pi * (s % (total_steps // args.cycles)) / (total_steps // args.cycles))) + 1e-3)
    samples_per_cycle = min(2, args.epochs)
    ring = CH.SampleRing(samples_per_cycle, fg.total, device)
    log = dict(rank=rank, sampler=args.sampler, params=fg.n_params, tensors=fg.nseg, decisions=[], test=[], gathered=[])

    t0 = time.perf_counter()
    prev_u = exact_potential_and_grad()
    opt.sample_momentum()
    opt.initial_step(calc_metrics=True, save_state=True)
    step = 0
    for cycle in range(args.cycles):
        ring.reset()
        for epoch in range(args.epochs):
            perm = torch.randperm(args.n_train, device=device, generator=g)
            for i in range(steps_per_epoch):
                idx = perm[i * args.batch:(i + 1) * args.batch]
                minibatch_potential_and_grad(x[idx], y[idx])
                step += 1
                opt.step(calc_metrics=(step % 10 == 0))
                sched.step()
            if epoch >= args.epochs - samples_per_cycle:                # a sampling epoch
                step += 1
                u = exact_potential_and_grad()
                opt.final_step(calc_metrics=True)
                de = opt.delta_energy(prev_u, u)
                rejected, log_acc = opt.maybe_reject(de)
                prev_u = prev_u if rejected else u
                log["decisions"].append(dict(step=step, delta_energy=round(de, 4), rejected=bool(rejected)))
                model.eval()
                res = evaluate_model(model, test_loader, {k: v.unsqueeze(0) for k, v in model.state_dict().items()},
                                     likelihood_eval=True, accuracy_eval=True, calibration_eval=False)
                model.train()
                log["test"].append({k: round(v, 4) for k, v in res.items() if k.endswith("last")})
                ring.push(fg.P, step=step, rejected=rejected)
                sched.step()
                if args.sampler == "HMC":
                    opt.sample_momentum()
                opt.initial_step(calc_metrics=False, save_state=True)   # same (restored or accepted) gradient
            if (epoch + 1) % 2 == 0:
                opt.update_preconditioner()
        samples, meta = ring.gather()                                   # the one collective of the cycle
        log["gathered"].append(list(samples.shape))
        assert torch.equal(samples[rank], ring.rows)
    torch.cuda.synchronize(device)
    log["seconds"] = round(time.perf_counter() - t0, 2)
    log["steps"] = step
    assert all(math.isfinite(d["delta_energy"]) for d in log["decisions"])
    print(json.dumps(log), flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
