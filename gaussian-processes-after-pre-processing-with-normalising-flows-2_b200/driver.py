"""Training-script pieces of the reference (marscf_main.py:216-247, 296-366) on top of the sharded trainer: evaluation
(mean bits/dim over a loader, batch-sharded), best-test-NLL checkpointing with the reference's NaN guard, sample
post-processing (NaN -> -0.5, clamp to [-0.5, 0.5]) and the epoch loop.  One process per GPU; with world size 1 it is
the reference's single-GPU script.  Data loaders are anything iterable that yields `(images, ...)` tuples or image
tensors in [-0.5, 0.5] (`utils.py:25` normalisation); synthetic loaders are used by the tests and the bench.
"""
import math
import os

import torch
import torch.distributed as dist

from . import sharding


def _images(item):
    return item[0] if isinstance(item, (tuple, list)) else item


def test_model(model, test_loader, device=None, reference_mode=False):
    """Mean bits/dim over the loader (marscf_main.py:233-246); every rank evaluates its shard of each batch and the
    (sum, count) pair is all-reduced once at the end.

    The reference never calls `.eval()`: its test NLL is computed with the conditioner dropout ACTIVE (train mode under
    no_grad).  `reference_mode=True` reproduces that; the default evaluates in eval mode (deterministic, and it lets
    the tcgen05 inference conditioner run) - a documented deviation."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    total = None
    was_training = model.training
    model.train(bool(reference_mode))
    with torch.no_grad():
        for item in test_loader:
            x = _images(item)
            if device is not None:
                x = x.to(device, non_blocking=True)
            x = sharding.shard_batch(x, rank, world)
            if x.shape[0] == 0:
                continue
            _, nll, _ = model(x, reverse=False)
            part = torch.stack([nll.double().sum(), torch.tensor(float(nll.numel()), device=nll.device, dtype=torch.float64)])
            total = part if total is None else total + part
    model.train(was_training)
    if total is None:
        return float("nan")
    if world > 1:
        dist.all_reduce(total)
    return float(total[0] / total[1])


class BestCheckpoint:
    """Keep the state dict of the best test NLL so far (marscf_main.py:357-364): NaN evaluations never replace it; only
    rank 0 writes."""

    def __init__(self, path, best=9999999.0):
        self.path = path
        self.best = best

    def update(self, model, test_nll):
        if math.isnan(test_nll) or not test_nll < self.best:
            return False
        self.best = test_nll
        if not dist.is_initialized() or dist.get_rank() == 0:
            os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
            module = model.module if hasattr(model, "module") else model
            torch.save(module.state_dict(), self.path)
        return True


def sample_images(model, samples=None, eps_std=1.0):
    """`save_samples` without the image writer (marscf_main.py:216-224): reverse pass from the prior, NaNs replaced by
    -0.5, clamped to the data range; returns the first `samples` images [n, C, H, W] in [-0.5, 0.5]."""
    with torch.no_grad():
        rev = model(None, None, reverse=True, eps_std=eps_std)
    rev = torch.where(torch.isnan(rev), torch.full_like(rev, -0.5), rev)
    rev = torch.clamp(rev, -0.5, 0.5)
    return rev if samples is None else rev[:samples]


def fit(model, train_loader, test_loader, epochs, checkpoint_path=None, lr=1e-4, warm_up=10000, device=None,
        test_epoch_interval=1, use_graph=None, log=None):
    """The reference's epoch loop (marscf_main.py:334-366): Adamax + LambdaLR(min(1, samples / warm_up)) inside
    `ShardedTrainer`, evaluation every `test_epoch_interval` epochs, best-NLL checkpoint.  Returns the history."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    trainer = None
    best = BestCheckpoint(checkpoint_path) if checkpoint_path else None
    history = []
    for epoch in range(epochs):
        model.train()
        last = float("nan")
        for item in train_loader:
            x = _images(item)
            if device is not None:
                x = x.to(device, non_blocking=True)
            if trainer is None:
                # Replica sync before the first step: run the ActNorm data-dependent init (first TRAINING forward,
                # common_modules.py:141-151) on this rank's shard, then copy rank 0's parameters and buffers to every
                # rank - the initial weights (InvConv draws from each process's numpy RNG) and the init statistics
                # (the reference's DataParallel keeps replica 0's, marscf_main.py:326).
                with torch.no_grad():
                    model(sharding.shard_batch(x, rank, world), reverse=False)
                sharding.broadcast_module(model)
                trainer = sharding.ShardedTrainer(model, lr=lr, warm_up=warm_up, global_batch=x.shape[0],
                                                  use_graph=use_graph)
            last = float(trainer.step(sharding.shard_batch(x, rank, world)))
        entry = {"epoch": epoch, "train_nll": last}
        if epoch % test_epoch_interval == 0 and test_loader is not None:
            entry["test_nll"] = test_model(model, test_loader, device)
            if best is not None:
                entry["saved"] = best.update(model, entry["test_nll"])
                entry["best_test_nll"] = best.best
        history.append(entry)
        if log is not None and rank == 0:
            log(entry)
    return history
