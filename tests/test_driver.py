"""flowk.driver: the reference training script's evaluation / checkpoint / sampling rules (marscf_main.py:216-247,
334-366).  Host logic on a toy model (CPU); the real model on the GPU."""
import math
import os

import pytest
import torch

import flowk  # noqa: F401  (registers the package alias)
from flowk import driver


class Toy(torch.nn.Module):
    """Same call contract as MarScfFlow: forward(x) -> (z, nll[B], None); reverse -> images."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([0.5, -0.25, 0.125]))

    def forward(self, x=None, z=None, eps_std=None, reverse=False):
        if reverse:
            out = torch.tensor([[float("nan"), 0.7], [-0.9, 0.1]]).view(2, 1, 1, 2)
            return out
        return x, ((x - self.w) ** 2).sum(1), None


def test_test_model_is_the_sample_weighted_mean():
    torch.manual_seed(0)
    m = Toy()
    batches = [torch.randn(4, 3), torch.randn(2, 3), torch.randn(5, 3)]
    got = driver.test_model(m, [(b, None) for b in batches])
    ref = torch.cat([m(b)[1] for b in batches]).mean().item()
    assert abs(got - ref) < 1e-6
    assert math.isnan(driver.test_model(m, []))


def test_best_checkpoint_ignores_nan_and_worse(tmp_path):
    m = Toy()
    path = os.path.join(tmp_path, "ckpt", "best.pt")
    best = driver.BestCheckpoint(path)
    assert best.update(m, 3.5) and os.path.exists(path)
    with torch.no_grad():
        m.w.add_(1.0)
    assert not best.update(m, float("nan"))
    assert not best.update(m, 3.6)
    assert torch.equal(torch.load(path)["w"], torch.tensor([0.5, -0.25, 0.125]))
    assert best.update(m, 3.4) and best.best == 3.4
    assert torch.equal(torch.load(path)["w"], m.w.detach())


def test_sample_images_replaces_nan_and_clamps():
    out = driver.sample_images(Toy(), samples=2)
    assert torch.equal(out.flatten(), torch.tensor([-0.5, 0.5, -0.5, 0.1]))


def test_fit_loop_on_cpu_toy(tmp_path):
    torch.manual_seed(1)
    m = Toy()
    data = [(torch.randn(8, 3), None) for _ in range(5)]
    hist = driver.fit(m, data, data[:2], epochs=3, checkpoint_path=os.path.join(tmp_path, "b.pt"), lr=0.05, warm_up=8,
                      use_graph=False)
    assert [h["epoch"] for h in hist] == [0, 1, 2]
    assert hist[-1]["test_nll"] < hist[0]["test_nll"]              # it learns
    assert hist[-1]["best_test_nll"] == min(h["test_nll"] for h in hist)


@pytest.mark.gpu
def test_fit_and_checkpoint_round_trip_on_gpu(tmp_path):
    from flowk.marscf import MarScfFlow
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    model = MarScfFlow(8, (16, 16, 3), "affine", 2, 2, 16).to(dev)
    g = torch.Generator().manual_seed(3)
    train = [(torch.rand(8, 3, 16, 16, generator=g) - 0.5, None) for _ in range(4)]
    test = [(torch.rand(8, 3, 16, 16, generator=g) - 0.5, None) for _ in range(2)]
    path = os.path.join(tmp_path, "best.pt")
    hist = driver.fit(model, train, test, epochs=2, checkpoint_path=path, lr=1e-3, warm_up=16, device=dev, use_graph=False)
    assert all(math.isfinite(h["test_nll"]) for h in hist) and os.path.exists(path)
    clone = MarScfFlow(8, (16, 16, 3), "affine", 2, 2, 16).to(dev)
    clone.load_state_dict(torch.load(path), strict=True)
    torch.manual_seed(5)
    a = driver.test_model(clone, test, dev)
    # the checkpoint is the best epoch's weights: re-evaluating it reproduces that epoch's score up to dequantisation noise
    assert abs(a - min(h["test_nll"] for h in hist)) < 0.05
    imgs = driver.sample_images(clone, samples=4)
    assert imgs.shape == (4, 3, 16, 16) and float(imgs.min()) >= -0.5 and float(imgs.max()) <= 0.5
