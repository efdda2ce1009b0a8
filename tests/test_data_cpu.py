"""CPU: host logic of the device-resident data hand-off and the evaluation metrics."""
import pytest
import torch

from vit_b200.data import DeviceDataset, EvalMetrics, epoch_indices


def test_epoch_indices_distributed_sampler_semantics():
    n, B = 103, 8
    for world in (1, 2, 4):
        shards = [epoch_indices(n, 5, seed=3, shuffle=True, rank=r, world=world, batch=B) for r in range(world)]
        assert len({s.numel() for s in shards}) == 1 and shards[0].numel() % B == 0
        seen = torch.cat(shards)
        assert set(seen.tolist()) == set(range(n))            # every sample is visited each epoch
        assert seen.numel() - n < world + world * B           # only the padding is repeated
    # same permutation on every rank, a different one per epoch, reproducible
    a = epoch_indices(n, 0, seed=1, batch=1)
    assert torch.equal(a, epoch_indices(n, 0, seed=1, batch=1)) and not torch.equal(a, epoch_indices(n, 1, seed=1, batch=1))
    assert sorted(a.tolist()) == list(range(n))
    r0 = epoch_indices(n, 0, seed=1, rank=0, world=2, batch=1)
    assert torch.equal(r0, torch.cat([a, a[:1]])[0::2])       # padded by wrapping, then strided by rank
    # no shuffle (debug mode, src/basemodule.py:82), drop_last
    d = epoch_indices(n, 0, shuffle=False, batch=B, tail="drop")
    assert torch.equal(d, torch.arange(96))
    w = epoch_indices(n, 0, shuffle=False, batch=B, tail="wrap")
    assert w.numel() == 104 and int(w[-1]) == 0
    tiny = epoch_indices(3, 0, shuffle=False, batch=8)         # dataset smaller than a batch: wraps repeatedly
    assert tiny.tolist() == [0, 1, 2, 0, 1, 2, 0, 1]
    with pytest.raises(ValueError):
        epoch_indices(0, 0)
    with pytest.raises(ValueError):
        epoch_indices(10, 0, rank=2, world=2)


def test_eval_metrics_formulas_match_torchmetrics_definitions():
    g = torch.Generator().manual_seed(0)
    n, C = 1000, 2
    y = torch.rand(n, C, generator=g).double()
    p = y + 0.1 * torch.randn(n, C, generator=g).double()
    e = p - y
    m = EvalMetrics(C, False, "cpu")
    m.acc[0], m.acc[1] = n, n * 0.25
    for c in range(C):
        m.acc[2 + 4 * c:6 + 4 * c] = torch.stack([e[:, c].abs().sum(), (e[:, c] ** 2).sum(), y[:, c].sum(), (y[:, c] ** 2).sum()])
    out = m.compute()
    assert out["n"] == n and abs(out["loss"] - 0.25) < 1e-12
    assert abs(out["mae"] - float(e.abs().mean())) < 1e-12          # MeanAbsoluteError: over all elements
    assert abs(out["mse"] - float((e ** 2).mean())) < 1e-12         # MeanSquaredError
    r2 = 1 - (e ** 2).sum(0) / ((y - y.mean(0)) ** 2).sum(0)         # R2Score, multioutput='uniform_average'
    assert abs(out["r2"] - float(r2.mean())) < 1e-9
    mc = EvalMetrics(5, True, "cpu")
    mc.acc[0], mc.acc[2] = 40, 30
    assert mc.compute()["acc"] == 0.75
    assert EvalMetrics(1, False, "cpu").compute()["n"] == 0        # empty epoch: no division by zero


def test_device_dataset_has_no_cpu_path():
    with pytest.raises(RuntimeError, match="no CPU path"):
        DeviceDataset(torch.rand(4, 16), torch.rand(4))
    with pytest.raises(ValueError, match="multiple of 4"):
        DeviceDataset(torch.rand(4, 18), torch.rand(4), device="cuda")
