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


def test_lightning_eval_hooks_single_forward_and_epoch_statistics():
    """validation_step / test_step keep the outputs of their ONE forward (the reference runs the model twice per batch,
    src/vit.py:127-150) and on_validation_epoch_end logs median bias / p90 / slope per output (src/vit.py:157-192)."""
    import numpy as np

    from vit_b200.lightning_module import ViTLModule
    from vit_b200.model import ModelOutput

    cfg = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                         num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"},
           "data": {"param": "a,b"}}

    class Dummy(torch.nn.Module):
        loss_name = "l2"

    lm = ViTLModule(model=Dummy(), config=cfg)
    calls = []
    g = torch.Generator().manual_seed(0)

    def fake_forward(flux, labels, loss_only=True):
        calls.append(flux.shape[0])
        logits = labels + 0.1 * torch.randn(labels.shape, generator=g)
        out = ModelOutput(loss=((logits - labels) ** 2).mean(), logits=logits, hidden_states=None, attentions=None)
        return out.loss if loss_only else out

    lm.forward = fake_forward
    lm.on_validation_start()
    ys, ps = [], []
    for i in range(3):
        y = torch.rand(5, 2, generator=g)
        lm.validation_step((torch.zeros(5, 256), torch.zeros(5, 256), y), i)
        ys.append(y)
        ps.append(lm._last_outputs.logits)
    assert calls == [5, 5, 5]                       # one forward per batch
    assert abs(float(lm._logged["val_mae"]) - float((ps[-1] - ys[-1]).abs().mean())) < 1e-6
    assert len(lm.val_dict["preds"]) == 3
    lm.on_validation_epoch_end()
    P, Y = torch.cat(ps).numpy(), torch.cat(ys).numpy()
    for i in range(2):
        r = P[:, i] - Y[:, i]
        assert abs(lm._logged[f"val_bias_median_{i}"] - float(np.median(r))) < 1e-6
        assert abs(lm._logged[f"val_p90_{i}"] - float(np.percentile(np.abs(r), 90))) < 1e-6
        assert abs(lm._logged[f"val_beta_{i}"] - float(np.polyfit(Y[:, i], P[:, i], 1)[0])) < 1e-6
    assert lm.val_dict == {"preds": [], "labels": []}
    lm.on_test_start()
    lm.test_step((torch.zeros(4, 256), torch.zeros(4, 256), torch.zeros(4, 256), torch.rand(4, 2, generator=g)), 0)
    assert len(lm.test_dict["preds"]) == 1 and calls[-1] == 4
    lm.on_test_epoch_end()                          # no src.viz here: returns quietly


def test_configure_optimizers_mirrors_optmodule():
    """Same optimizer / scheduler selection and Lightning dicts as OptModule (src/opt/optimizer.py:37-172)."""
    import copy

    from vit_b200.lightning_module import ViTLModule

    base = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                          num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"},
            "data": {"param": "g", "val_path": "v.h5", "num_samples": 1000}, "train": {"batch_size": 64, "ep": 7}}
    S = torch.optim.lr_scheduler

    def make(**opt):
        cfg = copy.deepcopy(base)
        cfg["opt"] = opt
        from vit_b200 import get_model

        return ViTLModule(model=get_model(copy.deepcopy(cfg), device="cpu"), config=cfg).configure_optimizers()

    o = make(type="AdamW", lr=2e-3)
    assert isinstance(o, torch.optim.AdamW) and o.param_groups[0]["lr"] == 2e-3 and o.param_groups[0]["weight_decay"] == 0
    assert isinstance(make(lr=1e-3), torch.optim.Adam)                       # default type 'adam'
    d = make(type="adamw", lr_sch="plateau", factor=0.8, patience=10)     # configs/config.yaml
    assert isinstance(d["lr_scheduler"]["scheduler"], S.ReduceLROnPlateau)
    assert d["lr_scheduler"]["monitor"] == "val_mae" and d["lr_scheduler"]["reduce_on_plateau"] and not d["lr_scheduler"]["strict"]
    assert d["lr_scheduler"]["scheduler"].factor == 0.8 and d["lr_scheduler"]["scheduler"].patience == 10
    d = make(type="adamw", lr_sch="cosine", ep=50, eta_min=1e-6)
    assert isinstance(d["lr_scheduler"]["scheduler"], S.CosineAnnealingLR) and d["lr_scheduler"]["scheduler"].T_max == 50
    assert d["lr_scheduler"]["interval"] == "epoch"
    d = make(type="adamw", lr=1e-3, lr_sch="onecycle", pct_start=0.2)
    sch = d["lr_scheduler"]["scheduler"]
    assert isinstance(sch, S.OneCycleLR) and sch.total_steps == 7 * 16 and d["lr_scheduler"]["interval"] == "step"
    d = make(type="sgd", lr_sch="constant", factor=0.5, total_iters=3)
    assert isinstance(d["optimizer"], torch.optim.SGD) and isinstance(d["lr_scheduler"]["scheduler"], S.ConstantLR)
    d = make(type="adamw", lr_sch="cosine", ep=20, warmup_ratio=0.1)
    assert isinstance(d["lr_scheduler"]["scheduler"], S.SequentialLR)
    with pytest.raises(ValueError, match="Unknown scheduler"):
        make(type="adamw", lr_sch="step")
    cfg = copy.deepcopy(base)
    cfg["data"].pop("val_path")                                            # plateau without validation data: disabled
    cfg["opt"] = {"type": "adamw", "lr_sch": "plateau"}
    from vit_b200 import get_model

    assert isinstance(ViTLModule(model=get_model(copy.deepcopy(cfg), device="cpu"), config=cfg).configure_optimizers(),
                      torch.optim.AdamW)


_GLOO_SAMPLER_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VIT_ROOT"])
from vit_b200 import dp
from vit_b200.data import epoch_indices
from oracle import vit_oracle as vo
rank, local, world = dp.init_from_env("gloo")
cfg = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                     num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}, "data": {"param": "g"}}
spec = vo.spec_from_config(cfg)
params = vo.init_params(spec, seed=5)
N, B = 21, 4
x, y = vo.synthetic_batch(N, 256, seed=3, kind="rand")
mine = epoch_indices(N, 2, seed=9, rank=rank, world=world, batch=B)
both = [epoch_indices(N, 2, seed=9, rank=r, world=world, batch=B) for r in range(world)]
assert mine.numel() % B == 0 and all(b.numel() == mine.numel() for b in both)
assert set(torch.cat(both).tolist()) == set(range(N))           # the ranks together visit every sample
def grads_of(idx):
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    vo.forward(p, x[idx], spec, labels=y[idx])["loss"].backward()
    return torch.cat([t.grad.reshape(-1) for k, t in p.items() if t.grad is not None])
for i in range(mine.numel() // B):
    g = grads_of(mine[i * B:(i + 1) * B])
    dist.all_reduce(g)                                           # SUM; the optimizer kernel's grad_scale is 1/world
    g /= world
    union = torch.cat([b[i * B:(i + 1) * B] for b in both])      # what one process with batch world*B would see
    full = grads_of(union)
    err = float((g - full).abs().max() / full.abs().max())
    assert err < 1e-5, (i, err)
if rank == 0: print("SAMPLER_OK")
dist.destroy_process_group()
'''


def test_fit_device_sharding_equals_global_batch_gloo_world2(tmp_path):
    """world_size 2 over gloo: stepping on rank r's share of the epoch permutation (what TrainStep.fit_device does) and
    averaging the gradients over ranks == one process stepping on the union batch (Lightning DDP + DistributedSampler)."""
    import os
    import subprocess
    import sys

    from conftest import ROOT

    script = tmp_path / "worker.py"
    script.write_text(_GLOO_SAMPLER_WORKER)
    env = dict(os.environ, VIT_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SAMPLER_OK" in r.stdout
