"""Multi-GPU (needs >= 2 CUDA devices; run under `gpurun --gpus 2`): data-parallel TrainStep == single-GPU step on the
concatenated batch (gradient mean over ranks == full-batch gradient), bucketed NCCL all-reduce overlapped with backward."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VIT_ROOT"])
from vit_b200 import dp, get_model
from vit_b200.step import TrainStep
from oracle import vit_oracle as vo
local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = {"model": dict(name="vit", task_type="reg", image_size=4096, patch_size=32, hidden_size=32, num_hidden_layers=3,
                     num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}, "data": {"param": "g"}}
prec = os.environ.get("VIT_PREC", "bf16-mixed")
use_graph = os.environ.get("VIT_GRAPH", "1") == "1"
B = 16
x, y = vo.synthetic_batch(B * world, 4096, seed=5, kind="rand")
# single-GPU reference on the concatenated batch, computed on every rank BEFORE any communicator exists
torch.manual_seed(7)
m1 = get_model(cfg, precision=prec, device=dev)
s1 = TrainStep(m1, B * world, use_graph=use_graph, world_size=1, train=False)
l1 = [float(s1.step(x.to(dev), y.to(dev))) for _ in range(3)]
ref_flat = m1._arena.data.clone()
del s1, m1
rank, local, world = dp.init_from_env("nccl")
torch.manual_seed(7)
m = get_model(cfg, precision=prec, device=dev)
dp.broadcast_parameters(m._arena.data)
p0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
step = TrainStep(m, B, use_graph=use_graph, world_size=world, train=False)
lo, hi = rank * B, (rank + 1) * B
losses = [float(step.step(x[lo:hi].to(dev), y[lo:hi].to(dev))) for _ in range(3)]
flat = m._arena.data.clone()
# every rank must hold identical parameters after the steps
ref = flat.clone(); dist.broadcast(ref, src=0)
assert torch.equal(ref, flat), "replicas diverged"
d = float((ref_flat - flat).abs().max() / ref_flat.abs().max())
tol = 2e-3 if prec == "32" else 2e-2
assert d < tol, d
# ... and the ORACLE on the union batch (CPU restatement of the reference step): mean of the rank losses, AdamW first
# moments (= clipped mean gradients) and, in fp32, the parameters themselves
lt = torch.tensor(losses, device=dev).reshape(1, -1)
alll = [torch.zeros_like(lt) for _ in range(world)]
dist.all_gather(alll, lt)
if rank == 0:
    spec = vo.spec_from_config(cfg)
    ora = vo.OracleTrainer(spec, p0, autocast_bf16=prec != "32")
    ol = [ora.step(x, y, train=False) for _ in range(3)]
    mine = torch.cat(alll, 0).mean(0).cpu()
    ltol = 1e-4 if prec == "32" else 2e-2
    for i in range(3):
        assert abs(float(mine[i]) - ol[i]) <= ltol * max(1.0, abs(ol[i])), (i, float(mine[i]), ol[i])
    lay, eng = m._arena.layout, step.eng
    gmax = max(float(v.abs().max()) for k, v in ora.m.items() if "pooler" not in k)
    merr = perr = 0.0
    for k, v in ora.m.items():
        e = lay.entries.get(k)
        if e is None or e.offset >= lay.n_opt:
            continue
        got = eng.exp_avg[e.offset:e.offset + e.numel].reshape(e.shape).cpu()
        merr = max(merr, float((got - v).abs().max()) / max(float(v.abs().max()), 1e-2 * gmax))
        gp = flat[e.offset:e.offset + e.numel].reshape(e.shape).cpu()
        perr = max(perr, float((gp - ora.params[k].detach()).abs().max()) / max(float(ora.params[k].abs().max()), 3e-3))
    assert merr < (2e-3 if prec == "32" else 4e-2), merr
    assert prec != "32" or perr < 1e-3, perr
    print("DP_OK", json.dumps({"world": world, "param_rel_diff_vs_single_gpu": d, "exp_avg_rel_err_vs_oracle": merr,
                               "param_rel_err_vs_oracle": perr, "losses_mean": mine.tolist(), "losses_oracle": ol}))
dist.barrier()
step.close()   # the graph holds captured NCCL collectives: it must go before the communicator
dist.destroy_process_group()
'''


@pytest.mark.parametrize("prec,graph", [("32", "0"), ("bf16-mixed", "1")])
def test_dp_matches_single_gpu(tmp_path, prec, graph):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    n = min(torch.cuda.device_count(), 8)
    env = dict(os.environ, VIT_ROOT=ROOT, VIT_PREC=prec, VIT_GRAPH=graph)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-5000:]
    assert "DP_OK" in r.stdout
    print(r.stdout[-600:])
