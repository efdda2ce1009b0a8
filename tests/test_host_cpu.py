"""CPU-only checks (no kernels are launched): C-ABI surface, host-side mirror of the reference interface,
arena layout, data-parallel helpers (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import GOLDEN_VARIANTS, ROOT
from oracle import vit_oracle as vo


def _header_functions():
    src = open(os.path.join(ROOT, "include", "vit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vitb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from vit_b200 import _lib

    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vit_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in vit_b200/_lib.py"
    assert lib.vitb200_version() == 100
    assert b"unsupported shape" in lib.vitb200_strerror(-2)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.vitb200_linear_fwd(None, None, None, None, None, 1, 1, 1, 0, 0, None) == -3
    assert lib.vitb200_attn_fwd(1, 1, 1, 48, 1, 1, None, None, 1, 4, 1, 48, 1.0, 0.0, None, 0, 0, None) == -2
    assert lib.vitb200_add_ln_fwd(1, None, None, 1, 1, 1, 1, 1, 4, 30, 0, 1e-12, 0.0, None, 0, 0, None) == -2


def test_product_has_no_cpu_fallback():
    from vit_b200 import get_model

    cfg = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                         num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}}
    m = get_model(cfg, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.rand(2, 256))
    # nothing in the product package may import the oracle
    for f in os.listdir(os.path.join(ROOT, "vit_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "vit_b200", f)).read().replace("# oracle", ""), f


@pytest.mark.parametrize("name", GOLDEN_VARIANTS)
def test_state_dict_matches_reference(golden, name):
    """Same keys, shapes and ORDER as the reference's state_dict; same model name string; loads strictly."""
    from vit_b200 import get_model

    fix = golden(name)
    m = get_model(fix["config"], device="cpu")
    sd = m.state_dict()
    assert list(sd) == list(fix["state_dict"])
    for k, v in fix["state_dict"].items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    assert m.name == fix["model_name"] and m.loss_name == fix["loss_name"]
    m.load_state_dict(fix["state_dict"], strict=True)
    for k, v in fix["state_dict"].items():
        assert torch.equal(m.state_dict()[k], v)
    # all parameters are views of one arena; q/k/v are adjacent so they form one [3H, H] matrix
    base = m._arena.data.data_ptr()
    for p in m.parameters():
        assert base <= p.data_ptr() < base + m._arena.data.numel() * 4
    a = m.vit.encoder.layer[0].attention.attention
    H = m.config.hidden_size
    assert a.key.weight.data_ptr() == a.query.weight.data_ptr() + 4 * H * H
    assert a.value.weight.data_ptr() == a.query.weight.data_ptr() + 8 * H * H
    assert a.value.bias.data_ptr() == a.query.bias.data_ptr() + 8 * H
    assert sum(p.numel() for p in m.parameters()) == sum(v.numel() for v in fix["state_dict"].values())


def test_config_quirks_match_reference_builder():
    from vit_b200.builder import get_vit_config

    cfg = {"model": dict(task_type="reg", image_size=1000, patch_size=48, hidden_size=32, num_hidden_layers=3,
                         num_attention_heads=2, stride_size=48, proj_fn="SW", num_labels=7),
           "data": {"param": "a, b ,c"}}
    c = get_vit_config(cfg)
    assert c.num_labels == 3 and cfg["model"]["num_labels"] == 3      # builder.py:206-225
    assert c.num_patches == 21 and c.n_valid == 20                    # ceil + one zero-padded window
    assert c.intermediate_size == 128 and c.layer_norm_eps == 1e-12
    cfg["model"]["proj_fn"] = "CNN"
    assert get_vit_config(cfg).num_patches == 20                      # floor, tokenization.py:63
    cfg["model"].update(proj_fn="SW", stride_size=None, stride_ratio=0.5)
    assert get_vit_config(cfg).stride == 24                           # embedding.py:26-27
    cfg["model"]["hidden_size"] = 48
    with pytest.raises(ValueError, match="head_dim"):                 # error at construction, never a fallback
        get_vit_config(cfg)
    for name in GOLDEN_VARIANTS:  # host mirror agrees with the oracle's shape logic on every variant
        from conftest import load_golden
        g = load_golden(name)["config"]
        import copy
        a, b = get_vit_config(copy.deepcopy(g)), vo.spec_from_config(copy.deepcopy(g))
        assert (a.num_patches, a.tokens, a.stride, a.num_labels) == (b.num_patches, b.tokens, b.stride, b.num_labels)


def test_loss_selection_quirk():
    from vit_b200 import get_model, _lib

    base = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                          num_attention_heads=2, stride_size=32, proj_fn="SW")}
    assert get_model({**base, "loss": {"name": "mae"}}, device="cpu")._loss_kind == _lib.LOSS_MSE
    assert get_model({**base, "loss": {"name": "l1"}}, device="cpu")._loss_kind == _lib.LOSS_L1
    assert get_model({**base, "loss": {"name": "smooth_L1"}}, device="cpu")._loss_kind == _lib.LOSS_L1
    m = get_model(base, device="cpu")
    assert m._loss_kind == _lib.LOSS_MSE and m.loss_name == "l2"


def test_shard_range_partitions():
    from vit_b200.dp import shard_range

    for n in (0, 1, 7, 64, 129):
        for w in (1, 2, 3, 8):
            got = [shard_range(n, r, w) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VIT_ROOT"])
from vit_b200 import dp
from vit_b200.arena import ParamLayout
from vit_b200.builder import get_vit_config
from oracle import vit_oracle as vo
rank, local, world = dp.init_from_env("gloo")
cfg = {"model": dict(task_type="reg", image_size=512, patch_size=32, hidden_size=32, num_hidden_layers=2,
                     num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}, "data": {"param": "g"}}
lay = ParamLayout(get_vit_config(cfg))
spec = vo.spec_from_config(cfg)
params = vo.init_params(spec, seed=5)
flat = torch.zeros(lay.n_total)
if rank == 0:
    for k, v in params.items():
        e = lay.entries[k]; flat[e.offset:e.offset + e.numel] = v.reshape(-1)
dp.broadcast_parameters(flat)
params = {k: flat[lay.entries[k].offset:lay.entries[k].offset + lay.entries[k].numel].view(lay.entries[k].shape).clone()
          for k in lay.entries}
B = 6
x, y = vo.synthetic_batch(B, 512, seed=3, kind="rand")
def grads_of(xs, ys):
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    vo.forward(p, xs, spec, labels=ys)["loss"].backward()
    g = torch.zeros(lay.n_total)
    for k, t in p.items():
        if t.grad is not None:
            e = lay.entries[k]; g[e.offset:e.offset + e.numel] = t.grad.reshape(-1)
    return g
lo, hi = dp.shard_range(B, rank, world)
g_local = grads_of(x[lo:hi], y[lo:hi])
dp.allreduce_all(g_local, lay.buckets)
g_local[:lay.n_opt] *= 1.0 / world          # the optimizer kernel's grad_scale
g_full = grads_of(x, y)
err = float((g_local - g_full).abs().max() / g_full.abs().max())
assert err < 1e-5, err
assert float(g_local[lay.n_opt:].abs().max()) == 0.0   # pooler: outside every bucket, never exchanged
covered = sum(e - s for _, s, e in lay.buckets)
assert covered == lay.n_opt
if rank == 0: print("GLOO_OK", err)
dist.destroy_process_group()
'''


def test_dp_gradient_mean_equals_full_batch_gloo_world2(tmp_path):
    """world_size 2 over gloo: bucketed SUM all-reduce of per-shard gradients x 1/world == full-batch gradients."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, VIT_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_OK" in r.stdout


def test_shadow_goes_stale_after_to_and_foreign_updates():
    """after model.to()/_apply the Parameters keep their own version counters; updates through them
    (torch.optim step, load_state_dict, in-place ops) must still mark the bf16 operand copy stale."""
    from vit_b200 import get_model

    cfg = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                         num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}}
    m = get_model(cfg, precision="bf16-mixed", device="cpu")
    m._apply(lambda t: t.clone())          # what .to(device) / .cuda() do
    ar = m._arena
    assert ar.shadow is not None
    ar.mark_shadow_fresh()
    assert not ar.shadow_stale()
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    assert ar.shadow_stale()
    ar.mark_shadow_fresh()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    assert ar.shadow_stale()
    ar.mark_shadow_fresh()
    m.load_state_dict(m.state_dict())
    assert ar.shadow_stale()


def test_fused_optimizer_state_dict_is_torch_adamw_format():
    """FusedClipAdamW.state_dict()/load_state_dict() carry the flat moments in torch.optim.AdamW's
    per-parameter layout (what Lightning's ModelCheckpoint stores and the reference resumes from)."""
    from vit_b200 import get_model
    from vit_b200.optim import FusedClipAdamW

    cfg = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                         num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}}
    m = get_model(cfg, device="cpu")
    ref = torch.optim.AdamW(m.parameters(), lr=2e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.0)
    torch.manual_seed(0)
    for _ in range(2):
        for n, p in m.named_parameters():
            p.grad = None if "pooler" in n else torch.randn_like(p)
        ref.step()
    want = ref.state_dict()
    opt = FusedClipAdamW(m, lr=1e-3)
    assert opt.state_dict()["state"] == {}
    opt.load_state_dict(want)
    got = opt.state_dict()
    assert sorted(got["state"]) == sorted(want["state"])
    for i, st in want["state"].items():
        assert float(got["state"][i]["step"]) == float(st["step"]) == 2.0
        assert torch.equal(got["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(got["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
    g = got["param_groups"][0]
    assert g["lr"] == 2e-3 and tuple(g["betas"]) == (0.8, 0.95) and g["eps"] == 1e-7
    # a stock AdamW accepts it
    torch.optim.AdamW(m.parameters()).load_state_dict({"state": got["state"], "param_groups": [
        {k: v for k, v in g.items() if k != "max_norm"}]})


def test_lightning_checkpoint_has_no_empty_loops_key():
    """Lightning's restore_loops() indexes ckpt['loops']['fit_loop'] whenever 'loops' exists."""
    from vit_b200 import get_model
    from vit_b200.checkpoint import save_lightning_checkpoint

    cfg = {"model": dict(task_type="reg", image_size=256, patch_size=32, hidden_size=32, num_hidden_layers=1,
                         num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"}}
    ck = save_lightning_checkpoint(get_model(cfg, device="cpu"), epoch=3, global_step=17,
                                   lr_schedulers=[{"best": 0.5, "num_bad_epochs": 2}])
    assert "loops" not in ck and ck["epoch"] == 3 and ck["global_step"] == 17
    assert ck["lr_schedulers"] == [{"best": 0.5, "num_bad_epochs": 2}]
