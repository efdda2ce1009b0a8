"""CPU: pin the oracle restatement (oracle/vit_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/gen_golden.py), and against the live reference when present."""
import pytest
import torch

from conftest import GOLDEN_VARIANTS, grad_floor, rel_err
from oracle import vit_oracle as vo


def _inputs(fix):
    cfg = fix["config"]
    x, _ = vo.synthetic_batch(fix["batch"], cfg["model"]["image_size"], seed=fix["x_seed"], kind=fix["x_kind"])
    return x, fix["labels"]


@pytest.mark.parametrize("name", GOLDEN_VARIANTS)
def test_oracle_forward_and_grads_fp32(golden, name):
    fix = golden(name)
    spec = vo.spec_from_config(fix["config"])
    x, y = _inputs(fix)
    params = {k: v.clone().requires_grad_(True) for k, v in fix["state_dict"].items()}
    assert set(params) == set(vo.param_shapes(spec)), "state_dict keys differ from the reference"
    for k, shp in vo.param_shapes(spec).items():
        assert tuple(params[k].shape) == shp, k
    out = vo.forward(params, x, spec, labels=y)
    out["loss"].backward()
    assert rel_err(out["loss"], fix["eval"]["loss"]) < 1e-5
    assert rel_err(out["logits"], fix["eval"]["logits"]) < 1e-5
    assert rel_err(out["last_hidden"][:, 0], fix["eval"]["last_hidden_cls"]) < 1e-5
    assert len(out["hidden_states"]) == len(fix["eval"]["hidden_states"])
    for mine, ref in zip(out["hidden_states"], fix["eval"]["hidden_states"]):
        assert rel_err(mine[0], ref) < 1e-5
    got = {k: p.grad for k, p in params.items() if p.grad is not None}
    assert set(got) == set(fix["grads"]), "set of params receiving a gradient differs (pooler must get none)"
    fl = grad_floor(fix["grads"])
    for k, g in fix["grads"].items():
        assert rel_err(got[k], g, fl) < 2e-4, k


@pytest.mark.parametrize("name", ["baseline", "rope", "cls", "h64multi"])
def test_oracle_bf16_autocast(golden, name):
    fix = golden(name)
    spec = vo.spec_from_config(fix["config"])
    x, y = _inputs(fix)
    params = {k: v.clone().requires_grad_(True) for k, v in fix["state_dict"].items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = vo.forward(params, x, spec, labels=y)
    out["loss"].backward()
    assert rel_err(out["loss"].float(), fix["bf16"]["loss"]) < 2e-2
    assert rel_err(out["logits"].float(), fix["bf16"]["logits"]) < 2e-2
    fl = grad_floor(fix["bf16"]["grads"])
    for k, g in fix["bf16"]["grads"].items():
        assert rel_err(params[k].grad, g, fl) < 5e-2, k


@pytest.mark.parametrize("name", ["baseline", "learned", "cls"])
def test_oracle_train_steps(golden, name):
    """clip_grad_norm_(0.5) + AdamW restated == torch's, over 3 steps (dropout off)."""
    fix = golden(name)
    spec = vo.spec_from_config(fix["config"])
    x, y = _inputs(fix)
    tr = vo.OracleTrainer(spec, fix["state_dict"])
    for i in range(3):
        loss = tr.step(x, y, train=False)
        assert abs(loss - float(fix["train3"]["losses"][i])) < 1e-5 * max(1.0, abs(loss))
        assert rel_err(tr.last_grad_norm, fix["train3"]["grad_norms"][i]) < 1e-4
    for k, v in fix["train3"]["state_dict"].items():
        assert rel_err(tr.params[k], v, 3e-3) < 1e-4, k   # floor = 3 steps x lr (key.bias has a zero gradient)


@pytest.mark.parametrize("name", ["baseline", "cls"])
def test_oracle_train_steps_bf16_autocast(golden, name):
    """The autocast-bf16 oracle trainer vs the reference's own 3 autocast steps: losses, grad norms and AdamW first
    moments (the quantity the GPU bf16 train-step tests pin; post-step weights are not comparable in bf16)."""
    fix = golden(name)
    spec = vo.spec_from_config(fix["config"])
    x, y = _inputs(fix)
    tr = vo.OracleTrainer(spec, fix["state_dict"], autocast_bf16=True)
    ref = fix["train3_bf16"]
    for i in range(3):
        loss = tr.step(x, y, train=False)
        assert abs(loss - float(ref["losses"][i])) < 2e-2 * max(1.0, abs(loss))
        assert rel_err(tr.last_grad_norm, ref["grad_norms"][i]) < 0.2
    gmax = max(float(v.abs().max()) for v in ref["exp_avg"].values())
    for k, v in ref["exp_avg"].items():
        assert rel_err(tr.m[k], v, 1e-2 * gmax) < 4e-2, (k, rel_err(tr.m[k], v, 1e-2 * gmax))


def test_spec_quirks():
    cfg = {"model": dict(task_type="reg", image_size=4096, patch_size=32, hidden_size=32, num_hidden_layers=3,
                         num_attention_heads=2, stride_size=32, proj_fn="SW"), "loss": {"name": "mae"},
           "data": {"param": "log_g"}}
    s = vo.spec_from_config(cfg)
    assert s.loss_kind == "mse"          # 'mae' selects MSELoss (specvit.py:52-53)
    assert s.tokens == 129 and s.intermediate == 128 and s.head_dim == 16
    assert sum(torch.Size(v).numel() for v in vo.param_shapes(s).values()) == 40353
    cfg["model"].update(patch_size=48, stride_size=48, image_size=1000)
    assert vo.spec_from_config(cfg).num_patches == 21   # ceil -> zero-padded tail window
    cfg["model"]["proj_fn"] = "CNN"
    assert vo.spec_from_config(cfg).num_patches == 20   # floor
    fwd, step = vo.flops_per_sample(s)
    assert abs(fwd / 1e6 - 16.16) < 0.01 and abs(step / 1e6 - 51.68) < 0.01


def test_oracle_vs_live_reference():
    from oracle import ref_shims

    if not ref_shims.reference_available():
        pytest.skip("/root/reference not present (GPU box)")
    cfg = {"model": dict(name="vit", task_type="reg", image_size=512, patch_size=16, hidden_size=32,
                         num_hidden_layers=2, num_attention_heads=4, stride_size=12, proj_fn="SW",
                         pos_encoding_type="rope"),
           "loss": {"name": "l1"}, "data": {"param": "a,b"}, "noise": {"noise_level": 0}}
    torch.manual_seed(0)
    model = ref_shims.reference_get_model(cfg).eval()
    spec = vo.spec_from_config(cfg)
    x = torch.rand(3, 512)
    y = torch.rand(3, 2)
    ref = model(x, labels=y)
    mine = vo.forward(dict(model.state_dict()), x, spec, labels=y)
    assert rel_err(mine["loss"], ref.loss) < 1e-5
    assert rel_err(mine["logits"], ref.logits) < 1e-5
