"""CPU: the checkpoint wire format -- Lightning `.ckpt` dicts of the reference <-> the flat arena
and the fused optimizer's flat AdamW moments, in both directions, checked against a real torch.optim.AdamW."""
import copy

import pytest
import torch

from vit_b200 import checkpoint as ck


def _model(golden, name="baseline"):
    from vit_b200 import get_model

    fix = golden(name)
    m = get_model(copy.deepcopy(fix["config"]), device="cpu")
    m.load_state_dict(fix["state_dict"])
    return m, fix


def _fake_grads(m, seed):
    g = torch.Generator().manual_seed(seed)
    for n, p in m.named_parameters():
        p.grad = None if "pooler" in n else torch.randn(p.shape, generator=g) * 1e-2


class _Eng:   # what load/save touch of a ViTEngine, on CPU
    def __init__(self, m):
        self.arena = m._arena
        n = m._arena.layout.n_total
        self.exp_avg, self.exp_avg_sq = torch.zeros(n), torch.zeros(n)
        self.state, self.hyper = torch.zeros(8), torch.zeros(8)
        self.shadow_refreshed = False

    def _ensure_opt_state(self):
        pass

    def set_lr(self, lr):
        self.hyper[0] = lr

    def refresh_shadow(self, force=False):
        self.shadow_refreshed = True


class _Step:
    def __init__(self, m):
        self.eng = _Eng(m)


def test_lightning_checkpoint_round_trip(golden, tmp_path):
    m, fix = _model(golden)
    opt = torch.optim.AdamW(m.parameters(), lr=8e-4, weight_decay=0)
    for s in range(2):
        _fake_grads(m, s)
        opt.step()
    ref_ckpt = {"epoch": 3, "global_step": 2, "pytorch-lightning_version": "2.5.0",
                "state_dict": {"model." + k: v.clone() for k, v in m.state_dict().items()},
                "optimizer_states": [copy.deepcopy(opt.state_dict())], "lr_schedulers": []}
    path = tmp_path / "ref.ckpt"
    torch.save(ref_ckpt, path)

    # reference .ckpt -> arena + flat moments
    m2, _ = _model(golden)
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    st = _Step(m2)
    info = ck.load_lightning_checkpoint(m2, str(path), train_step=st)
    assert info["epoch"] == 3 and info["step"] == 2 and not info["extra_optimizer_state"]
    for k, v in m.state_dict().items():
        assert torch.equal(m2.state_dict()[k], v), k
    assert float(st.eng.state[0]) == 2.0 and abs(float(st.eng.hyper[0]) - 8e-4) < 1e-9 and st.eng.shadow_refreshed
    lay = m2._arena.layout
    names = [n for n, _ in m.named_parameters()]
    for i, n in enumerate(names):
        e = lay.entries[n]
        sl = slice(e.offset, e.offset + e.numel)
        if "pooler" in n:
            assert i not in opt.state_dict()["state"]
            assert float(st.eng.exp_avg[sl].abs().max()) == 0.0
            continue
        s = opt.state_dict()["state"][i]
        assert torch.equal(st.eng.exp_avg[sl].view(e.shape), s["exp_avg"]), n
        assert torch.equal(st.eng.exp_avg_sq[sl].view(e.shape), s["exp_avg_sq"]), n

    # arena + flat moments -> a .ckpt the reference's torch optimizer accepts, bit-identical to what it wrote itself
    out = ck.save_lightning_checkpoint(m2, str(tmp_path / "ours.ckpt"), train_step=st, epoch=3)
    back = torch.load(tmp_path / "ours.ckpt", map_location="cpu", weights_only=False)
    assert list(back["state_dict"]) == list(ref_ckpt["state_dict"])
    assert back["global_step"] == 2 and out["epoch"] == 3
    a, b = back["optimizer_states"][0], opt.state_dict()
    assert set(a["state"]) == set(b["state"])
    for i in b["state"]:
        assert float(a["state"][i]["step"]) == float(b["state"][i]["step"])
        assert torch.equal(a["state"][i]["exp_avg"], b["state"][i]["exp_avg"])
        assert torch.equal(a["state"][i]["exp_avg_sq"], b["state"][i]["exp_avg_sq"])
    assert set(a["param_groups"][0]) == set(b["param_groups"][0])
    assert a["param_groups"][0]["params"] == b["param_groups"][0]["params"]
    assert a["param_groups"][0]["lr"] == b["param_groups"][0]["lr"] == 8e-4

    # and torch resumes from it exactly like from its own state
    m3, _ = _model(golden)
    m3.load_state_dict(ck.strip_prefix(back["state_dict"]))
    opt3 = torch.optim.AdamW(m3.parameters(), lr=1.0)
    opt3.load_state_dict(back["optimizer_states"][0])
    _fake_grads(m, 7)
    _fake_grads(m3, 7)
    opt.step()
    opt3.step()
    for (n, p), (_, q) in zip(m.named_parameters(), m3.named_parameters()):
        assert torch.equal(p, q), n


def test_checkpoint_errors(golden):
    m, fix = _model(golden)
    sd = {"model." + k: v for k, v in fix["state_dict"].items()}
    sd.pop("model.vit.layernorm.weight")
    with pytest.raises(RuntimeError, match="Missing key"):
        ck.load_lightning_checkpoint(m, {"state_dict": sd})
    ck.load_lightning_checkpoint(m, {"state_dict": sd}, strict=False)
    ck.load_lightning_checkpoint(m, fix["state_dict"])          # a bare (unprefixed) state_dict loads too
    names = [n for n, _ in m.named_parameters()]
    lay = m._arena.layout
    z = torch.zeros(lay.n_total)
    bad = {"state": {}, "param_groups": [{"params": list(range(len(names) - 1))}]}
    with pytest.raises(ValueError, match="optimizer covers"):
        ck.adam_state_from_torch(bad, names, lay, z, z.clone())
    two = {"state": {0: {"step": torch.tensor(1.0), "exp_avg": torch.zeros(32), "exp_avg_sq": torch.zeros(32)},
                     1: {"step": torch.tensor(2.0), "exp_avg": torch.zeros(1024), "exp_avg_sq": torch.zeros(1024)}},
           "param_groups": [{"params": list(range(len(names)))}]}
    with pytest.raises(ValueError, match="step counters differ"):
        ck.adam_state_from_torch(two, names, lay, z, z.clone())


def test_trainable_preprocessor_state_is_kept_aside(golden, tmp_path):
    """Optimizer state of parameters outside the arena (an unfrozen preprocessor matrix) is handed back by name."""
    from conftest import config_with_cov
    from vit_b200 import get_model
    from vit_b200.preprocessor import clear_cov_cache

    fix = golden("pre_zca_r32")
    clear_cov_cache()
    m = get_model(config_with_cov(fix, tmp_path), device="cpu")
    m.load_state_dict(fix["state_dict"])
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    _fake_grads(m, 1)
    opt.step()
    names = [n for n, _ in m.named_parameters()]
    lay = m._arena.layout
    ea, es = torch.zeros(lay.n_total), torch.zeros(lay.n_total)
    step, hyper, extra = ck.adam_state_from_torch(opt.state_dict(), names, lay, ea, es)
    assert step == 1 and set(extra) == {"preprocessor.linear.weight", "preprocessor.linear.bias"}
    back = ck.adam_state_to_torch(names, lay, ea, es, step, extra=extra)
    assert set(back["state"]) == set(opt.state_dict()["state"])
