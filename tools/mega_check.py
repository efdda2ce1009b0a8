# compare the whole-network forward with the per-op fused forward on the same model / inputs
import os, sys, json, copy, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from vit_b200 import get_model
from oracle import vit_oracle as vo
dev = torch.device("cuda:0")
def run(cfg, B, mega, train, seed=3):
    os.environ["VITB200_MEGA"] = "1" if mega else "0"
    torch.manual_seed(seed)
    m = get_model(copy.deepcopy(cfg), precision="bf16-mixed", device=dev)
    m.train(train)
    eng = m._engine(B)
    x, y = vo.synthetic_batch(B, cfg["model"]["image_size"], seed=5, kind="rand")
    if cfg["model"].get("task_type") == "cls":
        y = torch.randint(0, cfg["model"]["num_labels"], (B,))
    m._stage_raw(eng, x.to(dev), y.to(dev))
    eng.forward(train=train, with_labels=True)
    torch.cuda.synchronize()
    assert eng.mega == mega, (eng.mega, mega)
    out = dict(z=eng.z_all.clone(), hmid=eng.hmid_all.clone(), u=eng.u_all.float().clone(), u2=eng.u2_all.float().clone(),
               qkv=eng.qkv_all.float().clone(), ctx=eng.ctx_all.float().clone(), a=eng.a_all.float().clone(),
               m=eng.m_all.float().clone(), lse=eng.lse_all.clone(), stats=eng.stats.clone(), s_cls=eng.s_cls.float().clone(),
               logits=eng.logits.clone(), loss=eng.loss.clone())
    return out, m.config.tokens
base = dict(name="vit", task_type="reg", image_size=4096, patch_size=32, hidden_size=32, num_hidden_layers=3,
            num_attention_heads=2, stride_size=32, proj_fn="SW")
cases = [("baseline", base, 64), ("T33", dict(base, image_size=1024), 5), ("rope33", dict(base, image_size=1024, pos_encoding_type="rope"), 3),
         ("learned", dict(base, pos_encoding_type="learned", num_hidden_layers=1), 70), ("b200", dict(base, num_hidden_layers=2), 200)]
only = sys.argv[1:]
for name, mc, B in cases:
    if only and name not in only: continue
    cfg = {"model": mc, "loss": {"name": "mae"}, "data": {"param": "g"}}
    for train in (False, True):
        ref, T = run(cfg, B, False, train)
        got, _ = run(cfg, B, True, train)
        worst = {}
        for k in ref:
            a, b = got[k], ref[k]
            d = (a - b).abs()
            # split main rows / side row for row tensors
            worst[k] = float(d.max() / b.abs().max().clamp_min(1e-6))
        nbad = {k: int(((got[k] - ref[k]).abs() > 0).sum()) for k in ref}
        print(name, "train" if train else "eval", "T", T, {k: f"{v:.2e}" for k, v in worst.items()})
        print("   mismatching elements:", nbad)
        if T == 129:
            zz = (got["z"] - ref["z"]).abs().reshape(got["z"].shape[0], B, T, -1)
            print("   z diff main rows", float(zz[:, :, :128].max()), "side row", float(zz[:, :, 128].max()))
