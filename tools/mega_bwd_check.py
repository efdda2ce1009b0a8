# compare the whole-network backward with the per-op fused backward (same forward, same dropout masks)
import os, sys, copy, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from vit_b200 import get_model
from oracle import vit_oracle as vo
dev = torch.device("cuda:0")
def run(cfg, B, mega_bwd, train, seed=3, cls_only=False):
    os.environ["VITB200_MEGA_BWD"] = "1" if mega_bwd else "0"
    os.environ["VITB200_MEGA"] = "1" if mega_bwd else "0"   # reference = the per-op programs in both directions
    torch.manual_seed(seed)
    m = get_model(copy.deepcopy(cfg), precision="bf16-mixed", device=dev)
    m.train(train)
    eng = m._engine(B)
    x, y = vo.synthetic_batch(B, cfg["model"]["image_size"], seed=5, kind="rand")
    if cfg["model"].get("task_type") == "cls":
        y = torch.randint(0, cfg["model"]["num_labels"], (B,))
    m._stage_raw(eng, x.to(dev), y.to(dev))
    fh = eng.can_fuse_head
    eng.cls_only = cls_only
    eng.forward(train=train, with_labels=True, head_bwd=fh)
    eng.backward(train=train, skip_reduce=False, skip_head=fh)
    torch.cuda.synchronize()
    assert eng.mega_bwd == mega_bwd, (eng.mega_bwd, mega_bwd)
    loss = float(eng.loss[0])
    lay = eng.arena.layout
    g = {k: eng.arena.grad[e.offset:e.offset + e.numel].clone() for k, e in lay.entries.items() if e.offset < lay.n_opt}
    return g, eng.dzA.clone(), m.config.tokens, loss
base = dict(name="vit", task_type="reg", image_size=4096, patch_size=32, hidden_size=32, num_hidden_layers=3,
            num_attention_heads=2, stride_size=32, proj_fn="SW")
cases = [("baseline", base, 64), ("T33", dict(base, image_size=1024), 5), ("rope33", dict(base, image_size=1024, pos_encoding_type="rope"), 3),
         ("learned", dict(base, pos_encoding_type="learned", num_hidden_layers=1), 70), ("l2b16", dict(base, num_hidden_layers=2), 16),
         ("cls", dict(base, image_size=1024, task_type="cls", num_labels=3), 5), ("rope129", dict(base, pos_encoding_type="rope", num_hidden_layers=2), 4)]
only = sys.argv[1:]
for name, mc, B in cases:
    if only and name not in only: continue
    cfg = {"model": mc, "loss": {"name": "mae"}, "data": {"param": "g"}}
    for train in (False, True):
      for co in (False, True):
        ref, dz_ref, T, l_ref = run(cfg, B, False, train)
        got, dz_got, _, l_got = run(cfg, B, True, train, cls_only=co)
        gmax = max(float(v.abs().max()) for v in ref.values())
        worst = []
        for k in ref:
            d = float((got[k] - ref[k]).abs().max()) / max(float(ref[k].abs().max()), 1e-3 * gmax)
            worst.append((d, k))
        worst.sort(reverse=True)
        dzerr = float((dz_got - dz_ref).abs().max() / dz_ref.abs().max().clamp_min(1e-12))
        print(name, "train" if train else "eval", "cls_only" if co else "full", "loss", f"{l_ref:.6f} {l_got:.6f}", "T", T, "B", B, "worst:", [(f"{d:.2e}", k.replace("vit.encoder.layer.", "L").replace("attention", "att")) for d, k in worst[:4]], "dz0 err", f"{dzerr:.2e}")
