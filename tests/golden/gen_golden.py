"""Generate golden fixtures by running the UNMODIFIED reference (`/root/reference/src/models`,
imported through `oracle/ref_shims.py`) on seeded synthetic inputs.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/gen_golden.py

Writes `tests/golden/<variant>.pt`, each a dict:
    config        the reference config dict the model was built from
    state_dict    the reference model's weights (float32)
    batch, x_seed, x_kind   how to regenerate the input with oracle.vit_oracle.synthetic_batch
    labels        the labels used
    eval          fp32, dropout off: loss, logits, hidden_states (sample 0 only), last_hidden[:, 0]
    grads         fp32, dropout off: d loss / d param for every param that received a grad
    bf16          the same under torch.autocast('cpu', bfloat16): loss, logits, grads
    train3        3 steps of clip_grad_norm_(0.5) + torch.optim.AdamW(lr=1e-3) with dropout off:
                  losses, grad norms, AdamW first moments and the final weights
    train3_bf16   the same 3 steps under torch.autocast('cpu', bfloat16) (Lightning precision='bf16-mixed')
"""
from __future__ import annotations

import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle.vit_oracle import synthetic_batch  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def base_cfg(**model_over):
    m = dict(name="vit", task_type="reg", image_size=4096, patch_size=32, hidden_size=32,
             num_hidden_layers=3, num_attention_heads=2, stride_size=32, proj_fn="SW")
    m.update(model_over)
    return {"model": m, "loss": {"name": "mae"}, "data": {"param": "log_g"}, "noise": {"noise_level": 0},
            "opt": {"type": "AdamW", "lr": 0.001}, "train": {"batch_size": 64}}


def variants():
    v = {}
    v["baseline"] = (base_cfg(), 4, "dummy")                       # configs/exp/att_clp/baseline.yaml
    v["config2"] = (base_cfg(num_hidden_layers=2), 4, "rand")      # configs/config.yaml model
    v["rope"] = (base_cfg(image_size=1024, pos_encoding_type="rope"), 3, "rand")
    v["learned"] = (base_cfg(image_size=1024, pos_encoding_type="learned"), 3, "rand")
    v["cnn"] = (base_cfg(image_size=1000, proj_fn="CNN", stride_size=24), 3, "rand")   # floor patch count
    v["stride8"] = (base_cfg(image_size=1024, stride_size=8), 2, "rand")               # overlapping windows
    v["pad48"] = (base_cfg(image_size=1000, patch_size=48, stride_size=48), 3, "rand")  # zero-pad tail window
    c = base_cfg(image_size=1024, task_type="cls", num_labels=3)
    v["cls"] = (c, 5, "rand")
    c = base_cfg(image_size=1024)
    c["loss"]["name"] = "l1"
    v["l1"] = (c, 4, "rand")
    c = base_cfg(image_size=2048, hidden_size=64, num_attention_heads=4, num_hidden_layers=2)
    c["data"]["param"] = "t_eff,log_g"
    v["h64multi"] = (c, 3, "rand")
    c = base_cfg(image_size=2048, hidden_size=128, num_attention_heads=2, num_hidden_layers=1,
                 pos_encoding_type="rope", stride_size=16)
    v["h128d64rope"] = (c, 2, "rand")
    # head_dim 4 (hidden 32 over 8 heads, the small end of the head-count sweep), with RoPE on 2 rotation pairs
    v["heads8d4rope"] = (base_cfg(image_size=1024, num_attention_heads=8, pos_encoding_type="rope"), 3, "rand")
    # BASELINE config 4, the long-sequence sweep at its full spectrum length: x4 tokens (T = 510) and x16 (T = 2034)
    v["long510"] = (base_cfg(stride_size=8), 2, "rand")
    v["long2034"] = (base_cfg(stride_size=2, num_hidden_layers=2), 1, "rand")
    # input preprocessors (src/models/builder.py:43-133) on seeded covariance statistics of 256-pixel spectra
    for name, warm in (
        ("pre_zca_full", dict(preprocessor="zca", freeze_epochs=-1)),                           # frozen buffers
        ("pre_zca_r32", dict(preprocessor="zca", r=32, shrinkage=0.1, freeze_epochs=0)),        # trainable, low rank
        ("pre_pca_r128", dict(preprocessor="pca", r=128, freeze_epochs=2)),                     # image_size 256 -> 128
        ("pre_attn_r64", dict(preprocessor="attention", r=64)),                                 # q_lin only, trainable
    ):
        c = base_cfg(image_size=256, patch_size=16, stride_size=16)
        c["warmup"] = warm
        v[name] = (c, 4, "rand")
    return v


def cov_stats(dim: int, seed: int = 11) -> dict:
    """Seeded stand-in for the offline covariance statistics file (src/prepca/precompute_pca.py writes mean, cov,
    eigvals, eigvecs sorted by descending eigenvalue): smooth correlated 'spectra' so that the spectrum decays."""
    g = torch.Generator().manual_seed(seed)
    n = 4 * dim
    basis = torch.randn(24, dim, generator=g).cumsum(1) / dim ** 0.5
    X = torch.randn(n, 24, generator=g) @ basis + 0.05 * torch.randn(n, dim, generator=g) + 0.5
    mean = X.mean(0)
    Xc = (X - mean).double()
    cov = (Xc.t() @ Xc / (n - 1))
    lam, V = torch.linalg.eigh(cov)
    lam, V = lam.flip(0).clamp_min(0).float(), V.flip(1).float().contiguous()
    return dict(mean=mean.float(), cov=cov.float(), eigvals=lam, eigvecs=V)


def run_variant(name, cfg, batch, kind):
    torch.manual_seed(42)
    stats = None
    if cfg.get("warmup"):
        import tempfile

        stats = cov_stats(cfg["model"]["image_size"])
        tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
        torch.save(stats, tmp.name)
        cfg = copy.deepcopy(cfg)
        cfg["warmup"]["cov_path"] = tmp.name
    model = ref_shims.reference_get_model(cfg)
    if stats is not None:
        os.unlink(cfg["warmup"]["cov_path"])
        cfg["warmup"].pop("cov_path")
    model.eval()  # dropout off; grads still flow
    spec_len = cfg["model"]["image_size"]
    x, y = synthetic_batch(batch, spec_len, seed=7, kind=kind)
    if cfg["model"]["task_type"] == "cls":
        y = torch.randint(0, cfg["model"]["num_labels"], (batch,), generator=torch.Generator().manual_seed(3))
    elif model.config.num_labels > 1:
        y = torch.rand(batch, model.config.num_labels, generator=torch.Generator().manual_seed(3))
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    preprocessed = None
    if model.preprocessor is not None:   # with the INITIAL weights (the train3 section below updates a trainable matrix)
        with torch.no_grad():
            preprocessed = model.preprocessor(x).clone()

    out = model(x, labels=y, output_hidden_states=True)
    model.zero_grad()
    out.loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    ev = dict(loss=out.loss.detach().clone(), logits=out.logits.detach().clone(),
              hidden_states=[h[0].detach().clone() for h in out.hidden_states])
    with torch.no_grad():  # model.vit takes preprocessed pixels (src/viz/viz_callback.py:574-579)
        last = model.vit(x if model.preprocessor is None else model.preprocessor(x)).last_hidden_state
    ev["last_hidden_cls"] = last[:, 0].clone()

    model.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        outb = model(x, labels=y)
    outb.loss.backward()
    bf = dict(loss=outb.loss.detach().float().clone(), logits=outb.logits.detach().float().clone(),
              grads={k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})

    # 3 training steps, Lightning semantics restated (basemodule.py:244, optimizer.py:108)
    model.zero_grad()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0)
    losses, norms = [], []
    for _ in range(3):
        opt.zero_grad()
        loss = model(x, labels=y).loss
        loss.backward()
        norms.append(torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5).detach().clone())
        opt.step()
        losses.append(loss.detach().clone())
    names = [k for k, _ in model.named_parameters()]

    def moments(o):
        return {names[i]: st["exp_avg"].detach().clone() for i, st in o.state_dict()["state"].items()}

    tr = dict(losses=torch.stack(losses), grad_norms=torch.stack(norms), exp_avg=moments(opt),
              state_dict={k: v.detach().clone() for k, v in model.state_dict().items()})

    # the same 3 steps under Lightning's precision='bf16-mixed' (= torch.autocast(bf16) around forward), from the initial
    # weights again.  Adam normalises every update to ~lr, so post-step WEIGHTS of elements whose gradient sits at the bf16
    # noise level are not comparable across implementations; the first moments (clipped gradients, EMA) are.
    model.load_state_dict(sd0)
    model.zero_grad()
    optb = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0)
    lossesb, normsb = [], []
    for _ in range(3):
        optb.zero_grad()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            lossb = model(x, labels=y).loss
        lossb.backward()
        normsb.append(torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5).detach().clone())
        optb.step()
        lossesb.append(lossb.detach().float().clone())
    trb = dict(losses=torch.stack(lossesb), grad_norms=torch.stack(normsb), exp_avg=moments(optb),
               state_dict={k: v.detach().clone() for k, v in model.state_dict().items()})

    fix = dict(config=copy.deepcopy(cfg), state_dict=sd0, batch=batch, x_seed=7, x_kind=kind, labels=y,
               eval=ev, grads=grads, bf16=bf, train3=tr, train3_bf16=trb, model_name=model.name, loss_name=model.loss_name,
               torch_version=torch.__version__)
    if stats is not None:   # what the builder needs to rebuild the preprocessor (the covariance itself is not used)
        fix["stats"] = {k: stats[k].clone() for k in ("mean", "eigvals", "eigvecs")}
        fix["param_names"] = [k for k, _ in model.named_parameters()]
        fix["eval"]["preprocessed"] = preprocessed
    path = os.path.join(OUT, f"{name}.pt")
    torch.save(fix, path)
    print(f"{name}: loss={float(ev['loss']):.6f} bf16={float(bf['loss']):.6f} name={model.name} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (cfg, batch, kind) in variants().items():
        if only and name not in only:
            continue
        run_variant(name, cfg, batch, kind)
