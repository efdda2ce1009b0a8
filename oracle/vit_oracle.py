"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference hot path.

This file is the checker, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  Nothing under
`vit_b200/` imports it, and the product path raises if the CUDA library is missing.

What it restates (plain torch tensor ops on CPU, fp32 or fp64, HF-free, Lightning-free):

  * tokenizer            /root/reference/src/models/tokenization.py:31-69
  * embeddings           /root/reference/src/models/embedding.py:79-100
  * RoPE                 /root/reference/src/models/rope.py:37-98
                         /root/reference/src/models/vit_with_rope.py:43-84
  * encoder layer        transformers/models/vit/modeling_vit.py (third-party, the reference pins
                         transformers==4.56.0 in requirements.txt:57; 5.5.0 is installed here):
                         eager attention :171-196, ViTSelfAttention :199-251,
                         ViTSelfOutput/ViTIntermediate/ViTOutput :254-312, ViTLayer :315-346,
                         final LayerNorm :454-455
  * head + loss          /root/reference/src/models/specvit.py:68-94
  * preprocessor         /root/reference/src/models/layers.py:62-63, preprocessor.py:107-108
  * config -> shapes     /root/reference/src/models/builder.py:200-258
  * init                 modeling_vit.py:384-398 + embedding.py:47,65-67
  * train step           Lightning semantics restated (src/basemodule.py:237-251 ->
                         clip_grad_norm_(0.5)) + AdamW (src/opt/optimizer.py:108)
  * synthetic input      /root/reference/src/utils.py:131-139 (make_dummy_spectra)

PARITY PINNING: the reference ships no tests / golden vectors, so this
oracle is pinned against outputs of the *reference itself run in the build container*
(`tests/golden/gen_golden.py` imports the unmodified `/root/reference/src/models` through
`oracle/ref_shims.py` and writes `tests/golden/*.pt`); `tests/test_oracle_golden.py` checks the
oracle against those fixtures on every run, and against the live reference when it is present.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# config -> shapes (builder.py:200-258, embedding.py:24-27, tokenization.py:34-41,56-64)
# ----------------------------------------------------------------------------------------------
@dataclass
class VitSpec:
    image_size: int = 4096
    patch_size: int = 32
    stride: int = 32
    hidden: int = 32
    layers: int = 3
    heads: int = 2
    proj_fn: str = "SW"            # 'SW' | 'C1D' | 'CNN'
    pos_type: Optional[str] = None  # None | 'none' | 'learned' | 'rope'
    rope_base: float = 10000.0
    max_pos: int = 512
    num_labels: int = 1
    task: str = "reg"              # 'reg' | 'cls'
    loss_kind: str = "mse"         # 'mse' | 'l1' | 'ce'
    p_hidden: float = 0.1
    p_attn: float = 0.1
    eps: float = 1e-12
    intermediate: int = field(init=False)
    num_patches: int = field(init=False)

    def __post_init__(self):
        self.intermediate = 4 * self.hidden
        if self.proj_fn == "SW":
            self.num_patches = math.ceil((self.image_size - self.patch_size) / self.stride) + 1
        elif self.proj_fn in ("C1D", "CNN"):
            self.num_patches = (self.image_size - self.patch_size) // self.stride + 1
        else:
            raise ValueError(f"Unsupported proj_fn '{self.proj_fn}'")

    @property
    def tokens(self) -> int:
        return self.num_patches + 1

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def head_name(self) -> str:
        return "classifier" if self.task == "cls" else "regressor"


def spec_from_config(config: dict) -> VitSpec:
    """builder.py:200-258 (+ specvit.py:46-55 for the loss selection quirk: 'mae' -> MSE)."""
    m = config["model"]
    d = config.get("data", {}) or {}
    task = (m.get("task_type") or m.get("task") or "cls").lower()
    if task in ("reg", "regression"):
        p = d.get("param", None)
        num_labels = 1
        if isinstance(p, str) and len(p) > 0:
            plist = [s.strip() for s in p.split(",") if s.strip()]
            if len(plist) >= 1:
                num_labels = len(plist)
        elif isinstance(p, (list, tuple)) and len(p) > 0:
            num_labels = len(p)
    else:
        num_labels = int(m.get("num_labels", 1) or 1)
    task_type = m["task_type"]
    if task_type == "cls":
        loss_kind = "ce"
    elif task_type == "reg":
        loss_name = (config.get("loss", {}) or {}).get("name", None) or "l2"
        loss_kind = "l1" if "l1" in loss_name.lower() else "mse"
    else:
        raise ValueError(f"Unsupported task_type '{task_type}'")
    stride_size = m.get("stride_size", None)
    stride = stride_size if stride_size and stride_size > 0 else int(m.get("stride_ratio", 1) * m["patch_size"])
    return VitSpec(
        image_size=m["image_size"], patch_size=m["patch_size"], stride=stride,
        hidden=m["hidden_size"], layers=m["num_hidden_layers"], heads=m["num_attention_heads"],
        proj_fn=m["proj_fn"], pos_type=m.get("pos_encoding_type", None),
        rope_base=m.get("rope_base", 10000.0), max_pos=m.get("max_position_embeddings", 512),
        num_labels=num_labels, task=task_type, loss_kind=loss_kind,
    )


# ----------------------------------------------------------------------------------------------
# parameters (names = the reference's state_dict keys)
# ----------------------------------------------------------------------------------------------
def param_shapes(spec: VitSpec) -> Dict[str, tuple]:
    H, I, P = spec.hidden, spec.intermediate, spec.patch_size
    s: Dict[str, tuple] = {}
    s["vit.embeddings.cls_token"] = (1, 1, H)
    if spec.pos_type == "learned":
        s["vit.embeddings.position_embeddings"] = (1, spec.tokens, H)
    if spec.proj_fn == "SW":
        s["vit.embeddings.patch_embeddings.projection.weight"] = (H, P)
    else:
        s["vit.embeddings.patch_embeddings.projection.weight"] = (H, 1, P)
    s["vit.embeddings.patch_embeddings.projection.bias"] = (H,)
    for i in range(spec.layers):
        pre = f"vit.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[pre + f"attention.attention.{n}.weight"] = (H, H)
            s[pre + f"attention.attention.{n}.bias"] = (H,)
        s[pre + "attention.output.dense.weight"] = (H, H)
        s[pre + "attention.output.dense.bias"] = (H,)
        s[pre + "intermediate.dense.weight"] = (I, H)
        s[pre + "intermediate.dense.bias"] = (I,)
        s[pre + "output.dense.weight"] = (H, I)
        s[pre + "output.dense.bias"] = (H,)
        s[pre + "layernorm_before.weight"] = (H,)
        s[pre + "layernorm_before.bias"] = (H,)
        s[pre + "layernorm_after.weight"] = (H,)
        s[pre + "layernorm_after.bias"] = (H,)
    s["vit.layernorm.weight"] = (H,)
    s["vit.layernorm.bias"] = (H,)
    s["vit.pooler.dense.weight"] = (H, H)
    s["vit.pooler.dense.bias"] = (H,)
    s[spec.head_name + ".weight"] = (spec.num_labels, H)
    s[spec.head_name + ".bias"] = (spec.num_labels,)
    return s


def init_params(spec: VitSpec, seed: int = 42, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Same *distribution* as the reference init (modeling_vit.py:384-393: Linear/Conv weights
    trunc_normal(0, 0.02), biases 0, LN weight 1 / bias 0; embedding.py:47,65-67: cls/pos randn).
    Not the same random stream: parity tests copy weights from a state_dict instead."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(spec).items():
        if name.endswith("cls_token") or name.endswith("position_embeddings"):
            t = torch.randn(shape, generator=g)
        elif "layernorm" in name:
            t = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        elif name.endswith("bias"):
            t = torch.zeros(shape)
        else:
            t = torch.empty(shape)
            t.normal_(0.0, 0.02, generator=g).clamp_(-2.0, 2.0)
        out[name] = t.to(dtype)
    return out


# ----------------------------------------------------------------------------------------------
# input preprocessor (src/models/preprocessor.py:12-90, src/models/builder.py:43-133, src/models/attention.py:58-72)
# ----------------------------------------------------------------------------------------------
def zca_matrix(eigvecs, eigvals, eps=1e-5, r=None, shrinkage=0.0):
    """preprocessor.py:12-72.  Shrinkage towards the mean eigenvalue first; full rank: V diag(1/sqrt(lam+eps)) V^T;
    rank r: signal subspace whitened, orthogonal complement scaled by 1/sqrt(lam0+eps), lam0 = median of the tail
    eigenvalues (the r-th one if there is no tail), floored at 1e-3 * mean(lam[:r])."""
    lam = eigvals if shrinkage <= 0.0 else (1.0 - shrinkage) * eigvals + shrinkage * eigvals.mean()
    if r is None:
        return eigvecs @ torch.diag(1.0 / torch.sqrt(lam + eps)) @ eigvecs.t()
    Vr = eigvecs[:, :r]
    tail = lam[r:]
    lam0 = tail.median() if tail.numel() > 0 else lam[r - 1]
    lam0 = torch.clamp(lam0, min=1e-3 * lam[:r].mean())
    s_perp = 1.0 / torch.sqrt(lam0 + eps)
    ident = torch.eye(eigvecs.shape[0], dtype=eigvecs.dtype)
    return (Vr * torch.rsqrt(lam[:r] + eps)) @ Vr.t() + s_perp * (ident - Vr @ Vr.t())


def preprocessor_tensors(warmup: dict, stats: dict) -> dict:
    """warmup config + covariance statistics -> dict(tensors keyed by state_dict name, out_dim, prefix, frozen) following
    builder.py:43-133: 'zca' / 'pca' give `preprocessor.linear.{weight,bias}` with bias = -mean @ P^T (unless
    warmup.bias is false), frozen iff freeze_epochs != 0 (builder.py:168); 'attention' gives q_lin / k_lin prefilled with
    the (optionally 1/sqrt(lam+eps)-scaled) leading eigenvectors, always Parameters (but see the quirk below)."""
    kind = warmup["preprocessor"]
    V, mean, r = stats["eigvecs"], stats.get("mean"), warmup.get("r")
    fe = warmup.get("freeze_epochs", 0)
    fz = "perm" if fe == -1 else str(fe)
    if kind in ("zca", "pca"):
        use_bias = warmup.get("bias", True)
        if kind == "zca":
            sh = warmup.get("shrinkage", 0.0)
            P = zca_matrix(V, stats["eigvals"], eps=warmup.get("eps", 1e-5), r=r, shrinkage=sh)
            prefix = ("ZCA" if r is None else f"ZCA{r}") + f"_fz{fz}" + (f"_s{int(sh * 10)}" if sh > 0 else "")
        else:
            P = V.t() if r is None else V[:, :r].t()      # preprocessor.py:75-90
            prefix = ("PCA" if r is None else f"PCA{r}") + f"_fz{fz}"
        prefix += "" if use_bias else "_nobias"
        t = {"preprocessor.linear.weight": P.to(torch.float32)}
        if use_bias and mean is not None:
            t["preprocessor.linear.bias"] = (-mean @ P.t()).to(torch.float32)
        return dict(tensors=t, out_dim=P.shape[0], prefix=prefix, frozen=fe != 0)
    if kind == "attention":
        D = V.shape[0]
        rr = r if r is not None else V.shape[1]
        basis = V[:, :rr].t().contiguous()
        lam = stats.get("eigvals")
        scaled = warmup.get("scale_by_eigvals", True) and lam is not None
        if scaled:
            basis = basis * torch.rsqrt(lam[:rr] + warmup.get("eps", 1e-5)).unsqueeze(1)
        if rr < D:
            W = basis
        else:
            W = torch.zeros(D, D)
            W[:rr] = basis
        prefix = f"Attn{r if r else 'Full'}" + ("_scaled" if scaled else "") + f"_fz{fz}"
        # QUIRK (pinned by the pre_attn_r64 fixture): the prefill does not survive model construction.  MyViT.__init__
        # ends with self.init_weights() (specvit.py:57) and HF's _init_weights re-draws every nn.Linear below MyViT --
        # q_lin / k_lin / v_lin included -- from trunc_normal(0, initializer_range=0.02).  So the state_dict holds random
        # matrices (`reinit`), and `prefill` is only what attention.py:58-72 computed before that.
        return dict(tensors={}, prefill={"preprocessor.q_lin.weight": W.clone(), "preprocessor.k_lin.weight": W.clone()},
                    reinit=("preprocessor.q_lin.weight", "preprocessor.k_lin.weight", "preprocessor.v_lin.weight"),
                    out_dim=(r if r is not None else D), prefix=prefix, frozen=False)
    raise ValueError(f"Unknown preprocessor type: '{kind}'")


# ----------------------------------------------------------------------------------------------
# synthetic inputs (src/utils.py:131-139; datasets clip flux >= 0, src/dataloader/base.py:236)
# ----------------------------------------------------------------------------------------------
def make_dummy_spectra(n: int = 512, length: int = 4096, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, length, generator=g) * 0.05
    x = torch.arange(length)
    for c in (300, 600, 1200, 1600):
        base -= 0.8 * torch.exp(-0.5 * ((x - c) / 6.0) ** 2)[None, :]
    return base


def synthetic_batch(batch: int, length: int = 4096, seed: int = 0, kind: str = "dummy"):
    """(flux, labels): flux = make_dummy_spectra(...).clip(min=0) or U[0,1]; labels U[0,1]."""
    if kind == "dummy":
        x = make_dummy_spectra(batch, length, seed).clamp_(min=0)
    else:
        x = torch.rand(batch, length, generator=torch.Generator().manual_seed(seed))
    y = torch.rand(batch, generator=torch.Generator().manual_seed(seed + 1))
    return x, y


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def _dropout(x, p, train, masks, key):
    """masks: optional dict key -> {0,1} tensor (same shape as x) supplied by the caller so that
    a CUDA run's Philox masks can be replayed exactly; otherwise torch's own RNG."""
    if not train or p <= 0.0:
        return x
    if masks is not None and key in masks:
        return x * masks[key].to(x.dtype) / (1.0 - p)
    return F.dropout(x, p=p, training=True)


def tokenize(x: torch.Tensor, spec: VitSpec, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tokenization.py:43-50 (SW: unfold, zero-pad missing tail windows, Linear) and :66-69
    (CNN: Conv1d(1,H,P,stride) == the same windows, floor patch count, no padding)."""
    B = x.shape[0]
    P, S = spec.patch_size, spec.stride
    patches = x.unfold(1, P, S)
    if spec.proj_fn == "SW":
        if patches.size(1) < spec.num_patches:
            pad = torch.zeros(B, spec.num_patches - patches.size(1), P, dtype=x.dtype)
            patches = torch.cat([patches, pad], dim=1)
        Wm = W
    else:
        patches = patches[:, : spec.num_patches]
        Wm = W[:, 0, :]
    return F.linear(patches, Wm, b)


def rope_tables(spec: VitSpec, T: int, dtype=torch.float32):
    """rope.py:37,44-57: inv_freq = base^(-2j/d); emb = [freqs|freqs]; cos/sin cached."""
    d = spec.head_dim
    inv_freq = 1.0 / (spec.rope_base ** (torch.arange(0, d, 2).float() / d))
    t = torch.arange(T).type_as(inv_freq)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """rope.py:60-98: x*cos + rotate_half(x)*sin, rotate_half([x1|x2]) = [-x2|x1]; x [B,a,T,d]."""
    x1, x2 = x.chunk(2, dim=-1)
    rot = torch.cat([-x2, x1], dim=-1)
    return x * cos[None, None] + rot * sin[None, None]


def forward(
    params: Dict[str, torch.Tensor],
    x: torch.Tensor,
    spec: VitSpec,
    labels: Optional[torch.Tensor] = None,
    train: bool = False,
    masks: Optional[Dict[str, torch.Tensor]] = None,
    preproc: Optional[tuple] = None,
    keep: bool = False,
):
    """Returns dict(loss, logits, last_hidden (post final LN), hidden_states [emb, l0, l1, ...]
    (pre final LN, as HF's output_hidden_states), attn_probs per layer when keep=True).

    Run it under `torch.autocast('cpu', dtype=torch.bfloat16)` to get the reference's
    `bf16-mixed` semantics: the ops below are the same torch ops the
    reference/HF code calls, so autocast makes the same per-op dtype choices."""
    H, a, d = spec.hidden, spec.heads, spec.head_dim
    p = params
    if preproc is None:  # preprocessor tensors carried in `params` under their state_dict names
        if "preprocessor.linear.weight" in p:      # LinearPreprocessor (preprocessor.py:107-108)
            preproc = (p["preprocessor.linear.weight"], p.get("preprocessor.linear.bias"))
        elif "preprocessor.q_lin.weight" in p:     # PrefilledAttention, 2-D input: q_lin only (attention.py:81-82)
            preproc = (p["preprocessor.q_lin.weight"], None)
    if preproc is not None:  # layers.py:62-63
        x = F.linear(x, preproc[0], preproc[1])
    B = x.shape[0]
    # --- embeddings (embedding.py:85-100) ---
    tok = tokenize(x, spec, p["vit.embeddings.patch_embeddings.projection.weight"],
                   p["vit.embeddings.patch_embeddings.projection.bias"])
    cls = p["vit.embeddings.cls_token"].expand(B, -1, -1)
    z = torch.cat((cls, tok), dim=1)
    if spec.pos_type == "learned":
        z = z + p["vit.embeddings.position_embeddings"]
    z = _dropout(z, spec.p_hidden, train, masks, "emb")
    T = z.shape[1]
    hidden_states = [z]
    probs_all = []
    cos = sin = None
    if spec.pos_type == "rope":
        cos, sin = rope_tables(spec, max(T, spec.max_pos))
        cos, sin = cos[:T], sin[:T]
    # --- layers (modeling_vit.py:328-346) ---
    for i in range(spec.layers):
        pre = f"vit.encoder.layer.{i}."
        u = F.layer_norm(z, (H,), p[pre + "layernorm_before.weight"], p[pre + "layernorm_before.bias"], spec.eps)
        q = F.linear(u, p[pre + "attention.attention.query.weight"], p[pre + "attention.attention.query.bias"])
        k = F.linear(u, p[pre + "attention.attention.key.weight"], p[pre + "attention.attention.key.bias"])
        v = F.linear(u, p[pre + "attention.attention.value.weight"], p[pre + "attention.attention.value.bias"])
        q = q.view(B, T, a, d).transpose(1, 2)
        k = k.view(B, T, a, d).transpose(1, 2)
        v = v.view(B, T, a, d).transpose(1, 2)
        if spec.pos_type == "rope":  # vit_with_rope.py:59-67
            q = apply_rope(q, cos, sin)
            k = apply_rope(k, cos, sin)
            s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
        else:                        # modeling_vit.py:171-196 (scaling = d^-0.5)
            s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
        pr = F.softmax(s, dim=-1)
        if keep:
            probs_all.append(pr)
        prd = _dropout(pr, spec.p_attn, train, masks, f"attn{i}")
        c = torch.matmul(prd, v).transpose(1, 2).contiguous().view(B, T, H)
        ao = F.linear(c, p[pre + "attention.output.dense.weight"], p[pre + "attention.output.dense.bias"])
        ao = _dropout(ao, spec.p_hidden, train, masks, f"proj{i}")
        h = ao + z
        u2 = F.layer_norm(h, (H,), p[pre + "layernorm_after.weight"], p[pre + "layernorm_after.bias"], spec.eps)
        m = F.gelu(F.linear(u2, p[pre + "intermediate.dense.weight"], p[pre + "intermediate.dense.bias"]))
        mo = F.linear(m, p[pre + "output.dense.weight"], p[pre + "output.dense.bias"])
        mo = _dropout(mo, spec.p_hidden, train, masks, f"mlp{i}")
        z = mo + h
        hidden_states.append(z)
    # --- tail (modeling_vit.py:454-455; specvit.py:78-89). The pooler (modeling_vit.py:456) is
    # computed and discarded by the reference; it has no effect on any output, so it is omitted.
    sfin = F.layer_norm(z, (H,), p["vit.layernorm.weight"], p["vit.layernorm.bias"], spec.eps)
    logits = F.linear(sfin[:, 0, :], p[spec.head_name + ".weight"], p[spec.head_name + ".bias"])
    loss = None
    if labels is not None:
        if spec.task == "cls":
            loss = F.cross_entropy(logits.view(-1, spec.num_labels), labels.view(-1))
        elif spec.loss_kind == "l1":
            loss = F.l1_loss(logits.view(-1), labels.view(-1).float())
        else:
            loss = F.mse_loss(logits.view(-1), labels.view(-1).float())
    out = dict(loss=loss, logits=logits, last_hidden=sfin, hidden_states=hidden_states)
    if keep:
        out["attn_probs"] = probs_all
    return out


# ----------------------------------------------------------------------------------------------
# train step: zero_grad -> fwd -> bwd -> clip_grad_norm_(0.5) -> AdamW
# ----------------------------------------------------------------------------------------------
def clip_grad_norm(grads: Dict[str, torch.Tensor], max_norm: float = 0.5) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ restated: L2 over all grads, coef = max/(norm+1e-6) clamped to 1."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(torch.float32)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return total


def adamw_update(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.AdamW single-tensor math (torch/optim/adam.py _single_tensor_adam, decoupled wd)."""
    p.mul_(1.0 - lr * wd)
    m.lerp_(g, 1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


class OracleTrainer:
    """Lightning-free restatement of the reference training step on CPU (the CPU baseline)."""

    def __init__(self, spec: VitSpec, params: Dict[str, torch.Tensor], lr=1e-3, wd=0.0, clip=0.5,
                 autocast_bf16: bool = False, frozen=()):
        """frozen: name prefixes of tensors that are buffers, not parameters (a frozen preprocessor, layers.py:21-24)."""
        self.spec = spec
        self.params = {k: v.clone().requires_grad_(not any(k.startswith(f) for f in frozen)) for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.steps = {k: 0 for k in params}
        self.lr, self.wd, self.clip = lr, wd, clip
        self.autocast_bf16 = autocast_bf16
        self.last_grad_norm = None

    def loss_and_grads(self, x, y, train=True, masks=None):
        for t in self.params.values():
            t.grad = None
        if self.autocast_bf16:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                out = forward(self.params, x, self.spec, labels=y, train=train, masks=masks)
        else:
            out = forward(self.params, x, self.spec, labels=y, train=train, masks=masks)
        out["loss"].backward()
        grads = {k: t.grad for k, t in self.params.items() if t.grad is not None}
        return out, grads

    def step(self, x, y, train=True, masks=None) -> float:
        out, grads = self.loss_and_grads(x, y, train=train, masks=masks)
        with torch.no_grad():
            if self.clip is not None and self.clip > 0:
                self.last_grad_norm = clip_grad_norm(grads, self.clip)
            for k, g in grads.items():  # params without a grad (the pooler) are skipped, as torch does
                self.steps[k] += 1
                adamw_update(self.params[k], g, self.m[k], self.v[k], self.steps[k], lr=self.lr, wd=self.wd)
        return float(out["loss"].detach())


# FLOP model used by bench.py (SURVEY.md section 8d / BASELINE.md section 4)
def flops_per_sample(spec: VitSpec):
    Np, P, H, L, T = spec.num_patches, spec.patch_size, spec.hidden, spec.layers, spec.tokens
    fwd = 2 * Np * P * H + L * (24 * T * H * H + 4 * T * T * H) + 2 * H * spec.num_labels
    step = 3 * (2 * Np * P * H + 24 * L * T * H * H + 2 * H * spec.num_labels) + 3.5 * L * 4 * T * T * H
    return fwd, step
