"""Tensor-level wrappers over the C ABI, used by the modular (hook-friendly, eval / no-grad) path of
the module tree and by the per-op parity tests.  Every function launches vit_b200 kernels on the
current stream; none has a PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, BF16, F32


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _dt(owner) -> int:
    return BF16 if owner.precision == "bf16" else F32


def _act_dtype(owner) -> torch.dtype:
    return torch.bfloat16 if owner.precision == "bf16" else torch.float32


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("vit_b200 has no CPU path: tensors must live on a CUDA (sm_100a) device")


def _operand(owner, w: torch.Tensor) -> torch.Tensor:
    """GEMM operand copy of a weight: the arena's bf16 shadow view in bf16 mode, the master otherwise."""
    if owner.precision != "bf16":
        return w
    ar = owner._arena
    off = (w.data_ptr() - ar.data.data_ptr()) // 4
    if 0 <= off < ar.layout.n_total and ar.shadow is not None:
        if ar.shadow_stale():
            _lib.check(_lib.load().vitb200_cast_bf16(ar.data.data_ptr(), ar.shadow.data_ptr(), ar.layout.n_total,
                                                     _stream(w)), "cast")
            ar.mark_shadow_fresh()
        return ar.shadow[off:off + w.numel()].view(w.shape)
    return w.to(torch.bfloat16)


def linear(owner, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE):
    _require_cuda(x)
    lib = _lib.load()
    adt = _act_dtype(owner)
    w2 = weight.reshape(weight.shape[0], -1)
    K, N = w2.shape[1], w2.shape[0]
    xin = x.reshape(-1, K).to(adt).contiguous()
    M = xin.shape[0]
    wop = _operand(owner, weight).reshape(N, K)
    y = torch.empty(M, N, dtype=adt, device=x.device)
    y_act = torch.empty_like(y) if act == ACT_GELU else None
    _lib.check(lib.vitb200_linear_fwd(xin.data_ptr(), wop.data_ptr(), None if bias is None else bias.data_ptr(),
                                      y.data_ptr(), None if y_act is None else y_act.data_ptr(), M, N, K, act,
                                      _dt(owner), _stream(x)), "linear_fwd")
    out = y_act if act == ACT_GELU else y
    return out.view(*x.shape[:-1], N)


def gelu(owner, x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    adt = _act_dtype(owner)
    xin = x.to(adt).contiguous()
    y = torch.empty_like(xin)
    _lib.check(_lib.load().vitb200_gelu_fwd(xin.data_ptr(), y.data_ptr(), xin.numel(), _dt(owner), _stream(x)), "gelu")
    return y


def residual_add(owner, z: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    _require_cuda(z)
    zin = z.to(torch.float32).contiguous()
    d = delta.to(_act_dtype(owner)).contiguous()
    out = torch.empty_like(zin)
    _lib.check(_lib.load().vitb200_residual_add(zin.data_ptr(), d.data_ptr(), out.data_ptr(), zin.numel(), _dt(owner),
                                                _stream(z)), "residual_add")
    return out


def layer_norm(owner, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    _require_cuda(x)
    H = x.shape[-1]
    xin = x.to(torch.float32).reshape(-1, H).contiguous()
    M = xin.shape[0]
    u = torch.empty(M, H, dtype=_act_dtype(owner), device=x.device)
    stats = torch.empty(2, M, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().vitb200_add_ln_fwd(xin.data_ptr(), None, None, u.data_ptr(), stats[0].data_ptr(),
                                              stats[1].data_ptr(), weight.data_ptr(), bias.data_ptr(), M, H, 0,
                                              float(eps), 0.0, None, 0, _dt(owner), _stream(x)), "add_ln_fwd")
    return u.view(x.shape)


def embed(owner, x: torch.Tensor) -> torch.Tensor:
    """Eval-mode embeddings (no dropout): [B, L] -> [B, T, H] fp32."""
    _require_cuda(x)
    c = owner.config
    e = owner.vit.embeddings
    B = x.shape[0]
    xin = x.to(torch.float32).contiguous()
    z = torch.empty(B, c.tokens, c.hidden_size, dtype=torch.float32, device=x.device)
    w = e.patch_embeddings.projection.weight
    pos = e.position_embeddings
    _lib.check(_lib.load().vitb200_patch_embed_fwd(
        xin.data_ptr(), _operand(owner, w).data_ptr(), e.patch_embeddings.projection.bias.data_ptr(),
        e.cls_token.data_ptr(), None if pos is None else pos.data_ptr(), z.data_ptr(), B, c.image_size, c.patch_size,
        c.stride, c.num_patches, c.n_valid, c.hidden_size, 0.0, None, 0, _dt(owner), _stream(x)), "patch_embed_fwd")
    return z


def _rope_tables(owner, T: int, device):
    c = owner.config
    if c.pos_encoding_type != "rope":
        return None, None
    key = (T, str(device))
    cache = owner.__dict__.setdefault("_rope_cache", {})
    if key not in cache:
        d = c.head_dim
        inv_freq = 1.0 / (c.rope_base ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
        fr = torch.outer(torch.arange(T, dtype=torch.float32), inv_freq)
        cache[key] = (fr.cos().to(device).contiguous(), fr.sin().to(device).contiguous())
    return cache[key]


def attention(owner, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, want_probs: bool = False):
    """Eval-mode attention on separately projected q, k, v [B, T, H] -> (context [B,T,H], probs or None)."""
    _require_cuda(q)
    lib = _lib.load()
    c = owner.config
    B, T, H = q.shape
    adt = _act_dtype(owner)
    q, k, v = (t.to(adt).contiguous() for t in (q, k, v))
    ctx = torch.empty(B, T, H, dtype=adt, device=q.device)
    lse = torch.empty(B, c.num_attention_heads, T, dtype=torch.float32, device=q.device)
    cos, sin = _rope_tables(owner, T, q.device)
    scale = 1.0 / math.sqrt(c.head_dim)
    pc = None if cos is None else cos.data_ptr()
    ps = None if sin is None else sin.data_ptr()
    _lib.check(lib.vitb200_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), H, ctx.data_ptr(), lse.data_ptr(), pc, ps,
                                    B, T, c.num_attention_heads, c.head_dim, scale, 0.0, None, 0, _dt(owner),
                                    _stream(q)), "attn_fwd")
    probs = None
    if want_probs:
        probs = torch.empty(B, c.num_attention_heads, T, T, dtype=torch.float32, device=q.device)
        _lib.check(lib.vitb200_attn_probs(q.data_ptr(), k.data_ptr(), H, lse.data_ptr(), probs.data_ptr(), pc, ps, B, T,
                                          c.num_attention_heads, c.head_dim, scale, _dt(owner), _stream(q)),
                   "attn_probs")
    return ctx, probs


def head_loss(owner, s_cls: torch.Tensor, labels: Optional[torch.Tensor]):
    lib = _lib.load()
    c = owner.config
    head = owner.classifier if owner.task_type == "cls" else owner.regressor
    B, H = s_cls.shape
    s = s_cls.to(_act_dtype(owner)).contiguous()
    logits = torch.empty(B, c.num_labels, dtype=torch.float32, device=s.device)
    loss = torch.zeros(1, dtype=torch.float32, device=s.device)
    lab = None
    if labels is not None:
        lab = labels.reshape(-1).to(torch.int64 if owner._loss_kind == _lib.LOSS_CE else torch.float32).contiguous()
    _lib.check(lib.vitb200_head_loss_fwd(s.data_ptr(), _operand(owner, head.weight).data_ptr(), head.bias.data_ptr(),
                                         None if lab is None else lab.data_ptr(), logits.data_ptr(), loss.data_ptr(),
                                         B, H, c.num_labels, owner._loss_kind, _dt(owner), _stream(s)), "head_loss_fwd")
    return (loss[0] if labels is not None else None), logits


def modular_forward(model, x: torch.Tensor, labels: Optional[torch.Tensor], want_attn: bool):
    """Module-by-module eval forward: every inner module is *called*, so forward hooks fire."""
    probs = []
    handles = []
    if want_attn:
        for layer in model.vit.encoder.layer:
            handles.append(layer.attention.attention.register_forward_hook(lambda m, i, o: probs.append(o[1])))
    try:
        enc = model.vit.encoder(model.vit.embeddings(x.to(torch.float32)))
        last = model.vit.layernorm(enc.last_hidden_state)
    finally:
        for h in handles:
            h.remove()
    loss, logits = head_loss(model, last[:, 0, :], labels)
    return dict(loss=loss, logits=logits, hidden_states=enc.hidden_states, attentions=tuple(probs) if want_attn else None)
