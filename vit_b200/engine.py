"""The step engine: pre-allocated activations + the fixed kernel sequence of one ViT
forward / backward / clip+AdamW step, expressed as C-ABI calls on an explicit stream so that the whole
step can be captured in one CUDA graph (the configured shape is launch-latency bound).

Call sequence restated from the reference:
  forward   src/models/embedding.py:79-100 -> HF ViTLayer :328-346 x L -> HF :454-455 -> specvit.py:78-89
  backward  autograd of the above, written out explicitly per fused region
  update    clip_grad_norm_(0.5) (src/basemodule.py:244) + AdamW (src/opt/optimizer.py:108)
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Callable, List, Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, BF16, F32, SITE_EMB, site_attn, site_mlp, site_proj
from .arena import ParamArena

ROWS_SLOT = 99   # input slot of the device-resident dataset mode (ViTEngine.bind_rows)
HOST_SLOTS = 8   # regular input slots 0 .. 7 (TrainStep.fit_host: two groups of four steps)


def _normalize_precision(precision) -> str:
    p = str(precision).lower()
    if p in ("32", "32-true", "fp32", "float32"):
        return "fp32"
    if p in ("bf16", "bf16-mixed", "bfloat16", "bf16-true"):
        return "bf16"
    raise ValueError(f"vit_b200: unsupported precision '{precision}' (use '32' or 'bf16-mixed')")


class ViTEngine:
    """Owns every activation / scratch buffer for a fixed batch size and runs the kernel programs."""

    def __init__(self, cfg, arena: ParamArena, batch: int, precision: str, loss_kind: int, seed: int = 0,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_norm: float = 0.5, grad_scale: float = 1.0):
        if arena.data.device.type != "cuda":
            raise RuntimeError("vit_b200 kernels are CUDA-only (sm_100a); move the model to a CUDA device")
        self.lib = _lib.load()
        self.cfg, self.arena, self.B = cfg, arena, int(batch)
        self.precision = _normalize_precision(precision)
        self.dt = BF16 if self.precision == "bf16" else F32
        self.act_dtype = torch.bfloat16 if self.dt == BF16 else torch.float32
        if self.dt == BF16 and arena.shadow is None:
            raise RuntimeError("bf16 engine needs an arena with a bf16 shadow")
        self.loss_kind = loss_kind
        dev = arena.data.device
        self.device = dev
        _lib.check(self.lib.vitb200_init(dev.index if dev.index is not None else torch.cuda.current_device()), "init")
        c = cfg
        B, T, H, I, Lh = self.B, c.tokens, c.hidden_size, c.intermediate_size, c.num_hidden_layers
        M = B * T
        self.M = M
        f32 = dict(dtype=torch.float32, device=dev)
        act = dict(dtype=self.act_dtype, device=dev)
        # ---- inputs / outputs ----
        self.x = torch.zeros(B, c.image_size, **f32)
        if loss_kind == _lib.LOSS_CE:
            self.labels = torch.zeros(B, dtype=torch.int64, device=dev)
        else:
            self.labels = torch.zeros(B * c.num_labels, **f32)
        self.logits = torch.zeros(B, c.num_labels, **f32)
        self.loss = torch.zeros(1, **f32)
        # ---- saved activations: one layer-major tensor per kind (the whole-network kernels address them through ONE
        # 4-D tensor map [cols, T, B, layers]); the per-layer lists are views for the per-op programs ----
        Ln = max(Lh, 1)
        self.z_all = torch.empty(Lh + 1, M, H, **f32)
        self.hmid_all = torch.empty(Ln, M, H, **f32)
        self.u_all = torch.empty(Ln, M, H, **act)
        self.u2_all = torch.empty(Ln, M, H, **act)
        self.qkv_all = torch.empty(Ln, M, 3 * H, **act)
        self.ctx_all = torch.empty(Ln, M, H, **act)
        self.lse_all = torch.empty(Ln, B, c.num_attention_heads, T, **f32)
        self.a_all = torch.empty(Ln, M, I, **act)
        self.m_all = torch.empty(Ln, M, I, **act)
        self.z = [self.z_all[l] for l in range(Lh + 1)]
        self.hmid = [self.hmid_all[l] for l in range(Lh)]
        self.u = [self.u_all[l] for l in range(Ln)]
        self.u2 = [self.u2_all[l] for l in range(Lh)]
        self.stats = torch.empty(4 * max(Lh, 1) + 2, M, **f32)  # mean1,rstd1,mean2,rstd2 per layer + final
        self.qkv = [self.qkv_all[l] for l in range(Lh)]
        self.ctx = [self.ctx_all[l] for l in range(Lh)]
        self.lse = [self.lse_all[l] for l in range(Lh)]
        self.a = [self.a_all[l] for l in range(Lh)]
        self.m = [self.m_all[l] for l in range(Lh)]
        self.delta = torch.empty(M, H, **act)
        self.s_cls = torch.empty(B, H, **act)
        # ---- backward scratch (allocated lazily) ----
        self._bwd_ready = False
        # ---- device scalars ----
        self.rng = torch.tensor([seed, 0], dtype=torch.int64, device=dev)  # {seed, step}, read as uint64
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, max_norm, grad_scale, 0.0], **f32)
        self.state = torch.zeros(8, **f32)
        self.exp_avg = None
        self.exp_avg_sq = None
        # ---- rope tables (rope.py:37-57), [T, d/2] ----
        self.rope_cos = self.rope_sin = None
        if c.pos_encoding_type == "rope":
            d = c.head_dim
            inv_freq = 1.0 / (c.rope_base ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
            fr = torch.outer(torch.arange(T, dtype=torch.float32), inv_freq)
            self.rope_cos = fr.cos().to(dev).contiguous()
            self.rope_sin = fr.sin().to(dev).contiguous()
        # fused row-chain kernels (tcgen05 GEMM chains per 128-row tile): bf16, H in {32, 64}
        self.fused = bool(self.dt == BF16 and Lh >= 1 and I == 4 * H
                          and self.lib.vitb200_fused_supported(H, c.patch_size)
                          and os.environ.get("VITB200_FUSED", "1") != "0")
        self.fused_bwd = bool(self.fused and self.lib.vitb200_fused_bwd_supported(H)
                              and os.environ.get("VITB200_FUSED_BWD", "1") != "0")
        # whole-network kernels (one persistent CTA per sample): bf16, hidden 32, 2 heads, T <= 129
        self.mega = bool(self.fused and Lh >= 1 and self.lib.vitb200_mega_supported(
            H, c.num_attention_heads, T, c.patch_size, c.num_labels, Lh, 1 if c.pos_encoding_type == "rope" else 0)
            and os.environ.get("VITB200_MEGA", "1") != "0")
        # CTA pairs (one attention head per CTA, 2 SMs per sample) while every sample still gets its own pair in one wave
        self.mega_cluster = int(os.environ.get("VITB200_MEGA_CLUSTER", "2" if 2 * self.B <= 148 else "1"))
        # whole-network backward: one sample per CTA (pair), i.e. small batches -- the latency-bound regime it exists for
        self.mega_bwd = bool(self.mega and self.fused_bwd and os.environ.get("VITB200_MEGA_BWD", "1") != "0"
                             and self.lib.vitb200_mega_bwd_supported(H, c.num_attention_heads, T, c.patch_size, c.num_labels,
                                                                     Lh, self.B, self.mega_cluster))
        # cls_only: the caller needs logits / loss / gradients only (TrainStep, EvalStep, MyViT.forward without
        # output_hidden_states): the whole-network kernels then run the LAST layer for the CLS row alone -- the head reads
        # nothing else (specvit.py:78).  Callers that read every token's last hidden state leave it False.
        self.cls_only = False
        self._keep = []  # ctypes argument structs referenced by the cached programs
        ws_bytes = self._ws_bytes()
        self.ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
        self._progs = {}
        self.launches = {}

    # ------------------------------------------------------------------------------------------
    def _ws_bytes(self) -> int:
        c, lib, M = self.cfg, self.lib, self.M
        H, I = c.hidden_size, c.intermediate_size
        n = [
            lib.vitb200_add_ln_bwd_ws_bytes(M, H),
            lib.vitb200_linear_wgrad_ws_bytes(M, 3 * H, H),
            lib.vitb200_linear_wgrad_ws_bytes(M, H, H),
            lib.vitb200_linear_wgrad_ws_bytes(M, I, H),
            lib.vitb200_linear_wgrad_ws_bytes(M, H, I),
            lib.vitb200_patch_embed_bwd_ws_bytes(self.B, c.num_patches, c.patch_size, H),
            lib.vitb200_grad_norm_ws_bytes(self.arena.layout.n_opt),
        ]
        return int(max(n)) + 4096

    def _alloc_backward(self):
        if self._bwd_ready:
            return
        c, dev = self.cfg, self.device
        B, T, H, I = self.B, c.tokens, c.hidden_size, c.intermediate_size
        M = self.M
        f32 = dict(dtype=torch.float32, device=dev)
        act = dict(dtype=self.act_dtype, device=dev)
        self.ds_cls = torch.empty(B, H, **act)
        self.dzA = torch.empty(M, H, **f32)
        self.dzB = torch.empty(M, H, **f32)
        self.ddelta = torch.empty(M, H, **act)
        self.dbig = torch.empty(M, I, **act)
        self.du = torch.empty(M, H, **act)
        self.dqkv = torch.empty(M, 3 * H, **act)
        self.dctx = torch.empty(M, H, **act)
        self.dsum = torch.empty(B, c.num_attention_heads, T, **f32)
        self._bwd_ready = True

    # ---- pointer helpers ----------------------------------------------------------------------
    def _p(self, name: str) -> int:  # fp32 master parameter
        return self.arena.data.data_ptr() + 4 * self.arena.layout.off(name)

    def _g(self, name: str) -> int:  # fp32 gradient
        return self.arena.grad.data_ptr() + 4 * self.arena.layout.off(name)

    def _w(self, name: str) -> int:  # GEMM operand copy of a weight (bf16 shadow in bf16 mode)
        if self.dt == BF16:
            return self.arena.shadow.data_ptr() + 2 * self.arena.layout.off(name)
        return self._p(name)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
        return None if t is None else t.data_ptr()

    def _stat(self, i: int) -> int:
        return self.stats[i].data_ptr()

    # ---- programs -----------------------------------------------------------------------------
    # ---- attention call selection ----------------------------------------------------------------
    def _attn_fwd_call(self, l: int, pa: float, es: int):
        c, P_ = self.cfg, self._ptr
        H, T = c.hidden_size, c.tokens
        qkv = self.qkv[l].data_ptr()
        scale = 1.0 / math.sqrt(c.head_dim)
        rng = self.rng.data_ptr()
        if (self.dt == BF16 and os.environ.get("VITB200_FLASH", "1") != "0"
                and not self.lib.vitb200_attn_tc_supported(T, c.head_dim, 3 * H, H)
                and self.lib.vitb200_attn_flash_supported(T, c.head_dim, 3 * H, H)):
            # longer sequences / head_dim 64: the in-kernel key/value loop (one launch, online softmax)
            return (self.lib.vitb200_attn_flash_fwd, (
                qkv, P_(self.ctx[l]), P_(self.lse[l]), P_(self.rope_cos), P_(self.rope_sin), self.B, T,
                c.num_attention_heads, c.head_dim, scale, pa, rng, site_attn(l)))
        return (self.lib.vitb200_attn_fwd, (
            qkv, qkv + H * es, qkv + 2 * H * es, 3 * H, P_(self.ctx[l]), P_(self.lse[l]),
            P_(self.rope_cos), P_(self.rope_sin), self.B, T, c.num_attention_heads, c.head_dim, scale, pa, rng,
            site_attn(l), self.dt))

    def _attn_bwd_call(self, l: int, pa: float, es: int):
        c, P_ = self.cfg, self._ptr
        H, T = c.hidden_size, c.tokens
        qkv = self.qkv[l].data_ptr()
        dqkv = self.dqkv.data_ptr()
        scale = 1.0 / math.sqrt(c.head_dim)
        rng = self.rng.data_ptr()
        if (self.dt == BF16 and os.environ.get("VITB200_FLASH", "1") != "0"
                and not self.lib.vitb200_attn_tc_supported(T, c.head_dim, 3 * H, H)
                and self.lib.vitb200_attn_flash_supported(T, c.head_dim, 3 * H, H)):
            # longer sequences / head_dim 64: one launch over (key block, head, sample), dK / dV resident in TMEM
            if not hasattr(self, "flash_ws"):
                nbytes = int(self.lib.vitb200_attn_flash_bwd_ws_bytes(self.B, T, c.num_attention_heads, c.head_dim))
                self.flash_ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            return (self.lib.vitb200_attn_flash_bwd, (
                qkv, P_(self.ctx[l]), P_(self.dctx), P_(self.lse[l]), dqkv, P_(self.rope_cos), P_(self.rope_sin),
                self.B, T, c.num_attention_heads, c.head_dim, scale, pa, rng, site_attn(l), P_(self.flash_ws)))
        return (self.lib.vitb200_attn_bwd, (
            qkv, qkv + H * es, qkv + 2 * H * es, 3 * H, P_(self.ctx[l]), P_(self.dctx), P_(self.lse[l]),
            P_(self.dsum), dqkv, dqkv + H * es, dqkv + 2 * H * es, 3 * H, P_(self.rope_cos), P_(self.rope_sin),
            self.B, T, c.num_attention_heads, c.head_dim, scale, pa, rng, site_attn(l), self.dt))

    def input_slot(self, slot: int):
        """(pixels, labels) input buffers of `slot`.  Slot 0 is `self.x` / `self.labels`; slots 1 .. HOST_SLOTS - 1 are
        further pairs the whole-network programs can be built on, so that a host pipeline uploads the next batches while
        the running steps still read theirs (TrainStep.fit_host) -- no device-to-device staging copy on the step's
        critical path."""
        if slot == 0:
            return self.x, self.labels
        if slot == ROWS_SLOT:
            return self._rows["x"], self._rows["labels"]
        if not 0 < slot < HOST_SLOTS:
            raise ValueError(f"input slot {slot} does not exist")
        if not hasattr(self, "_slots"):
            self._slots = {}
        if slot not in self._slots:
            self._slots[slot] = (torch.zeros_like(self.x), torch.zeros_like(self.labels))
        return self._slots[slot]

    def bind_rows(self, x_all: torch.Tensor, labels_all: torch.Tensor, rows: torch.Tensor, loss_log: torch.Tensor) -> None:
        """Device-resident dataset mode of the whole-network kernels (input slot ROWS_SLOT): x_all [N, L] f32 and labels_all
        are the whole dataset, rows (int64) the epoch's permutation, loss_log [>= steps per epoch] f32.  The kernels read
        row rows[pos * B + b] with pos = rng[1] - rows_base (vitb200_mega_fwd_args.rows), so a captured step needs no
        gather launch and no staging copy.  `start_rows()` restarts pos at 0."""
        if x_all.dtype != torch.float32 or not x_all.is_contiguous() or x_all.shape[1] != self.cfg.image_size:
            raise ValueError("bind_rows: x_all must be a contiguous [N, image_size] float32 tensor")
        if labels_all.dtype != self.labels.dtype or not labels_all.is_contiguous() or rows.dtype != torch.int64:
            raise ValueError("bind_rows: labels / rows dtype mismatch")
        if labels_all.numel() != x_all.shape[0] * (self.labels.numel() // self.B):
            raise ValueError("bind_rows: labels do not match the dataset rows")
        self._rows = dict(x=x_all, labels=labels_all, rows=rows, loss_log=loss_log,
                          base=torch.zeros(1, dtype=torch.int64, device=self.device))
        for k in [k for k in self._progs if ROWS_SLOT in k[1:] and k[0] in ("fwd", "bwd")]:
            del self._progs[k]

    def start_rows(self) -> None:
        """The next step reads rows[0 : B] (stream-ordered: rows_base <- the device step counter)."""
        self._rows["base"].copy_(self.rng[1:2])

    def _mega_fwd_args(self, train: bool, with_labels: bool, slot: int = 0, defer_loss: bool = False):
        c, lay, P_ = self.cfg, self.arena.layout, self._ptr
        x_in, lab_in = self.input_slot(slot)
        if not hasattr(self, "mega_ws"):
            self.mega_ws = torch.zeros(int(self.lib.vitb200_mega_ws_bytes()), dtype=torch.uint8, device=self.device)
        emb, L0 = "vit.embeddings.", "vit.encoder.layer.0."
        base0 = lay.off(L0 + "layernorm_before.weight")
        stride = (lay.off("vit.encoder.layer.1.layernorm_before.weight") - base0) if c.num_hidden_layers > 1 else 0
        rel = lambda n: lay.off(L0 + n) - base0   # noqa: E731
        a = _lib.MegaFwdArgs(
            B=self.B, L=c.image_size, P=c.patch_size, S=c.stride, Np=c.num_patches, n_valid=c.n_valid,
            layers=c.num_hidden_layers, C=c.num_labels, loss_kind=self.loss_kind, cluster=self.mega_cluster,
            cls_only=1 if self.cls_only else 0,
            eps=float(c.layer_norm_eps), p_hidden=float(c.hidden_dropout_prob) if train else 0.0,
            p_attn=float(c.attention_probs_dropout_prob) if train else 0.0, rng=self.rng.data_ptr(), x=P_(x_in),
            labels=P_(lab_in) if with_labels else None, params=self.arena.data.data_ptr(),
            shadow=self.arena.shadow.data_ptr(), off_cls=lay.off(emb + "cls_token"),
            off_pos=lay.off(emb + "position_embeddings") if c.pos_encoding_type == "learned" else -1,
            off_wp=lay.off(emb + "patch_embeddings.projection.weight"), off_bp=lay.off(emb + "patch_embeddings.projection.bias"),
            off_layer0=base0, layer_stride=stride, o_ln1g=0, o_ln1b=rel("layernorm_before.bias"),
            o_wqkv=rel("attention.attention.query.weight"), o_bqkv=rel("attention.attention.query.bias"),
            o_wo=rel("attention.output.dense.weight"), o_bo=rel("attention.output.dense.bias"),
            o_ln2g=rel("layernorm_after.weight"), o_ln2b=rel("layernorm_after.bias"),
            o_w1=rel("intermediate.dense.weight"), o_b1=rel("intermediate.dense.bias"),
            o_w2=rel("output.dense.weight"), o_b2=rel("output.dense.bias"),
            off_lnfg=lay.off("vit.layernorm.weight"), off_lnfb=lay.off("vit.layernorm.bias"),
            off_wh=lay.off(lay.head_name + ".weight"), off_bh=lay.off(lay.head_name + ".bias"),
            rope_cos=P_(self.rope_cos), rope_sin=P_(self.rope_sin), z=P_(self.z_all), hmid=P_(self.hmid_all),
            u=P_(self.u_all), u2=P_(self.u2_all), qkv=P_(self.qkv_all), ctx=P_(self.ctx_all), a=P_(self.a_all),
            m=P_(self.m_all), stats=P_(self.stats), lse=P_(self.lse_all), s_cls=P_(self.s_cls), logits=P_(self.logits),
            loss=P_(self.loss), ws=P_(self.mega_ws))
        a.defer_loss = 1 if (defer_loss and with_labels) else 0   # the backward kernel of the step finishes the loss
        if slot == ROWS_SLOT:
            a.rows, a.rows_base = P_(self._rows["rows"]), P_(self._rows["base"])
            a.loss_log = P_(self._rows["loss_log"])
        elif with_labels:
            # the loss also lands in page-locked HOST memory (unified addressing: the kernel stores through the host
            # pointer), so a host loop reads it after the step's event without a device-to-host copy launch
            if not hasattr(self, "loss_pinned"):
                self.loss_pinned = torch.zeros(HOST_SLOTS, dtype=torch.float32, pin_memory=True)
            a.loss_log = self.loss_pinned.data_ptr() + 4 * slot
        self._keep.append(a)
        return a

    def _build_forward_fused(self, train: bool, with_labels: bool, head_bwd: bool = False, slot: int = 0) -> List[Tuple[Callable, tuple]]:
        """embed -> [attention, fused layer] x L -> head: 2 + 2L (+1 loss) launches instead of 4 + 7L;
        whole-network kernel when the shape allows: ONE launch (+ the head backward for training steps)."""
        c, lib, dt, P_ = self.cfg, self.lib, self.dt, self._ptr
        if slot != 0 and not (self.mega and self.mega_bwd):
            raise RuntimeError("vit_b200: input slots other than 0 exist for the whole-network programs only")
        if self.mega:
            a = self._mega_fwd_args(train, with_labels, slot, defer_loss=bool(head_bwd and self.mega_bwd))
            prog = [(lib.vitb200_mega_fwd, (ctypes.addressof(a),))]
            if head_bwd and not self.mega_bwd:
                self._alloc_backward()
                self._ensure_dz_cls()
                T, H, Lh, hd = c.tokens, c.hidden_size, c.num_hidden_layers, self.arena.layout.head_name
                fin = 4 * max(Lh, 1)
                prog.append((lib.vitb200_head_fused_bwd, (
                    P_(self.s_cls), self._w(hd + ".weight"), P_(self.logits), P_(self.labels), None, P_(self.z[Lh]), T * H,
                    self._stat(fin), self._stat(fin + 1), self._p("vit.layernorm.weight"), P_(self.dz_cls),
                    self._g("vit.layernorm.weight"), self._g("vit.layernorm.bias"), self._g(hd + ".weight"),
                    self._g(hd + ".bias"), self.B, H, c.num_labels, self.loss_kind, 0, dt)))
            return prog
        B, T, H, Lh, M = self.B, c.tokens, c.hidden_size, c.num_hidden_layers, self.M
        ph = float(c.hidden_dropout_prob) if train else 0.0
        pa = float(c.attention_probs_dropout_prob) if train else 0.0
        eps = float(c.layer_norm_eps)
        rng = self.rng.data_ptr()
        scale = 1.0 / math.sqrt(c.head_dim)
        emb = "vit.embeddings."
        L0 = "vit.encoder.layer.0."
        ea = _lib.EmbedFwdArgs(
            B=B, L=c.image_size, P=c.patch_size, S=c.stride, Np=c.num_patches, n_valid=c.n_valid, H=H, eps=eps,
            p_drop=ph, rng=rng, x=P_(self.x), w_p=self._w(emb + "patch_embeddings.projection.weight"),
            b_p=self._p(emb + "patch_embeddings.projection.bias"), cls=self._p(emb + "cls_token"),
            pos=self._p(emb + "position_embeddings") if c.pos_encoding_type == "learned" else None,
            ln_g=self._p(L0 + "layernorm_before.weight"), ln_b=self._p(L0 + "layernorm_before.bias"),
            w_qkv=self._w(L0 + "attention.attention.query.weight"), b_qkv=self._p(L0 + "attention.attention.query.bias"),
            z0=P_(self.z[0]), u=P_(self.u[0]), mean=self._stat(0), rstd=self._stat(1), qkv=P_(self.qkv[0]))
        self._keep.append(ea)
        prog = [(lib.vitb200_fused_embed_fwd, (ctypes.addressof(ea),))]
        fin = 4 * max(Lh, 1)
        for l in range(Lh):
            pre = f"vit.encoder.layer.{l}."
            last = l == Lh - 1
            nxt = "" if last else f"vit.encoder.layer.{l + 1}."
            qkv = self.qkv[l].data_ptr()
            prog.append(self._attn_fwd_call(l, pa, 2))
            la = _lib.LayerFwdArgs(
                B=B, T=T, H=H, last=1 if last else 0, eps=eps, p_drop=ph, rng=rng, site_proj=site_proj(l),
                site_mlp=site_mlp(l), ctx=P_(self.ctx[l]), z_in=P_(self.z[l]),
                w_o=self._w(pre + "attention.output.dense.weight"), w_1=self._w(pre + "intermediate.dense.weight"),
                w_2=self._w(pre + "output.dense.weight"),
                w_qkv=None if last else self._w(nxt + "attention.attention.query.weight"),
                b_o=self._p(pre + "attention.output.dense.bias"), ln2_g=self._p(pre + "layernorm_after.weight"),
                ln2_b=self._p(pre + "layernorm_after.bias"), b_1=self._p(pre + "intermediate.dense.bias"),
                b_2=self._p(pre + "output.dense.bias"),
                lnn_g=self._p("vit.layernorm.weight" if last else nxt + "layernorm_before.weight"),
                lnn_b=self._p("vit.layernorm.bias" if last else nxt + "layernorm_before.bias"),
                b_qkv=None if last else self._p(nxt + "attention.attention.query.bias"),
                hmid=P_(self.hmid[l]), u2=P_(self.u2[l]), mean2=self._stat(4 * l + 2), rstd2=self._stat(4 * l + 3),
                a=P_(self.a[l]), m=P_(self.m[l]), z_out=P_(self.z[l + 1]),
                u_next=P_(self.s_cls) if last else P_(self.u[l + 1]),
                mean_n=self._stat(fin) if last else self._stat(4 * (l + 1)),
                rstd_n=self._stat(fin + 1) if last else self._stat(4 * (l + 1) + 1),
                qkv_next=None if last else P_(self.qkv[l + 1]))
            self._keep.append(la)
            prog.append((lib.vitb200_fused_layer_fwd, (ctypes.addressof(la),)))
        hd = self.arena.layout.head_name
        if head_bwd:  # training step: logits, loss AND the head / final-LayerNorm backward in one launch
            self._alloc_backward()
            self._ensure_dz_cls()
            prog.append((lib.vitb200_head_fused_fwd_bwd, (
                P_(self.s_cls), self._w(hd + ".weight"), self._p(hd + ".bias"), P_(self.labels), P_(self.logits),
                P_(self.loss), P_(self.z[Lh]), T * H, self._stat(fin), self._stat(fin + 1),
                self._p("vit.layernorm.weight"), P_(self.dz_cls), self._g("vit.layernorm.weight"),
                self._g("vit.layernorm.bias"), self._g(hd + ".weight"), self._g(hd + ".bias"), B, H, c.num_labels,
                self.loss_kind, dt)))
            return prog
        head_fn = lib.vitb200_head_fused_fwd if lib.vitb200_head_fused_supported(H, c.num_labels) else lib.vitb200_head_loss_fwd
        prog.append((head_fn, (
            P_(self.s_cls), self._w(hd + ".weight"), self._p(hd + ".bias"),
            P_(self.labels) if with_labels else None, P_(self.logits), P_(self.loss), B, H, c.num_labels,
            self.loss_kind, dt)))
        return prog

    def _ensure_dz_cls(self) -> None:
        if not hasattr(self, "dz_cls"):
            c = self.cfg
            self.dz_cls = torch.zeros(self.B, c.hidden_size, dtype=torch.float32, device=self.device)

    @property
    def can_fuse_head(self) -> bool:
        """forward's last launch can also do the head backward (fused programs, CLS-row top gradient, known dloss = 1)."""
        c = self.cfg
        return bool(self.fused and self.fused_bwd and self.lib.vitb200_head_fused_supported(c.hidden_size, c.num_labels)
                    and self.loss_kind in (_lib.LOSS_MSE, _lib.LOSS_L1, _lib.LOSS_CE))

    def _build_forward(self, train: bool, with_labels: bool) -> List[Tuple[Callable, tuple]]:
        if self.fused:
            return self._build_forward_fused(train, with_labels)
        c, lib, dt, P_ = self.cfg, self.lib, self.dt, self._ptr
        B, T, H, I, Lh, M = self.B, c.tokens, c.hidden_size, c.intermediate_size, c.num_hidden_layers, self.M
        ph = float(c.hidden_dropout_prob) if train else 0.0
        pa = float(c.attention_probs_dropout_prob) if train else 0.0
        eps = float(c.layer_norm_eps)
        rng = self.rng.data_ptr()
        es = 2 if dt == BF16 else 4
        scale = 1.0 / math.sqrt(c.head_dim)
        emb = "vit.embeddings."
        pos = self._p(emb + "position_embeddings") if c.pos_encoding_type == "learned" else None
        prog = []
        prog.append((lib.vitb200_patch_embed_fwd, (
            P_(self.x), self._w(emb + "patch_embeddings.projection.weight"),
            self._p(emb + "patch_embeddings.projection.bias"), self._p(emb + "cls_token"), pos, P_(self.z[0]),
            B, c.image_size, c.patch_size, c.stride, c.num_patches, c.n_valid, H, ph, rng, SITE_EMB, dt)))
        fin = 4 * max(Lh, 1)
        if Lh == 0:
            prog.append((lib.vitb200_add_ln_fwd, (
                P_(self.z[0]), None, None, P_(self.s_cls), self._stat(fin), self._stat(fin + 1),
                self._p("vit.layernorm.weight"), self._p("vit.layernorm.bias"), M, H, T, eps, 0.0, rng, 0, dt)))
        else:
            pre = "vit.encoder.layer.0."
            prog.append((lib.vitb200_add_ln_fwd, (
                P_(self.z[0]), None, None, P_(self.u[0]), self._stat(0), self._stat(1),
                self._p(pre + "layernorm_before.weight"), self._p(pre + "layernorm_before.bias"),
                M, H, 0, eps, 0.0, rng, 0, dt)))
        for l in range(Lh):
            pre = f"vit.encoder.layer.{l}."
            qkv = self.qkv[l].data_ptr()
            prog.append((lib.vitb200_linear_fwd, (
                P_(self.u[l]), self._w(pre + "attention.attention.query.weight"),
                self._p(pre + "attention.attention.query.bias"), qkv, None, M, 3 * H, H, ACT_NONE, dt)))
            prog.append(self._attn_fwd_call(l, pa, es))
            prog.append((lib.vitb200_linear_fwd, (
                P_(self.ctx[l]), self._w(pre + "attention.output.dense.weight"),
                self._p(pre + "attention.output.dense.bias"), P_(self.delta), None, M, H, H, ACT_NONE, dt)))
            prog.append((lib.vitb200_add_ln_fwd, (
                P_(self.z[l]), P_(self.delta), P_(self.hmid[l]), P_(self.u2[l]), self._stat(4 * l + 2),
                self._stat(4 * l + 3), self._p(pre + "layernorm_after.weight"), self._p(pre + "layernorm_after.bias"),
                M, H, 0, eps, ph, rng, site_proj(l), dt)))
            prog.append((lib.vitb200_linear_fwd, (
                P_(self.u2[l]), self._w(pre + "intermediate.dense.weight"), self._p(pre + "intermediate.dense.bias"),
                P_(self.a[l]), P_(self.m[l]), M, I, H, ACT_GELU, dt)))
            prog.append((lib.vitb200_linear_fwd, (
                P_(self.m[l]), self._w(pre + "output.dense.weight"), self._p(pre + "output.dense.bias"),
                P_(self.delta), None, M, H, I, ACT_NONE, dt)))
            if l < Lh - 1:
                nxt = f"vit.encoder.layer.{l + 1}."
                prog.append((lib.vitb200_add_ln_fwd, (
                    P_(self.hmid[l]), P_(self.delta), P_(self.z[l + 1]), P_(self.u[l + 1]), self._stat(4 * (l + 1)),
                    self._stat(4 * (l + 1) + 1), self._p(nxt + "layernorm_before.weight"),
                    self._p(nxt + "layernorm_before.bias"), M, H, 0, eps, ph, rng, site_mlp(l), dt)))
            else:
                prog.append((lib.vitb200_add_ln_fwd, (
                    P_(self.hmid[l]), P_(self.delta), P_(self.z[Lh]), P_(self.s_cls), self._stat(fin),
                    self._stat(fin + 1), self._p("vit.layernorm.weight"), self._p("vit.layernorm.bias"),
                    M, H, T, eps, ph, rng, site_mlp(l), dt)))
        hd = self.arena.layout.head_name
        head_fn = lib.vitb200_head_fused_fwd if lib.vitb200_head_fused_supported(H, c.num_labels) else lib.vitb200_head_loss_fwd
        prog.append((head_fn, (
            P_(self.s_cls), self._w(hd + ".weight"), self._p(hd + ".bias"),
            P_(self.labels) if with_labels else None, P_(self.logits), P_(self.loss), B, H, c.num_labels,
            self.loss_kind, dt)))
        return prog

    def _build_backward_fused(self, train: bool, gloss_ptr: Optional[int], given: bool, skip_reduce: bool = False,
                              skip_head: bool = False, slot: int = 0, streamed: bool = False):
        """head -> final LN -> [upper, attention bwd, lower] x L -> embed -> reduce of the per-CTA partials."""
        self._alloc_backward()
        c, lib, dt, P_ = self.cfg, self.lib, self.dt, self._ptr
        B, T, H, Lh, M = self.B, c.tokens, c.hidden_size, c.num_hidden_layers, self.M
        ph = float(c.hidden_dropout_prob) if train else 0.0
        pa = float(c.attention_probs_dropout_prob) if train else 0.0
        rng = self.rng.data_ptr()
        scale = 1.0 / math.sqrt(c.head_dim)
        ws = self.ws.data_ptr()
        lay = self.arena.layout
        hd = lay.head_name
        fin = 4 * max(Lh, 1)
        G = int(lib.vitb200_fused_bwd_grid(M))
        if not hasattr(self, "gpart"):
            slots = max(G, B) if self.mega_bwd else G
            self.gpart = torch.zeros(slots, lay.n_opt, dtype=torch.float32, device=self.device)  # padding stays zero
        gp = self.gpart.data_ptr()
        prog = []
        if given and not hasattr(self, "dlogits"):
            self.dlogits = torch.zeros(B, c.num_labels, dtype=torch.float32, device=self.device)
        lab_ptr = P_(self.dlogits) if given else P_(self.input_slot(slot)[1])
        kind = _lib.LOSS_GIVEN if given else self.loss_kind
        gl = None if given else gloss_ptr
        if self.mega_bwd:
            # the whole backward is ONE launch: head, final LN, every layer, embeddings; one gradient set per sample
            fa = self._mega_fwd_args(train, True, slot, defer_loss=bool(skip_head and not given))
            ba = _lib.MegaBwdArgs(f=fa, labels=lab_ptr, gloss=gl, loss_kind=kind, n_opt=lay.n_opt, gpart=gp,
                                  dz0=P_(self.dzA))
            if streamed:
                ba.done = self.grad_stream().done
            self._keep.append(ba)
            prog.append((lib.vitb200_mega_bwd, (ctypes.addressof(ba),)))
            self._red = (B, 0, lay.n_opt)
            if not skip_reduce:
                prog.append((lib.vitb200_grad_reduce, (gp, B, lay.n_opt, 0, lay.n_opt, self.arena.grad.data_ptr())))
            return prog
        cur, other = self.dzA, self.dzB
        top_cls = bool(lib.vitb200_head_fused_supported(H, c.num_labels))
        if top_cls and skip_head:
            self._ensure_dz_cls()          # written by forward's vitb200_head_fused_fwd_bwd
        elif top_cls:
            # head backward + final-LayerNorm backward of the CLS rows in one launch; only those rows carry a gradient
            self._ensure_dz_cls()
            prog.append((lib.vitb200_head_fused_bwd, (
                P_(self.s_cls), self._w(hd + ".weight"), P_(self.logits), lab_ptr, gl, P_(self.z[Lh]), T * H,
                self._stat(fin), self._stat(fin + 1), self._p("vit.layernorm.weight"), P_(self.dz_cls),
                self._g("vit.layernorm.weight"), self._g("vit.layernorm.bias"), self._g(hd + ".weight"),
                self._g(hd + ".bias"), B, H, c.num_labels, kind, 0, dt)))
        else:
            prog.append((lib.vitb200_head_loss_bwd, (
                P_(self.s_cls), self._w(hd + ".weight"), P_(self.logits), lab_ptr, gl, P_(self.ds_cls),
                self._g(hd + ".weight"), self._g(hd + ".bias"), B, H, c.num_labels, kind, 0, dt)))
            prog.append((lib.vitb200_add_ln_bwd, (
                P_(self.ds_cls), P_(self.z[Lh]), self._stat(fin), self._stat(fin + 1), self._p("vit.layernorm.weight"),
                None, P_(cur), None, self._g("vit.layernorm.weight"), self._g("vit.layernorm.bias"), M, H, T, 0.0,
                rng, 0, 0, dt, ws)))
        for l in range(Lh - 1, -1, -1):
            pre = f"vit.encoder.layer.{l}."
            qkv = self.qkv[l].data_ptr()
            dqkv = self.dqkv.data_ptr()
            from_cls = top_cls and l == Lh - 1
            ua = _lib.LayerBwdUpperArgs(
                B=B, T=T, H=H, p_drop=ph, rng=rng, site_proj=site_proj(l), site_mlp=site_mlp(l),
                dz=None if from_cls else P_(cur), dz_cls=P_(self.dz_cls) if from_cls else None,
                m=P_(self.m[l]), a=P_(self.a[l]), u2=P_(self.u2[l]), ctx=P_(self.ctx[l]), hmid=P_(self.hmid[l]),
                mean2=self._stat(4 * l + 2), rstd2=self._stat(4 * l + 3), ln2_g=self._p(pre + "layernorm_after.weight"),
                w_2=self._w(pre + "output.dense.weight"), w_1=self._w(pre + "intermediate.dense.weight"),
                w_o=self._w(pre + "attention.output.dense.weight"), dh=P_(other), dctx=P_(self.dctx), gpart=gp,
                n_opt=lay.n_opt, off_w2=lay.off(pre + "output.dense.weight"), off_b2=lay.off(pre + "output.dense.bias"),
                off_w1=lay.off(pre + "intermediate.dense.weight"), off_b1=lay.off(pre + "intermediate.dense.bias"),
                off_ln2g=lay.off(pre + "layernorm_after.weight"), off_ln2b=lay.off(pre + "layernorm_after.bias"),
                off_wo=lay.off(pre + "attention.output.dense.weight"), off_bo=lay.off(pre + "attention.output.dense.bias"))
            self._keep.append(ua)
            prog.append((lib.vitb200_fused_layer_bwd_upper, (ctypes.addressof(ua),)))
            prog.append(self._attn_bwd_call(l, pa, 2))
            la = _lib.LayerBwdLowerArgs(
                B=B, T=T, H=H, dqkv=dqkv, u=P_(self.u[l]), z=P_(self.z[l]), mean1=self._stat(4 * l),
                rstd1=self._stat(4 * l + 1), ln1_g=self._p(pre + "layernorm_before.weight"), dh=P_(other),
                w_qkv=self._w(pre + "attention.attention.query.weight"), dz=P_(cur), gpart=gp, n_opt=lay.n_opt,
                off_wqkv=lay.off(pre + "attention.attention.query.weight"),
                off_bqkv=lay.off(pre + "attention.attention.query.bias"),
                off_ln1g=lay.off(pre + "layernorm_before.weight"), off_ln1b=lay.off(pre + "layernorm_before.bias"))
            self._keep.append(la)
            prog.append((lib.vitb200_fused_layer_bwd_lower, (ctypes.addressof(la),)))
        emb = "vit.embeddings."
        learned = c.pos_encoding_type == "learned"
        start = lay.buckets[1][1]          # first encoder layer
        if lib.vitb200_fused_embed_bwd_supported(H, c.patch_size, 1 if learned else 0):
            eb = _lib.EmbedBwdArgs(
                B=B, L=c.image_size, P=c.patch_size, S=c.stride, Np=c.num_patches, n_valid=c.n_valid, H=H, p_drop=ph,
                rng=rng, dz0=P_(cur), x=P_(self.x), gpart=gp, n_opt=lay.n_opt,
                off_wp=lay.off(emb + "patch_embeddings.projection.weight"),
                off_bp=lay.off(emb + "patch_embeddings.projection.bias"), off_cls=lay.off(emb + "cls_token"))
            self._keep.append(eb)
            prog.append((lib.vitb200_fused_embed_bwd, (ctypes.addressof(eb),)))
            start = lay.buckets[0][1]      # the embeddings bucket is reduced from the CTA partials as well
        else:
            dpos = self._g(emb + "position_embeddings") if learned else None
            prog.append((lib.vitb200_patch_embed_bwd, (
                P_(cur), P_(self.x), self._g(emb + "patch_embeddings.projection.weight"),
                self._g(emb + "patch_embeddings.projection.bias"), self._g(emb + "cls_token"), dpos, B, c.image_size,
                c.patch_size, c.stride, c.num_patches, c.n_valid, H, ph, rng, SITE_EMB, 0, dt, ws)))
        end = lay.buckets[Lh][2]           # end of the last encoder layer
        self._red = (G, start, end)        # partial slots and the arena range they cover
        if not skip_reduce:                # (TrainStep folds this reduction into the one-launch optimizer tail)
            prog.append((lib.vitb200_grad_reduce, (gp, G, lay.n_opt, start, end, self.arena.grad.data_ptr())))
        return prog

    def _build_backward(self, train: bool, gloss_ptr: Optional[int] = None,
                        given: bool = False, skip_reduce: bool = False, skip_head: bool = False, slot: int = 0,
                        streamed: bool = False) -> List[Tuple[Callable, tuple]]:
        if self.fused_bwd:
            return self._build_backward_fused(train, gloss_ptr, given, skip_reduce, skip_head, slot, streamed)
        self._alloc_backward()
        c, lib, dt, P_ = self.cfg, self.lib, self.dt, self._ptr
        B, T, H, I, Lh, M = self.B, c.tokens, c.hidden_size, c.intermediate_size, c.num_hidden_layers, self.M
        ph = float(c.hidden_dropout_prob) if train else 0.0
        pa = float(c.attention_probs_dropout_prob) if train else 0.0
        rng = self.rng.data_ptr()
        es = 2 if dt == BF16 else 4
        scale = 1.0 / math.sqrt(c.head_dim)
        ws = self.ws.data_ptr()
        acc = 0
        hd = self.arena.layout.head_name
        fin = 4 * max(Lh, 1)
        prog = []
        if given:
            if not hasattr(self, "dlogits"):
                self.dlogits = torch.zeros(B, c.num_labels, dtype=torch.float32, device=self.device)
            prog.append((lib.vitb200_head_loss_bwd, (
                P_(self.s_cls), self._w(hd + ".weight"), P_(self.logits), P_(self.dlogits), None, P_(self.ds_cls),
                self._g(hd + ".weight"), self._g(hd + ".bias"), B, H, c.num_labels, _lib.LOSS_GIVEN, acc, dt)))
        else:
            prog.append((lib.vitb200_head_loss_bwd, (
                P_(self.s_cls), self._w(hd + ".weight"), P_(self.logits), P_(self.labels), gloss_ptr,
                P_(self.ds_cls), self._g(hd + ".weight"), self._g(hd + ".bias"), B, H, c.num_labels,
                self.loss_kind, acc, dt)))
        cur, other = self.dzA, self.dzB
        if Lh == 0:
            prog.append((lib.vitb200_add_ln_bwd, (
                P_(self.ds_cls), P_(self.z[0]), self._stat(fin), self._stat(fin + 1), self._p("vit.layernorm.weight"),
                None, P_(cur), None, self._g("vit.layernorm.weight"), self._g("vit.layernorm.bias"), M, H, T, 0.0,
                rng, 0, acc, dt, ws)))
        else:
            prog.append((lib.vitb200_add_ln_bwd, (
                P_(self.ds_cls), P_(self.z[Lh]), self._stat(fin), self._stat(fin + 1), self._p("vit.layernorm.weight"),
                None, P_(cur), P_(self.ddelta), self._g("vit.layernorm.weight"), self._g("vit.layernorm.bias"),
                M, H, T, ph, rng, site_mlp(Lh - 1), acc, dt, ws)))
        for l in range(Lh - 1, -1, -1):
            pre = f"vit.encoder.layer.{l}."
            qkv = self.qkv[l].data_ptr()
            dqkv = self.dqkv.data_ptr()
            # MLP down:  delta2 = m . W2^T + b2
            prog.append((lib.vitb200_linear_wgrad, (
                P_(self.ddelta), P_(self.m[l]), self._g(pre + "output.dense.weight"), self._g(pre + "output.dense.bias"),
                M, H, I, acc, dt, ws)))
            prog.append((lib.vitb200_linear_dgrad, (
                P_(self.ddelta), self._w(pre + "output.dense.weight"), P_(self.a[l]), P_(self.dbig), M, H, I, dt)))
            # MLP up:  a = u2 . W1^T + b1   (dbig now holds da = dm * gelu'(a))
            prog.append((lib.vitb200_linear_wgrad, (
                P_(self.dbig), P_(self.u2[l]), self._g(pre + "intermediate.dense.weight"),
                self._g(pre + "intermediate.dense.bias"), M, I, H, acc, dt, ws)))
            prog.append((lib.vitb200_linear_dgrad, (
                P_(self.dbig), self._w(pre + "intermediate.dense.weight"), None, P_(self.du), M, I, H, dt)))
            # LN2 + residual + proj dropout
            prog.append((lib.vitb200_add_ln_bwd, (
                P_(self.du), P_(self.hmid[l]), self._stat(4 * l + 2), self._stat(4 * l + 3),
                self._p(pre + "layernorm_after.weight"), P_(cur), P_(other), P_(self.ddelta),
                self._g(pre + "layernorm_after.weight"), self._g(pre + "layernorm_after.bias"), M, H, 0, ph, rng,
                site_proj(l), acc, dt, ws)))
            cur, other = other, cur
            # attention output projection
            prog.append((lib.vitb200_linear_wgrad, (
                P_(self.ddelta), P_(self.ctx[l]), self._g(pre + "attention.output.dense.weight"),
                self._g(pre + "attention.output.dense.bias"), M, H, H, acc, dt, ws)))
            prog.append((lib.vitb200_linear_dgrad, (
                P_(self.ddelta), self._w(pre + "attention.output.dense.weight"), None, P_(self.dctx), M, H, H, dt)))
            prog.append(self._attn_bwd_call(l, pa, es))
            # fused QKV projection
            prog.append((lib.vitb200_linear_wgrad, (
                dqkv, P_(self.u[l]), self._g(pre + "attention.attention.query.weight"),
                self._g(pre + "attention.attention.query.bias"), M, 3 * H, H, acc, dt, ws)))
            prog.append((lib.vitb200_linear_dgrad, (
                dqkv, self._w(pre + "attention.attention.query.weight"), None, P_(self.du), M, 3 * H, H, dt)))
            # LN1 + residual (+ the previous layer's MLP dropout)
            if l > 0:
                prog.append((lib.vitb200_add_ln_bwd, (
                    P_(self.du), P_(self.z[l]), self._stat(4 * l), self._stat(4 * l + 1),
                    self._p(pre + "layernorm_before.weight"), P_(cur), P_(other), P_(self.ddelta),
                    self._g(pre + "layernorm_before.weight"), self._g(pre + "layernorm_before.bias"), M, H, 0, ph,
                    rng, site_mlp(l - 1), acc, dt, ws)))
            else:
                prog.append((lib.vitb200_add_ln_bwd, (
                    P_(self.du), P_(self.z[0]), self._stat(0), self._stat(1),
                    self._p(pre + "layernorm_before.weight"), P_(cur), P_(other), None,
                    self._g(pre + "layernorm_before.weight"), self._g(pre + "layernorm_before.bias"), M, H, 0, 0.0,
                    rng, 0, acc, dt, ws)))
            cur, other = other, cur
        emb = "vit.embeddings."
        dpos = self._g(emb + "position_embeddings") if c.pos_encoding_type == "learned" else None
        prog.append((lib.vitb200_patch_embed_bwd, (
            P_(cur), P_(self.x), self._g(emb + "patch_embeddings.projection.weight"),
            self._g(emb + "patch_embeddings.projection.bias"), self._g(emb + "cls_token"), dpos, B, c.image_size,
            c.patch_size, c.stride, c.num_patches, c.n_valid, H, ph, rng, SITE_EMB, acc, dt, ws)))
        return prog

    def _run(self, key, builder) -> None:
        prog = self._progs.get(key)
        if prog is None:
            prog = builder()
            self._progs[key] = prog
            self.launches[key] = len(prog)
        st = torch.cuda.current_stream(self.device).cuda_stream
        for fn, args in prog:
            rc = fn(*args, st)
            if rc != 0:
                _lib.check(rc, fn.__name__)

    # ---- public steps -------------------------------------------------------------------------
    def refresh_shadow(self, force: bool = False) -> None:
        """bf16 GEMM operands <- fp32 master weights (after load_state_dict / an external optimizer)."""
        ar = self.arena
        if ar.shadow is None or not (force or ar.shadow_stale()):
            return
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vitb200_cast_bf16(ar.data.data_ptr(), ar.shadow.data_ptr(), ar.layout.n_total, st), "cast")
        ar.mark_shadow_fresh()

    def forward(self, train: bool, with_labels: bool = True, head_bwd: bool = False, slot: int = 0) -> None:
        """head_bwd (training steps, needs can_fuse_head): the last launch also runs the head / final-LayerNorm backward
        with dloss = 1; the caller must then use backward(skip_head=True)."""
        self.refresh_shadow()
        co = bool(self.cls_only and self.mega)
        if head_bwd:
            assert with_labels and self.can_fuse_head
            key = ("fwd", train, True, True, co) if slot == 0 else ("fwd", train, True, True, co, slot)
            self._run(key, lambda: self._build_forward_fused(train, True, head_bwd=True, slot=slot))
        elif slot != 0:
            self._run(("fwd", train, with_labels, co, slot), lambda: self._build_forward_fused(train, with_labels, slot=slot))
        else:
            self._run(("fwd", train, with_labels, co), lambda: self._build_forward(train, with_labels))

    def backward(self, train: bool, gloss: Optional[torch.Tensor] = None, skip_reduce: bool = False,
                 skip_head: bool = False, slot: int = 0, streamed: bool = False) -> None:
        """skip_reduce (fused programs only): leave the layer / embedding gradients as per-CTA partials; the caller
        must follow with optimizer_step(fused_reduce=True), which sums them inside the optimizer kernel.
        streamed (whole-network backward only): the kernel also signals every finished gradient group; the NEXT launch on
        the stream must then be optimizer_step(fused_reduce=True, streamed=True), which consumes the signals."""
        gp = None if gloss is None else gloss.data_ptr()
        skip = bool(skip_reduce and self.fused_bwd)
        sh = bool(skip_head and self.fused_bwd)
        co = bool(self.cls_only and self.mega)
        key = ("bwd", train, gp, skip, sh, co) if (skip or sh) else ("bwd", train, gp, co)
        streamed = bool(streamed and skip and self.mega_bwd)
        if slot != 0 or streamed:
            key = key + (slot,)
        if streamed:
            key = key + ("streamed",)
        self._run(key, lambda: self._build_backward(train, gp, skip_reduce=skip, skip_head=sh, slot=slot, streamed=streamed))

    def backward_from_dlogits(self, train: bool, dlogits: torch.Tensor) -> None:
        """Backward when the caller computed its own loss from `logits` (labels=None forward)."""
        key = ("bwd_given", train, bool(self.cls_only and self.mega))
        if key not in self._progs:
            self._progs[key] = self._build_backward(train, None, given=True)
            self.launches[key] = len(self._progs[key])
        self.dlogits.copy_(dlogits.reshape(self.dlogits.shape))
        self._run(key, lambda: None)

    def pixel_grad(self, train: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """d loss / d pixel_values [B, L] (fp32) of the backward that just ran: the embedding kernels' input gradient,
        needed only when a trainable preprocessor (src/models/layers.py:51-60) produced the pixels.  Both backward
        programs leave d loss / d z0 in `dzA`; the dropout mask of the embedding site is regenerated from the same
        (seed, step)."""
        c = self.cfg
        dx = out if out is not None else torch.empty(self.B, c.image_size, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        ph = float(c.hidden_dropout_prob) if train else 0.0
        _lib.check(self.lib.vitb200_patch_embed_dgrad(
            self.dzA.data_ptr(), self._w("vit.embeddings.patch_embeddings.projection.weight"), dx.data_ptr(), self.B,
            c.image_size, c.patch_size, c.stride, c.num_patches, c.n_valid, c.hidden_size, ph, self.rng.data_ptr(),
            SITE_EMB, self.dt, st), "patch_embed_dgrad")
        return dx

    def _ensure_opt_state(self):
        if self.exp_avg is None:
            n = self.arena.layout.n_total
            self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)

    def grad_norm(self) -> None:
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vitb200_grad_norm(self.arena.grad.data_ptr(), self.arena.layout.n_opt,
                                              self.hyper.data_ptr(), self.state.data_ptr(), self.ws.data_ptr(), st),
                   "grad_norm")

    def adamw(self, advance_rng: bool = True) -> None:
        self._ensure_opt_state()
        ar = self.arena
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vitb200_adamw(
            ar.data.data_ptr(), ar.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            None if ar.shadow is None else ar.shadow.data_ptr(), ar.layout.n_opt, self.hyper.data_ptr(),
            self.state.data_ptr(), self.rng.data_ptr() if advance_rng else None, st), "adamw")
        if ar.shadow is not None:
            ar.mark_shadow_fresh()

    def grad_stream(self):
        """vitb200_grad_stream of this engine: one counter per gradient group (everything in front of layer 0, each
        encoder layer, final LayerNorm + head), raised by the whole-network backward kernel and consumed by the streamed
        optimizer kernel."""
        if not hasattr(self, "_gs"):
            c, lay = self.cfg, self.arena.layout
            Lh = c.num_hidden_layers
            base0 = lay.off("vit.encoder.layer.0.layernorm_before.weight")
            stride = (lay.off("vit.encoder.layer.1.layernorm_before.weight") - base0) if Lh > 1 else (lay.off("vit.layernorm.weight") - base0)
            cuts = [0] + [base0 + l * stride for l in range(Lh + 1)] + [lay.n_opt]
            assert cuts == sorted(cuts) and cuts[Lh + 1] <= lay.off("vit.layernorm.weight") and all(v % 4 == 0 for v in cuts)
            self.grad_done = torch.zeros(Lh + 2, dtype=torch.int32, device=self.device)
            gs = _lib.GradStream(done=self.grad_done.data_ptr(), expect=self.B * self.mega_cluster, n_groups=Lh + 2)
            for i in range(Lh + 2):
                gs.lo[i], gs.hi[i] = cuts[i], cuts[i + 1]
            self._gs = gs
        return self._gs

    def optimizer_step(self, fused_reduce: bool = False, streamed: bool = False) -> None:
        """clip_grad_norm_ + AdamW.step in ONE launch (vitb200_clip_adamw_fused).  fused_reduce: the backward ran with
        skip_reduce=True, so the kernel first sums the per-CTA gradient partials.  streamed: the backward ran with
        streamed=True and is the previous launch on the stream: the kernel starts on each gradient group as soon as the
        backward kernel has signalled it (and, data parallel, exchanges it over NVLink) instead of waiting for its end."""
        self._ensure_opt_state()
        ar = self.arena
        if not hasattr(self, "tail_ws"):
            self.tail_ws = torch.zeros(int(self.lib.vitb200_clip_adamw_fused_ws_bytes()), dtype=torch.uint8, device=self.device)
        slots, start, end = (self._red if (fused_reduce and self.fused_bwd) else (0, 0, 0))
        st = torch.cuda.current_stream(self.device).cuda_stream
        if streamed and fused_reduce and self.mega_bwd:
            pr = getattr(self, "peer", None)
            _lib.check(self.lib.vitb200_clip_adamw_fused_streamed(
                ar.data.data_ptr(), ar.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                None if ar.shadow is None else ar.shadow.data_ptr(), ar.layout.n_opt, self.hyper.data_ptr(),
                self.state.data_ptr(), self.rng.data_ptr(), self.gpart.data_ptr(), slots, ar.layout.n_opt, start, end,
                (pr.tail_ws if pr is not None else self.tail_ws).data_ptr(), ctypes.addressof(self.grad_stream()),
                None if pr is None else pr.table.data_ptr(), 0 if pr is None else pr.rank, 1 if pr is None else pr.world,
                st), "clip_adamw_fused_streamed")
            if ar.shadow is not None:
                ar.mark_shadow_fresh()
            return
        if getattr(self, "peer", None) is not None:   # data parallel: gradient all-reduce over peer memory, in-kernel
            pr = self.peer
            tail_ws = pr.tail_ws
            _lib.check(self.lib.vitb200_clip_adamw_fused_dp(
                ar.data.data_ptr(), ar.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                None if ar.shadow is None else ar.shadow.data_ptr(), ar.layout.n_opt, self.hyper.data_ptr(),
                self.state.data_ptr(), self.rng.data_ptr(), self.gpart.data_ptr() if slots else None, slots,
                ar.layout.n_opt, start, end, tail_ws.data_ptr(), pr.table.data_ptr(), pr.rank, pr.world, st),
                "clip_adamw_fused_dp")
            if ar.shadow is not None:
                ar.mark_shadow_fresh()
            return
        _lib.check(self.lib.vitb200_clip_adamw_fused(
            ar.data.data_ptr(), ar.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            None if ar.shadow is None else ar.shadow.data_ptr(), ar.layout.n_opt, self.hyper.data_ptr(),
            self.state.data_ptr(), self.rng.data_ptr(), self.gpart.data_ptr() if slots else None, slots,
            ar.layout.n_opt, start, end, self.tail_ws.data_ptr(), st), "clip_adamw_fused")
        if ar.shadow is not None:
            ar.mark_shadow_fresh()

    def advance_rng(self) -> None:
        self.rng[1] += 1

    def set_lr(self, lr: float) -> None:
        self.hyper[0] = lr

    def set_grad_scale(self, s: float) -> None:
        self.hyper[6] = s

    def last_hidden_state(self) -> torch.Tensor:
        """Final LayerNorm over ALL rows (HF:455) -- only needed by callers that ask for it."""
        c = self.cfg
        M, H, Lh = self.M, c.hidden_size, c.num_hidden_layers
        out = torch.empty(M, H, dtype=self.act_dtype, device=self.device)
        tmp = torch.empty(2, M, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vitb200_add_ln_fwd(
            self.z[Lh].data_ptr(), None, None, out.data_ptr(), tmp[0].data_ptr(), tmp[1].data_ptr(),
            self._p("vit.layernorm.weight"), self._p("vit.layernorm.bias"), M, H, 0, float(c.layer_norm_eps), 0.0,
            self.rng.data_ptr(), 0, self.dt, st), "final ln")
        return out.view(self.B, c.tokens, H)

    def attention_probs(self, l: int) -> torch.Tensor:
        c = self.cfg
        B, T, H = self.B, c.tokens, c.hidden_size
        es = 2 if self.dt == BF16 else 4
        probs = torch.empty(B, c.num_attention_heads, T, T, dtype=torch.float32, device=self.device)
        qkv = self.qkv[l].data_ptr()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vitb200_attn_probs(
            qkv, qkv + H * es, 3 * H, self.lse[l].data_ptr(), probs.data_ptr(), self._ptr(self.rope_cos),
            self._ptr(self.rope_sin), B, T, c.num_attention_heads, c.head_dim, 1.0 / math.sqrt(c.head_dim), self.dt,
            st), "attn_probs")
        return probs

    def dropout_mask(self, site: int, n: int, p: float) -> torch.Tensor:
        """Test support: the keep-mask the kernels use for `site` at the CURRENT rng step."""
        out = torch.empty(n, dtype=torch.uint8, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.vitb200_dropout_mask(out.data_ptr(), n, p, self.rng.data_ptr(), site, st), "mask")
        return out

    def kernel_launches(self, train: bool = True, fused_tail: bool = False) -> int:
        """Kernel launches of one fwd+bwd+update step (for bench.py's gpu_launches), counted from the programs.
        fused_tail: the TrainStep single-GPU sequence (gradient-partial reduction folded into the optimizer kernel)."""
        skip = bool(fused_tail and self.fused_bwd)
        sh = bool(skip and self.can_fuse_head)
        co = bool(self.cls_only and self.mega)
        fkey = ("fwd", train, True, True, co) if sh else ("fwd", train, True, co)
        if fkey not in self._progs:
            self._progs[fkey] = self._build_forward_fused(train, True, head_bwd=True) if sh else self._build_forward(train, True)
        bkey = ("bwd", train, None, skip, sh, co) if (skip or sh) else ("bwd", train, None, co)
        if bkey not in self._progs:
            self._progs[bkey] = self._build_backward(train, None, skip_reduce=skip, skip_head=sh)
        per_call = {"vitb200_head_loss_fwd": 2, "vitb200_head_loss_bwd": 2, "vitb200_patch_embed_bwd": 2}
        n = 1  # clip + AdamW (one launch)
        for key in (fkey, bkey):
            for fn, args in self._progs[key]:
                k = per_call.get(fn.__name__, 1)
                if fn.__name__ == "vitb200_linear_wgrad" and self.dt == BF16 and \
                        self.lib.vitb200_tc_supported(args[4], args[5], args[6]):
                    k = 2  # column-sum (bias gradient) kernel + tensor-core wgrad
                if fn.__name__.startswith("vitb200_attn_tc_blocked"):
                    nb = (self.cfg.tokens + 127) // 128
                    k = nb + 1 if fn.__name__.endswith("fwd") else nb   # one launch per key block (+ the merge kernel)
                if fn.__name__ == "vitb200_attn_bwd":
                    tc = self.dt == BF16 and self.lib.vitb200_attn_tc_supported(
                        self.cfg.tokens, self.cfg.head_dim, 3 * self.cfg.hidden_size, self.cfg.hidden_size)
                    k = 1 if tc else 2  # tcgen05: one kernel; SIMT: dq pass + dk/dv pass
                n += k
        return n
