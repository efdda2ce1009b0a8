"""Flat parameter arena: every parameter of the model lives in ONE contiguous fp32 buffer (plus a
matching gradient buffer and, in bf16 mode, a bf16 shadow used as GEMM operands).

Why: one clip+AdamW launch over the whole model, one (or L+2 bucketed) NCCL all-reduce(s) instead of
57, Q/K/V weights adjacent so the three projections run as one [3H, H] GEMM while `state_dict` still
exposes the reference's separate `query/key/value` tensors.
The unused `vit.pooler.dense.*` parameters (never receive a gradient in the reference) sit after the
optimised range so they are neither all-reduced nor updated.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

ALIGN = 64  # elements: 256 B in fp32, 128 B in bf16


def _round_up(x: int, a: int = ALIGN) -> int:
    return (x + a - 1) // a * a


@dataclass
class Entry:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int


class ParamLayout:
    def __init__(self, cfg):
        H, I, P, T = cfg.hidden_size, cfg.intermediate_size, cfg.patch_size, cfg.tokens
        self.entries: Dict[str, Entry] = {}
        self.buckets: List[Tuple[str, int, int]] = []  # (name, start, end) in forward order
        self._cur = 0

        def add(name, shape, align=True):
            if align:
                self._cur = _round_up(self._cur)
            n = 1
            for s in shape:
                n *= s
            self.entries[name] = Entry(name, tuple(shape), self._cur, n)
            self._cur += n

        def bucket(name, start):
            self._cur = _round_up(self._cur)
            self.buckets.append((name, start, self._cur))

        s0 = 0
        emb = "vit.embeddings."
        add(emb + "cls_token", (1, 1, H))
        if cfg.pos_encoding_type == "learned":
            add(emb + "position_embeddings", (1, T, H))
        add(emb + "patch_embeddings.projection.weight", (H, P) if cfg.proj_fn == "SW" else (H, 1, P))
        add(emb + "patch_embeddings.projection.bias", (H,))
        bucket("embeddings", s0)
        for i in range(cfg.num_hidden_layers):
            s0 = self._cur
            pre = f"vit.encoder.layer.{i}."
            add(pre + "layernorm_before.weight", (H,))
            add(pre + "layernorm_before.bias", (H,))
            add(pre + "attention.attention.query.weight", (H, H))
            add(pre + "attention.attention.key.weight", (H, H), align=False)
            add(pre + "attention.attention.value.weight", (H, H), align=False)
            add(pre + "attention.attention.query.bias", (H,))
            add(pre + "attention.attention.key.bias", (H,), align=False)
            add(pre + "attention.attention.value.bias", (H,), align=False)
            add(pre + "attention.output.dense.weight", (H, H))
            add(pre + "attention.output.dense.bias", (H,))
            add(pre + "layernorm_after.weight", (H,))
            add(pre + "layernorm_after.bias", (H,))
            add(pre + "intermediate.dense.weight", (I, H))
            add(pre + "intermediate.dense.bias", (I,))
            add(pre + "output.dense.weight", (H, I))
            add(pre + "output.dense.bias", (H,))
            bucket(f"layer{i}", s0)
        s0 = self._cur
        add("vit.layernorm.weight", (H,))
        add("vit.layernorm.bias", (H,))
        head = "classifier" if cfg.task_type == "cls" else "regressor"
        add(head + ".weight", (cfg.num_labels, H))
        add(head + ".bias", (cfg.num_labels,))
        bucket("head", s0)
        self.n_opt = self._cur  # everything before here receives gradients and is optimised
        add("vit.pooler.dense.weight", (H, H))
        add("vit.pooler.dense.bias", (H,))
        self.n_total = _round_up(self._cur)
        self.head_name = head

    def off(self, name: str) -> int:
        return self.entries[name].offset


class ParamArena:
    """data / grad (fp32) and shadow (bf16, optional) flat buffers + per-parameter views."""

    def __init__(self, layout: ParamLayout, device, with_shadow: bool):
        self.layout = layout
        self.device = torch.device(device)
        self.data = torch.zeros(layout.n_total, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(layout.n_total, dtype=torch.float32, device=self.device)
        self.shadow = (torch.zeros(layout.n_total, dtype=torch.bfloat16, device=self.device)
                       if with_shadow else None)
        self._shadow_version = -1
        self.watch = []   # the module Parameters that alias `data` (set by MyViT._bind_arena)

    def view(self, name: str, buf: torch.Tensor | None = None) -> torch.Tensor:
        e = self.layout.entries[name]
        buf = self.data if buf is None else buf
        return buf[e.offset:e.offset + e.numel].view(e.shape)

    def grad_view(self, name: str) -> torch.Tensor:
        return self.view(name, self.grad)

    def to(self, device, with_shadow: bool | None = None) -> "ParamArena":
        new = ParamArena(self.layout, device, self.shadow is not None if with_shadow is None else with_shadow)
        new.data.copy_(self.data)
        new.grad.copy_(self.grad)
        return new

    # ---- bf16 shadow maintenance -------------------------------------------------------------
    def _version(self) -> int:
        # Parameters rebound with `p.data = view` (MyViT._apply after .to()/.cuda()) keep their OWN version counters:
        # an in-place update through such a Parameter (torch.optim step, load_state_dict) does not bump data._version.
        # The sum over both moves whenever any alias is written.
        return self.data._version + sum(p._version for p in self.watch)

    def shadow_stale(self) -> bool:
        return self.shadow is not None and self._shadow_version != self._version()

    def mark_shadow_fresh(self) -> None:
        self._shadow_version = self._version()
