"""FusedClipAdamW -- clip_grad_norm_ + torch.optim.AdamW as two kernels over the flat arena.

Replaces Lightning's `gradient_clip_val` (src/basemodule.py:244 -> torch.nn.utils.clip_grad_norm_) followed
by `torch.optim.AdamW.step` (src/opt/optimizer.py:108).  It is a `torch.optim.Optimizer`, so LR schedulers
(ReduceLROnPlateau etc., src/opt/optimizer.py:94-99) and Lightning drive it through `param_groups[0]['lr']`.
Pass `max_norm=0` when the trainer already clips (Lightning's gradient_clip_val).
"""
from __future__ import annotations

import torch

from . import _lib


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_norm: float = 0.5):
        self.model = model
        params = [p for p in model.parameters()]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm))
        self._state = None

    def _buffers(self):
        ar = self.model._arena
        if self._state is None or self._state["arena"] is not ar:
            dev = ar.data.device
            n = ar.layout.n_total
            self._state = dict(
                arena=ar, m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
                hyper=torch.zeros(8, device=dev), st=torch.zeros(8, device=dev),
                ws=torch.zeros(int(_lib.load().vitb200_grad_norm_ws_bytes(n)) + 4096, dtype=torch.uint8, device=dev))
        return self._state

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        s = self._buffers()
        ar = s["arena"]
        inside = {id(p) for p in self.model._param_list}
        outside = [p for p in self.param_groups[0]["params"] if id(p) not in inside and p.requires_grad]
        if outside:
            raise NotImplementedError(
                "vit_b200: FusedClipAdamW updates the flat parameter arena only; this model has trainable parameters "
                "outside it (an unfrozen input preprocessor) -- use torch.optim.AdamW over model.parameters(), or "
                "freeze the preprocessor (model.set_preprocessor_trainable(False))")
        if ar.data.device.type != "cuda":
            raise RuntimeError("vit_b200 has no CPU path")
        g = self.param_groups[0]
        lay = ar.layout
        skipped = []
        # gather p.grad into the flat gradient arena (no-op copy when p.grad already aliases it)
        for name, p in zip(self.model._param_names, self.model._param_list):
            e = lay.entries[name]
            if e.offset >= lay.n_opt:
                continue
            dst = ar.grad[e.offset:e.offset + e.numel]
            if p.grad is None or not p.requires_grad:
                # torch.optim.AdamW skips such a parameter entirely (no decay, no moment update).  A zero gradient with
                # weight_decay == 0 moves nothing either as long as its moments are zero, which they stay; with weight
                # decay the parameter is restored after the launch (below).
                dst.zero_()
                skipped.append((e, p))
            elif p.grad.data_ptr() != dst.data_ptr():
                dst.copy_(p.grad.reshape(-1))
        keep = [(e, ar.data[e.offset:e.offset + e.numel].clone(), s["m"][e.offset:e.offset + e.numel].clone(),
                 s["v"][e.offset:e.offset + e.numel].clone()) for e, _ in skipped]
        s["hyper"].copy_(torch.tensor([g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"],
                                       g["max_norm"], 1.0, 0.0]), non_blocking=True)
        st = torch.cuda.current_stream(ar.data.device).cuda_stream
        _lib.check(lib.vitb200_grad_norm(ar.grad.data_ptr(), lay.n_opt, s["hyper"].data_ptr(), s["st"].data_ptr(),
                                         s["ws"].data_ptr(), st), "grad_norm")
        _lib.check(lib.vitb200_adamw(ar.data.data_ptr(), ar.grad.data_ptr(), s["m"].data_ptr(), s["v"].data_ptr(),
                                     None if ar.shadow is None else ar.shadow.data_ptr(), lay.n_opt,
                                     s["hyper"].data_ptr(), s["st"].data_ptr(), None, st), "adamw")
        for e, pv, mv, vv in keep:   # parameters torch would have skipped: nothing about them changes
            ar.data[e.offset:e.offset + e.numel].copy_(pv)
            s["m"][e.offset:e.offset + e.numel].copy_(mv)
            s["v"][e.offset:e.offset + e.numel].copy_(vv)
        if keep and ar.shadow is not None:
            for e, pv, _, _ in keep:
                ar.shadow[e.offset:e.offset + e.numel].copy_(pv)
        if ar.shadow is not None:
            ar.mark_shadow_fresh()
        return loss

    # ---- checkpointing: the moments live in flat buffers, the wire format is torch.optim.AdamW's -------------------
    def state_dict(self) -> dict:
        """Same structure `torch.optim.AdamW(model.parameters())` produces (per-parameter step / exp_avg / exp_avg_sq), so
        Lightning's ModelCheckpoint stores the real optimizer state and the reference can resume from it."""
        from .checkpoint import adam_state_to_torch

        g = self.param_groups[0]
        names = [n for n, _ in self.model.named_parameters()]
        if self._state is None:
            step, m, v = 0, None, None
        else:
            step, m, v = int(float(self._state["st"][0])), self._state["m"], self._state["v"]
        ar = self.model._arena
        if m is None:
            m = v = torch.zeros(1)
        sd = adam_state_to_torch(names, ar.layout, m, v, step, lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"],
                                 weight_decay=g["weight_decay"])
        sd["param_groups"][0]["max_norm"] = g["max_norm"]
        return sd

    def load_state_dict(self, state_dict: dict) -> None:
        from .checkpoint import adam_state_from_torch

        s = self._buffers() if self.model._arena.data.device.type == "cuda" else None
        names = [n for n, _ in self.model.named_parameters()]
        ar = self.model._arena
        if s is None:   # CPU-side bookkeeping only (tests): keep flat copies so that state_dict() round-trips
            n = ar.layout.n_total
            self._state = s = dict(arena=ar, m=torch.zeros(n), v=torch.zeros(n), hyper=torch.zeros(8), st=torch.zeros(8),
                                   ws=None)
        step, hyper, _ = adam_state_from_torch(state_dict, names, ar.layout, s["m"], s["v"])
        s["st"][0] = float(step)
        g = self.param_groups[0]
        for k in ("lr", "eps", "weight_decay", "max_norm"):
            if k in hyper:
                g[k] = hyper[k]
        if "betas" in hyper:
            g["betas"] = tuple(hyper["betas"])

    @property
    def last_grad_norm(self) -> torch.Tensor:
        return self._buffers()["st"][1]
