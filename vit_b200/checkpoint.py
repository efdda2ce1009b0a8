"""Checkpoint wire format: Lightning `.ckpt` files of the reference <-> the flat arena.

The reference saves through Lightning's ModelCheckpoint (src/vit.py:387-414) and resumes with
`trainer.fit(..., ckpt_path=...)` / `scripts/run.py --ckpt` / `scripts/test.py --ckpt` (src/vit.py:464, scripts/test.py:48).
A `.ckpt` is a `torch.save`d dict:

    state_dict        LightningModule.state_dict(): the model's keys under the `model.` prefix (src/basemodule.py:146)
    optimizer_states  [torch.optim.AdamW.state_dict()]: {'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}},
                      'param_groups': [{'lr', 'betas', 'eps', 'weight_decay', ..., 'params': [0, 1, ...]}]}
                      where i indexes `LightningModule.parameters()` in order; parameters that never received a
                      gradient (vit.pooler.dense.*) have no entry
    epoch, global_step, lr_schedulers, callbacks, loops, hyper_parameters, pytorch-lightning_version

Here the parameters are views into one flat fp32 arena and AdamW's moments are two flat buffers of the same layout
(`ViTEngine.exp_avg / exp_avg_sq`), with ONE step counter (`ViTEngine.state[0]`; every optimised tensor is updated on every
step, so torch's per-tensor `step` values are all equal to it).  The functions below move state between the two
layouts, in both directions, without touching kernels (so they are testable on CPU).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

PREFIX = "model."


def strip_prefix(state_dict: Dict[str, torch.Tensor], prefix: str = PREFIX) -> Dict[str, torch.Tensor]:
    """LightningModule keys -> model keys (non-matching keys, e.g. torchmetrics buffers, are dropped)."""
    return {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}


def adam_state_from_torch(opt_state: dict, param_names: List[str], layout, exp_avg: torch.Tensor,
                          exp_avg_sq: torch.Tensor) -> Tuple[int, dict, Dict[str, dict]]:
    """torch.optim.Adam(W) state_dict -> flat moment buffers.  `param_names`: the model's `named_parameters()` order (the
    index space of the optimizer state).  Returns (step, param_group hyper-parameters, state of parameters that live
    outside the arena, e.g. a trainable preprocessor, keyed by name)."""
    state = opt_state.get("state", {})
    groups = opt_state.get("param_groups", [])
    if len(groups) != 1:
        raise ValueError(f"expected ONE param group (src/opt/optimizer.py:108), found {len(groups)}")
    order = list(groups[0]["params"])
    if len(order) != len(param_names):
        raise ValueError(f"optimizer covers {len(order)} parameters, the model has {len(param_names)}")
    exp_avg.zero_()
    exp_avg_sq.zero_()
    steps, extra = set(), {}
    for pos, idx in enumerate(order):
        st = state.get(idx, state.get(str(idx)))
        if not st:
            continue                       # never received a gradient (pooler): torch keeps no state for it
        name = param_names[pos]
        e = layout.entries.get(name)
        if e is None:
            extra[name] = st
            continue
        if e.offset >= layout.n_opt:
            raise ValueError(f"{name} has optimizer state but lies outside the optimised range of the arena")
        if st["exp_avg"].numel() != e.numel:
            raise ValueError(f"{name}: optimizer state has {st['exp_avg'].numel()} elements, expected {e.numel}")
        exp_avg[e.offset:e.offset + e.numel].copy_(st["exp_avg"].reshape(-1))
        exp_avg_sq[e.offset:e.offset + e.numel].copy_(st["exp_avg_sq"].reshape(-1))
        steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"per-parameter step counters differ ({sorted(steps)}): the fused optimizer keeps ONE counter")
    hyper = {k: v for k, v in groups[0].items() if k != "params"}
    return (steps.pop() if steps else 0), hyper, extra


def adam_state_to_torch(param_names: List[str], layout, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
                        lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                        extra: Optional[Dict[str, dict]] = None) -> dict:
    """Flat moment buffers -> the state_dict a `torch.optim.AdamW` over `model.parameters()` would produce, so that the
    reference (Lightning) can resume from it.  Arena parameters outside the optimised range (pooler) get no entry."""
    proto = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=lr, betas=tuple(betas), eps=eps,
                              weight_decay=weight_decay).state_dict()["param_groups"][0]
    group = dict(proto)
    group["params"] = list(range(len(param_names)))
    state = {}
    for i, name in enumerate(param_names):
        e = layout.entries.get(name)
        if e is None:
            if extra and name in extra:
                state[i] = extra[name]
            continue
        if e.offset >= layout.n_opt or step <= 0:
            continue
        state[i] = {"step": torch.tensor(float(step)),
                    "exp_avg": exp_avg[e.offset:e.offset + e.numel].detach().reshape(e.shape).cpu().clone(),
                    "exp_avg_sq": exp_avg_sq[e.offset:e.offset + e.numel].detach().reshape(e.shape).cpu().clone()}
    return {"state": state, "param_groups": [group]}


def load_lightning_checkpoint(model, ckpt, train_step=None, strict: bool = True) -> dict:
    """Load a reference `.ckpt` (path or already-loaded dict) into `model` (MyViT); with `train_step` (TrainStep) also
    the AdamW moments, step counter and learning rate, so that training resumes where the reference stopped.
    Returns {'epoch', 'global_step', 'step', 'extra_optimizer_state'}."""
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd = ckpt.get("state_dict", ckpt)
    if any(k.startswith(PREFIX) for k in sd):
        sd = strip_prefix(sd)
    model.load_state_dict(sd, strict=strict)
    info = {"epoch": ckpt.get("epoch"), "global_step": ckpt.get("global_step"), "step": None,
            "extra_optimizer_state": {}, "lr_schedulers": list(ckpt.get("lr_schedulers") or [])}
    opt_states = ckpt.get("optimizer_states") or []
    if train_step is not None and opt_states:
        eng = train_step.eng
        eng._ensure_opt_state()
        names = [n for n, _ in model.named_parameters()]
        step, hyper, extra = adam_state_from_torch(opt_states[0], names, eng.arena.layout, eng.exp_avg, eng.exp_avg_sq)
        eng.state[0] = float(step)
        h = model._opt_hyper
        if "lr" in hyper:
            h["lr"] = float(hyper["lr"])
            eng.set_lr(h["lr"])
        if "betas" in hyper:
            h["betas"] = (float(hyper["betas"][0]), float(hyper["betas"][1]))
            eng.hyper[1], eng.hyper[2] = h["betas"]
        if "eps" in hyper:
            h["eps"] = float(hyper["eps"])
            eng.hyper[3] = h["eps"]
        if "weight_decay" in hyper:
            h["weight_decay"] = float(hyper["weight_decay"])
            eng.hyper[4] = h["weight_decay"]
        eng.refresh_shadow(force=True)
        pre_tr = getattr(train_step, "_pre_tr", None)
        if pre_tr is not None:      # a trainable preprocessor trained inside the step: its moments live next to the step
            pre_tr.load_optimizer_state(extra)
        info.update(step=step, extra_optimizer_state=extra)
    return info


def save_lightning_checkpoint(model, path=None, train_step=None, epoch: int = 0, global_step: Optional[int] = None,
                              hyper_parameters: Optional[dict] = None, lr_schedulers: Optional[list] = None) -> dict:
    """Write (and return) a dict the reference's Lightning trainer can load: `model.`-prefixed state_dict, and -- with
    `train_step` -- the AdamW state in torch's per-parameter layout."""
    sd = {PREFIX + k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ckpt = {"epoch": int(epoch), "global_step": 0 if global_step is None else int(global_step),
            "pytorch-lightning_version": "2.0.0", "state_dict": sd, "callbacks": {},
            "optimizer_states": [], "lr_schedulers": list(lr_schedulers or []),
            "hyper_parameters": {"config": hyper_parameters or {}}}
    # No "loops" key on purpose: Lightning's _CheckpointConnector.restore_loops() reads ckpt["loops"]["fit_loop"] whenever
    # the key exists, and falls back to the plain epoch / global_step fields (the pre-1.6 format) when it does not.
    if train_step is not None:
        eng = train_step.eng
        eng._ensure_opt_state()
        h = model._opt_hyper   # the python-side copies of the device scalars (exact, not fp32-rounded)
        step = int(float(eng.state[0]))
        names = [n for n, _ in model.named_parameters()]
        pre_tr = getattr(train_step, "_pre_tr", None)
        ckpt["optimizer_states"] = [adam_state_to_torch(names, eng.arena.layout, eng.exp_avg, eng.exp_avg_sq, step,
                                                        lr=h["lr"], betas=tuple(h["betas"]), eps=h["eps"],
                                                        weight_decay=h["weight_decay"],
                                                        extra=None if pre_tr is None else pre_tr.optimizer_state(step))]
        if global_step is None:
            ckpt["global_step"] = step
    if path is not None:
        torch.save(ckpt, path)
    return ckpt
