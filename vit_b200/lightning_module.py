"""ViTLModule -- the reference's LightningModule surface (src/vit.py:58-215, src/basemodule.py:143-196)
around the B200 MyViT.  `lightning` / `torchmetrics` are optional: without them the class is a plain
nn.Module with the same methods, so `training_step` / `validation_step` / `configure_optimizers` can be
driven by any loop (and by the tests).
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:  # pragma: no cover - lightning is not installed in the build image
    import lightning as L

    _Base = L.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _Base = nn.Module
    HAVE_LIGHTNING = False

from .builder import get_model


def _normalize_task(config):
    m = (config.get("model", {}) or {})
    task = (m.get("task_type") or m.get("task") or "cls").lower()
    return "cls" if task in ("classification", "cls", "class") else "reg"


class ViTLModule(_Base):
    def __init__(self, model=None, config=None):
        super().__init__()
        config = config if config is not None else {}
        self.model = model or self.get_model(config)
        self.loss_name = self.model.loss_name
        self.callbacks = []
        self.config = config
        self.sweep = False
        self.task_type = _normalize_task(config)
        self.noise_level = (config.get("noise", {}) or {}).get("noise_level", 0.0)
        self.monitor_metric = "acc" if self.task_type == "cls" else "mae"
        self._logged = {}
        if HAVE_LIGHTNING:
            self.save_hyperparameters(ignore=["model"])

    def get_model(self, config):
        return get_model(config)

    if not HAVE_LIGHTNING:
        def log(self, name, value, **kw):  # Lightning's self.log stand-in: keeps the last value
            self._logged[name] = value

    def forward(self, flux, labels, loss_only=True):  # src/vit.py:78-81
        outputs = self.model(flux, labels=labels)
        return outputs.loss if loss_only else outputs

    def training_step(self, batch, batch_idx):  # src/vit.py:83-92
        flux, error, labels = batch
        if self.noise_level > 0:
            noisy = flux + torch.randn_like(flux) * error * self.noise_level
            loss = self(noisy, labels, loss_only=True)
        else:
            loss = self(flux, labels, loss_only=True)
        self.log(f"{self.loss_name}_loss", loss, on_step=True, on_epoch=True, prog_bar=True)
        return loss

    def _shared_eval_step(self, batch, prefix):  # src/vit.py:94-125 (metrics computed on the device)
        if len(batch) == 4:
            noisy, flux, error, labels = batch
            outputs = self.forward(noisy if self.noise_level > 0 else flux, labels, loss_only=False)
        else:
            flux, error, labels = batch
            outputs = self.forward(flux, labels, loss_only=False)
        loss = outputs.loss
        self.log(f"{prefix}_{self.loss_name}_loss", loss, on_step=False, on_epoch=True)
        if self.task_type == "cls":
            acc = (outputs.logits.argmax(-1) == labels).float().mean()
            self.log(f"{prefix}_acc", acc, on_step=False, on_epoch=True, prog_bar=True)
        else:
            preds = outputs.logits.squeeze()
            lab = labels.reshape(preds.shape).float()
            err = preds - lab
            self.log(f"{prefix}_mae", err.abs().mean(), on_step=False, on_epoch=True)
            self.log(f"{prefix}_mse", (err * err).mean(), on_step=False, on_epoch=True)
            ss_tot = ((lab - lab.mean()) ** 2).sum().clamp_min(1e-12)
            self.log(f"{prefix}_r2", 1.0 - (err * err).sum() / ss_tot, on_step=False, on_epoch=True)
        self._last_outputs = outputs
        return loss

    def validation_step(self, batch, batch_idx):
        return self._shared_eval_step(batch, "val")

    def test_step(self, batch, batch_idx):
        return self._shared_eval_step(batch, "test")

    def configure_optimizers(self):  # src/basemodule.py:152-182 + src/opt/optimizer.py:37-172
        opt = {**(self.config.get("opt", {}) or {})}
        lr = float(opt.get("lr", 1e-3))
        wd = opt.get("weight_decay", 0)
        kind = str(opt.get("type", "adam")).lower()
        if opt.get("fused", False) and kind == "adamw":
            from .optim import FusedClipAdamW

            optimizer = FusedClipAdamW(self.model, lr=lr, weight_decay=wd, max_norm=0.0)
        else:
            fns = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD,
                   "rmsprop": torch.optim.RMSprop, "adagrad": torch.optim.Adagrad}
            optimizer = fns[kind](self.model.parameters(), lr=lr, weight_decay=wd)
        sch = str(opt.get("lr_sch", "") or "").lower()
        has_val = bool((self.config.get("data", {}) or {}).get("val_path"))
        if not sch or ("plateau" in sch and not has_val):
            return optimizer
        if "plateau" in sch:
            scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=opt.get("factor", 0.1),
                                                                   patience=opt.get("patience", 10))
            cfg = {"scheduler": scheduler, "monitor": f"val_{self.monitor_metric}", "reduce_on_plateau": True,
                   "strict": False}
        elif "cosine" in sch:
            scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(
                optimizer, T_max=opt.get("T_max", opt.get("ep", 100)), eta_min=opt.get("eta_min", 0.0))
            cfg = {"scheduler": scheduler, "monitor": f"val_{self.monitor_metric}", "interval": "epoch", "frequency": 1}
        else:
            return optimizer
        return {"optimizer": optimizer, "lr_scheduler": cfg}
