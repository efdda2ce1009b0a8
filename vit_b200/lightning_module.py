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
            # torchmetrics.R2Score() (src/vit.py:72,120): R^2 per output column, uniformly averaged
            e2, l2 = (err.reshape(len(err), -1), lab.reshape(len(lab), -1)) if err.dim() > 0 else (err.reshape(1, 1), lab.reshape(1, 1))
            ss_res = (e2 * e2).sum(0)
            ss_tot = ((l2 - l2.mean(0, keepdim=True)) ** 2).sum(0).clamp_min(1e-12)
            self.log(f"{prefix}_r2", (1.0 - ss_res / ss_tot).mean(), on_step=False, on_epoch=True)
        self._last_outputs = outputs
        return loss

    # The reference's validation_step / test_step run the model a SECOND time per batch just to get the predictions for
    # the epoch-level statistics (src/vit.py:127-150,194-215).  Here the outputs of the one forward in
    # `_shared_eval_step` are kept, so every batch costs one forward.
    def _collect(self, store: str, batch) -> None:
        if self.task_type != "reg" or not hasattr(self, store):
            return
        preds = self._last_outputs.logits.squeeze()
        getattr(self, store)["preds"].append(preds.detach().cpu())
        getattr(self, store)["labels"].append(batch[-1].detach().cpu())

    def validation_step(self, batch, batch_idx):
        loss = self._shared_eval_step(batch, "val")
        self._collect("val_dict", batch)
        return loss

    def on_validation_start(self):  # src/vit.py:152-155
        if self.task_type == "reg":
            self.val_dict = {"preds": [], "labels": []}

    def on_validation_epoch_end(self):
        """Epoch-level regression statistics per output (src/vit.py:157-192): median residual, 90th percentile of
        |residual|, slope of the linear fit pred = a + beta * label."""
        if self.task_type != "reg" or not getattr(self, "val_dict", None) or not self.val_dict["preds"]:
            return
        import numpy as np

        preds = torch.cat([p.reshape(p.shape[0] if p.dim() else 1, -1) for p in self.val_dict["preds"]], dim=0).numpy()
        labels = torch.cat([l.reshape(l.shape[0] if l.dim() else 1, -1) for l in self.val_dict["labels"]], dim=0).numpy()
        for i in range(preds.shape[1]):
            res = preds[:, i] - labels[:, i]
            beta = float(np.polyfit(labels[:, i], preds[:, i], 1)[0])
            suffix = "" if preds.shape[1] == 1 else f"_{i}"
            self.log(f"val_bias_median{suffix}", float(np.median(res)), on_epoch=True)
            self.log(f"val_p90{suffix}", float(np.percentile(np.abs(res), 90)), on_epoch=True)
            self.log(f"val_beta{suffix}", beta, on_epoch=True)
        self.val_dict = {"preds": [], "labels": []}

    def on_test_start(self):  # src/vit.py:194-197
        if self.task_type == "reg":
            self.test_dict = {"preds": [], "labels": []}

    def test_step(self, batch, batch_idx):
        loss = self._shared_eval_step(batch, "test")
        self._collect("test_dict", batch)
        return loss

    def on_test_epoch_end(self):
        """src/vit.py:216-296: hand the collected predictions to the reference's RegressionPlotter when this module runs
        inside the reference repository (plotting itself is the reference's code, not part of the hot path)."""
        if self.task_type != "reg" or not getattr(self, "test_dict", None) or not self.test_dict["preds"]:
            return
        try:
            from src.viz import RegressionPlotter  # the reference package, when present
        except Exception:  # noqa: BLE001
            return
        preds = torch.cat(self.test_dict["preds"], dim=0).numpy()
        labels = torch.cat(self.test_dict["labels"], dim=0).numpy()
        n_out = preds.shape[1] if preds.ndim > 1 else 1
        names = (self.config.get("model", {}) or {}).get("param_names", ["Teff", "log_g", "M_H"][:n_out])
        norm = {}
        ds = getattr(getattr(getattr(self, "trainer", None), "datamodule", None), "test", None)
        if ds is not None:
            norm["label_norm"] = getattr(ds, "label_norm", None)
            for k in ("label_mean", "label_std", "label_min", "label_max"):
                norm[k] = getattr(ds, k, None)
        RegressionPlotter(predictions=preds, labels=labels, param_names=names, logger=getattr(self, "logger", None),
                          save_dir=(self.config.get("train", {}) or {}).get("plot_dir", "plots"), **norm
                          ).generate_all_plots(quick_mode=(self.config.get("plotting", {}) or {}).get("quick_mode", False))

    def configure_optimizers(self):
        """src/basemodule.py:152-182 + OptModule (src/opt/optimizer.py:37-172): same config keys (`opt.type`, `lr`,
        `weight_decay`, `lr_sch` in {plateau, cosine*, onecycle, constant*}, `warmup`/`warmup_ratio`/`warmup_epochs`, scheduler
        arguments) and the same Lightning scheduler dicts.  `opt.fused: true` with AdamW selects the one-arena fused
        optimizer instead of torch's."""
        opt = {**(self.config.get("opt", {}) or {})}
        data_cfg = self.config.get("data", {}) or {}
        lr = opt.get("lr", 1e-3)
        wd = opt.get("weight_decay", 0)
        kind = str(opt.get("type", "adam")).lower()
        sch = str(opt.get("lr_sch", "") or "").lower()
        if "plateau" in sch and not bool(data_cfg.get("val_path")):
            print("[WARNING] ReduceLROnPlateau requires validation data ('data.val_path' in config) but none configured.")
            print("[WARNING] Disabling learning rate scheduler.")
            sch = ""
        if opt.get("fused", False) and kind == "adamw":
            from .optim import FusedClipAdamW

            optimizer = FusedClipAdamW(self.model, lr=lr, weight_decay=wd, max_norm=0.0)
        else:
            fns = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD,
                   "rmsprop": torch.optim.RMSprop, "adadelta": torch.optim.Adadelta, "adagrad": torch.optim.Adagrad,
                   "adamax": torch.optim.Adamax, "asgd": torch.optim.ASGD, "lbfgs": torch.optim.LBFGS,
                   "rprop": torch.optim.Rprop, "sparseadam": torch.optim.SparseAdam}   # src/opt/optimizer.py:14-26
            # (like the reference, every optimizer is built with lr + weight_decay: the ones whose constructor has no
            #  weight_decay -- lbfgs, rprop, sparseadam -- raise TypeError there too)
            optimizer = fns[kind](self.model.parameters(), lr=lr, weight_decay=wd)
        if not sch:
            return optimizer
        S = torch.optim.lr_scheduler
        table = {"cosine": S.CosineAnnealingLR, "cosineannealing": S.CosineAnnealingLR, "cosineannealinglr": S.CosineAnnealingLR,
                 "onecycle": S.OneCycleLR,
                 "constant": S.ConstantLR, "constantlr": S.ConstantLR, "plateau": S.ReduceLROnPlateau}
        if sch not in table:
            raise ValueError(f"Unknown scheduler: {sch}")
        kw = {}
        if "cosine" in sch:
            kw["T_max"] = opt.get("T_max", opt.get("ep", 100))
            if "eta_min" in opt:
                kw["eta_min"] = opt["eta_min"]
        elif "onecycle" in sch:   # steps_per_epoch / epochs from the train + data config (src/basemodule.py:167-179)
            train_cfg = self.config.get("train", {}) or {}
            bs = train_cfg.get("batch_size", 64)
            kw.update(max_lr=lr, steps_per_epoch=(data_cfg.get("num_samples", 32000) + bs - 1) // bs,
                      epochs=train_cfg.get("ep", 100))
            for k in ("pct_start", "div_factor", "final_div_factor"):
                if k in opt:
                    kw[k] = opt[k]
        elif "constant" in sch:
            kw.update(factor=opt.get("factor", 1.0), total_iters=opt.get("total_iters", 1))
        elif "plateau" in sch:
            kw.update(factor=opt.get("factor", 0.1), patience=opt.get("patience", 10))
            if "mode" in opt:
                kw["mode"] = opt["mode"]
        wcfg = opt.get("warmup", {}) or {}
        w_ratio = wcfg.get("ratio", opt.get("warmup_ratio", 0.0))
        w_epochs = wcfg.get("epochs", opt.get("warmup_epochs", None))
        if (w_ratio > 0 or w_epochs is not None) and "onecycle" not in sch:
            if w_epochs is None:
                w_epochs = max(1, int(kw.get("T_max", kw.get("epochs", 100)) * w_ratio))
            scheduler = S.SequentialLR(optimizer, schedulers=[S.LinearLR(optimizer, start_factor=0.1, total_iters=w_epochs),
                                                               table[sch](optimizer, **kw)], milestones=[w_epochs])
            print(f"[Warmup] Using {w_epochs} warmup epochs before {sch}")
        else:
            scheduler = table[sch](optimizer, **kw)
        cfg = {"scheduler": scheduler, "monitor": f"val_{self.monitor_metric}"}
        if "plateau" in sch:
            cfg.update(reduce_on_plateau=True, strict=False)
        elif "onecycle" in sch:
            cfg.update(interval="step", frequency=1)
        else:
            cfg.update(interval="epoch", frequency=1)
        return {"optimizer": optimizer, "lr_scheduler": cfg}
