"""Device-resident dataset hand-off.

The reference keeps the whole training set in host RAM as fp32 tensors (src/dataloader/base.py:219-245), and every step
pays DataLoader collation in Python plus the host->device copy of `flux` AND `error` (32 KB per sample, the `error` half
unused when noise_level = 0; src/dataloader/base.py:299-300, src/basemodule.py:76-85).  With a step of ~150 us that
hand-off is the bottleneck, so here the tensors live in HBM (4096-pixel spectra: 16 KB per sample -> 10 M samples fit in
180 GB) and a batch is assembled by ONE kernel (`vitb200_gather_batch`): row gather by a device index vector, optional
noise injection `flux + N(0,1) * error * noise_level` (src/vit.py:86-88), labels alongside -- written straight into the
engine's input buffers.  `error` is only touched when noise is on.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def epoch_indices(n: int, epoch: int, seed: int = 0, shuffle: bool = True, rank: int = 0, world: int = 1,
                  batch: int = 1, tail: str = "wrap") -> torch.Tensor:
    """Sample order of one epoch for one rank (CPU int64), DistributedSampler semantics (Lightning DDP,
    src/hardware_utils.py:95): one permutation per epoch, identical on every rank (seeded by seed + epoch), padded by
    wrapping to a multiple of `world`, rank r takes elements r::world.  The rank's list is then made a multiple of
    `batch` (the step runs a fixed-shape CUDA graph): tail='wrap' re-uses the first samples of the list, tail='drop' drops
    the remainder (DataLoader drop_last=True)."""
    if n <= 0 or batch <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("epoch_indices: need n > 0, batch > 0, 0 <= rank < world")
    if shuffle:
        order = torch.randperm(n, generator=torch.Generator().manual_seed(int(seed) + int(epoch)))
    else:
        order = torch.arange(n)
    if world > 1:
        total = -(-n // world) * world
        if total > n:
            order = torch.cat([order, order[: total - n]])
        order = order[rank::world]
    m = order.numel()
    rem = m % batch
    if rem:
        if tail == "drop":
            order = order[: m - rem]
        elif tail == "wrap":
            reps = -(-(batch - rem) // m)
            order = torch.cat([order, order.repeat(reps)[: batch - rem]])
        else:
            raise ValueError("tail must be 'wrap' or 'drop'")
    return order.contiguous()


class DeviceDataset:
    """flux [N, L] fp32 (+ error [N, L]) and labels ([N] / [N, C] fp32 for regression, [N] int64 for classification)
    resident on the device."""

    def __init__(self, flux: torch.Tensor, labels: Optional[torch.Tensor], error: Optional[torch.Tensor] = None,
                 device=None):
        if flux.dim() != 2:
            raise ValueError("flux must be [N, L]")
        if flux.shape[1] % 4:
            raise ValueError("vit_b200: spectrum length must be a multiple of 4")
        dev = torch.device(device) if device is not None else flux.device
        if dev.type != "cuda":
            raise RuntimeError("vit_b200 has no CPU path: a DeviceDataset lives on a CUDA (sm_100a) device")
        self.flux = flux.to(dev, torch.float32).contiguous()
        self.device = dev = self.flux.device      # normalised ('cuda' -> 'cuda:0')
        self.error = None if error is None else error.to(dev, torch.float32).contiguous()
        if self.error is not None and self.error.shape != self.flux.shape:
            raise ValueError("error must have the shape of flux")
        self.labels = None
        self.label_bytes = 0
        if labels is not None:
            if labels.shape[0] != flux.shape[0]:
                raise ValueError("labels and flux disagree on N")
            lab = labels.to(dev)
            lab = lab.to(torch.int64) if lab.dtype in (torch.int64, torch.int32) else lab.to(torch.float32)
            self.labels = lab.reshape(flux.shape[0], -1).contiguous()
            self.label_bytes = self.labels.shape[1] * self.labels.element_size()

    def __len__(self) -> int:
        return self.flux.shape[0]

    @property
    def length(self) -> int:
        return self.flux.shape[1]

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.flux, self.error, self.labels) if t is not None)

    def gather(self, idx: Optional[torch.Tensor], x_out: torch.Tensor, y_out: Optional[torch.Tensor],
               noise_level: float = 0.0, rng: Optional[torch.Tensor] = None) -> None:
        """x_out[i] = flux[idx[i]] (+ noise), y_out[i] = labels[idx[i]] on the current stream.  idx: int64 on the device
        (None: the first len(x_out) rows).  rng: the engine's device {seed, step} pair (keys the noise)."""
        B = x_out.shape[0]
        if x_out.dtype != torch.float32 or not x_out.is_contiguous() or x_out.shape[1] != self.length:
            raise ValueError("x_out must be a contiguous fp32 [B, L] tensor")
        if idx is not None:
            if idx.dtype != torch.int64 or idx.device != self.device or idx.numel() != B or not idx.is_contiguous():
                raise ValueError("idx must be a contiguous int64 [B] tensor on the dataset's device")
        elif B > len(self):
            raise ValueError("batch larger than the dataset")
        want_y = y_out is not None and self.labels is not None
        if want_y and y_out.numel() * y_out.element_size() != B * self.label_bytes:
            raise ValueError("y_out does not match the label rows")
        if want_y and (y_out.dtype != self.labels.dtype):
            raise ValueError(f"y_out dtype {y_out.dtype} != label dtype {self.labels.dtype}")
        noisy = noise_level > 0 and self.error is not None
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.load().vitb200_gather_batch(
            self.flux.data_ptr(), self.error.data_ptr() if noisy else None,
            self.labels.data_ptr() if want_y else None, None if idx is None else idx.data_ptr(), x_out.data_ptr(),
            y_out.data_ptr() if want_y else None, B, self.length, self.label_bytes if want_y else 0, len(self),
            float(noise_level) if noisy else 0.0, None if rng is None else rng.data_ptr(), st), "gather_batch")


class EvalMetrics:
    """Running regression / classification metrics on the device (replaces the per-batch torchmetrics updates of
    src/vit.py:94-125; one `vitb200_eval_metrics_accum` launch per batch, no host synchronisation until `compute`)."""

    def __init__(self, num_labels: int, is_cls: bool, device):
        self.C, self.is_cls = int(num_labels), bool(is_cls)
        self.acc = torch.zeros(2 + 4 * max(self.C, 1), dtype=torch.float64, device=device)

    def reset(self) -> None:
        self.acc.zero_()

    def update(self, logits: torch.Tensor, labels: torch.Tensor, loss: Optional[torch.Tensor] = None) -> None:
        B = logits.shape[0]
        st = torch.cuda.current_stream(logits.device).cuda_stream
        _lib.check(_lib.load().vitb200_eval_metrics_accum(
            logits.data_ptr(), labels.data_ptr(), None if loss is None else loss.data_ptr(), self.acc.data_ptr(), B,
            self.C, 1 if self.is_cls else 0, st), "eval_metrics_accum")

    def compute(self) -> dict:
        """One device->host read.  MAE / MSE over all elements, R2 = uniform average over outputs of
        1 - SS_res / SS_tot (torchmetrics MeanAbsoluteError / MeanSquaredError / R2Score defaults), accuracy for cls."""
        a = self.acc.cpu().tolist()
        n = a[0]
        out = {"n": int(n), "loss": a[1] / n if n else float("nan")}
        if n == 0:
            return out
        if self.is_cls:
            out["acc"] = a[2] / n
            return out
        C = self.C
        out["mae"] = sum(a[2 + 4 * c] for c in range(C)) / (n * C)
        out["mse"] = sum(a[3 + 4 * c] for c in range(C)) / (n * C)
        r2 = []
        for c in range(C):
            ss_res, sy, syy = a[3 + 4 * c], a[4 + 4 * c], a[5 + 4 * c]
            ss_tot = syy - sy * sy / n
            r2.append(1.0 - ss_res / ss_tot if ss_tot > 0 else float("nan"))
        out["r2"] = sum(r2) / C
        return out
