"""Input preprocessors of the reference (SURVEY.md 8a rows a2/a2', 8f rank 3) on the B200 kernels.

Same classes, attribute names and `state_dict` keys as the reference:

  * `compute_zca_matrix`, `compute_pca_matrix`   src/models/preprocessor.py:12-90   (one-time, at model build)
  * `PrefilledLinear`                            src/models/layers.py:12-63        (buffer when frozen, Parameter when not)
  * `LinearPreprocessor`                         src/models/preprocessor.py:93-111
  * `PrefilledAttention`                         src/models/attention.py:13-124    (2-D inputs: `q_lin(x)` only)
  * `load_cov_stats`                             src/utils.py:17-71

The per-step work -- `x @ P^T + b` on `[B, D_in]` -- runs on the B200 kernels: fp32 mode `vitb200_linear_fwd` (SIMT fp32
GEMM); bf16-mixed `vitb200_tc_prelinear_fwd`, a tcgen05/TMA GEMM on a bf16 copy of the matrix whose contraction is split
over one wave of CTAs (the few-rows x large-matrix shape is weight-streaming bound) and whose fp32 output holds
bf16-rounded values, as autocast's bf16 Linear output does.  A trainable preprocessor adds `vitb200_linear_wgrad` /
`vitb200_linear_dgrad` in backward.  There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import math
import os
from pathlib import Path
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import ACT_NONE, BF16, F32

__all__ = ["LinearPreprocessor", "PrefilledLinear", "PrefilledAttention", "compute_zca_matrix", "compute_pca_matrix",
           "load_cov_stats", "clear_cov_cache"]


# ----------------------------------------------------------------------------------------------
# statistics -> matrices (host side, once per model build)
# ----------------------------------------------------------------------------------------------
def compute_zca_matrix(eigvecs: torch.Tensor, eigvals: torch.Tensor, eps: float = 1e-5, r: int | None = None,
                       shrinkage: float = 0.1) -> torch.Tensor:
    """ZCA whitening matrix (preprocessor.py:12-72).  Full rank: V diag((lam_hat + eps)^-1/2) V^T.  Low rank r:
    (Vr * (lam_hat_r + eps)^-1/2) Vr^T + s_perp (I - Vr Vr^T) with s_perp from the median tail eigenvalue, floored at
    1e-3 x the mean of the leading r."""
    lam = eigvals
    if shrinkage > 0.0:
        lam = (1.0 - shrinkage) * eigvals + shrinkage * eigvals.mean()
    if r is None:
        return eigvecs @ torch.diag(1.0 / torch.sqrt(lam + eps)) @ eigvecs.t()
    Vr = eigvecs[:, :r]
    inv_sqrt_r = torch.rsqrt(lam[:r] + eps)
    tail = lam[r:]
    lam0 = tail.median() if tail.numel() > 0 else lam[r - 1]
    lam0 = torch.clamp(lam0, min=1e-3 * lam[:r].mean())
    s_perp = 1.0 / torch.sqrt(lam0 + eps)
    D = eigvecs.shape[0]
    eye = torch.eye(D, dtype=eigvecs.dtype, device=eigvecs.device)
    return (Vr * inv_sqrt_r) @ Vr.t() + s_perp * (eye - Vr @ Vr.t())


def zca_lowrank_factors(eigvecs: torch.Tensor, eigvals: torch.Tensor, eps: float, r: int, shrinkage: float):
    """(Vr [D, r], g [r], s_perp) with compute_zca_matrix(..., r) == s_perp I + (Vr * g) Vr^T: the factored form of the
    low-rank ZCA matrix (preprocessor.py:40-72), g = 1/sqrt(lam_r + eps) - s_perp."""
    lam = eigvals
    if shrinkage > 0.0:
        lam = (1.0 - shrinkage) * eigvals + shrinkage * eigvals.mean()
    Vr = eigvecs[:, :r]
    inv_sqrt_r = torch.rsqrt(lam[:r] + eps)
    tail = lam[r:]
    lam0 = tail.median() if tail.numel() > 0 else lam[r - 1]
    lam0 = torch.clamp(lam0, min=1e-3 * lam[:r].mean())
    s_perp = 1.0 / torch.sqrt(lam0 + eps)
    return Vr.contiguous(), (inv_sqrt_r - s_perp).contiguous(), float(s_perp)


def compute_pca_matrix(eigvecs: torch.Tensor, r: int | None = None) -> torch.Tensor:
    """PCA projection V[:, :r]^T (preprocessor.py:75-90)."""
    return eigvecs.t() if r is None else eigvecs[:, :r].t()


_COV_CACHE: Dict[Path, dict] = {}


def load_cov_stats(cov_path) -> dict:
    """src/utils.py:17-71: a torch-saved dict with 'mean', 'cov', 'eigvals', 'eigvecs' (cached per path)."""
    path = Path(cov_path).resolve()
    if path in _COV_CACHE:
        return _COV_CACHE[path]
    if not path.exists():
        raise FileNotFoundError(f"Covariance file not found: {path}")
    stats = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(stats, dict):
        raise ValueError(f"Expected dict from {cov_path}, got {type(stats)}")
    missing = {"mean", "cov", "eigvals", "eigvecs"} - set(stats.keys())
    if missing:
        raise ValueError(f"Missing required keys in {cov_path}: {missing}")
    _COV_CACHE[path] = stats
    return stats


def clear_cov_cache() -> None:
    _COV_CACHE.clear()


# ----------------------------------------------------------------------------------------------
# the GEMM on the device
# ----------------------------------------------------------------------------------------------
def _precision_of(mod: nn.Module) -> str:
    owner = mod.__dict__.get("_owner")
    return getattr(owner, "precision", None) or mod.__dict__.get("_precision", "fp32")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _aligned(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _bf16_copy(mod: nn.Module, weight: torch.Tensor) -> torch.Tensor:
    """bf16 GEMM operand of the matrix, refreshed when the fp32 tensor was written (optimizer step, load_state_dict,
    .to()): keyed by storage pointer + version counter, like the parameter arena's shadow."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape))
    cached = mod.__dict__.get("_w16")
    if cached is None or cached[0] != key:
        w32 = _aligned(weight.detach().to(torch.float32))
        w16 = torch.empty(w32.shape, dtype=torch.bfloat16, device=w32.device)
        _lib.check(_lib.load().vitb200_cast_bf16(w32.data_ptr(), w16.data_ptr(), w32.numel(), _stream(w32)), "cast_bf16")
        cached = (key, w16)
        mod.__dict__["_w16"] = cached
    return cached[1]


def _check_input(x: torch.Tensor, weight: torch.Tensor) -> None:
    if not x.is_cuda or not weight.is_cuda:
        raise RuntimeError("vit_b200 has no CPU path: the preprocessor and its input must live on a CUDA (sm_100a) device")
    if x.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise ValueError(f"expected input of shape [B, {weight.shape[1]}], got {tuple(x.shape)}")
    if weight.shape[0] % 4 or weight.shape[1] % 4:
        raise ValueError("vit_b200 preprocessor: input and output dimensions must be multiples of 4")


def linear_forward(mod: nn.Module, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                   out: Optional[torch.Tensor] = None):
    """y[B, N] (fp32) = x[B, K] . weight[N, K]^T + bias.  Returns (y, x_operand) -- the operand copy of x is what backward
    needs.  `out`: write into this fp32 [B, N] buffer (the engine's pixel buffer) instead of allocating; that is the
    no-grad step path, which also re-uses its bf16 staging buffer from call to call."""
    _check_input(x, weight)
    lib = _lib.load()
    B, K = x.shape
    N = weight.shape[0]
    st = _stream(x)
    xin = _aligned(x.detach().to(torch.float32))
    y = out if out is not None else torch.empty(B, N, dtype=torch.float32, device=x.device)
    bptr = None if bias is None else _aligned(bias.detach().to(torch.float32)).data_ptr()
    if _precision_of(mod) == "bf16":
        xb = mod.__dict__.get("_x16") if out is not None else None
        if xb is None or xb.shape != (B, K) or xb.device != x.device:
            xb = torch.empty(B, K, dtype=torch.bfloat16, device=x.device)
            if out is not None:
                mod.__dict__["_x16"] = xb
        _lib.check(lib.vitb200_cast_bf16(xin.data_ptr(), xb.data_ptr(), B * K, st), "cast_bf16")
        wb = _bf16_copy(mod, weight)
        if N % 8 == 0 and K % 8 == 0 and y.data_ptr() % 16 == 0:
            # few rows x a large matrix: split-K tcgen05 GEMM sized to one wave of CTAs, fp32 (bf16-rounded) output
            need = int(lib.vitb200_tc_prelinear_ws_bytes(B, N, K))
            ws = mod.__dict__.get("_fwd_ws")
            if ws is None or ws.numel() < need or ws.device != x.device:
                ws = torch.zeros(need, dtype=torch.uint8, device=x.device)
                mod.__dict__["_fwd_ws"] = ws
            _lib.check(lib.vitb200_tc_prelinear_fwd(xb.data_ptr(), wb.data_ptr(), bptr, y.data_ptr(), B, N, K,
                                                    ws.data_ptr(), st), "preprocessor prelinear_fwd")
            return y, xb
        yb = torch.empty(B, N, dtype=torch.bfloat16, device=x.device)
        _lib.check(lib.vitb200_linear_fwd(xb.data_ptr(), wb.data_ptr(), bptr, yb.data_ptr(), None, B, N, K, ACT_NONE, BF16,
                                          st), "preprocessor linear_fwd")
        _lib.check(lib.vitb200_cast_f32(yb.data_ptr(), y.data_ptr(), B * N, st), "cast_f32")
        return y, xb
    w32 = _aligned(weight.detach().to(torch.float32))
    _lib.check(lib.vitb200_linear_fwd(xin.data_ptr(), w32.data_ptr(), bptr, y.data_ptr(), None, B, N, K, ACT_NONE, F32, st),
               "preprocessor linear_fwd")
    return y, xin


class _PreLinearFunction(torch.autograd.Function):
    """Autograd node of a TRAINABLE preprocessor matrix: dW = dy^T x, db = column sums of dy (vitb200_linear_wgrad) and,
    if the raw input itself requires a gradient, dx = dy W (vitb200_linear_dgrad)."""

    @staticmethod
    def forward(ctx, mod, x, weight, bias):
        y, x_op = linear_forward(mod, x, weight, bias)
        ctx.mod, ctx.x_op = mod, x_op
        ctx.save_for_backward(weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        mod = ctx.mod
        (weight,) = ctx.saved_tensors
        lib = _lib.load()
        x_op = ctx.x_op
        B, K = x_op.shape
        N = weight.shape[0]
        st = _stream(dy)
        bf = x_op.dtype == torch.bfloat16
        dt = BF16 if bf else F32
        dy32 = _aligned(dy.to(torch.float32))
        if bf:
            dy_op = torch.empty(B, N, dtype=torch.bfloat16, device=dy.device)
            _lib.check(lib.vitb200_cast_bf16(dy32.data_ptr(), dy_op.data_ptr(), B * N, st), "cast_bf16")
        else:
            dy_op = dy32
        dw = db = dx = None
        if ctx.needs_input_grad[2] or (ctx.has_bias and ctx.needs_input_grad[3]):
            ws = mod.__dict__.get("_wgrad_ws")
            need = int(lib.vitb200_linear_wgrad_ws_bytes(B, N, K)) + 4096
            if ws is None or ws.numel() < need or ws.device != dy.device:
                ws = torch.zeros(need, dtype=torch.uint8, device=dy.device)
                mod.__dict__["_wgrad_ws"] = ws
            dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
            db = torch.empty(N, dtype=torch.float32, device=dy.device) if ctx.has_bias else None
            _lib.check(lib.vitb200_linear_wgrad(dy_op.data_ptr(), x_op.data_ptr(), dw.data_ptr(),
                                                None if db is None else db.data_ptr(), B, N, K, 0, dt, ws.data_ptr(), st),
                       "preprocessor linear_wgrad")
            if not ctx.needs_input_grad[2]:
                dw = None
            if not (ctx.has_bias and ctx.needs_input_grad[3]):
                db = None
        if ctx.needs_input_grad[1]:
            w_op = _bf16_copy(mod, weight) if bf else _aligned(weight.detach().to(torch.float32))
            dx_op = torch.empty(B, K, dtype=x_op.dtype, device=dy.device)
            _lib.check(lib.vitb200_linear_dgrad(dy_op.data_ptr(), w_op.data_ptr(), None, dx_op.data_ptr(), B, N, K, dt, st),
                       "preprocessor linear_dgrad")
            if bf:
                dx = torch.empty(B, K, dtype=torch.float32, device=dy.device)
                _lib.check(lib.vitb200_cast_f32(dx_op.data_ptr(), dx.data_ptr(), B * K, st), "cast_f32")
            else:
                dx = dx_op
        return None, dx, dw, db


def _lowrank_state(mod: nn.Module, weight: torch.Tensor):
    """The factored form of a FROZEN low-rank ZCA matrix, or None.  The factors come from the builder (they are not part
    of the state_dict, which stays the reference's); they are trusted only while the dense matrix still equals
    s_perp I + (Vr g) Vr^T -- re-checked (one D x D product) whenever the matrix tensor was replaced or written, e.g. by
    load_state_dict of a trained matrix, which silently falls back to the dense kernel."""
    lr = mod.__dict__.get("_lowrank")
    if lr is None or not getattr(mod, "_is_frozen", False) or weight.requires_grad:
        return None
    if os.environ.get("VITB200_ZCA_LOWRANK", "1") == "0":
        return None
    key = (weight.data_ptr(), weight._version, weight.device)
    if lr.get("key") != key:
        lr["key"] = key
        lr["ok"] = False
        D, r = lr["Vr"].shape
        R = 32 if r <= 32 else 64
        if r <= 64 and weight.shape == (D, D) and _lib.load().vitb200_zca_lowrank_supported(D, R):
            dev = weight.device
            Vr = torch.zeros(D, R, dtype=torch.float32, device=dev)
            Vr[:, :r] = lr["Vr"].to(dev, torch.float32)
            g = torch.zeros(R, dtype=torch.float32, device=dev)
            g[:r] = lr["g"].to(dev, torch.float32)
            dense = (Vr * g) @ Vr.t()
            dense.diagonal().add_(lr["s_perp"])
            w = weight.detach().to(torch.float32)
            if float((dense - w).abs().max()) <= 1e-5 * max(float(w.abs().max()), 1e-30):
                lr.update(ok=True, R=R, Vr32=Vr.contiguous(), Vr16=Vr.to(torch.bfloat16).contiguous(), g32=g)
            del dense
    return lr if lr["ok"] else None


def lowrank_forward(mod: nn.Module, lr: dict, x: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None):
    """y = s_perp x + ((x Vr) o g) Vr^T + bias in one launch (vitb200_zca_lowrank_fwd)."""
    if not x.is_cuda:
        raise RuntimeError("vit_b200 has no CPU path: the preprocessor and its input must live on a CUDA (sm_100a) device")
    B, D = x.shape
    if D != lr["Vr32"].shape[0]:
        raise ValueError(f"expected input of shape [B, {lr['Vr32'].shape[0]}], got {tuple(x.shape)}")
    xin = _aligned(x.detach().to(torch.float32))
    y = out if out is not None else torch.empty(B, D, dtype=torch.float32, device=x.device)
    bf = _precision_of(mod) == "bf16"
    bptr = None if bias is None else _aligned(bias.detach().to(torch.float32)).data_ptr()
    _lib.check(_lib.load().vitb200_zca_lowrank_fwd(
        xin.data_ptr(), (lr["Vr16"] if bf else lr["Vr32"]).data_ptr(), lr["g32"].data_ptr(), lr["s_perp"], bptr,
        y.data_ptr(), B, D, lr["R"], BF16 if bf else F32, _stream(x)), "zca_lowrank_fwd")
    return y


def _apply_linear(mod: nn.Module, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    needs = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or
                                         (bias is not None and bias.requires_grad))
    if needs:
        return _PreLinearFunction.apply(mod, x, weight, bias)
    lr = _lowrank_state(mod, weight)
    if lr is not None:
        return lowrank_forward(mod, lr, x, bias)
    return linear_forward(mod, x, weight, bias)[0]


# ----------------------------------------------------------------------------------------------
# modules (reference surface)
# ----------------------------------------------------------------------------------------------
class PrefilledLinear(nn.Module):
    """Linear layer initialised from a matrix (ZCA, PCA, ...), optional centering bias (layers.py:12-63).  Frozen: weight
    and bias are buffers; unfrozen: Parameters -- `freeze()` converts between the two at run time (layers.py:36-60)."""

    def __init__(self, matrix: torch.Tensor, bias: torch.Tensor | None = None, freeze: bool = True) -> None:
        super().__init__()
        self._is_frozen = bool(freeze)
        self._install("weight", matrix.to(torch.float32), as_param=not freeze)
        # a missing bias is registered as a None buffer, so `state_dict` has no bias key (layers.py:34-35)
        self._install("bias", None if bias is None else bias.to(torch.float32), as_param=not freeze)

    def _install(self, name: str, value: Optional[torch.Tensor], as_param: bool) -> None:
        """(Re-)register `name` as a Parameter (trainable) or as a buffer (frozen / absent)."""
        if name in self._parameters:
            del self._parameters[name]
        if name in self._buffers:
            del self._buffers[name]
        if value is not None and as_param:
            self.register_parameter(name, nn.Parameter(value))
        else:
            self.register_buffer(name, value)

    @property
    def in_features(self) -> int:
        return self.weight.shape[1]

    @property
    def out_features(self) -> int:
        return self.weight.shape[0]

    def freeze(self, freeze: bool = True) -> None:
        """Run-time switch between buffers and Parameters (layers.py:36-60): the tensors keep their values but change
        their role, so `parameters()` -- and what an optimizer built afterwards sees -- changes with it."""
        if bool(freeze) == self._is_frozen:
            return
        for name in ("weight", "bias"):
            cur = getattr(self, name)
            self._install(name, None if cur is None else cur.detach().clone(), as_param=not freeze)
        self._is_frozen = bool(freeze)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _apply_linear(self, x, self.weight, self.bias)


class LinearPreprocessor(nn.Module):
    """x -> x P^T + bias: ZCA whitening (P [D, D]) or PCA projection (P [r, D]) (preprocessor.py:93-111)."""

    def __init__(self, matrix: torch.Tensor, bias: torch.Tensor | None = None, freeze: bool = True) -> None:
        super().__init__()
        self.linear = PrefilledLinear(matrix, bias=bias, freeze=freeze)

    @property
    def in_features(self) -> int:
        return self.linear.in_features

    @property
    def out_features(self) -> int:
        return self.linear.out_features

    @property
    def is_frozen(self) -> bool:
        return self.linear._is_frozen

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.linear(x)

    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> None:
        """No-grad forward straight into the engine's pixel buffer (TrainStep / EvalStep with a frozen preprocessor)."""
        lr = _lowrank_state(self.linear, self.linear.weight)
        if lr is not None:
            lowrank_forward(self.linear, lr, x, self.linear.bias, out=out)
            return
        linear_forward(self.linear, x, self.linear.weight, self.linear.bias, out=out)

    def set_lowrank_factors(self, Vr: torch.Tensor, g: torch.Tensor, s_perp: float) -> None:
        """Builder hook: the matrix is s_perp I + (Vr g) Vr^T (zca_lowrank_factors).  While it stays frozen and unchanged
        the forward runs the factored kernel instead of streaming the dense matrix."""
        self.linear.__dict__["_lowrank"] = dict(Vr=Vr.detach().clone(), g=g.detach().clone(), s_perp=float(s_perp))

    def freeze(self, freeze: bool = True) -> None:
        self.linear.freeze(freeze)


class _KernelLinear(nn.Linear):
    """nn.Linear parameter holder whose forward runs the vit_b200 GEMM (bias-free q_lin of PrefilledAttention)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _apply_linear(self, x, self.weight, self.bias)


class PrefilledAttention(nn.Module):
    """Attention layer whose query/key projections are prefilled from an eigenbasis (attention.py:13-124).  The model feeds
    it 2-D inputs `[B, D]`, for which the reference evaluates `q_lin(x)` only (attention.py:81-82); `k_lin` / `v_lin` exist
    for `state_dict` compatibility (and get the reference's initialisation), the 3-D attention branch is not on this path."""

    def __init__(self, input_dim: int, eigvecs: torch.Tensor, eigvals: torch.Tensor | None = None, r: int | None = None,
                 low_rank: bool | None = None, scale_by_eigvals: bool = True, eps: float = 1e-5) -> None:
        super().__init__()
        self.input_dim = input_dim
        self.r = r if r is not None else eigvecs.shape[1]
        self.low_rank = low_rank if low_rank is not None else (self.r < input_dim)
        self.scale_by_eigvals = scale_by_eigvals and eigvals is not None
        out = self.r if self.low_rank else input_dim
        self.q_lin = _KernelLinear(input_dim, out, bias=False)
        self.k_lin = nn.Linear(input_dim, out, bias=False)
        self.v_lin = nn.Linear(input_dim, input_dim, bias=False)
        V = eigvecs[:, :self.r].t().contiguous()
        if self.scale_by_eigvals:
            V = V * torch.rsqrt(eigvals[:self.r] + eps).unsqueeze(1)
        with torch.no_grad():
            for layer in (self.q_lin, self.k_lin):
                if self.low_rank:
                    layer.weight.copy_(V)
                else:
                    layer.weight.zero_()
                    layer.weight[:V.shape[0], :].copy_(V)
        nn.init.kaiming_uniform_(self.v_lin.weight, a=math.sqrt(5))
        self.softmax = nn.Softmax(dim=-1)

    @property
    def in_features(self) -> int:
        return self.input_dim

    @property
    def out_features(self) -> int:
        return self.q_lin.weight.shape[0]

    @property
    def is_frozen(self) -> bool:
        return not self.q_lin.weight.requires_grad

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            return self.q_lin(x)
        raise NotImplementedError("vit_b200: PrefilledAttention is only evaluated on 2-D inputs [B, D] by the model "
                                  "(src/models/attention.py:81-82); the 3-D attention branch is not on the hot path")

    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> None:
        linear_forward(self.q_lin, x, self.q_lin.weight, None, out=out)

    def set_qk_trainable(self, trainable: bool = True) -> None:
        for p in self.q_lin.parameters():
            p.requires_grad = trainable
        for p in self.k_lin.parameters():
            p.requires_grad = trainable
