"""Config -> model, mirroring the reference builder (same function names, config keys, quirks).

    get_vit_config(config)   /root/reference/src/models/builder.py:200-258
    get_model(config)        /root/reference/src/models/builder.py:136-197

`VitConfig` plays the role of `transformers.ViTConfig` for the attributes the reference reads
(`config.num_labels`, `hidden_size`, ...), without importing transformers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Optional


@dataclass
class VitConfig:
    task_type: str = "reg"
    image_size: int = 4096
    patch_size: int = 32
    num_channels: int = 1
    hidden_size: int = 32
    num_hidden_layers: int = 3
    num_attention_heads: int = 2
    intermediate_size: int = 128
    stride_ratio: float = 1
    stride_size: Optional[int] = None
    proj_fn: str = "SW"
    hidden_act: str = "gelu"
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    initializer_range: float = 0.02
    layer_norm_eps: float = 1e-12
    qkv_bias: bool = True
    num_labels: int = 1
    pos_encoding_type: Optional[str] = None
    max_position_embeddings: int = 512
    rope_base: float = 10000.0
    use_return_dict: bool = True
    # derived
    stride: int = field(init=False, default=0)
    num_patches: int = field(init=False, default=0)
    n_valid: int = field(init=False, default=0)

    def __post_init__(self):
        # embedding.py:26-27
        s = self.stride_size
        self.stride = s if s and s > 0 else int(self.stride_ratio * self.patch_size)
        if self.stride <= 0:
            raise ValueError("stride must be positive")
        L, P, S = self.image_size, self.patch_size, self.stride
        if P > L:
            raise ValueError(f"patch_size {P} > image_size {L}")
        n_unfold = (L - P) // S + 1
        if self.proj_fn == "SW":          # tokenization.py:40 (ceil; missing windows are zero patches)
            self.num_patches = math.ceil((L - P) / S) + 1
        elif self.proj_fn in ("C1D", "CNN"):  # tokenization.py:63 (floor)
            self.num_patches = n_unfold
        else:
            raise ValueError(f"Unsupported proj_fn '{self.proj_fn}'")
        self.n_valid = min(n_unfold, self.num_patches)
        if self.hidden_size % self.num_attention_heads != 0:
            raise ValueError(
                f"The hidden size {self.hidden_size} is not a multiple of the number of attention heads "
                f"{self.num_attention_heads}.")
        if self.task_type not in ("cls", "reg"):
            raise ValueError(f"Unsupported task_type '{self.task_type}'")
        if self.pos_encoding_type not in (None, "none", "learned", "rope"):
            raise ValueError(
                f"Unsupported pos_encoding_type '{self.pos_encoding_type}'. "
                f"Choose from: 'rope', 'learned', 'none', or None")
        # kernel support envelope -- reported at construction, never a silent fallback
        H, d = self.hidden_size, self.head_dim
        if H % 4 != 0 or H > 1024:
            raise ValueError(f"vit_b200: hidden_size must be a multiple of 4 and <= 1024 (got {H})")
        if d not in (4, 8, 16, 32, 64, 128):
            raise ValueError(f"vit_b200: head_dim must be one of 4,8,16,32,64,128 (got {d})")

    @property
    def tokens(self) -> int:
        return self.num_patches + 1

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def to_dict(self) -> dict:
        return {k: getattr(self, k) for k in self.__dataclass_fields__}


def get_vit_config(config: dict) -> VitConfig:
    """builder.py:200-258, including: num_labels derived from data.param for regression (and written
    back into config['model']['num_labels']), intermediate = 4*hidden, dropouts 0.1, eps 1e-12."""
    m = config["model"]
    d = config.get("data", {}) or {}
    task = (m.get("task_type") or m.get("task") or "cls").lower()
    if task in ("reg", "regression"):
        p = d.get("param", None)
        num_labels = 1
        if isinstance(p, str) and len(p) > 0:
            plist = [x.strip() for x in p.split(",") if x.strip()]
            if len(plist) >= 1:
                num_labels = len(plist)
        elif isinstance(p, (list, tuple)) and len(p) > 0:
            num_labels = len(p)
        cfg_nl = m.get("num_labels")
        if cfg_nl is not None and int(cfg_nl) != num_labels:
            print(f"Warning: model.num_labels={cfg_nl} conflicts with data.param (which implies {num_labels} "
                  f"labels). Using {num_labels} from data.param.")
        m["num_labels"] = num_labels
    else:
        num_labels = int(m.get("num_labels", 1) or 1)
    return VitConfig(
        task_type=m["task_type"], image_size=m["image_size"], patch_size=m["patch_size"], num_channels=1,
        hidden_size=m["hidden_size"], num_hidden_layers=m["num_hidden_layers"],
        num_attention_heads=m["num_attention_heads"], intermediate_size=4 * m["hidden_size"],
        stride_ratio=m.get("stride_ratio", 1), stride_size=m.get("stride_size", None), proj_fn=m["proj_fn"],
        num_labels=num_labels, pos_encoding_type=m.get("pos_encoding_type", None),
        max_position_embeddings=m.get("max_position_embeddings", 512), rope_base=m.get("rope_base", 10000.0),
    )


def build_model_name(config: VitConfig, model_prefix: str = "ViT", full_config: dict | None = None) -> str:
    """src/models/model_utils.py:9-41."""
    stride_used = getattr(config, "stride_size", None)
    stride_tag = int(stride_used) if (stride_used is not None and stride_used) else config.stride_ratio
    name = (f"{model_prefix}_p{config.patch_size}_h{config.hidden_size}_l{config.num_hidden_layers}_"
            f"a{config.num_attention_heads}_s{stride_tag}_p{config.proj_fn}")
    if full_config is not None:
        noise_level = (full_config.get("noise", {}) or {}).get("noise_level", 0)
        if noise_level > 0:
            name += f"_nz{str(noise_level).replace('.', '')}"
    return name


def get_model(config: dict, precision: Any = None, device: Any = None):
    """Build the B200 MyViT for `config` (builder.py:136-150).  `precision` defaults to
    config['train']['precision'] (reference default '32', src/basemodule.py:233)."""
    from .model import MyViT

    warmup_cfg = config.get("warmup", {}) or {}
    loss_name = (config.get("loss", {}) or {}).get("name", None)
    preproc_type = warmup_cfg.get("preprocessor", None)
    if precision is None:
        precision = str((config.get("train", {}) or {}).get("precision", "32"))
    if preproc_type is None or str(preproc_type).lower() in ("none", "null"):
        vit_config = get_vit_config(config)
        model = MyViT(vit_config, loss_name=loss_name, model_name="ViT", full_config=config,
                      precision=precision, device=device)
        print("[builder] Created vanilla ViT model")
        return model
    # ---- input preprocessor (builder.py:152-197) ----
    from .preprocessor import load_cov_stats

    cov_path = warmup_cfg.get("cov_path", None)
    if cov_path is None:
        raise ValueError(f"preprocessor='{preproc_type}' requires 'cov_path' in warmup config")
    stats = load_cov_stats(cov_path)
    input_dim = stats["eigvecs"].shape[0]
    original_image_size = config["model"]["image_size"]
    if input_dim != original_image_size:
        raise ValueError(f"Mismatch: eigvecs dimension {input_dim} != image_size {original_image_size}")
    freeze_epochs = warmup_cfg.get("freeze_epochs", 0)
    preprocessor, output_dim, name_prefix, desc = build_preprocessor(
        preproc_type, warmup_cfg, stats, input_dim, initial_freeze=freeze_epochs != 0)
    if output_dim != original_image_size:
        print(f"[builder] Auto-adjusting image_size: {original_image_size} -> {output_dim}")
        config["model"]["image_size"] = output_dim
    vit_config = get_vit_config(config)
    print(f"[builder] Created {desc} preprocessor")
    if freeze_epochs == -1:
        print("[builder] Preprocessor will be PERMANENTLY FROZEN (never trained)")
    elif freeze_epochs > 0:
        print(f"[builder] Preprocessor will be frozen for first {freeze_epochs} epochs")
    else:
        print("[builder] Preprocessor is trainable from start")
    model = MyViT(vit_config, loss_name=loss_name, model_name=f"{name_prefix}_ViT", preprocessor=preprocessor,
                  full_config=config, precision=precision, device=device)
    print(f"[builder] Created {model._model_name} with {preproc_type} preprocessor")
    return model


def build_preprocessor(preproc_type: str, warmup_cfg: dict, stats: dict, input_dim: int, initial_freeze: bool):
    """(preprocessor, output_dim, model-name prefix, description) for warmup.preprocessor in {'zca','pca','attention'}
    (builder.py:43-133): matrices from the covariance statistics, centering bias = -mean @ P^T unless warmup.bias is
    false, names `ZCA{r}_fz{..}[_s{..}][_nobias]`, `PCA{r}_fz{..}[_nobias]`, `Attn{r|Full}[_scaled]_fz{..}`."""
    from .preprocessor import LinearPreprocessor, PrefilledAttention, compute_pca_matrix, compute_zca_matrix

    eigvecs = stats["eigvecs"]
    mean = stats.get("mean", None)
    r = warmup_cfg.get("r", None)
    fe = warmup_cfg.get("freeze_epochs", 0)
    fz = "perm" if fe == -1 else str(fe)
    kind = str(preproc_type)
    if kind in ("zca", "pca"):
        use_bias = warmup_cfg.get("bias", True)
        if kind == "zca":
            eps = warmup_cfg.get("eps", 1e-5)
            shrinkage = warmup_cfg.get("shrinkage", 0.0)
            P = compute_zca_matrix(eigvecs, stats["eigvals"], eps=eps, r=r, shrinkage=shrinkage)
            shrink = f"_s{int(shrinkage * 10)}" if shrinkage > 0 else ""
            prefix = f"{'ZCA' + str(r) if r is not None else 'ZCA'}_fz{fz}{shrink}{'' if use_bias else '_nobias'}"
            desc = f"{'low-rank' if r else 'full-rank'} ZCA, eps={eps}, shrinkage={shrinkage}, bias={use_bias}"
        else:
            P = compute_pca_matrix(eigvecs, r=r)
            prefix = f"{'PCA' + str(r) if r is not None else 'PCA'}_fz{fz}{'' if use_bias else '_nobias'}"
            desc = f"PCA with r={r}, bias={use_bias}" if r else f"full-rank PCA, bias={use_bias}"
        bias = -mean @ P.t() if (use_bias and mean is not None) else None   # y = (x - mean) P^T
        pre = LinearPreprocessor(P, bias=bias, freeze=initial_freeze)
        if kind == "zca" and r is not None:
            from .preprocessor import zca_lowrank_factors

            pre.set_lowrank_factors(*zca_lowrank_factors(eigvecs, stats["eigvals"], eps, int(r), shrinkage))
        return pre, P.shape[0], prefix, desc
    if kind == "attention":
        eigvals = stats.get("eigvals", None)
        scale = warmup_cfg.get("scale_by_eigvals", True)
        pre = PrefilledAttention(input_dim=input_dim, eigvecs=eigvecs, eigvals=eigvals, r=r, scale_by_eigvals=scale,
                                 eps=warmup_cfg.get("eps", 1e-5))
        prefix = f"Attn{r if r else 'Full'}{'_scaled' if scale and eigvals is not None else ''}_fz{fz}"
        return pre, (r if r is not None else input_dim), prefix, f"Attention preprocessor with r={r}, scale_by_eigvals={scale}"
    raise ValueError(f"Unknown preprocessor type: '{preproc_type}'")
