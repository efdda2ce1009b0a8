"""MyViT -- the reference's model interface (src/models/specvit.py:17-115) on the B200 engine.

Same module tree and `state_dict` keys as the reference (`vit.embeddings.*`, `vit.encoder.layer.i.*`,
`vit.layernorm.*`, `vit.pooler.dense.*`, `regressor|classifier.*`), same
`forward(pixel_values, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None)`
signature and output fields, same `.name / .loss_name / .config / .vit / .preprocessor /
.set_preprocessor_trainable`.  Every parameter is a view into one flat arena (arena.py).

Two execution paths, same kernels:
  * fused   (default): the whole forward is one autograd node backed by `ViTEngine`; its backward runs
            the explicit backward kernel sequence.  Forward hooks on inner modules do not fire here.
  * modular (eval / no-grad, chosen when forward hooks are registered on inner modules or attention
            probabilities are requested): each leaf module (`query`, `dense`, `layernorm_*`, ...) runs
            its own kernel, so the reference's viz/CKA callbacks (src/viz/viz_callback.py:220-235,
            src/viz/cka_utils.py:156-172) see `(context, attention_probs)` etc.
There is no PyTorch fallback for the math: without the CUDA library or a CUDA device, forward raises.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .arena import ParamArena, ParamLayout
from .builder import VitConfig, build_model_name
from .engine import ViTEngine, _normalize_precision


class ModelOutput(dict):
    """Attribute + mapping + tuple-style access, like transformers' ModelOutput."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __getitem__(self, k):
        if isinstance(k, int):
            return [v for v in self.values() if v is not None][k]
        return dict.__getitem__(self, k)


# ----------------------------------------------------------------------------------------------
# leaf modules (parameters are arena views; forward = modular no-grad path)
# ----------------------------------------------------------------------------------------------
def _rt(model) -> "MyViT":
    return model.__dict__["_owner"]


class ArenaLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_parameter("weight", None)
        self.register_parameter("bias", None)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias=True"

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import ops

        return ops.linear(self.__dict__["_owner"], x, self.weight, self.bias)


class ArenaConv1d(ArenaLinear):
    """nn.Conv1d(1, H, kernel_size=P, stride=S) parameter holder (weight [H,1,P]); tokenization.py:64."""

    def __init__(self, patch: int, hidden: int, stride: int):
        super().__init__(patch, hidden)
        self.kernel_size, self.stride = (patch,), (stride,)


class ArenaLayerNorm(nn.Module):
    def __init__(self, hidden: int, eps: float):
        super().__init__()
        self.normalized_shape, self.eps = (hidden,), eps
        self.register_parameter("weight", None)
        self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import ops

        return ops.layer_norm(self.__dict__["_owner"], x, self.weight, self.bias, self.eps)


class PatchTokenizer(nn.Module):
    """SlidingWindowTokenizer / Conv1DPatchTokenizer (tokenization.py:31-69)."""

    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.image_size, self.patch_size, self.stride_size = cfg.image_size, cfg.patch_size, cfg.stride
        self.num_patches = cfg.num_patches
        if cfg.proj_fn == "SW":
            self.projection = ArenaLinear(cfg.patch_size, cfg.hidden_size)
        else:
            self.num_channels = 1
            self.projection = ArenaConv1d(cfg.patch_size, cfg.hidden_size, cfg.stride)


class SpectraEmbeddings(nn.Module):
    """embedding.py:15-106."""

    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.patch_embeddings = PatchTokenizer(cfg)
        self.num_patches = cfg.num_patches
        self.register_parameter("cls_token", None)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)
        self.pos_encoding_type = cfg.pos_encoding_type
        self.register_parameter("position_embeddings", None)
        self.rope = None

    def forward(self, x, bool_masked_pos=None, interpolate_pos_encoding=False):
        from . import ops

        return ops.embed(self.__dict__["_owner"], x)

    def set_patch_proj_trainable(self, trainable: bool = True) -> None:
        for p in self.patch_embeddings.projection.parameters():
            p.requires_grad = trainable


class ViTSelfAttention(nn.Module):
    def __init__(self, cfg: VitConfig, layer: int):
        super().__init__()
        H = cfg.hidden_size
        self.num_attention_heads = cfg.num_attention_heads
        self.attention_head_size = cfg.head_dim
        self.all_head_size = H
        self.dropout_prob = cfg.attention_probs_dropout_prob
        self.scaling = cfg.head_dim ** -0.5
        self.use_rope = cfg.pos_encoding_type == "rope"
        self.layer_index = layer
        self.query = ArenaLinear(H, H)
        self.key = ArenaLinear(H, H)
        self.value = ArenaLinear(H, H)

    def forward(self, hidden_states, head_mask=None, output_attentions=False, **kw):
        from . import ops

        q, k, v = self.query(hidden_states), self.key(hidden_states), self.value(hidden_states)
        return ops.attention(self.__dict__["_owner"], q, k, v, want_probs=True)


class ViTSelfOutput(nn.Module):
    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.dense = ArenaLinear(cfg.hidden_size, cfg.hidden_size)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor=None):
        return self.dense(hidden_states)  # eval path: dropout is the identity


class ViTAttention(nn.Module):
    def __init__(self, cfg: VitConfig, layer: int):
        super().__init__()
        self.attention = ViTSelfAttention(cfg, layer)
        self.output = ViTSelfOutput(cfg)

    def forward(self, hidden_states, **kw):
        ctx, _ = self.attention(hidden_states)
        return self.output(ctx, hidden_states)


class ViTIntermediate(nn.Module):
    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.dense = ArenaLinear(cfg.hidden_size, cfg.intermediate_size)

    def forward(self, hidden_states):
        from . import ops

        return ops.gelu(self.__dict__["_owner"], self.dense(hidden_states))


class ViTOutput(nn.Module):
    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.dense = ArenaLinear(cfg.intermediate_size, cfg.hidden_size)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        from . import ops

        return ops.residual_add(self.__dict__["_owner"], input_tensor, self.dense(hidden_states))


class ViTLayer(nn.Module):
    def __init__(self, cfg: VitConfig, layer: int):
        super().__init__()
        self.attention = ViTAttention(cfg, layer)
        self.intermediate = ViTIntermediate(cfg)
        self.output = ViTOutput(cfg)
        self.layernorm_before = ArenaLayerNorm(cfg.hidden_size, cfg.layer_norm_eps)
        self.layernorm_after = ArenaLayerNorm(cfg.hidden_size, cfg.layer_norm_eps)

    def forward(self, hidden_states, **kw):  # HF:328-346
        from . import ops

        attn = self.attention(self.layernorm_before(hidden_states))
        hidden_states = ops.residual_add(self.__dict__["_owner"], hidden_states, attn)
        out = self.intermediate(self.layernorm_after(hidden_states))
        return self.output(out, hidden_states)


class ViTEncoder(nn.Module):
    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.layer = nn.ModuleList([ViTLayer(cfg, i) for i in range(cfg.num_hidden_layers)])

    def forward(self, hidden_states, **kw):
        states = [hidden_states]
        for layer in self.layer:
            hidden_states = layer(hidden_states)
            states.append(hidden_states)
        return ModelOutput(last_hidden_state=hidden_states, hidden_states=tuple(states))


class ViTPooler(nn.Module):
    """Present for state_dict compatibility; the reference computes it and discards the result
    (HF:456, specvit.py:76-78), so it is never evaluated here and never receives a gradient."""

    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.dense = ArenaLinear(cfg.hidden_size, cfg.hidden_size)


class ViTModel(nn.Module):
    def __init__(self, cfg: VitConfig):
        super().__init__()
        self.config = cfg
        self.embeddings = SpectraEmbeddings(cfg)
        self.encoder = ViTEncoder(cfg)
        self.layernorm = ArenaLayerNorm(cfg.hidden_size, cfg.layer_norm_eps)
        self.pooler = ViTPooler(cfg)

    def forward(self, pixel_values=None, bool_masked_pos=None, interpolate_pos_encoding=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, **kw):
        """Eval / no-grad forward returning `last_hidden_state` [B,T,H] (src/viz/viz_callback.py:574-579)."""
        owner: MyViT = self.__dict__["_owner"]
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("vit_b200: call MyViT.forward for training; MyViT.vit(...) is the eval/no-grad path "
                               "(wrap it in torch.no_grad())")
        if owner._has_inner_hooks() or output_attentions:
            x = pixel_values.to(torch.float32)
            enc = self.encoder(self.embeddings(x))
            last = self.layernorm(enc.last_hidden_state)
            return ModelOutput(last_hidden_state=last, pooler_output=None,
                               hidden_states=enc.hidden_states if output_hidden_states else None, attentions=None)
        eng = owner._engine(pixel_values.shape[0])
        owner._stage_inputs(eng, pixel_values, None)
        eng.cls_only = False   # last_hidden_state of every token is the result here
        eng.forward(train=False, with_labels=False)
        hs = None
        if output_hidden_states:
            hs = tuple(z.view(eng.B, self.config.tokens, -1).clone() for z in eng.z)
        return ModelOutput(last_hidden_state=eng.last_hidden_state().float(), pooler_output=None, hidden_states=hs,
                           attentions=None)


# ----------------------------------------------------------------------------------------------
# fused autograd node
# ----------------------------------------------------------------------------------------------
class _FusedViTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, eng, x, labels, train, names, *params):
        ctx.set_materialize_grads(False)
        owner._stage_inputs(eng, x, labels)
        if train:
            eng.advance_rng()
        eng.forward(train=train, with_labels=labels is not None)
        eng._fwd_token = getattr(eng, "_fwd_token", 0) + 1
        ctx.owner, ctx.eng, ctx.train, ctx.names = owner, eng, train, names
        ctx.token = eng._fwd_token
        ctx.has_labels = labels is not None
        ctx.n_params = len(params)
        ctx.want_dx = bool(ctx.needs_input_grad[2])   # a trainable input preprocessor sits in front of the encoder
        loss = eng.loss[0].clone() if labels is not None else eng.loss.new_zeros(())
        return loss, eng.logits.clone()

    @staticmethod
    def backward(ctx, gloss, glogits):
        eng: ViTEngine = ctx.eng
        if ctx.token != eng._fwd_token:
            raise RuntimeError("vit_b200: the activations of this forward were overwritten by a later forward with "
                               "the same batch size; call backward before running the model again")
        if ctx.has_labels:
            if glogits is not None:
                raise NotImplementedError("vit_b200: differentiate through either the loss or the logits, not both")
            if gloss is None:
                return (None,) * (6 + ctx.n_params)
            eng.gloss.copy_(gloss.reshape(1).to(torch.float32))
            eng.backward(train=ctx.train, gloss=eng.gloss)
        else:
            if glogits is None:
                return (None,) * (6 + ctx.n_params)
            eng.backward_from_dlogits(ctx.train, glogits)
        flat = eng.arena.grad.clone()
        lay = eng.arena.layout
        grads = []
        for n in ctx.names:
            e = lay.entries[n]
            grads.append(flat[e.offset:e.offset + e.numel].view(e.shape) if e.offset < lay.n_opt else None)
        dx = eng.pixel_grad(ctx.train) if ctx.want_dx else None
        return (None, None, dx, None, None, None, *grads)


# ----------------------------------------------------------------------------------------------
# MyViT
# ----------------------------------------------------------------------------------------------
class MyViT(nn.Module):
    """Vision Transformer for 1-D spectra with optional input preprocessor (specvit.py:17)."""

    def __init__(self, config: VitConfig, loss_name: str = "", model_name: str = "ViT",
                 preprocessor: nn.Module | None = None, full_config: dict | None = None,
                 precision: Any = "32", device: Any = None, seed: int | None = None):
        super().__init__()
        self.config = config
        self.precision = _normalize_precision(precision)
        self.vit = ViTModel(config)
        self.preprocessor = preprocessor
        self.task_type = config.task_type
        if self.task_type == "cls":
            self.classifier = ArenaLinear(config.hidden_size, config.num_labels)
            loss_name = "ce"
            self._loss_kind = _lib.LOSS_CE
        elif self.task_type == "reg":
            self.regressor = ArenaLinear(config.hidden_size, config.num_labels)
            loss_name = loss_name or "l2"
            # specvit.py:52-53 -- only names containing 'l1' select L1; 'mae' therefore means MSE
            self._loss_kind = _lib.LOSS_L1 if "l1" in loss_name.lower() else _lib.LOSS_MSE
        else:
            raise ValueError(f"Unsupported task_type '{self.task_type}'")
        self._model_name = build_model_name(config, model_name, full_config=full_config)
        self._loss_name = loss_name
        print(f"Creating {self._model_name} model with {self._loss_name} loss")
        self._layout = ParamLayout(config)
        self._engines: Dict[int, ViTEngine] = {}
        self._seed = int(torch.initial_seed() % (2 ** 31)) if seed is None else int(seed)
        self._opt_hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.5)
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self._bind_arena(ParamArena(self._layout, device, with_shadow=True), init=True)
        if self.preprocessor is not None:
            self.preprocessor.to(device)
        for m in self.modules():
            m.__dict__["_owner"] = self

    # ---- reference surface ------------------------------------------------------------------
    @property
    def name(self):
        return self._model_name

    @property
    def loss_name(self):
        return self._loss_name

    @property
    def input_dim(self) -> int:
        """Length of the raw spectra `forward` takes: the preprocessor's input size, else config.image_size."""
        pre = self.preprocessor
        return int(pre.in_features) if pre is not None and hasattr(pre, "in_features") else int(self.config.image_size)

    def compute_loss(self, *args, **kwargs):
        return self.forward(*args, **kwargs).loss

    def log_outputs(self, outputs, log_fn=print, stage: str = ""):
        loss = outputs.get("loss") if isinstance(outputs, dict) else getattr(outputs, "loss", None)
        if loss is not None:
            log_fn({f"{self.loss_name}_loss": loss})

    def set_preprocessor_trainable(self, trainable: bool) -> None:  # specvit.py:104-115
        if self.preprocessor is None:
            return
        if hasattr(self.preprocessor, "set_qk_trainable"):
            self.preprocessor.set_qk_trainable(trainable)
        elif hasattr(self.preprocessor, "freeze"):
            self.preprocessor.freeze(not trainable)
        else:
            for param in self.preprocessor.parameters():
                param.requires_grad = trainable

    # ---- arena plumbing ---------------------------------------------------------------------
    def _param_slots(self):
        """(state_dict name, module, attribute) for every arena-backed parameter."""
        out = []
        for name in self._layout.entries:
            mod_path, attr = name.rsplit(".", 1)
            mod = self.get_submodule(mod_path)
            out.append((name, mod, attr))
        return out

    def _bind_arena(self, arena: ParamArena, init: bool) -> None:
        self._arena = arena
        self._engines = {}
        for name, mod, attr in self._param_slots():
            view = arena.view(name)
            old = getattr(mod, attr, None)
            if isinstance(old, nn.Parameter):
                old.data = view
                if old.grad is not None:
                    old.grad = None
            else:
                setattr(mod, attr, nn.Parameter(view))
        self._param_names = [n for n, _ in self.named_parameters() if n in self._layout.entries]
        self._param_list = [p for n, p in self.named_parameters() if n in self._layout.entries]
        arena.watch = self._param_list
        if init:
            self.init_weights()

    @torch.no_grad()
    def init_weights(self) -> None:
        """HF ViTPreTrainedModel._init_weights (HF:384-393): Linear/Conv weights trunc_normal(0, 0.02),
        biases 0, LayerNorm 1/0; cls_token / position_embeddings keep their randn init
        (embedding.py:47,65-67 -- `_init_weights` only matches HF's own ViTEmbeddings)."""
        std = self.config.initializer_range
        for name, p in self.named_parameters():
            if name not in self._layout.entries:
                continue
            if name.endswith("cls_token") or name.endswith("position_embeddings"):
                p.copy_(torch.randn(p.shape))
            elif "layernorm" in name:
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            elif name.endswith("bias"):
                p.zero_()
            else:
                w = torch.empty(p.shape)
                nn.init.trunc_normal_(w, mean=0.0, std=std)
                p.copy_(w)
        if self.preprocessor is not None:
            # Reference quirk, reproduced on purpose: MyViT.__init__ ends with self.init_weights() (specvit.py:57), and HF's
            # _init_weights reaches EVERY nn.Linear below MyViT -- including PrefilledAttention's q_lin / k_lin / v_lin, whose
            # eigenvector prefill (attention.py:58-72) is therefore overwritten by trunc_normal(0, 0.02).  PrefilledLinear
            # (ZCA / PCA) is not an nn.Linear and keeps its matrix.
            for mod in self.preprocessor.modules():
                if isinstance(mod, nn.Linear):
                    nn.init.trunc_normal_(mod.weight, mean=0.0, std=std)
                    if mod.bias is not None:
                        mod.bias.zero_()

    def _apply(self, fn, recurse=True):
        super()._apply(fn)
        # nn.Module._apply gave every parameter its own storage; re-flatten into a fresh arena
        ref = self.vit.layernorm.weight
        if ref.dtype != torch.float32:
            raise RuntimeError("vit_b200: parameters stay fp32 (master weights); select bf16 compute with "
                               "precision='bf16-mixed' instead of casting the module")
        new = ParamArena(self._layout, ref.device, with_shadow=True)
        for name, mod, attr in self._param_slots():
            new.view(name).copy_(getattr(mod, attr).data)
        self._bind_arena(new, init=False)
        return self

    def _engine(self, batch: int) -> ViTEngine:
        eng = self._engines.get(batch)
        if eng is None:
            if self._arena.data.device.type != "cuda":
                raise RuntimeError("vit_b200 has no CPU path: move the model to a CUDA (sm_100a) device first")
            h = self._opt_hyper
            eng = ViTEngine(self.config, self._arena, batch, self.precision, self._loss_kind, seed=self._seed,
                            lr=h["lr"], betas=h["betas"], eps=h["eps"], weight_decay=h["weight_decay"],
                            max_norm=h["max_norm"])
            eng.gloss = torch.ones(1, dtype=torch.float32, device=eng.device)
            if len(self._engines) >= 4:  # keep at most a few batch sizes resident
                self._engines.pop(next(iter(self._engines)))
            self._engines[batch] = eng
        return eng

    def _stage_inputs(self, eng: ViTEngine, x: torch.Tensor, labels: Optional[torch.Tensor]) -> None:
        """(already preprocessed) pixels and labels -> the engine's input buffers."""
        c = self.config
        if x.dim() != 2 or x.shape[1] != c.image_size:
            raise ValueError(f"expected pixel_values of shape [B, {c.image_size}], got {tuple(x.shape)}")
        eng.x.copy_(x, non_blocking=True)
        self._stage_labels(eng, labels)

    def _stage_labels(self, eng: ViTEngine, labels: Optional[torch.Tensor]) -> None:
        if labels is not None:
            if self._loss_kind == _lib.LOSS_CE:
                eng.labels.copy_(labels.reshape(-1), non_blocking=True)
            else:
                lab = labels.reshape(-1)
                if lab.numel() != eng.labels.numel():
                    raise ValueError(f"labels have {lab.numel()} elements, expected {eng.labels.numel()}")
                eng.labels.copy_(lab, non_blocking=True)  # copy_ casts to float like labels.view(-1).float()

    def _raw_buffer(self, eng: ViTEngine) -> torch.Tensor:
        """Where RAW spectra go for `eng`: the pixel buffer itself, or -- with a preprocessor in front -- a
        [B, input_dim] staging buffer that the preprocessor kernels turn into pixels."""
        pre = self.preprocessor
        if pre is None:
            return eng.x
        if not hasattr(pre, "forward_into"):
            raise NotImplementedError(f"vit_b200: no kernel path for preprocessor {type(pre).__name__}")
        raw = eng.__dict__.get("x_raw")
        if raw is None:
            raw = eng.x_raw = torch.empty(eng.B, self.input_dim, dtype=torch.float32, device=eng.device)
        return raw

    def _stage_raw(self, eng: ViTEngine, x: torch.Tensor, labels: Optional[torch.Tensor], run_pre: bool = True) -> None:
        """RAW spectra -> [preprocessor kernel ->] the engine's pixel buffer (TrainStep / EvalStep, which drive the engine
        without autograd).  run_pre=False: only stage the raw spectra (TrainStep with a TRAINABLE matrix runs the
        preprocessor inside its captured step, next to its weight gradient and optimizer update)."""
        if self.preprocessor is None:
            return self._stage_inputs(eng, x, labels)
        raw = self._raw_buffer(eng)
        if x.dim() != 2 or x.shape[1] != self.input_dim:
            raise ValueError(f"expected spectra of shape [B, {self.input_dim}], got {tuple(x.shape)}")
        raw.copy_(x, non_blocking=True)
        if run_pre:
            self.preprocessor.forward_into(raw, eng.x)
        self._stage_labels(eng, labels)

    def _has_inner_hooks(self) -> bool:
        for m in self.modules():
            if m is self:
                continue
            if m._forward_hooks or m._forward_pre_hooks:
                return True
        return False

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, pixel_values, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        if self.preprocessor is not None:
            pixel_values = self.preprocessor(pixel_values)
        B = pixel_values.shape[0]
        c = self.config
        grad_needed = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        modular = (not grad_needed) and (self._has_inner_hooks() or bool(output_attentions))
        hidden_states = attentions = None
        if modular:
            from . import ops

            out = ops.modular_forward(self, pixel_values, labels, want_attn=bool(output_attentions))
            loss, logits = out["loss"], out["logits"]
            hidden_states = out["hidden_states"] if output_hidden_states else None
            attentions = out["attentions"] if output_attentions else None
        else:
            eng = self._engine(B)
            eng.cls_only = not output_hidden_states   # every token's last hidden state is only computed when asked for
            if grad_needed:
                loss, logits = _FusedViTFunction.apply(self, eng, pixel_values, labels, self.training,
                                                       self._param_names, *self._param_list)
                if labels is None:
                    loss = None
            else:
                self._stage_inputs(eng, pixel_values, labels)
                if self.training:
                    eng.advance_rng()
                eng.forward(train=self.training, with_labels=labels is not None)
                logits = eng.logits.clone()
                loss = eng.loss[0].clone() if labels is not None else None
            if output_hidden_states:
                hidden_states = tuple(z.view(B, c.tokens, c.hidden_size).clone() for z in eng.z)
        return ModelOutput(loss=loss, logits=logits, hidden_states=hidden_states, attentions=attentions)
