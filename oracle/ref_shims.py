"""TEST INFRASTRUCTURE ONLY -- import shims for running the *unmodified* reference
(`/root/reference/src/models`) inside this container.

Only `tests/golden/gen_golden.py` and the `-m "not gpu"` oracle-validation tests use
this file, and only when `/root/reference` exists (it does not exist on the GPU box).
Nothing under `vit_b200/` may import it.

Why shims are needed:
  1. `src/models/specvit.py:7` imports `src.basemodule`, which imports `lightning`
     (`src/basemodule.py:4`) and `src.dataloader` -> `h5py` (`src/dataloader/base.py:15`).
     Neither is installed; neither is used by the model arithmetic.
  2. The reference pins transformers==4.56.0 (`requirements.txt:57`); the image has 5.5.0,
     where `PreTrainedModel.init_weights()` expects `all_tied_weights_keys`, which
     `MyViT.__init__` never sets because it does not call `post_init()`
     (`src/models/specvit.py:57`).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VIT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def _stub_lightning() -> None:
    if "lightning" in sys.modules:
        return
    import torch.nn as nn

    L = types.ModuleType("lightning")

    class _Trainer:  # never instantiated by the model path
        def __init__(self, *a, **k):
            pass

    class _Callback:
        pass

    class _LM(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class _DM:
        def __init__(self, *a, **k):
            pass

    L.LightningModule = _LM
    L.LightningDataModule = _DM
    L.Trainer = _Trainer
    L.Callback = _Callback
    L.seed_everything = lambda *a, **k: None
    pl = types.ModuleType("lightning.pytorch")
    cb = types.ModuleType("lightning.pytorch.callbacks")
    cb.Callback = _Callback
    cb.ModelCheckpoint = _Callback
    cb.EarlyStopping = _Callback
    pl.callbacks = cb
    L.pytorch = pl
    sys.modules["lightning"] = L
    sys.modules["lightning.pytorch"] = pl
    sys.modules["lightning.pytorch.callbacks"] = cb


def _stub_h5py() -> None:
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")


def import_reference_models():
    """Return the reference's `src.models` package (unmodified code, shimmed imports)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _stub_lightning()
    _stub_h5py()
    import transformers.modeling_utils as mu

    if not hasattr(mu.PreTrainedModel, "all_tied_weights_keys"):
        mu.PreTrainedModel.all_tied_weights_keys = {}
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.models as ref_models  # noqa: E402

    return ref_models


def reference_get_model(config: dict):
    """`src/models/builder.py:136` get_model on a deep copy of `config`."""
    import copy

    return import_reference_models().get_model(copy.deepcopy(config))
