"""vit_b200 -- B200-native (sm_100a) ViT encoder step behind the ViskaWei/VIT model interface.

Public surface (mirrors the reference):
    get_model(config) -> MyViT            (src/models/builder.py:136)
    get_vit_config(config)                (src/models/builder.py:200)
    MyViT, ViTLModule, FusedClipAdamW, TrainStep
"""
__version__ = "0.1.0"

_LAZY = {
    "get_model": "vit_b200.builder",
    "get_vit_config": "vit_b200.builder",
    "VitConfig": "vit_b200.builder",
    "MyViT": "vit_b200.model",
    "ViTLModule": "vit_b200.lightning_module",
    "FusedClipAdamW": "vit_b200.optim",
    "TrainStep": "vit_b200.step",
    "EvalStep": "vit_b200.step",
    "ViTEngine": "vit_b200.engine",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(name)
