import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_VARIANTS = ["baseline", "config2", "rope", "learned", "cnn", "stride8", "pad48", "cls", "l1",
                   "h64multi", "h128d64rope", "heads8d4rope", "long510", "long2034"]
# input-preprocessor variants (src/models/builder.py:43-133): the fixture also carries the covariance statistics
PRE_VARIANTS = ["pre_zca_full", "pre_zca_r32", "pre_pca_r128", "pre_attn_r64"]


def config_with_cov(fix, tmp_path):
    """The fixture's config with warmup.cov_path pointing at a statistics file written from the fixture."""
    import copy

    import torch

    cfg = copy.deepcopy(fix["config"])
    st = dict(fix["stats"])
    d = st["eigvecs"].shape[0]
    st["cov"] = torch.zeros(d, d)   # required key (src/utils.py:64); the builder never reads it
    path = os.path.join(str(tmp_path), "cov.pt")
    torch.save(st, path)
    cfg["warmup"]["cov_path"] = path
    return cfg


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get


def rel_err(a, b, floor=1e-30):
    """max |a-b| / max(|b|, tiny): the 'max relative error' of BASELINE.json's north_star, taken
    relative to the largest magnitude of the reference tensor (per-element relative error is
    meaningless for entries that are ~0).  `floor` bounds the denominator from below for tensors that
    are mathematically zero (e.g. d loss / d key.bias: softmax is shift-invariant)."""
    import torch

    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(float(b.abs().max()), floor)
    return float((a - b).abs().max()) / denom


def grad_floor(grads):
    """1e-3 x the largest gradient magnitude in the model: the scale below which a gradient tensor is noise."""
    return 1e-3 * max(float(g.detach().abs().max()) for g in grads.values())
