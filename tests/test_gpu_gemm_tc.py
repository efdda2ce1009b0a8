"""GPU: the tcgen05/TMA GEMM kernels (csrc/gemm_tc.cu) against fp32 matmuls of the same bf16 operands, and
against the SIMT kernels they replace in bf16 mode.  Direct C-ABI calls."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K)
    (8256, 96, 32), (8256, 32, 32), (8256, 128, 32), (8256, 32, 128),  # baseline.yaml: QKV, proj, MLP up/down
    (645, 64, 48), (129, 8, 8), (1000, 768, 768), (300, 3072, 768), (517, 200, 136),
]


def _setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vit_b200 import _lib

    return _lib, _lib.load(), torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_linear_fwd(M, N, K, act):
    _lib, lib, dev = _setup()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev)
    y = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    ya = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16) if act else None
    _lib.check(lib.vitb200_tc_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                         ya.data_ptr() if act else None, M, N, K, act, _st()), "tc fwd")
    ref = x.float() @ w.float().t() + b
    assert rel_err(y, ref) < 1e-2
    if act:
        assert rel_err(ya, torch.nn.functional.gelu(ref.bfloat16().float())) < 1e-2
    # and the SIMT kernel agrees (same operands, same epilogue)
    y2 = torch.empty_like(y)
    ya2 = torch.empty_like(y) if act else None
    old = lib.vitb200_set_gemm_mode(1)
    try:
        _lib.check(lib.vitb200_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y2.data_ptr(),
                                          ya2.data_ptr() if act else None, M, N, K, act, _lib.BF16, _st()), "simt fwd")
    finally:
        lib.vitb200_set_gemm_mode(old)
    assert rel_err(y, y2) < 1e-2


@pytest.mark.parametrize("with_pre", [False, True])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_linear_dgrad(M, N, K, with_pre):
    _lib, lib, dev = _setup()
    g = torch.Generator(device="cpu").manual_seed(M * 3 + N + K)
    dy = torch.randn(M, N, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / N ** 0.5).to(dev).bfloat16()
    pre = torch.randn(M, K, generator=g).to(dev).bfloat16() if with_pre else None
    dx = torch.full((M, K), float("nan"), device=dev, dtype=torch.bfloat16)
    _lib.check(lib.vitb200_tc_linear_dgrad(dy.data_ptr(), w.data_ptr(), pre.data_ptr() if with_pre else None,
                                           dx.data_ptr(), M, N, K, _st()), "tc dgrad")
    ref = dy.float() @ w.float()
    if with_pre:
        p = pre.float().requires_grad_(True)
        torch.nn.functional.gelu(p).backward(ref.bfloat16().float())
        ref = p.grad
    assert rel_err(dx, ref) < 1e-2


@pytest.mark.parametrize("accumulate", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_linear_wgrad(M, N, K, accumulate):
    _lib, lib, dev = _setup()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + K)
    dy = torch.randn(M, N, generator=g).to(dev).bfloat16()
    x = torch.randn(M, K, generator=g).to(dev).bfloat16()
    dw0 = torch.randn(N, K, generator=g).to(dev)
    db0 = torch.randn(N, generator=g).to(dev)
    dw, db = dw0.clone(), db0.clone()
    ws = torch.zeros(int(lib.vitb200_tc_linear_wgrad_ws_bytes(M, N, K)) + 4096, dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(2):  # twice: the tickets must reset themselves, and the result must be bitwise reproducible
        dw.copy_(dw0); db.copy_(db0)
        _lib.check(lib.vitb200_tc_linear_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, N, K,
                                               accumulate, ws.data_ptr(), _st()), "tc wgrad")
        outs.append((dw.clone(), db.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref_w = dy.float().t() @ x.float() + (dw0 if accumulate else 0)
    ref_b = dy.float().sum(0) + (db0 if accumulate else 0)
    assert rel_err(dw, ref_w) < 2e-4
    assert rel_err(db, ref_b) < 2e-4
