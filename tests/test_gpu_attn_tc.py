"""GPU: tcgen05 attention kernels (csrc/attention_tc.cu) vs. the SIMT kernels (same C-ABI entry point, mode
switched) and vs. a plain fp32 torch attention built from the same bf16 inputs and the same dropout mask."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ref(q, k, v, scale, mask, p):
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    pr = torch.softmax(s, dim=-1)
    prd = pr * mask / (1 - p) if p > 0 else pr
    return torch.matmul(prd, v)


@pytest.mark.parametrize("rope", [False, True])
@pytest.mark.parametrize("B,T,heads,d,p", [(3, 129, 2, 16, 0.1), (2, 33, 4, 16, 0.0), (2, 160, 2, 32, 0.1),
                                            (5, 128, 2, 16, 0.1), (1, 130, 1, 32, 0.0)])
def test_attn_tc_matches_simt_and_torch(B, T, heads, d, p, rope):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vit_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    H = heads * d
    assert lib.vitb200_attn_tc_supported(T, d, 3 * H, H) == 1
    g = torch.Generator(device="cpu").manual_seed(T * 7 + d)
    qkv = torch.randn(B * T, 3 * H, generator=g).to(dev).bfloat16()
    dctx = torch.randn(B * T, H, generator=g).to(dev).bfloat16()
    rng = torch.tensor([1234, 5], dtype=torch.int64, device=dev)
    cos = sin = None
    if rope:
        inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
        fr = torch.outer(torch.arange(T, dtype=torch.float32), inv)
        cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
    scale = 1.0 / math.sqrt(d)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    es = 2

    def run(mode):
        old = lib.vitb200_set_attn_mode(mode)
        try:
            ctx = torch.full((B * T, H), float("nan"), device=dev, dtype=torch.bfloat16)
            lse = torch.zeros(B, heads, T, device=dev)
            dsum = torch.zeros(B, heads, T, device=dev)
            dq = torch.full((B * T, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
            q = qkv.data_ptr()
            _lib.check(lib.vitb200_attn_fwd(q, q + H * es, q + 2 * H * es, 3 * H, ctx.data_ptr(), lse.data_ptr(), P(cos),
                                            P(sin), B, T, heads, d, scale, p, rng.data_ptr(), 4, _lib.BF16, st), "fwd")
            dqp = dq.data_ptr()
            _lib.check(lib.vitb200_attn_bwd(q, q + H * es, q + 2 * H * es, 3 * H, ctx.data_ptr(), dctx.data_ptr(),
                                            lse.data_ptr(), dsum.data_ptr(), dqp, dqp + H * es, dqp + 2 * H * es, 3 * H,
                                            P(cos), P(sin), B, T, heads, d, scale, p, rng.data_ptr(), 4, _lib.BF16, st), "bwd")
            torch.cuda.synchronize()
            return ctx.float(), lse, dq.float()
        finally:
            lib.vitb200_set_attn_mode(old)

    ctx_tc, lse_tc, dq_tc = run(0)
    ctx_tc2, _, dq_tc2 = run(0)
    assert torch.equal(ctx_tc, ctx_tc2) and torch.equal(dq_tc, dq_tc2), "tcgen05 attention must be deterministic"
    ctx_s, lse_s, dq_s = run(1)
    assert rel_err(ctx_tc, ctx_s) < 2e-2 and rel_err(lse_tc, lse_s) < 1e-3 and rel_err(dq_tc, dq_s) < 3e-2
    # fp32 torch reference with the kernels' own dropout mask
    Tpad = (T + 7) // 8 * 8
    mask = torch.ones(B, heads, T, T, device=dev)
    if p > 0:
        m = torch.empty(B * heads * T * Tpad, dtype=torch.uint8, device=dev)
        _lib.check(lib.vitb200_dropout_mask(m.data_ptr(), m.numel(), p, rng.data_ptr(), 4, st), "mask")
        mask = m.view(B, heads, T, Tpad)[..., :T].float()
    x = qkv.float().view(B, T, 3, heads, d).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q_, k_, v_ = x[0], x[1], x[2]
    if rope:
        def rot(t):
            t1, t2 = t.chunk(2, dim=-1)
            c = torch.cat([cos, cos], -1)[None, None]
            s = torch.cat([sin, sin], -1)[None, None]
            return t * c + torch.cat([-t2, t1], -1) * s
        q_, k_ = rot(q_), rot(k_)
    out = _ref(q_, k_, v_, scale, mask, p)                                  # [B, heads, T, d]
    out.backward(dctx.float().view(B, T, heads, d).permute(0, 2, 1, 3))
    ref_ctx = out.permute(0, 2, 1, 3).reshape(B * T, H)
    ref_dqkv = x.grad.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * H)
    assert rel_err(ctx_tc, ref_ctx) < 2e-2
    assert rel_err(dq_tc, ref_dqkv) < 3e-2


@pytest.mark.parametrize("rope", [False, True])
@pytest.mark.parametrize("B,T,heads,d,p", [(2, 510, 2, 16, 0.1), (1, 513, 2, 16, 0.1), (2, 200, 2, 32, 0.1),
                                            (1, 257, 1, 16, 0.0), (1, 129, 2, 16, 0.1), (1, 1030, 2, 16, 0.0)])
def test_attn_tc_key_blocked_matches_torch(B, T, heads, d, p, rope):
    """Flash forward + the per-key-block tcgen05 backward (the cross-check of the flash backward) for any T vs. a plain
    fp32 torch attention built from the same bf16 inputs and the kernels' own dropout mask; run twice for bitwise
    reproducibility."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vit_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    H = heads * d
    assert lib.vitb200_attn_tc_blocked_supported(T, d, 3 * H, H) == 1
    g = torch.Generator(device="cpu").manual_seed(T * 3 + d)
    qkv = torch.randn(B * T, 3 * H, generator=g).to(dev).bfloat16()
    dctx = torch.randn(B * T, H, generator=g).to(dev).bfloat16()
    rng = torch.tensor([77, 3], dtype=torch.int64, device=dev)
    cos = sin = None
    if rope:
        inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
        fr = torch.outer(torch.arange(T, dtype=torch.float32), inv)
        cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
    scale = 1.0 / math.sqrt(d)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    ws = torch.zeros(int(lib.vitb200_attn_tc_blocked_ws_bytes(B, T, heads, d)), dtype=torch.uint8, device=dev)

    def run():
        ctx = torch.full((B * T, H), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B, heads, T, device=dev)
        dq = torch.full((B * T, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
        _lib.check(lib.vitb200_attn_flash_fwd(qkv.data_ptr(), ctx.data_ptr(), lse.data_ptr(), P(cos), P(sin), B, T,
                                              heads, d, scale, p, rng.data_ptr(), 4, st), "fwd")
        _lib.check(lib.vitb200_attn_tc_blocked_bwd(qkv.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(),
                                                   dq.data_ptr(), P(cos), P(sin), B, T, heads, d, scale, p,
                                                   rng.data_ptr(), 4, ws.data_ptr(), st), "bwd")
        torch.cuda.synchronize()
        return ctx.float(), lse, dq.float()

    ctx_a, lse_a, dq_a = run()
    ctx_b, lse_b, dq_b = run()
    assert torch.equal(ctx_a, ctx_b) and torch.equal(dq_a, dq_b) and torch.equal(lse_a, lse_b)
    assert torch.isfinite(ctx_a).all() and torch.isfinite(dq_a).all()
    Tpad = (T + 7) // 8 * 8
    mask = torch.ones(B, heads, T, T, device=dev)
    if p > 0:
        m = torch.empty(B * heads * T * Tpad, dtype=torch.uint8, device=dev)
        _lib.check(lib.vitb200_dropout_mask(m.data_ptr(), m.numel(), p, rng.data_ptr(), 4, st), "mask")
        mask = m.view(B, heads, T, Tpad)[..., :T].float()
    x = qkv.float().view(B, T, 3, heads, d).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q_, k_, v_ = x[0], x[1], x[2]
    if rope:
        def rot(t):
            t1, t2 = t.chunk(2, dim=-1)
            c = torch.cat([cos, cos], -1)[None, None]
            s = torch.cat([sin, sin], -1)[None, None]
            return t * c + torch.cat([-t2, t1], -1) * s
        q_, k_ = rot(q_), rot(k_)
    s_ = torch.matmul(q_, k_.transpose(-1, -2)) * scale
    out = _ref(q_, k_, v_, scale, mask, p)
    out.backward(dctx.float().view(B, T, heads, d).permute(0, 2, 1, 3))
    ref_ctx = out.permute(0, 2, 1, 3).reshape(B * T, H)
    ref_dqkv = x.grad.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * H)
    assert rel_err(lse_a, torch.logsumexp(s_, dim=-1).detach()) < 1e-2
    assert rel_err(ctx_a, ref_ctx) < 2e-2, rel_err(ctx_a, ref_ctx)
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        e = rel_err(dq_a[:, sl], ref_dqkv[:, sl])
        assert e < 3e-2, (name, e)


@pytest.mark.parametrize("rope", [False, True])
@pytest.mark.parametrize("B,T,heads,d,p", [(2, 510, 2, 16, 0.1), (1, 2034, 2, 16, 0.1), (2, 200, 2, 32, 0.1), (1, 257, 1, 16, 0.0),
                                            (3, 129, 2, 16, 0.1), (2, 128, 2, 64, 0.1), (1, 300, 2, 64, 0.0), (1, 1030, 4, 32, 0.1)])
def test_attn_flash_fwd_matches_torch_and_blocked_backward(B, T, heads, d, p, rope):
    """One-launch flash attention forward (in-kernel key/value loop, online softmax) vs. a plain fp32 torch attention
    built from the same bf16 inputs and the kernels' own dropout mask; its lse / ctx feed the existing backward kernels
    (key-blocked tcgen05 for d in {16, 32}) whose dq / dk / dv are checked against torch autograd; bitwise reproducible."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vit_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    H = heads * d
    assert lib.vitb200_attn_flash_supported(T, d, 3 * H, H) == 1
    g = torch.Generator(device="cpu").manual_seed(T * 5 + d)
    qkv = torch.randn(B * T, 3 * H, generator=g).to(dev).bfloat16()
    dctx = torch.randn(B * T, H, generator=g).to(dev).bfloat16()
    rng = torch.tensor([99, 2], dtype=torch.int64, device=dev)
    cos = sin = None
    if rope:
        inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
        fr = torch.outer(torch.arange(T, dtype=torch.float32), inv)
        cos, sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
    scale = 1.0 / math.sqrt(d)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: None if t is None else t.data_ptr()  # noqa: E731

    def run():
        ctx = torch.full((B * T, H), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B, heads, T, device=dev)
        _lib.check(lib.vitb200_attn_flash_fwd(qkv.data_ptr(), ctx.data_ptr(), lse.data_ptr(), P(cos), P(sin), B, T, heads, d,
                                              scale, p, rng.data_ptr(), 4, st), "flash fwd")
        torch.cuda.synchronize()
        return ctx, lse

    ctx_a, lse_a = run()
    ctx_b, lse_b = run()
    assert torch.equal(ctx_a, ctx_b) and torch.equal(lse_a, lse_b)
    assert torch.isfinite(ctx_a.float()).all()
    Tpad = (T + 7) // 8 * 8
    mask = torch.ones(B, heads, T, T, device=dev)
    if p > 0:
        m = torch.empty(B * heads * T * Tpad, dtype=torch.uint8, device=dev)
        _lib.check(lib.vitb200_dropout_mask(m.data_ptr(), m.numel(), p, rng.data_ptr(), 4, st), "mask")
        mask = m.view(B, heads, T, Tpad)[..., :T].float()
    x = qkv.float().view(B, T, 3, heads, d).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    q_, k_, v_ = x[0], x[1], x[2]
    if rope:
        def rot(t):
            t1, t2 = t.chunk(2, dim=-1)
            c = torch.cat([cos, cos], -1)[None, None]
            s = torch.cat([sin, sin], -1)[None, None]
            return t * c + torch.cat([-t2, t1], -1) * s
        q_, k_ = rot(q_), rot(k_)
    s_ = torch.matmul(q_, k_.transpose(-1, -2)) * scale
    out = _ref(q_, k_, v_, scale, mask, p)
    out.backward(dctx.float().view(B, T, heads, d).permute(0, 2, 1, 3))
    ref_ctx = out.permute(0, 2, 1, 3).reshape(B * T, H)
    ref_dqkv = x.grad.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * H)
    assert rel_err(lse_a, torch.logsumexp(s_, dim=-1).detach()) < 1e-2
    assert rel_err(ctx_a.float(), ref_ctx) < 2e-2, rel_err(ctx_a.float(), ref_ctx)
    # backward from the flash forward's ctx / lse
    dq = torch.full((B * T, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
    if d in (16, 32):
        ws = torch.zeros(int(lib.vitb200_attn_tc_blocked_ws_bytes(B, T, heads, d)), dtype=torch.uint8, device=dev)
        _lib.check(lib.vitb200_attn_tc_blocked_bwd(qkv.data_ptr(), ctx_a.data_ptr(), dctx.data_ptr(), lse_a.data_ptr(),
                                                   dq.data_ptr(), P(cos), P(sin), B, T, heads, d, scale, p,
                                                   rng.data_ptr(), 4, ws.data_ptr(), st), "bwd")
    else:
        dsum = torch.zeros(B, heads, T, device=dev)
        q, dqp, es = qkv.data_ptr(), dq.data_ptr(), 2
        _lib.check(lib.vitb200_attn_bwd(q, q + H * es, q + 2 * H * es, 3 * H, ctx_a.data_ptr(), dctx.data_ptr(),
                                        lse_a.data_ptr(), dsum.data_ptr(), dqp, dqp + H * es, dqp + 2 * H * es, 3 * H,
                                        P(cos), P(sin), B, T, heads, d, scale, p, rng.data_ptr(), 4, _lib.BF16, st), "bwd")
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        e = rel_err(dq.float()[:, sl], ref_dqkv[:, sl])
        assert e < 3e-2, (name, e)
    # the one-launch flash backward (dK / dV resident in TMEM over the query loop, dQ partials summed in key-block
    # order): vs torch autograd, bitwise reproducible, and -- same products, same accumulation order -- bit-identical
    # to the per-key-block launches where those run the tcgen05 tiles for every row (T not of the form 128 n + 1)
    fws = torch.zeros(int(lib.vitb200_attn_flash_bwd_ws_bytes(B, T, heads, d)), dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(2):
        df = torch.full((B * T, 3 * H), float("nan"), device=dev, dtype=torch.bfloat16)
        _lib.check(lib.vitb200_attn_flash_bwd(qkv.data_ptr(), ctx_a.data_ptr(), dctx.data_ptr(), lse_a.data_ptr(),
                                              df.data_ptr(), P(cos), P(sin), B, T, heads, d, scale, p, rng.data_ptr(), 4,
                                              fws.data_ptr(), st), "flash bwd")
        torch.cuda.synchronize()
        outs.append(df)
    assert torch.equal(outs[0], outs[1])
    assert torch.isfinite(outs[0].float()).all()
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        e = rel_err(outs[0].float()[:, sl], ref_dqkv[:, sl])
        assert e < 3e-2, ("flash " + name, e)
    if d in (16, 32) and (T % 128 != 1 or rope):
        assert torch.equal(outs[0], dq)
