"""GPU parity tests of the rows either side of the encoder step (SURVEY.md 8a a2/a2', 8f ranks 1-4): the input
preprocessor stage, the device-resident data hand-off, on-device evaluation metrics, checkpoint resume.  Everything goes
through the C ABI; the checkers are the golden fixtures generated from the unmodified reference (tests/golden/pre_*.pt),
the oracle and plain torch ops on the CPU."""
import copy

import pytest
import torch

from conftest import PRE_VARIANTS, config_with_cov, grad_floor, rel_err
from oracle import vit_oracle as vo

pytestmark = pytest.mark.gpu

TOL = {"32": 1e-4, "bf16-mixed": 2e-2}
GTOL = {"32": 2e-4, "bf16-mixed": 4e-2}

BASE = {"model": dict(name="vit", task_type="reg", image_size=512, patch_size=32, hidden_size=32, num_hidden_layers=2,
                      num_attention_heads=2, stride_size=32, proj_fn="SW"),
        "loss": {"name": "mae"}, "data": {"param": "log_g"}}


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _st():
    return torch.cuda.current_stream().cuda_stream


def _build_pre(fix, precision, dev, tmp_path):
    from vit_b200 import get_model
    from vit_b200.preprocessor import clear_cov_cache

    clear_cov_cache()
    m = get_model(config_with_cov(fix, tmp_path), precision=precision, device=dev)
    m.load_state_dict(fix["state_dict"], strict=True)
    return m


def _inputs(fix, dev):
    x, _ = vo.synthetic_batch(fix["batch"], fix["config"]["model"]["image_size"], seed=fix["x_seed"], kind=fix["x_kind"])
    return x.to(dev), fix["labels"].to(dev)


# ------------------------------------------------------------------------------------------------
# kernels, directly
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["32", "bf16-mixed"])
@pytest.mark.parametrize("L,P,S,proj", [(256, 32, 32, "SW"), (256, 32, 8, "SW"), (1000, 48, 48, "SW"), (1000, 32, 24, "CNN")])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_patch_embed_dgrad_vs_autograd(L, P, S, proj, precision, p_drop):
    """d loss / d pixel: the unfold + Linear backward w.r.t. the spectrum, incl. overlapping windows, the zero-padded
    tail window (no pixel behind it) and the embedding dropout mask regenerated from (seed, step)."""
    from vit_b200 import _lib
    from vit_b200.builder import VitConfig

    dev = _cuda()
    lib = _lib.load()
    c = VitConfig(image_size=L, patch_size=P, stride_size=S, proj_fn=proj)
    B, H, Np, T1 = 3, 32, c.num_patches, c.num_patches + 1
    g = torch.Generator().manual_seed(L + P + S)
    dz = torch.randn(B, T1, H, generator=g)
    w = torch.randn(H, P, generator=g) * 0.2
    bf = precision == "bf16-mixed"
    dt = _lib.BF16 if bf else _lib.F32
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
    mask = torch.ones(B * T1 * H, dtype=torch.uint8, device=dev)
    if p_drop > 0:
        _lib.check(lib.vitb200_dropout_mask(mask.data_ptr(), mask.numel(), p_drop, rng.data_ptr(), _lib.SITE_EMB, _st()), "mask")
    keep = mask.view(B, T1, H).cpu().float() / (1.0 - p_drop)
    wd = w.to(dev).bfloat16().contiguous() if bf else w.to(dev).contiguous()
    dzd = dz.to(dev).contiguous()
    dx = torch.full((B, L), float("nan"), device=dev)
    _lib.check(lib.vitb200_patch_embed_dgrad(dzd.data_ptr(), wd.data_ptr(), dx.data_ptr(), B, L, P, S, Np, c.n_valid, H,
                                             p_drop, rng.data_ptr(), _lib.SITE_EMB, dt, _st()), "dgrad")
    # reference: autograd through unfold + linear on the CPU (same operand rounding in bf16 mode)
    x = torch.zeros(B, L, requires_grad=True)
    wr = w.bfloat16().float() if bf else w
    gtok = (dz * keep)[:, 1:1 + c.n_valid]
    if bf:
        gtok = gtok.bfloat16().float()
    tok = torch.nn.functional.linear(x.unfold(1, P, S)[:, :c.n_valid], wr)
    (tok * gtok).sum().backward()
    assert torch.isfinite(dx).all()
    assert rel_err(dx, x.grad) < (2e-2 if bf else 1e-5)
    if c.n_valid < Np or (c.n_valid - 1) * S + P < L:   # pixels no window covers get an exact zero
        covered = torch.zeros(L, dtype=torch.bool)
        for n in range(c.n_valid):
            covered[n * S:n * S + P] = True
        assert float(dx.cpu()[:, ~covered].abs().max()) == 0.0 if (~covered).any() else True


def test_cast_f32_round_trip():
    from vit_b200 import _lib

    dev = _cuda()
    lib = _lib.load()
    x = torch.randn(3, 1000, device=dev)
    xb = torch.empty(3, 1000, dtype=torch.bfloat16, device=dev)
    y = torch.empty(3, 1000, device=dev)
    _lib.check(lib.vitb200_cast_bf16(x.data_ptr(), xb.data_ptr(), x.numel(), _st()), "cast_bf16")
    _lib.check(lib.vitb200_cast_f32(xb.data_ptr(), y.data_ptr(), x.numel(), _st()), "cast_f32")
    assert torch.equal(y, x.bfloat16().float())
    assert lib.vitb200_cast_f32(xb.data_ptr(), y.data_ptr(), 3, _st()) == -2


@pytest.mark.parametrize("M,N,K", [(64, 4096, 4096), (4, 256, 256), (300, 136, 520), (64, 1000, 4096), (129, 8, 8)])
@pytest.mark.parametrize("with_bias", [True, False])
def test_tc_prelinear_split_k_vs_fp32_matmul(M, N, K, with_bias):
    """Skinny-M, weight-streaming Linear of the preprocessor stage: bf16 operands on tcgen05, contraction split across one
    wave of CTAs, fixed-order split sum (bitwise reproducible, tickets reset themselves), fp32 output = bf16-rounded."""
    from vit_b200 import _lib

    dev = _cuda()
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev) if with_bias else None
    ws = torch.zeros(int(lib.vitb200_tc_prelinear_ws_bytes(M, N, K)), dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(2):
        y = torch.full((M, N), float("nan"), device=dev)
        _lib.check(lib.vitb200_tc_prelinear_fwd(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(),
                                                y.data_ptr(), M, N, K, ws.data_ptr(), _st()), "prelinear")
        outs.append(y)
    assert torch.equal(outs[0], outs[1])
    ref = x.float() @ w.float().t() + (b if with_bias else 0)
    assert torch.isfinite(outs[0]).all() and rel_err(outs[0], ref) < 1e-2
    assert torch.equal(outs[0], outs[0].bfloat16().float())          # values are bf16-representable
    # and it agrees with the generic bf16 Linear (same operands) to the last bf16 bit or one ulp
    y2 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.vitb200_linear_fwd(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), y2.data_ptr(), None,
                                      M, N, K, 0, _lib.BF16, _st()), "linear_fwd")
    assert rel_err(outs[0], y2.float()) < 1e-2


# ------------------------------------------------------------------------------------------------
# preprocessor stage vs the reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["32", "bf16-mixed"])
@pytest.mark.parametrize("name", PRE_VARIANTS)
def test_preprocessor_forward_backward_vs_reference_golden(golden, name, precision, tmp_path):
    """ZCA (full / low-rank, frozen / trainable), PCA (changes image_size) and the attention preprocessor's q_lin: the
    preprocessed pixels, loss, logits, hidden states and EVERY gradient -- including d loss / d P and d loss / d bias
    of a trainable matrix, which flow through vitb200_patch_embed_dgrad and vitb200_linear_wgrad."""
    dev = _cuda()
    fix = golden(name)
    m = _build_pre(fix, precision, dev, tmp_path).eval()
    x, y = _inputs(fix, dev)
    tol, gtol = TOL[precision], GTOL[precision]
    with torch.no_grad():
        pre = m.preprocessor(x)
    assert pre.dtype == torch.float32 and rel_err(pre, fix["eval"]["preprocessed"]) < tol
    out = m(x, labels=y, output_hidden_states=True)
    out.loss.backward()
    assert rel_err(out.loss, fix["eval"]["loss"]) < tol
    # logits that nearly cancel (pre_pca_r128: |logits| ~ 1e-2) are held to the reference's OWN bf16-vs-fp32 deviation
    own = rel_err(fix["bf16"]["logits"], fix["eval"]["logits"]) if precision == "bf16-mixed" else 0.0
    assert rel_err(out.logits, fix["eval"]["logits"]) < max(tol, 1.5 * own)
    for mine, ref in zip(out.hidden_states, fix["eval"]["hidden_states"]):
        assert rel_err(mine[0], ref) < tol
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(fix["grads"])
    fl = grad_floor(fix["grads"])
    for k, g in fix["grads"].items():
        assert rel_err(got[k], g, fl) < gtol, (k, rel_err(got[k], g, fl))
    if precision == "bf16-mixed":
        assert rel_err(out.loss, fix["bf16"]["loss"]) < tol
        assert rel_err(out.logits, fix["bf16"]["logits"]) < tol


@pytest.mark.parametrize("precision", ["32", "bf16-mixed"])
def test_lowrank_zca_factored_kernel_vs_reference_golden(golden, precision, tmp_path, monkeypatch):
    """A FROZEN low-rank ZCA matrix runs as y = s_perp x + ((x Vr) o g) Vr^T + b (vitb200_zca_lowrank_fwd, one launch)
    instead of the dense [D, D] Linear: preprocessed pixels vs the reference's, vs the dense kernel, bitwise reproducible,
    through TrainStep's staging path; a matrix that no longer equals its factors (trained / overwritten) silently takes
    the dense path again."""
    from vit_b200 import preprocessor as vp
    from vit_b200.step import TrainStep

    dev = _cuda()
    fix = golden("pre_zca_r32")
    m = _build_pre(fix, precision, dev, tmp_path).eval()
    m.preprocessor.freeze(True)
    lin = m.preprocessor.linear
    assert vp._lowrank_state(lin, lin.weight) is not None          # factors attached by the builder and still valid
    x, y = _inputs(fix, dev)
    tol = TOL[precision]
    with torch.no_grad():
        a = m.preprocessor(x)
        b = m.preprocessor(x)
        monkeypatch.setenv("VITB200_ZCA_LOWRANK", "0")
        dense = m.preprocessor(x)
        monkeypatch.delenv("VITB200_ZCA_LOWRANK")
    assert torch.equal(a, b) and a.dtype == torch.float32
    assert rel_err(a, fix["eval"]["preprocessed"]) < tol
    assert rel_err(a, dense) < tol
    if precision == "bf16-mixed":
        assert torch.equal(a, a.bfloat16().float())                   # bf16-representable values, like autocast's output
    # the step path: raw spectra -> factored kernel -> pixel buffer -> training step
    step = TrainStep(m, fix["batch"], use_graph=True, train=False)
    l0 = float(step.step(x, y))
    assert abs(l0 - float(fix["eval"]["loss"])) < 5 * tol * max(1.0, abs(float(fix["eval"]["loss"])))
    assert rel_err(step.eng.x.view(fix["batch"], -1), fix["eval"]["preprocessed"]) < tol
    step.close()
    # overwritten matrix: the factors no longer describe it
    with torch.no_grad():
        lin.weight.mul_(1.5)
        assert vp._lowrank_state(lin, lin.weight) is None
        assert rel_err(m.preprocessor(x), 1.5 * (dense - lin.bias) + lin.bias if lin.bias is not None else 1.5 * dense) < 5 * tol


@pytest.mark.parametrize("precision", ["32", "bf16-mixed"])
@pytest.mark.parametrize("name", ["pre_zca_full", "pre_pca_r128"])
def test_train_steps_with_frozen_preprocessor(golden, name, precision, tmp_path):
    """TrainStep (one CUDA graph per step) behind a FROZEN preprocessor: raw spectra in, 3 steps vs the reference."""
    from vit_b200.step import EvalStep, TrainStep

    dev = _cuda()
    fix = golden(name)
    m = _build_pre(fix, precision, dev, tmp_path)
    x, y = _inputs(fix, dev)
    step = TrainStep(m, fix["batch"], use_graph=True, train=False)
    tol = TOL[precision]
    for i in range(3):
        loss = float(step.step(x, y))
        ref = float(fix["train3"]["losses"][i])
        assert abs(loss - ref) < 5 * tol * max(1.0, abs(ref)), (i, loss, ref)
        assert rel_err(step.eng.state[1], fix["train3"]["grad_norms"][i]) < 20 * tol
    for k, v in fix["state_dict"].items():
        if k.startswith("preprocessor."):      # buffers: never touched
            assert torch.equal(m.state_dict()[k].cpu(), v), k
    if precision == "32":
        for k, v in fix["train3"]["state_dict"].items():
            assert rel_err(m.state_dict()[k], v, 3e-3) < 2e-3, k
    # the end-to-end calls take RAW spectra of the preprocessor's input size
    assert step.h_x.shape == (fix["batch"], 256)
    float(step.step_host(x.cpu(), y.cpu()))
    ev = EvalStep(m.eval(), fix["batch"], use_graph=True)
    with torch.no_grad():
        ref_logits = m(x).logits
    assert rel_err(ev.forward(x), ref_logits) < 1e-6
    # unfreezing makes the graph path refuse (the matrix is outside the fused arena), loudly
    m.set_preprocessor_trainable(True)
    with pytest.raises(NotImplementedError, match="FROZEN preprocessor"):
        step.step(x, y)


@pytest.mark.parametrize("name", ["pre_zca_r32", "pre_attn_r64"])
def test_trainable_preprocessor_with_torch_optimizer(golden, name, tmp_path):
    """Unfrozen preprocessor: stock loss.backward() + clip_grad_norm_ + torch.optim.AdamW over model.parameters()
    (Lightning's loop) reproduces the reference's 3 steps, preprocessor matrix included."""
    dev = _cuda()
    fix = golden(name)
    m = _build_pre(fix, "32", dev, tmp_path).eval()
    x, y = _inputs(fix, dev)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0)
    for i in range(3):
        opt.zero_grad()
        loss = m(x, labels=y).loss
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        opt.step()
        assert abs(float(loss) - float(fix["train3"]["losses"][i])) < 1e-3 * max(1.0, abs(float(loss))), i
        assert rel_err(norm, fix["train3"]["grad_norms"][i]) < 2e-3
    for k, v in fix["train3"]["state_dict"].items():
        assert rel_err(m.state_dict()[k], v, 3e-3) < 2e-3, k


@pytest.mark.parametrize("precision", ["32", "bf16-mixed"])
def test_trainable_preprocessor_inside_the_graph_step(golden, precision, tmp_path):
    """TrainStep with an UNFROZEN ZCA matrix (builder.py:168, the default freeze_epochs = 0): the captured step runs the
    preprocessor GEMM, its pixel / weight / bias gradients, one clip over ALL parameters and AdamW on the matrix next to the
    encoder's three launches.  3 steps vs the reference's: losses, total gradient norm, and in fp32 every post-step
    tensor incl. the preprocessor matrix; bit-identical to running the same kernels eagerly; the AdamW moments of the
    matrix travel through the .ckpt wire format."""
    from vit_b200.checkpoint import load_lightning_checkpoint, save_lightning_checkpoint
    from vit_b200.step import TrainStep

    dev = _cuda()
    fix = golden("pre_zca_r32")
    x, y = _inputs(fix, dev)
    tol = TOL[precision]
    runs = []
    for use_graph in (True, False):
        m = _build_pre(fix, precision, dev, tmp_path)
        assert not m.preprocessor.is_frozen
        step = TrainStep(m, fix["batch"], use_graph=use_graph, train=False)
        assert step._pre_tr is not None
        losses = []
        for i in range(3):
            losses.append(float(step.step(x, y)))
            if use_graph:
                ref = float(fix["train3"]["losses"][i])
                assert abs(losses[-1] - ref) < 5 * tol * max(1.0, abs(ref)), (i, losses[-1], ref)
                assert rel_err(step.eng.state[1], fix["train3"]["grad_norms"][i]) < 20 * tol
        runs.append((losses, {k: v.detach().clone() for k, v in m.state_dict().items()}))
        if use_graph:
            assert float(step.eng.state[5]) == 0.0                       # the extra squared norm was consumed
            if precision == "32":
                for k, v in fix["train3"]["state_dict"].items():
                    assert rel_err(m.state_dict()[k], v, 3e-3) < 2e-3, k
            moved = rel_err(m.state_dict()["preprocessor.linear.weight"], fix["state_dict"]["preprocessor.linear.weight"])
            assert moved > 1e-5                                           # the matrix really was trained
            # eval through the module's own forward sees the trained matrix (no stale bf16 operand copy)
            with torch.no_grad():
                a = m.eval()(x, labels=y).loss
            assert torch.isfinite(a)
            # the matrix's AdamW moments in the checkpoint, and back
            ck = save_lightning_checkpoint(m, train_step=step)
            names = [n for n, _ in m.named_parameters()]
            iw = names.index("preprocessor.linear.weight")
            assert rel_err(ck["optimizer_states"][0]["state"][iw]["exp_avg"], step._pre_tr.m_w) == 0.0
            m2 = _build_pre(fix, precision, dev, tmp_path)
            s2 = TrainStep(m2, fix["batch"], use_graph=True, train=False)
            load_lightning_checkpoint(m2, ck, train_step=s2)
            assert torch.equal(s2._pre_tr.m_w, step._pre_tr.m_w) and torch.equal(s2._pre_tr.v_w, step._pre_tr.v_w)
            l4a, l4b = float(step.step(x, y)), float(s2.step(x, y))
            assert abs(l4a - l4b) < 1e-6 * max(1.0, abs(l4a))
            s2.close()
            runs[-1] = (losses, runs[-1][1])
        step.close()
    assert runs[0][0] == runs[1][0]
    for k, v in runs[0][1].items():
        assert torch.equal(v, runs[1][1][k]), k


# ------------------------------------------------------------------------------------------------
# device-resident dataset hand-off
# ------------------------------------------------------------------------------------------------
def test_gather_batch_rows_labels_and_noise():
    from vit_b200.data import DeviceDataset

    dev = _cuda()
    g = torch.Generator().manual_seed(5)
    N, L, B = 300, 4096, 64
    flux = torch.rand(N, L, generator=g)
    lab = torch.rand(N, 2, generator=g)
    ds = DeviceDataset(flux, lab, error=torch.ones(N, L), device=dev)
    idx = torch.randint(0, N, (B,), generator=g).to(dev)
    x = torch.full((B, L), float("nan"), device=dev)
    y = torch.full((B * 2,), float("nan"), device=dev)
    ds.gather(idx, x, y)
    assert torch.equal(x.cpu(), flux[idx.cpu()]) and torch.equal(y.cpu().view(B, 2), lab[idx.cpu()])
    ds.gather(None, x, y)                                                   # no index vector: the first B rows
    assert torch.equal(x.cpu(), flux[:B])
    # ragged shapes: tiny batch, short rows, int64 class labels
    ds2 = DeviceDataset(torch.rand(10, 8, generator=g), torch.arange(10) % 3, device=dev)
    x2 = torch.empty(3, 8, device=dev)
    y2 = torch.empty(3, dtype=torch.int64, device=dev)
    i2 = torch.tensor([9, 0, 4], device=dev)
    ds2.gather(i2, x2, y2)
    assert torch.equal(x2, ds2.flux[i2]) and y2.tolist() == [0, 0, 1]
    # noise injection (src/vit.py:86-88): x = flux + N(0,1) * error * noise_level, keyed by the engine's {seed, step}
    rng = torch.tensor([99, 3], dtype=torch.int64, device=dev)
    xa, xb, xc = (torch.empty(B, L, device=dev) for _ in range(3))
    ds.gather(idx, xa, None, noise_level=0.5, rng=rng)
    ds.gather(idx, xb, None, noise_level=0.5, rng=rng)
    assert torch.equal(xa, xb)                                              # same (seed, step): same noise
    rng[1] += 1
    ds.gather(idx, xc, None, noise_level=0.5, rng=rng)
    assert not torch.equal(xa, xc)
    z = ((xa - ds.flux[idx]) / 0.5).double().flatten()
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z ** 4).mean()) - 3.0) < 0.1 and abs(float((z ** 3).mean())) < 0.05
    zc = ((xc - ds.flux[idx]) / 0.5).double().flatten()
    assert abs(float((z * zc).mean())) < 0.01                               # steps are uncorrelated
    with pytest.raises(ValueError):
        ds.gather(idx[:5], x, y)


@pytest.mark.parametrize("precision", ["32", "bf16-mixed"])
def test_fit_device_matches_stepping_by_hand(precision):
    """fit_device (gather kernel + step graph, losses kept on the device) == TrainStep.step on the same batches."""
    from vit_b200 import get_model
    from vit_b200.data import DeviceDataset, epoch_indices
    from vit_b200.step import TrainStep

    dev = _cuda()
    B, N = 8, 29
    x, y = vo.synthetic_batch(N, 512, seed=21, kind="rand")
    models = []
    for _ in range(2):
        torch.manual_seed(3)
        models.append(get_model(copy.deepcopy(BASE), precision=precision, device=dev))
    models[1].load_state_dict(models[0].state_dict())
    a, b = (TrainStep(m, B, use_graph=True, train=False) for m in models)
    ds = DeviceDataset(x, y, device=dev)
    got = a.fit_device(ds, epochs=2, shuffle=True, seed=4)
    assert len(got) == 2 and got[0].shape == (4,)            # 29 samples -> 4 batches of 8 (3 wrapped)
    want = []
    for ep in range(2):
        order = epoch_indices(N, ep, seed=4, batch=B)
        for i in range(order.numel() // B):
            sel = order[i * B:(i + 1) * B]
            want.append(float(b.step(x[sel].to(dev), y[sel].to(dev))))
    mine = torch.cat(got).cpu()
    assert rel_err(mine, torch.tensor(want)) < 1e-6
    for k, v in models[1].state_dict().items():
        assert rel_err(models[0].state_dict()[k], v, 1e-3) < 1e-5, k


@pytest.mark.parametrize("task", ["reg", "cls"])
def test_fit_device_rows_mode_equals_gather_mode(task, monkeypatch):
    """Device-resident dataset mode of the whole-network kernels (they pick the rows of the epoch permutation themselves,
    steps replayed 8 per graph) == the gather-kernel path, bit for bit: per-step losses and final weights; dropout ON,
    a remainder of single steps, max_steps, two epochs (fresh permutation uploaded between them)."""
    from vit_b200 import get_model
    from vit_b200.data import DeviceDataset
    from vit_b200.step import TrainStep

    dev = _cuda()
    cfg = copy.deepcopy(BASE)
    cfg["model"].update(image_size=4096, patch_size=32, stride_size=32)
    B, N = 8, 83
    x, y = vo.synthetic_batch(N, 4096, seed=5, kind="rand")
    if task == "cls":
        cfg["model"].update(task_type="cls", num_labels=4)
        y = torch.randint(0, 4, (N,), generator=torch.Generator().manual_seed(2))
    ds = DeviceDataset(x, y, device=dev)
    res = []
    for rows in ("1", "0"):
        monkeypatch.setenv("VITB200_ROWS", rows)
        torch.manual_seed(3)
        m = get_model(copy.deepcopy(cfg), precision="bf16-mixed", device=dev).train()
        st = TrainStep(m, B, use_graph=True, train=True)
        assert st.two_slots
        got = st.fit_device(ds, epochs=2, shuffle=True, seed=9)          # 83 -> 11 steps per epoch = 8 + 3
        got += st.fit_device(ds, epochs=1, seed=9, start_epoch=2, max_steps=6)
        assert [g.numel() for g in got] == [11, 11, 6]
        # a second dataset (other tensors, other length): the kernels are re-bound, the graphs re-captured
        ds2 = DeviceDataset(x[:40].flip(0).contiguous(), y[:40].flip(0).contiguous(), device=dev)
        got += st.fit_device(ds2, epochs=1, shuffle=True, seed=3)
        assert got[-1].numel() == 5
        assert (len(st._graph_rows) > 0) == (rows == "1")
        res.append((torch.cat(got).cpu(), {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}))
        st.close()
    assert torch.equal(res[0][0], res[1][0])
    for k, v in res[0][1].items():
        assert torch.equal(v, res[1][1][k]), k


@pytest.mark.parametrize("task", ["reg", "cls"])
def test_evaluate_metrics_on_device(task):
    """EvalStep.evaluate: one forward per batch, metrics accumulated by vitb200_eval_metrics_accum, exact tail batch."""
    from vit_b200 import get_model
    from vit_b200.data import DeviceDataset
    from vit_b200.step import EvalStep

    dev = _cuda()
    cfg = copy.deepcopy(BASE)
    N, B = 21, 8
    x, _ = vo.synthetic_batch(N, 512, seed=8, kind="rand")
    if task == "cls":
        cfg["model"].update(task_type="cls", num_labels=4)
        y = torch.randint(0, 4, (N,), generator=torch.Generator().manual_seed(2))
    else:
        cfg["data"]["param"] = "a,b"
        y = torch.rand(N, 2, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(9)
    m = get_model(cfg, precision="32", device=dev).eval()
    out = EvalStep(m, B, use_graph=True).evaluate(DeviceDataset(x, y, device=dev))
    with torch.no_grad():
        chunks = [(lo, min(N, lo + B)) for lo in range(0, N, B)]
        ref = [m(x[lo:hi].to(dev), labels=y[lo:hi].to(dev)) for lo, hi in chunks]
    logits = torch.cat([r.logits for r in ref]).cpu().double()
    loss = sum(float(r.loss) * (hi - lo) for r, (lo, hi) in zip(ref, chunks)) / N
    assert out["n"] == N and abs(out["loss"] - loss) < 1e-6 * max(1.0, abs(loss))
    assert rel_err(out["preds"], logits) < 1e-6
    if task == "cls":
        assert abs(out["acc"] - float((logits.argmax(-1) == y).double().mean())) < 1e-12
    else:
        e = logits - y.double()
        assert abs(out["mae"] - float(e.abs().mean())) < 1e-6
        assert abs(out["mse"] - float((e * e).mean())) < 1e-6
        yd = y.double()
        r2 = (1 - (e * e).sum(0) / ((yd - yd.mean(0)) ** 2).sum(0)).mean()
        assert abs(out["r2"] - float(r2)) < 1e-5 * max(1.0, abs(float(r2)))


def test_checkpoint_resume_on_device(tmp_path):
    """save_lightning_checkpoint / load_lightning_checkpoint around the fused optimizer: a resumed run continues exactly,
    and torch.optim.AdamW accepts the exported optimizer state (what the reference's trainer would load)."""
    from vit_b200 import get_model
    from vit_b200.checkpoint import load_lightning_checkpoint, save_lightning_checkpoint, strip_prefix
    from vit_b200.step import TrainStep

    dev = _cuda()
    B = 8
    x, y = vo.synthetic_batch(B, 512, seed=31, kind="rand")
    x, y = x.to(dev), y.to(dev)
    torch.manual_seed(5)
    m = get_model(copy.deepcopy(BASE), precision="32", device=dev)
    s = TrainStep(m, B, lr=2e-3, use_graph=True, train=False)
    for _ in range(2):
        s.step(x, y)
    ck = save_lightning_checkpoint(m, str(tmp_path / "a.ckpt"), train_step=s, epoch=1)
    assert ck["global_step"] == 2 and ck["optimizer_states"][0]["param_groups"][0]["lr"] == 2e-3
    nxt = [float(s.step(x, y)) for _ in range(2)]
    torch.manual_seed(6)
    m2 = get_model(copy.deepcopy(BASE), precision="32", device=dev)
    s2 = TrainStep(m2, B, use_graph=True, train=False)
    info = load_lightning_checkpoint(m2, str(tmp_path / "a.ckpt"), train_step=s2)
    assert info["step"] == 2 and info["epoch"] == 1
    res = [float(s2.step(x, y)) for _ in range(2)]
    assert rel_err(torch.tensor(res), torch.tensor(nxt)) < 1e-6
    for k, v in m.state_dict().items():
        assert rel_err(m2.state_dict()[k], v, 1e-3) < 1e-5, k
    # the exported state drives a stock torch optimizer on the oracle's parameters to the same third step
    spec = vo.spec_from_config(BASE)
    params = {k: v.clone().requires_grad_(True) for k, v in strip_prefix(ck["state_dict"]).items()}
    names = [n for n, _ in m.named_parameters()]
    opt = torch.optim.AdamW([params[n] for n in names], lr=1.0)
    opt.load_state_dict(ck["optimizer_states"][0])
    out = vo.forward(params, x.cpu(), spec, labels=y.cpu())
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_([params[n] for n in names], 0.5)
    opt.step()
    assert abs(float(out["loss"]) - nxt[0]) < 1e-4 * max(1.0, abs(nxt[0]))
