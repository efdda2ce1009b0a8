#!/usr/bin/env python
"""bench.py -- train samples/s of the ViT step (fwd + bwd + clip + AdamW) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`: "train samples/sec ... (baseline.yaml shape)"):
configs/exp/att_clp/baseline.yaml -- spectrum 4096, patch/stride 32 -> T=129, H=32, 2 heads, 3 layers,
per-GPU batch 64 (Lightning DDP semantics: weak scaling), precision bf16-mixed, dropout 0.1 (as configured),
grad clip 0.5, AdamW lr 1e-3.  Synthetic data, random-init weights.

Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BASELINE_CFG = {
    "model": dict(name="vit", task_type="reg", image_size=4096, patch_size=32, hidden_size=32,
                  num_hidden_layers=3, num_attention_heads=2, stride_size=32, proj_fn="SW"),
    "train": {"batch_size": 64, "precision": "bf16-mixed"},
    "loss": {"name": "mae"}, "opt": {"type": "AdamW", "lr": 0.001},
    "data": {"param": "log_g"}, "noise": {"noise_level": 0},
}
WORKLOAD = "configs/exp/att_clp/baseline.yaml ViT (L=4096,P=S=32,T=129,H=32,heads=2,layers=3), per-GPU batch 64, bf16-mixed, dropout 0.1, clip 0.5 + AdamW"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of the reference step) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, batch: int = 64):
    import torch
    from oracle import vit_oracle as vo

    try:
        ncpu = len(os.sched_getaffinity(0))   # the cores this process may actually use (cgroup / affinity aware)
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncpu))       # torchrun exports OMP_NUM_THREADS=1: override it for the CPU arm
    spec = vo.spec_from_config(BASELINE_CFG)
    tr = vo.OracleTrainer(spec, vo.init_params(spec, seed=42))
    x, y = vo.synthetic_batch(batch, 4096, seed=0, kind="dummy")
    torch.manual_seed(0)
    for _ in range(warmup):
        tr.step(x, y, train=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(x, y, train=True)
    dt = time.perf_counter() - t0
    return dict(value=batch * steps / dt, ms_per_step=1e3 * dt / steps, cores=torch.get_num_threads(),
                sample=f"{steps} steps (after {warmup} warm-up) of the same workload at batch {batch}: oracle port of the "
                       f"reference step (fp32, dropout 0.1, clip 0.5, AdamW), torch CPU ops on {torch.get_num_threads()} threads")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 100))
    r = cpu_reference_run(steps, max(3, min(args.warmup, 5)))
    line = {
        "impl": "reference", "metric": "train samples/sec (baseline.yaml shape)", "value": r["value"],
        "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(3, min(args.warmup, 5)),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD.replace("bf16-mixed", "fp32 (reference default)")},
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# per-kernel timing (CUDA events around a graph of R back-to-back launches of ONE call)
# ------------------------------------------------------------------------------------------------
def call_work(name: str, args: tuple, es: int):
    """(algorithmic flops, algorithmic bytes) of one C-ABI call; dims are read from its argument list
    (see include/vit_b200.h).  Bytes = compulsory reads + writes of its operands."""
    if name == "vitb200_linear_fwd":
        M, N, K, act = args[5], args[6], args[7], args[8]
        return 2.0 * M * N * K, (M * K + N * K + M * N * (2 if act else 1)) * es + 4 * N
    if name == "vitb200_linear_dgrad":
        M, N, K = args[4], args[5], args[6]
        return 2.0 * M * N * K, (M * N + N * K + M * K * (2 if args[2] else 1)) * es
    if name == "vitb200_linear_wgrad":
        M, N, K = args[4], args[5], args[6]
        return 2.0 * M * N * K, (M * N + M * K) * es + 4 * (N * K + N)
    if name == "vitb200_attn_fwd":
        B, T, h, d = args[8], args[9], args[10], args[11]
        return 4.0 * B * h * T * T * d, 4 * B * T * h * d * es + 4 * B * h * T
    if name == "vitb200_attn_bwd":
        B, T, h, d = args[14], args[15], args[16], args[17]
        return 10.0 * B * h * T * T * d, 8 * B * T * h * d * es + 8 * B * h * T
    if name == "vitb200_add_ln_fwd":
        M, H, cls_T = args[8], args[9], args[10]
        if args[1] is None:
            return 8.0 * M * H, M * H * (4 + es)
        rows_ln = M // cls_T if cls_T else M
        return 10.0 * M * H, M * H * (4 + es + 4) + rows_ln * H * es
    if name == "vitb200_add_ln_bwd":
        M, H, cls_T = args[10], args[11], args[12]
        rows_ln = M // cls_T if cls_T else M
        b = rows_ln * H * (es + 4) + M * H * 4 + (M * H * 4 if args[5] else 0) + (M * H * es if args[7] else 0)
        return 14.0 * rows_ln * H, b
    if name == "vitb200_patch_embed_fwd":
        B, L, P, S, Np, nv, H = args[6:13]
        return 2.0 * B * Np * P * H, 4 * B * L + 4 * B * (Np + 1) * H + es * H * P
    if name == "vitb200_patch_embed_bwd":
        B, L, P, S, Np, nv, H = args[6:13]
        return 2.0 * B * Np * P * H, 4 * B * L + 4 * B * (Np + 1) * H + 4 * H * P
    if name.startswith("vitb200_fused_") or name.startswith("vitb200_head_fused") or name.startswith("vitb200_mega_"):
        return fused_call_work(name, args)
    if name == "vitb200_clip_adamw_fused":
        n, slots, start, end = args[5], args[10], args[12], args[13]
        return 12.0 * n + slots * (end - start), 4.0 * (slots + 1) * (end - start) + 30.0 * n
    if name == "vitb200_grad_reduce":
        slots, start, end = args[1], args[3], args[4]
        return float(slots * (end - start)), 4.0 * (slots + 1) * (end - start)
    if name == "vitb200_head_loss_fwd":
        B, H, C = args[6], args[7], args[8]
        return 2.0 * B * H * C, B * H * es + 8 * B * C
    if name == "vitb200_head_loss_bwd":
        B, H, C = args[8], args[9], args[10]
        return 4.0 * B * H * C, 2 * B * H * es + 8 * B * C
    return 0.0, 0.0


def fused_call_work(name: str, args: tuple):
    """Algorithmic (flops, bytes) of the fused row-chain kernels; dims come from their C argument structs."""
    import ctypes
    from vit_b200 import _lib

    def st(tp):
        return ctypes.cast(args[0], ctypes.POINTER(tp)).contents

    if name in ("vitb200_mega_fwd", "vitb200_mega_bwd"):
        # whole-network kernels: algorithmic work of ONE launch = all samples.  Bytes = compulsory HBM traffic: the input
        # spectra, the parameters once, the tensors saved for backward written once (forward) / read once (backward), and
        # the per-sample gradient partials (backward).  With cls_only the last layer saves / reads only what its CLS row
        # and the keys / values of every token need.
        a = st(_lib.MegaBwdArgs).f if name.endswith("bwd") else st(_lib.MegaFwdArgs)
        B, T, H, L, P, Np, C = a.B, a.Np + 1, 32, a.layers, a.P, a.Np, a.C
        I = 4 * H
        row_full = 2 * I + 2 * I + 2 * H + 4 * H + 2 * H + 6 * H + 2 * H + 4 * H      # m a u2 hmid ctx qkv u z   (bytes / token / layer)
        row_top = 6 * H + 2 * H + 4 * H                                               # qkv u z                   (CLS-only last layer)
        n_full = L - 1 if a.cls_only else L
        saved = B * T * (n_full * row_full + (row_top if a.cls_only else 0)) + B * T * 4 * H * (0 if a.cls_only else 1)
        n_par = 3 * H + H * P + L * (12 * H * H + 13 * H) + 2 * H + C * H + C
        gemm = 2.0 * Np * P * H + (n_full * 24.0 * T * H * H) + (24.0 * H * H + 6.0 * T * H * H if a.cls_only else 0.0) + 2.0 * H * C
        attn = n_full * 4.0 * T * T * H + (4.0 * T * H if a.cls_only else 0.0)
        if name.endswith("fwd"):
            return B * (gemm + attn), float(4 * B * a.L + saved + 6 * n_par + 4 * B * C)
        return B * (2.0 * gemm + 2.5 * attn), float(4 * B * a.L + saved + 2 * n_par + 4 * B * n_par)
    if name == "vitb200_fused_layer_fwd":
        a = st(_lib.LayerFwdArgs)
        M, H = a.B * a.T, a.H
        I = 4 * H
        fl = 2.0 * M * (H * H + 2 * H * I + (0 if a.last else 3 * H * H))
        by = M * (2 * H + 4 * H + 4 * H + 2 * H + 2 * I + 2 * I + 4 * H) + (0 if a.last else M * (2 * H + 6 * H)) + 2 * (H * H + 2 * H * I + 3 * H * H)
        return fl, float(by)
    if name == "vitb200_fused_embed_fwd":
        a = st(_lib.EmbedFwdArgs)
        M, H = a.B * (a.Np + 1), a.H
        return 2.0 * M * (a.P * H + 3 * H * H), float(4 * a.B * a.L + M * (4 * H + 2 * H + 6 * H))
    if name == "vitb200_fused_layer_bwd_upper":
        a = st(_lib.LayerBwdUpperArgs)
        M, H = a.B * a.T, a.H
        I = 4 * H
        fl = 2.0 * M * 2 * (2 * H * I + H * H)
        by = M * ((0 if a.dz_cls else 4 * H) + 2 * I + 2 * I + 2 * H + 2 * H + 4 * H + 4 * H + 2 * H)
        return fl, float(by)
    if name == "vitb200_fused_layer_bwd_lower":
        a = st(_lib.LayerBwdLowerArgs)
        M, H = a.B * a.T, a.H
        return 2.0 * M * 2 * 3 * H * H, float(M * (6 * H + 2 * H + 4 * H + 4 * H + 4 * H))
    if name == "vitb200_fused_embed_bwd":
        a = st(_lib.EmbedBwdArgs)
        M, H = a.B * (a.Np + 1), a.H
        return 2.0 * M * a.P * H, float(4 * a.B * a.L + 4 * M * H)
    if name == "vitb200_head_fused_fwd":
        B, H, C = args[6], args[7], args[8]
        return 2.0 * B * H * C, float(B * H * 2 + 8 * B * C)
    if name == "vitb200_head_fused_fwd_bwd":
        B, H, C = args[16], args[17], args[18]
        return 8.0 * B * H * C + 10.0 * B * H, float(B * H * (2 + 4 + 4) + 8 * B * C)
    if name == "vitb200_head_fused_bwd":
        B, H, C = args[15], args[16], args[17]
        return 6.0 * B * H * C + 10.0 * B * H, float(B * H * (2 + 4 + 4) + 8 * B * C)
    return 0.0, 0.0


def tc_gemm_probe(dev, peaks):
    """Tensor-pipe evidence at a scaled shape (hidden 768, the config's documented range): the tcgen05 Linear forward
    alone, M = 128 x 129 rows, timed with CUDA events."""
    import torch
    from vit_b200 import _lib

    lib = _lib.load()
    out = {}
    for (M, N, K) in ((16512, 3072, 768), (16512, 768, 3072)):
        x = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        b = torch.zeros(N, device=dev)
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        st = torch.cuda.current_stream(dev)
        for _ in range(3):
            _lib.check(lib.vitb200_tc_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, M, N, K, 0, st.cuda_stream), "probe")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            lib.vitb200_tc_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, M, N, K, 0, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / 10
        tf = 2.0 * M * N * K / sec / 1e12
        out[f"linear_fwd_{M}x{N}x{K}"] = {"tflops": tf, "frac_of_bf16_burst_peak": tf / peaks["tf_burst"], "us": sec * 1e6}
    return out


def time_calls(eng, progs, repeats: int = 20, iters: int = 5):
    """Average device time of every call of the step, each measured alone as a CUDA graph of `repeats`
    back-to-back launches (CUDA events on the launching stream)."""
    import torch
    from vit_b200 import _lib

    out = []
    stream = torch.cuda.current_stream(eng.device)
    for prog in progs:
        for fn, args in prog:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st = torch.cuda.current_stream(eng.device).cuda_stream
                for _ in range(repeats):
                    rc = fn(*args, st)
                    if rc != 0:
                        _lib.check(rc, fn.__name__)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                g.replay()
            e1.record(stream)
            torch.cuda.synchronize()
            out.append((fn.__name__, args, e0.elapsed_time(e1) * 1e-3 / (iters * repeats)))
            del g
    return out


def kernel_table(eng, train: bool):
    import torch

    es = 2 if eng.act_dtype == torch.bfloat16 else 4
    eng.peer = None   # (data-parallel runs: the per-kernel table is a rank-0-only measurement, so no peer exchange here)
    fh = eng.can_fuse_head
    eng.forward(train=train, with_labels=True, head_bwd=fh)
    eng.backward(train=train, skip_reduce=True, skip_head=fh)
    eng.optimizer_step(fused_reduce=True)
    torch.cuda.synchronize()
    co = bool(eng.cls_only and eng.mega)
    bkey = ("bwd", train, None, True, fh, co) if eng.fused_bwd else ("bwd", train, None, co)
    fkey = ("fwd", train, True, True, co) if fh else ("fwd", train, True, co)
    ar = eng.arena
    slots, start, end = eng._red if eng.fused_bwd else (0, 0, 0)
    tail = (eng.lib.vitb200_clip_adamw_fused, (
        ar.data.data_ptr(), ar.grad.data_ptr(), eng.exp_avg.data_ptr(), eng.exp_avg_sq.data_ptr(),
        None if ar.shadow is None else ar.shadow.data_ptr(), ar.layout.n_opt, eng.hyper.data_ptr(), eng.state.data_ptr(),
        eng.rng.data_ptr(), eng.gpart.data_ptr() if slots else None, slots, ar.layout.n_opt, start, end,
        eng.tail_ws.data_ptr()))
    progs = [eng._progs[fkey], eng._progs[bkey], [tail]]
    rows = []
    for name, args, sec in time_calls(eng, progs):
        fl, by = call_work(name, args, es)
        rows.append(dict(call=name.replace("vitb200_", ""), us=sec * 1e6, flops=fl, bytes=by))
    return rows


def dp_parity_check(dev, rank, world, precision, B=16, steps=2):
    """Data-parallel correctness, proven inside the run that prints the scaling numbers: the same TrainStep the timed
    loop uses (in-kernel peer exchange, CUDA graph), dropout off, `steps` steps with rank r holding shard r of a global
    batch; then (1) every replica's parameter arena must be BIT-identical, (2) on rank 0 the oracle takes the same steps
    on the UNION batch and the mean of the rank losses / the AdamW first moments (= the clipped mean gradient) must agree
    within the precision's tolerance."""
    import torch
    import torch.distributed as dist
    from oracle import vit_oracle as vo
    from vit_b200 import dp, get_model
    from vit_b200.step import TrainStep

    cfg = json.loads(json.dumps(BASELINE_CFG))
    torch.manual_seed(7)
    m = get_model(cfg, precision=precision, device=dev).eval()
    dp.broadcast_parameters(m._arena.data)
    p0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    st = TrainStep(m, B, use_graph=True, world_size=world, train=False)
    x, y = vo.synthetic_batch(B * world, 4096, seed=5, kind="rand")
    lo, hi = rank * B, (rank + 1) * B
    xs, ys = x[lo:hi].to(dev), y[lo:hi].to(dev)
    losses = [st.step(xs, ys).clone() for _ in range(steps)]
    torch.cuda.synchronize()
    flat = m._arena.data[:m._arena.layout.n_opt]
    # (1) bitwise replica equality: a 64-bit checksum of the raw bits + a max over |p - p_rank0|
    bits = flat.view(torch.int32).to(torch.int64)
    chk = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
    allchk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allchk, chk)
    identical = all(bool(torch.equal(c, allchk[0])) for c in allchk)
    lt = torch.stack(losses).reshape(1, steps)
    alll = [torch.zeros_like(lt) for _ in range(world)]
    dist.all_gather(alll, lt)
    out = {"world": world, "shard_batch": B, "steps": steps, "replicas_bit_identical": identical}
    if rank == 0:
        tol = 2e-2 if "bf16" in precision else 1e-4
        spec = vo.spec_from_config(cfg)
        ref = vo.OracleTrainer(spec, p0, autocast_bf16="bf16" in precision)
        ref_losses = [ref.step(x, y, train=False) for _ in range(steps)]
        mine = torch.cat(alll, 0).mean(0).cpu()
        lerr = max(abs(float(mine[i]) - ref_losses[i]) / max(1.0, abs(ref_losses[i])) for i in range(steps))
        eng, lay = st.eng, m._arena.layout
        gmax = max(float(v.abs().max()) for k, v in ref.m.items() if "pooler" not in k)
        merr = 0.0
        for k, v in ref.m.items():
            e = lay.entries.get(k)
            if e is None or e.offset >= lay.n_opt:
                continue
            got = eng.exp_avg[e.offset:e.offset + e.numel].reshape(e.shape).cpu()
            merr = max(merr, float((got - v).abs().max()) / max(float(v.abs().max()), 1e-2 * gmax))
        out.update(oracle="CPU oracle, union batch of %d samples, %s" % (B * world, "autocast bf16" if "bf16" in precision else "fp32"),
                   loss_rel_err=lerr, exp_avg_rel_err=merr, tol_loss=tol, tol_exp_avg=2 * tol if "bf16" in precision else 2e-3,
                   peer_exchange=getattr(eng, "peer", None) is not None)
        out["ok"] = bool(identical and lerr < out["tol_loss"] and merr < out["tol_exp_avg"])
    dist.barrier()
    st.close()
    del st, m
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (train.batch_size)")
    ap.add_argument("--precision", default="bf16-mixed")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--kernels-json", default=None, help="write the per-kernel timing table here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from oracle import vit_oracle as vo  # only for the synthetic-input generator, the FLOP model and cpu_baseline
    from vit_b200 import dp, get_model
    from vit_b200.step import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local, world = dp.init_from_env("nccl") if world > 1 else (0, 0, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    B = args.batch

    cpu = None   # (measured AFTER the GPU regions: its 16 busy host threads must not sit next to the end-to-end loop)

    dp_parity = dp_parity_check(dev, rank, world, args.precision) if world > 1 else None

    torch.manual_seed(42)
    cfg = json.loads(json.dumps(BASELINE_CFG))
    model = get_model(cfg, precision=args.precision, device=dev)
    model.train()
    if world > 1:
        dp.broadcast_parameters(model._arena.data)
    step = TrainStep(model, B, lr=1e-3, grad_clip=0.5, use_graph=not args.no_graph, world_size=world, train=True)
    spec = vo.spec_from_config(BASELINE_CFG)
    fwd_flops, step_flops = vo.flops_per_sample(spec)

    # input pool larger than L2 (126 MB): every step's inputs come from HBM, not from a warm L2
    pool_n = max(8, int(160e6 // (B * 4096 * 4)) + 1)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    pool_x = torch.rand(pool_n, B, 4096, generator=g).to(dev)
    pool_y = torch.rand(pool_n, B, generator=g).to(dev)
    host_x = torch.rand(8, B, 4096, generator=g).pin_memory()
    host_y = torch.rand(8, B, generator=g).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    # The pool is a device-resident dataset (vit_b200.data.DeviceDataset): TrainStep.fit_device walks a
    # shuffled permutation of its rows, batch after batch.  With the whole-network kernels the step graph reads the rows
    # itself (no gather launch, no staging copy); otherwise a gather kernel fills the engine's input buffers first.
    from vit_b200.data import DeviceDataset

    pool_ds = DeviceDataset(pool_x.view(-1, 4096), pool_y.view(-1), device=dev)
    epoch_no = [0]

    def run_steps(n):
        """Exactly n training steps over the device-resident pool (a new shuffled epoch whenever one is used up)."""
        done = 0
        while done < n:
            got = step.fit_device(pool_ds, epochs=1, shuffle=True, seed=1234 + rank, start_epoch=epoch_no[0],
                                  max_steps=n - done)
            epoch_no[0] += 1
            done += int(got[0].numel())

    run_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # The ranks leave the host barrier up to ~1 ms apart.  In a data-parallel run every step ends at the slowest rank
    # (gradient exchange), so timing from the barrier would charge that one-off host skew to the K timed steps.  A few
    # untimed steps first let the in-kernel exchange align the GPUs (and fill the launch queue); the events then bracket
    # exactly K steps of steady state on the launching stream.  Same code path at N = 1.
    lead = 8
    run_steps(lead)
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    loss_last = float(step.eng.loss[0])
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    value = world * B * 1e3 / ms_step
    # the same through TrainStep.step(x, y) on device tensors (one staging copy of the inputs per step)
    for i in range(lead):
        step.step(pool_x[i % pool_n], pool_y[i % pool_n])
    e0.record(stream)
    for i in range(lead, lead + args.steps):
        step.step(pool_x[i % pool_n], pool_y[i % pool_n])
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_api_value = world * B * 1e3 / (float(t[0]) / args.steps)

    # ---- end to end through the public API: host batches -> H2D -> step -> D2H loss, every step ----
    # TrainStep.fit_host is the training loop a user calls with an iterable of host batches; it pipelines the copies
    # around the steps (next batch uploads on a copy stream, the loss is read one step late).  step_host is the
    # blocking single-step call (H2D, step, D2H, sync); it is reported beside it.
    def host_batches(n):
        for i in range(n):
            yield host_x[i % 8], host_y[i % 8]

    import gc

    step.fit_host(host_batches(max(8, args.warmup)))
    gc.collect()
    gc.disable()      # a collection inside a 3 ms wall-clock window would be 10 % of it
    barrier()
    t0 = time.perf_counter()
    e2e_losses = step.fit_host(host_batches(args.steps))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    assert len(e2e_losses) == args.steps
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te[0])
    nb = max(10, args.steps // 4)
    for i in range(3):
        step.step_host(host_x[i % 8], host_y[i % 8])
    barrier()
    t0 = time.perf_counter()
    for i in range(nb):
        step.step_host(host_x[i % 8], host_y[i % 8])
    torch.cuda.synchronize()
    tb = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    e2e_blocking = world * B * nb / float(tb[0])
    # the timed regions last ~0.1 s, nvidia-smi samples every 100 ms: keep the SAME step running (untimed) for about one
    # more second so that the clock / throttle-reason samples describe this load.  The count derives from the all-reduced
    # step time, so every rank runs the same number of steps (the in-kernel gradient exchange needs that).
    extra = max(0, min(8000, int(1.0 / max(ms_step * 1e-3, 1e-6))))
    run_steps(extra)
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    if clocks is not None:
        clocks["window"] = f"timed regions + {extra} more untimed steps of the same workload"
    # ---- BASELINE config 5 at N GPUs: scripts/test.py semantics (eval mode, forward only, large batch), N independent
    # replicas over a batch-sharded set, no collective on the data path; device time, max over ranks ----
    eval_replicas = None
    try:
        from vit_b200.step import EvalStep
        eb = 1024
        es_ = EvalStep(model.eval(), eb, use_graph=not args.no_graph)
        ex = torch.rand(4, eb, 4096, generator=g).to(dev)
        for i in range(3):
            es_.forward(ex[i % 4])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_ev = 20
        a0.record(stream)
        for i in range(n_ev):
            es_.forward(ex[i % 4])
        a1.record(stream)
        barrier()
        tev = torch.tensor([a0.elapsed_time(a1) / n_ev], device=dev)
        if world > 1:
            dist.all_reduce(tev, op=dist.ReduceOp.MAX)
        eval_replicas = {"samples_per_s": world * eb * 1e3 / float(tev[0]), "ms_per_batch": float(tev[0]),
                         "per_gpu_batch": eb, "replicas": world, "collective": "none (replicas only)"}
        model.train()
        del es_, ex
    except Exception as ex_:  # noqa: BLE001
        eval_replicas = {"error": str(ex_)[:200]}
        model.train()
    launches_all = step.kernel_launches()
    if world > 1:
        dist.barrier()
        peer = getattr(step.eng, "peer", None)
        if peer is not None:      # collective: every rank unmaps together, then rank 0 measures its kernels alone
            step.graph = None     # (the captured graph launches the peer-exchange kernel)
            peer.close()
            step.eng.peer = None

    if rank != 0:
        if world > 1:
            dist.barrier()
            step.close()   # a live graph with captured NCCL collectives would block destroy_process_group()
            dist.destroy_process_group()
        return

    # ---- per-kernel table + roofline of the dominant kernel (rank 0, after the timed regions) ----
    launches_per_step = launches_all
    h2d_b, d2h_b = step.h2d_bytes_per_step, step.d2h_bytes_per_step
    rows = kernel_table(step.eng, train=True)
    ksum_us = sum(r["us"] for r in rows)
    # dominant kernel = the call with the largest share of the step (summed over its launches); its roofline numbers are
    # per launch (mean over its launches)
    tot = {}
    for r in rows:
        t = tot.setdefault(r["call"], dict(us=0.0, n=0, flops=0.0, bytes=0.0))
        t["us"] += r["us"]; t["n"] += 1; t["flops"] += r["flops"]; t["bytes"] += r["bytes"]
    tname = max(tot, key=lambda k: tot[k]["us"])
    top = dict(call=tname, us=tot[tname]["us"] / tot[tname]["n"], flops=tot[tname]["flops"] / tot[tname]["n"],
               bytes=tot[tname]["bytes"] / tot[tname]["n"], launches=tot[tname]["n"], us_total=tot[tname]["us"])
    ridge = peaks["tf_sustained"] * 1e12 / (peaks["hbm"] * 1e9)
    ai = top["flops"] / max(top["bytes"], 1.0)
    if ai >= ridge:
        bound, achieved, peak, unit = "tensor", top["flops"] / (top["us"] * 1e-6) / 1e12, peaks["tf_sustained"], "TFLOP/s"
    else:
        bound, achieved, peak, unit = "hbm", top["bytes"] / (top["us"] * 1e-6) / 1e9, peaks["hbm"], "GB/s"
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if B == 64 and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {}).get(top["call"])
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes": top["bytes"], "algorithmic_flops": top["flops"],
                "kernel": top["call"], "kernel_us": top["us"], "peak_source": peaks["source"],
                "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
                "launches_per_step": top["launches"], "share_of_step": top["us_total"] / max(ksum_us, 1e-9),
                "note": "algorithmic bytes (or flops) of the call / its mean device time, timed alone as a CUDA graph "
                        "of 20 back-to-back launches with CUDA events"}
    agg = {}
    for r in rows:
        a = agg.setdefault(r["call"], dict(us=0.0, n=0, flops=0.0, bytes=0.0))
        a["us"] += r["us"]; a["n"] += 1; a["flops"] += r["flops"]; a["bytes"] += r["bytes"]
    kernels = sorted(({"call": k, "launches": v["n"], "us_total": round(v["us"], 2),
                       "gbs": round(v["bytes"] / (v["us"] * 1e-6) / 1e9, 1),
                       "tflops": round(v["flops"] / (v["us"] * 1e-6) / 1e12, 3)} for k, v in agg.items()),
                     key=lambda d: -d["us_total"])
    if args.kernels_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.kernels_json)), exist_ok=True)
        json.dump({"rows": rows, "aggregate": kernels, "kernel_time_sum_us": ksum_us, "step_us": ms_step * 1e3},
                  open(args.kernels_json, "w"), indent=1)

    probe = None
    if not args.no_sweep and world == 1:
        try:
            probe = tc_gemm_probe(dev, peaks)
        except Exception as ex:  # noqa: BLE001
            probe = {"error": str(ex)[:200]}
    sweep = None
    if not args.no_sweep and world == 1:
        sweep = {}
        for b2 in (1024, 8192):
            try:
                m2 = get_model(json.loads(json.dumps(BASELINE_CFG)), precision=args.precision, device=dev).train()
                s2 = TrainStep(m2, b2, use_graph=not args.no_graph, train=True)
                x2 = torch.rand(b2, 4096, device=dev); y2 = torch.rand(b2, device=dev)
                for _ in range(3):
                    s2.step(x2, y2)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n2 = 20 if b2 == 1024 else 5
                a0.record(stream)
                for _ in range(n2):
                    s2.step(x2, y2)
                a1.record(stream)
                torch.cuda.synchronize()
                msb = a0.elapsed_time(a1) / n2
                sweep[f"batch_{b2}"] = {"samples_per_s": b2 * 1e3 / msb, "ms_per_step": msb,
                                        "tflops_algorithmic": b2 * step_flops / (msb * 1e-3) / 1e12}
                del s2, m2, x2, y2
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                sweep[f"batch_{b2}"] = {"error": str(ex)[:200]}

    # ---- the other BASELINE.json configs, measured beside the headline (N = 1 only; not the bench line's value) ----
    other = None
    if not args.no_sweep and world == 1:
        from vit_b200.step import EvalStep
        other = {}

        def timed(fn, n):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(n):
                fn()
            a1.record(stream)
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / n

        def variant(**kw):
            c = json.loads(json.dumps(BASELINE_CFG))
            c["model"].update(kw)
            return c

        cases = [
            ("fp32_train_b64", variant(), 64, "train32", 50),
            ("config_yaml_2layers_train_b64", variant(num_hidden_layers=2), 64, "train", 50),
            ("eval_forward_b1024", variant(), 1024, "eval", 20),
            ("eval_forward_b8192", variant(), 8192, "eval", 5),
            ("long_seq_x4_stride8_T510_train_b64", variant(stride_size=8), 64, "train", 5),
            ("long_seq_x16_stride2_T2034_train_b64", variant(stride_size=2), 64, "train", 2),
        ]
        for name, cfg2, b2, mode, n2 in cases:
            try:
                m2 = get_model(cfg2, precision="32" if mode == "train32" else args.precision, device=dev)
                x2 = torch.rand(b2, 4096, device=dev); y2 = torch.rand(b2, device=dev)
                if mode in ("train", "train32"):
                    s2 = TrainStep(m2.train(), b2, use_graph=not args.no_graph, train=True)
                    msb = timed(lambda: s2.step(x2, y2), n2)
                else:
                    s2 = EvalStep(m2.eval(), b2, use_graph=not args.no_graph)
                    msb = timed(lambda: s2.forward(x2), n2)
                eng2 = s2.eng
                other[name] = {"samples_per_s": b2 * 1e3 / msb, "ms": msb, "tokens": int(m2.config.tokens),
                               "fused_fwd": bool(eng2.fused), "fused_bwd": bool(eng2.fused_bwd)}
                del s2, m2, x2, y2, eng2
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                other[name] = {"error": str(ex)[:200]}

    # ---- rows either side of the step: device-resident hand-off, frozen ZCA stage, on-device evaluation ----
    if other is not None:
        from vit_b200.data import DeviceDataset
        from vit_b200.step import EvalStep

        def ev_timed(fn):
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            r = fn()
            a1.record(stream)
            torch.cuda.synchronize()
            return a0.elapsed_time(a1), r

        try:   # one epoch over a device-resident dataset: gather kernel + step graph per batch, no host traffic
            nds = 12800
            gd = torch.Generator(device="cpu").manual_seed(7)
            ds = DeviceDataset(torch.rand(nds, 4096, generator=gd), torch.rand(nds, generator=gd), device=dev)
            m3 = get_model(json.loads(json.dumps(BASELINE_CFG)), precision=args.precision, device=dev).train()
            s3 = TrainStep(m3, B, use_graph=not args.no_graph, train=True)
            s3.fit_device(ds, epochs=1, seed=0)
            ms3, _ = ev_timed(lambda: s3.fit_device(ds, epochs=1, seed=1, start_epoch=1))
            other["fit_device_resident_dataset_b64"] = {"samples_per_s": nds * 1e3 / ms3, "ms_per_step": ms3 / (nds // B),
                                                        "dataset_mb": ds.nbytes / 1e6, "steps": nds // B}
            e3 = EvalStep(m3.eval(), 1024, use_graph=not args.no_graph)
            e3.evaluate(ds, return_preds=True)
            ms4, r4 = ev_timed(lambda: e3.evaluate(ds, return_preds=True))
            other["evaluate_on_device_metrics_b1024"] = {"samples_per_s": nds * 1e3 / ms4, "ms": ms4, "n": r4["n"],
                                                         "mae": r4.get("mae"), "r2": r4.get("r2")}
            del ds, m3, s3, e3
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            other["fit_device_resident_dataset_b64"] = {"error": str(ex)[:200]}
        try:   # frozen 4096 x 4096 ZCA-shaped preprocessor in front of the same model (2.1 GFLOP / step at B = 64)
            from vit_b200.builder import get_vit_config
            from vit_b200.model import MyViT
            from vit_b200.preprocessor import LinearPreprocessor

            gd = torch.Generator(device="cpu").manual_seed(8)
            Pm = torch.randn(4096, 4096, generator=gd) / 64.0
            cz = json.loads(json.dumps(BASELINE_CFG))
            m5 = MyViT(get_vit_config(cz), loss_name="mae", model_name="ZCA_fzperm_ViT",
                       preprocessor=LinearPreprocessor(Pm, bias=torch.zeros(4096), freeze=True), full_config=cz,
                       precision=args.precision, device=dev).train()
            s5 = TrainStep(m5, B, use_graph=not args.no_graph, train=True)
            x5 = torch.rand(B, 4096, device=dev); y5 = torch.rand(B, device=dev)
            ms5 = timed(lambda: s5.step(x5, y5), 50)
            raw = m5._raw_buffer(s5.eng)
            # device time of the preprocessor alone: 20 back-to-back calls in one CUDA graph (no host launch overhead)
            g5 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g5):
                for _ in range(20):
                    m5.preprocessor.forward_into(raw, s5.eng.x)
            ms6 = timed(g5.replay, 10) / 20
            wbytes = 4096 * 4096 * (2 if "bf16" in args.precision else 4)
            other["zca4096_frozen_train_b64"] = {"samples_per_s": B * 1e3 / ms5, "ms": ms5, "preprocessor_ms": ms6,
                                                 "preprocessor_gbs": wbytes / (ms6 * 1e-3) / 1e9,
                                                 "preprocessor_tflops": 2.0 * B * 4096 * 4096 / (ms6 * 1e-3) / 1e12}
            del m5, s5, Pm
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            other["zca4096_frozen_train_b64"] = {"error": str(ex)[:200]}
        try:   # the same stage for a frozen LOW-RANK ZCA (warmup.r = 32): factored kernel, 2 r D values instead of D^2
            from vit_b200.preprocessor import compute_zca_matrix, zca_lowrank_factors

            gd = torch.Generator(device="cpu").manual_seed(9)
            ev = torch.linalg.qr(torch.randn(4096, 4096, generator=gd).to(dev))[0]
            lam = torch.sort(torch.rand(4096, generator=gd) + 0.01, descending=True)[0].to(dev)
            Pz = compute_zca_matrix(ev, lam, eps=1e-5, r=32, shrinkage=0.0)
            pre6 = LinearPreprocessor(Pz, bias=torch.zeros(4096), freeze=True)
            pre6.set_lowrank_factors(*zca_lowrank_factors(ev, lam, 1e-5, 32, 0.0))
            cz = json.loads(json.dumps(BASELINE_CFG))
            m6 = MyViT(get_vit_config(cz), loss_name="mae", model_name="ZCA32_fzperm_ViT", preprocessor=pre6, full_config=cz,
                       precision=args.precision, device=dev).train()
            s6 = TrainStep(m6, B, use_graph=not args.no_graph, train=True)
            x6 = torch.rand(B, 4096, device=dev); y6 = torch.rand(B, device=dev)
            ms7 = timed(lambda: s6.step(x6, y6), 50)
            from vit_b200 import preprocessor as _vp
            active = _vp._lowrank_state(m6.preprocessor.linear, m6.preprocessor.linear.weight) is not None
            raw6 = m6._raw_buffer(s6.eng)
            g6 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g6):
                for _ in range(20):
                    m6.preprocessor.forward_into(raw6, s6.eng.x)
            ms8 = timed(g6.replay, 10) / 20
            es6 = 2 if "bf16" in args.precision else 4
            other["zca4096_r32_lowrank_frozen_train_b64"] = {
                "samples_per_s": B * 1e3 / ms7, "ms": ms7, "preprocessor_ms": ms8, "factored_kernel": bool(active),
                "preprocessor_gbs": (2 * B * 4096 * 4 + 4096 * 32 * es6) / (ms8 * 1e-3) / 1e9}
            del m6, s6, Pz, ev, pre6
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            other["zca4096_r32_lowrank_frozen_train_b64"] = {"error": str(ex)[:200]}

    if world == 1 and not args.no_cpu_baseline:   # reported at N=1 only
        r = cpu_reference_run(steps=60, warmup=3)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {
        "metric": "train samples/sec (baseline.yaml shape)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if "bf16" in args.precision else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"dp{world}",
                   "cuda_graph": not args.no_graph,
                   "grad_allreduce": ("none (1 GPU)" if world == 1 else
                                      "in-kernel over NVLink peer memory (vitb200_clip_adamw_fused_dp), no NCCL call per step"),
                   "l2": f"inputs = shuffled rows of a device-resident dataset of {pool_n * B} spectra "
                         f"({pool_n * B * 4096 * 4 / 1e6:.0f} MB > 126 MB L2), TrainStep.fit_device"},
        "step_api_samples_per_s": step_api_value,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_b,
                "d2h_bytes_per_step": d2h_b, "ms_per_step": 1e3 * float(te[0]) / args.steps,
                "api": "TrainStep.fit_host(iterable of host batches): per step H2D of the inputs from pinned memory into one of "
                       "the engine's eight input slots (copy stream, one group of four steps ahead), steps run four per CUDA graph, the "
                       "forward kernel stores each step's loss into pinned host memory (read one group late)",
                "blocking_step_host_samples_per_s": e2e_blocking},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "step_tflops_algorithmic": world * B * step_flops / (ms_step * 1e-3) / 1e12,
        "frac_of_bf16_sustained_peak": world * B * step_flops / (ms_step * 1e-3) / 1e12 / (world * peaks["tf_sustained"]),
        "kernel_time_sum_us": ksum_us,
        "kernels": kernels[:8],
        "sweep": sweep,
        "other_configs": other,
        "tc_gemm_probe": probe,
        "final_loss": loss_last,
    }
    line["eval_forward_replicas"] = eval_replicas
    if dp_parity is not None:
        line["dp_parity"] = dp_parity
        if not dp_parity.get("ok", False):
            print("bench.py: DATA-PARALLEL PARITY CHECK FAILED: " + json.dumps(dp_parity), file=sys.stderr, flush=True)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        step.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
