"""Per-phase cycle counts of the fused kernels (debug build with clock64 stamps).

Reading the numbers: a stamp that follows a __syncthreads() is taken when thread 0 ARRIVES at the barrier, not when the
barrier releases (BAR.SYNC.DEFER_BLOCKING lets the warp run on until its next memory instruction, and the clock read is
not one), so the waiting time for slower warps shows up in the FOLLOWING phase.  Sums over phases are exact.

    VITB200_TIMELINE=1 python -m vit_b200.build          # builds vit_b200/libvitb200_tl.so (here, no GPU needed)
    VITB200_TIMELINE=1 python tools/timeline.py          # on the GPU box

Runs a few eager steps of the bench workload, then reads the stamps of the LAST launch of each instrumented
kernel and prints, per phase, the mean / max over CTAs of the cycle deltas between consecutive stamps.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
assert os.environ.get("VITB200_TIMELINE") == "1", "set VITB200_TIMELINE=1"

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import BASELINE_CFG  # noqa: E402
from vit_b200 import _lib, get_model  # noqa: E402
from vit_b200.step import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = get_model(json.loads(json.dumps(BASELINE_CFG)), precision="bf16-mixed", device=dev).train()
st = TrainStep(m, B, lr=1e-3, grad_clip=0.5, use_graph=True, train=True)
x = torch.rand(B, 4096, device=dev)
y = torch.rand(B, device=dev)
for _ in range(5):
    st.step(x, y)
torch.cuda.synchronize()
lib = _lib.load()
SLOTS, CTAS = 32, 512
names = {
    "layer_fwd": ["prologue", "pdl_wait", "tmem+sync", "ctx TMA+MMA1", "epi1 (drop,LN)+sync", "MMA2", "GELU+sync", "MMA3",
                  "epi3 (drop,LN)", "sync+MMA4", "epi4 (qkv)+sync"],
    "bwd_upper": ["prologue", "pdl_wait", "tmem+sync", "dz,drop->sD +sync", "tile TMA+MMA1", "gelu'+sync", "MMA2",
                  "LN bwd+sync", "MMA3", "dctx epi", "grad tail"],
    "tail": ["pdl_wait", "reduce partial slots", "block sum + ticket + wait", "norm, coef, bias corrections", "AdamW"],
    "mega_fwd": ["prologue+pdl_wait+weights", "embed+QKV(0)", "S MMA wait", "max pass", "exp pass", "PV (+2nd head)", "O epi+exchange",
                 "hop1", "hop2", "hop3", "layers 1..", "exit"],
    "mega_bwd": ["prologue", "pdl_wait", "head + layers above the stamped one + stage 1 (ddelta2)", "dm MMA + gelu'", "du2 MMA + LN2 bwd",
                 "dctx MMA + epi", "grad tail (MLP half)", "qkv wait + S/dP MMA + P,dS", "dQ/dK/dV MMA", "attn epi + exchange",
                 "lower (du MMA, LN1) + staging", "layers below", "embedding"],
    "attn_fwd": ["prologue", "pdl_wait", "tmem+sync", "K/V/Q TMA", "rope+bar+MMA S", "max pass+bar", "exp pass+bar", "MMA PV",
                 "O epi"],
}
for key, labels in names.items():
    buf = (ctypes.c_longlong * (SLOTS * CTAS))()
    if not hasattr(lib, f"vitb200_tl_{key}"):
        continue
    rc = getattr(lib, f"vitb200_tl_{key}")(buf)
    assert rc == 0
    a = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, SLOTS)
    n = len(labels) + 1
    live = a[:, 0] != 0
    if not live.any():
        continue          # kernel not launched by this program (e.g. the per-op kernels when the whole-network kernel runs)
    a = a[live][:, :n]
    d = np.diff(a, axis=1)
    print(f"== {key}: {a.shape[0]} CTAs, total mean {d.sum(1).mean():.0f} cycles (max {d.sum(1).max()}) "
          f"= {d.sum(1).mean() / 1.965e3:.2f} us @1.965 GHz")
    for i, lab in enumerate(labels):
        print(f"   {lab:<24} mean {d[:, i].mean():8.0f}   max {d[:, i].max():8d}")
    if key == "mega_fwd" and os.environ.get("VITB200_TL_EXTRA") == "1":   # main thread 0 vs side thread 0, relative to stamp 2
        full = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, SLOTS)[live]
        t0 = full[:, 2]
        for nm, a_, b_ in (("layer top", 2, 18), ("attention work done", 27, 28), ("after exchange barrier", 7, 23),
                           ("after hop 1", 8, 24), ("after hop 2", 9, 25), ("after hop 3", 10, 26)):
            print(f"      {nm:<24} main +{(full[:, a_] - t0).mean():8.0f}   side +{(full[:, b_] - t0).mean():8.0f}")
    if key == "mega_bwd" and os.environ.get("VITB200_TL_EXTRA") == "1":   # finer stamps (debug): slot -> offset from stamp 5 / 8 / 9
        full = np.frombuffer(buf, dtype=np.int64).reshape(CTAS, SLOTS)[live]
        for a_, b_, what in ((5, 16, "dctx: tid0 waits done"), (5, 17, "dctx: CTA barrier"), (5, 6, "dctx: cluster arrive"),
                             (8, 14, "attn: after MMA issue"), (8, 15, "attn: after mma wait"), (8, 9, "attn: TMA issued"),
                             (9, 18, "epi: cluster wait"), (9, 19, "epi: pieces written"), (9, 10, "epi: cluster sync")):
            dd = full[:, b_] - full[:, a_]
            print(f"      stamp {a_} -> {b_}  {what:<26} mean {dd.mean():8.0f}  even CTAs {dd[0::2].mean():8.0f}  odd CTAs {dd[1::2].mean():8.0f}"
                  f"  min {dd.min()} max {dd.max()}")
