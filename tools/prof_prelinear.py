"""The preprocessor GEMM alone (M = 64 rows x a [4096, 4096] bf16 matrix): CUDA-event timing, and the target of
`ncu --set full` for its DRAM traffic.

    python tools/prof_prelinear.py                       # prints us / GB/s per call (graph of 20 calls, CUDA events)
    ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 4 --launch-count 2 \
        -o gpurun_out/prelinear python tools/prof_prelinear.py --eager 8
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vit_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=64)
ap.add_argument("--dim", type=int, default=4096)
ap.add_argument("--eager", type=int, default=0, help="only run this many eager calls (for ncu)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
M, N, K = a.rows, a.dim, a.dim
g = torch.Generator().manual_seed(0)
x32 = torch.rand(M, K, generator=g).to(dev)
x = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).bfloat16()
b = torch.zeros(N, device=dev)
y = torch.empty(M, N, device=dev)
ws = torch.zeros(int(lib.vitb200_tc_prelinear_ws_bytes(M, N, K)), dtype=torch.uint8, device=dev)
# something larger than the 126 MB L2 to evict the matrix between calls (otherwise the "stream" comes from L2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def call(st, with_cast=True):
    if with_cast:
        _lib.check(lib.vitb200_cast_bf16(x32.data_ptr(), x.data_ptr(), M * K, st), "cast")
    _lib.check(lib.vitb200_tc_prelinear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, ws.data_ptr(),
                                            st), "prelinear")


st = torch.cuda.current_stream().cuda_stream
if a.eager:
    for _ in range(a.eager):
        flush.zero_()
        call(st)
    torch.cuda.synchronize()
    print("eager calls done")
    sys.exit(0)
for _ in range(3):
    call(st)
torch.cuda.synchronize()
ref = x.float() @ w.float().t()
print("max rel err vs fp32 matmul:", float((y - ref).abs().max() / ref.abs().max()))
res = {}
for name, cast, fl in (("gemm_only_L2_warm", False, False), ("cast+gemm_L2_warm", True, False), ("gemm_only_L2_flushed", False, True)):
    gr = torch.cuda.CUDAGraph()
    reps = 20 if not fl else 1
    with torch.cuda.graph(gr):
        s2 = torch.cuda.current_stream().cuda_stream
        for _ in range(reps):
            call(s2, cast)
    tot = 0.0
    n = 10
    for _ in range(n):
        if fl:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / n / reps * 1e3
    res[name] = us
    print(f"{name}: {us:.2f} us/call  {2.0 * N * K / us / 1e3:.0f} GB/s of matrix  {2.0 * M * N * K / us / 1e6:.1f} TFLOP/s")
