"""Time vitb200_linear_wgrad (fp32 SIMT) alone on the baseline shapes.  VITB200_SPLITS=<cap> python tools/wgrad_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vit_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
M = 8256
for N, K in ((128, 32), (32, 128), (96, 32), (32, 32)):
    dy = torch.randn(M, N, device=dev)
    x = torch.randn(M, K, device=dev)
    dw = torch.empty(N, K, device=dev)
    db = torch.empty(N, device=dev)
    ws = torch.zeros(int(lib.vitb200_linear_wgrad_ws_bytes(M, N, K)) + 4096, dtype=torch.uint8, device=dev)
    call = lambda: _lib.check(lib.vitb200_linear_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, N, K, 0,  # noqa: E731
                                                       _lib.F32, ws.data_ptr(), torch.cuda.current_stream().cuda_stream), "wgrad")
    for _ in range(3):
        call()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            call()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ref = dy.double().t() @ x.double()
    err = float((dw.double() - ref).abs().max() / ref.abs().max())
    print(f"N={N} K={K}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us per call, rel err {err:.1e}")
