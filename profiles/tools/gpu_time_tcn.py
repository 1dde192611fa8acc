"""Microbenchmark of upd_stg_tcn_ln (fused causal-TCN + LayerNorm front half of a ResidualBlock)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200 import _lib
DEV = torch.device("cuda:0")
L = _lib.lib()
for (N, CI, C, T) in ((32768, 16, 16, 200), (32768, 32, 16, 200), (32768, 8, 8, 400), (32768, 4, 8, 400), (32768, 12, 4, 400),
                      (10000, 16, 16, 50), (10000, 8, 8, 100), (10000, 16, 16, 1000), (10000, 8, 8, 2000)):
    x = torch.randn(N, CI, T, device=DEV)
    w1, b1 = torch.randn(C, CI, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    w2, b2 = torch.randn(C, C, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    g, be = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    a3 = torch.empty(N, 3 * C * T + 8, dtype=torch.float16, device=DEV)
    st = _lib.stream_ptr(DEV)
    def run():
        rc = L.upd_stg_tcn_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g), _lib.ptr(be),
                              N, CI, C, T, None, _lib.ptr(a3), None, None, st)
        assert rc == 0, rc
    for _ in range(3): run()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flops = 2.0 * N * T * 3 * C * (CI + C)
    byts = N * T * (4.0 * CI + 6.0 * C)
    print("tcn_ln N=%d CI=%d C=%d T=%d: %.3f ms  %.1f TFLOP/s fp32  %.0f GB/s" % (N, CI, C, T, ms, flops / ms / 1e9, byts / ms / 1e6))
