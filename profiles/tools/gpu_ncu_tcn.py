"""One launch of upd_stg_tcn_ln per shape for ncu (kernel filter: stg_tcn)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200 import _lib
DEV = torch.device("cuda:0")
L = _lib.lib()
shapes = ((32768, 16, 16, 200), (32768, 32, 16, 200), (32768, 8, 8, 400))
for (N, CI, C, T) in shapes:
    x = torch.randn(N, CI, T, device=DEV)
    w1, b1 = torch.randn(C, CI, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    w2, b2 = torch.randn(C, C, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    g, be = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    a3 = torch.empty(N, 3 * C * T + 8, dtype=torch.float16, device=DEV)
    rc = L.upd_stg_tcn_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g), _lib.ptr(be),
                          N, CI, C, T, None, _lib.ptr(a3), None, None, _lib.stream_ptr(DEV))
    assert rc == 0, rc
    torch.cuda.synchronize()
print("ok")
