"""One launch each of the graph-sampler kernels at representative shapes, for `ncu --set full`
(profiles/r01_graph_kernels_ncu_full_metrics.json):
    ncu --set full --clock-control none --import-source on -k 'regex:stg_tcn_ln|stg_gated_aggregate|nsx_step' \
        -o gpurun_out/prof_graph python profiles/tools/gpu_ncu_graph_kernels.py
    python profiles/make_family_summaries.py r01 gpurun_out/prof_graph.ncu-rep graph_kernels
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200 import _lib, schedules
DEV = torch.device("cuda:0")
L = _lib.lib()
st = _lib.stream_ptr(DEV)
torch.manual_seed(0)
# fused causal-TCN + LayerNorm: DiffSTG resolutions 1 / 2, NsDiff_spatial's T = 50 (scalar path), a 2000-long row (4 segments)
for (N, CI, C, T) in ((32768, 16, 16, 200), (32768, 8, 8, 400), (10000, 16, 16, 50), (10000, 8, 8, 2000)):
    x = torch.randn(N, CI, T, device=DEV)
    w1, b1 = torch.randn(C, CI, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    w2, b2 = torch.randn(C, C, 3, device=DEV) * 0.3, torch.randn(C, device=DEV)
    g, be = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    a3 = torch.empty(N, 3 * C * T + 8, dtype=torch.float16, device=DEV)
    rc = L.upd_stg_tcn_ln(_lib.ptr(x), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), _lib.ptr(g), _lib.ptr(be),
                          N, CI, C, T, None, _lib.ptr(a3), None, None, st)
    assert rc == 0
    torch.cuda.synchronize()
    del x, a3
# gated graph aggregation: 163 replicas of a 100-node graph with ~21 in-neighbours, C = Td_h * c = 160
import networkx as nx
from updgm_b200.diffstg import graph_csr
G = nx.barabasi_albert_graph(100, 12, seed=0)
ei = torch.tensor(list(G.to_directed().edges)).t().contiguous()
rowptr, col = [t.to(DEV) for t in graph_csr(ei, 100)]
N, V, C = 16300, 100, 160
kqvs = torch.randn(N, 4 * C, device=DEV)
bias = torch.randn(C, device=DEV)
out = torch.empty(N, C, device=DEV)
assert L.upd_stg_gated_aggregate(_lib.ptr(kqvs), _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(bias), N, V, C, 1, _lib.ptr(out), st) == 0
# heads + NsDiff posterior step on [16000, 100, 1]
N, DH, T, nf, steps = 16000, 4, 100, 1, 20
rows = schedules.stack_rows(schedules.nsdiff_tables("linear", steps, 1e-4, 0.02), schedules.NSDIFF_ROWS).to(DEV)
e = torch.randn(N, DH, T, device=DEV)
w4, b4, ws, bs = [torch.randn(*s, device=DEV) * 0.3 for s in ((nf, DH), (nf,), (nf, DH), (nf,))]
y, yT, z = [torch.randn(N, T, nf, device=DEV) for _ in range(3)]
gx = torch.rand(N, T, nf, device=DEV) + 0.2
o = torch.empty_like(y)
assert L.upd_nsx_step(_lib.ptr(e), _lib.ptr(w4), _lib.ptr(b4), _lib.ptr(ws), _lib.ptr(bs), _lib.ptr(y), _lib.ptr(yT), _lib.ptr(gx),
                      _lib.ptr(z), _lib.ptr(rows), steps, 7, N, DH, T, nf, _lib.ptr(o), None, None, st) == 0
torch.cuda.synchronize()
print("graph kernels ok", bool(torch.isfinite(out).all()))
