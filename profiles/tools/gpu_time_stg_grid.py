"""DiffSTG at the corners of BASELINE config 5's sensitivity grid (T_h, T_p in {100, 500}); random-init weights."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200.diffstg import DiffSTG
import networkx as nx
DEV = "cuda:0"
g = np.load("tests/golden/stg_yaml_evalstep.npz"); base = json.loads(str(g["cfg"]))
G = nx.barabasi_albert_graph(100, 12, seed=0)
ei = torch.tensor(list(G.to_directed().edges)).t().contiguous()
for T_h, T_p in ((100, 100), (500, 100), (100, 500), (500, 500)):
    cfg = dict(base, T_h=T_h, T_p=T_p, parallel_sampling=10, sequential_sampling=10, device=DEV)
    torch.manual_seed(1)
    m = DiffSTG(cfg).eval()
    with torch.no_grad():
        m.scaler_std.fill_(1.0)
    win = torch.randn(1, 100, T_h, 1, device=DEV).cumsum(2) * 0.1
    m.sample_windows(win, ei, 100, seed=1, window_base=0)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); out = m.sample_windows(win, ei, 100, seed=1, window_base=0); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3
    print("STG grid T_h=%d T_p=%d: %d node-trajectories in %.2f s -> %.0f traj/s; finite %s; peak mem %.1f GB" % (
        T_h, T_p, out.shape[0] * out.shape[1], t, out.shape[0] * out.shape[1] / t, bool(torch.isfinite(out).all()),
        torch.cuda.max_memory_allocated() / 2 ** 30))
    del m, out
    torch.cuda.empty_cache()
