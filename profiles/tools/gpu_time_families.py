"""Timing of the DiffusionTS / DiffSTG samplers on BASELINE configs 4 / 5 shapes (scratch; numbers go to DESIGN.md)."""
import json, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from oracle import diffusionts_oracle as dto
DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "both"
prof = len(sys.argv) > 2 and sys.argv[2] == "prof"

def timed(fn, n=1):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / 1e3, r

if which in ("dts", "both"):
    from updgm_b200.diffusionts import DiffusionTS_model
    g = np.load("tests/golden/dts_yaml_steps.npz"); cfg = json.loads(str(g["cfg"])); shapes = json.loads(str(g["keys"]))
    m = DiffusionTS_model(dict(cfg, device=DEV)).eval()
    m.load_state_dict(dto.synth_state_dict(shapes, int(g["seed"])), strict=False)
    B = 100
    torch.manual_seed(0)
    win = torch.tanh(torch.randn(1, B, 100, 1, device=DEV).cumsum(2) * 0.1)
    x = torch.randn(1000, 200, 1, device=DEV)
    with torch.no_grad():
        m.predict_x0(x, 50)
    t_f, _ = timed(lambda: m.predict_x0(x, 50).sum().item() if False else torch.no_grad()(lambda: m.predict_x0(x, 50))(), 5)
    def fb():
        p = x.clone().requires_grad_(True)
        (gr,) = torch.autograd.grad((m.predict_x0(p, 50) ** 2).sum(), p)
        return gr
    fb(); t_fb, _ = timed(fb, 5)
    print("DTS forward 1000 rows: %.2f ms; fwd+bwd: %.2f ms; peak mem %.1f GB" % (t_f * 1e3, t_fb * 1e3, torch.cuda.max_memory_allocated() / 2**30))
    m.n_z_samples, m.parallel_sample = 20, 10          # 2 chunks of 1000 rows = one launch of 2000 rows
    m.configs.n_z_samples = 20
    m.rows_per_launch = 2000
    t, out = timed(lambda: m.sample_windows(win, seed=1, window_base=0))
    n_traj = out.shape[0] * out.shape[1]
    print("DTS sample: %d trajectories in %.2f s -> %.1f traj/s ; out finite %s, var %.4f" % (n_traj, t, n_traj / t, bool(torch.isfinite(out).all()), float(out.var(dim=1).mean())))
    if prof:
        from torch.profiler import profile, ProfilerActivity
        m.sampling_timesteps = 5
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            m.sample_windows(win, seed=1, window_base=0); torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

if which in ("stg", "both"):
    from updgm_b200.diffstg import DiffSTG
    import networkx as nx
    g = np.load("tests/golden/stg_yaml_evalstep.npz"); cfg = json.loads(str(g["cfg"])); shapes = json.loads(str(g["keys"]))
    cfg = dict(cfg, parallel_sampling=10, sequential_sampling=10)
    m = DiffSTG(dict(cfg, device=DEV)).eval()
    sd = dto.synth_state_dict(shapes, int(g["seed"]))
    for k in shapes:
        if ".net.0." in k: sd[k] = sd[k.replace(".net.0.", ".conv.")]
    sd.update(scaler_mean=torch.zeros(1), scaler_std=torch.ones(1))
    m.load_state_dict(sd, strict=True)
    G = nx.barabasi_albert_graph(100, 12, seed=0)
    ei = torch.tensor(list(G.to_directed().edges)).t().contiguous()
    print("graph edges (directed):", ei.shape[1])
    W = 4
    torch.manual_seed(0)
    win = (torch.randn(W, 100, 100, 1, device=DEV).cumsum(2) * 0.1)
    m.sample_windows(win[:1], ei, 100, seed=1, window_base=0)
    t_first, _ = timed(lambda: m.sample_windows(win, ei, 100, seed=1, window_base=0))
    t, out = timed(lambda: m.sample_windows(win, ei, 100, seed=1, window_base=0))
    n_traj = out.shape[0] * out.shape[1]
    print("STG first call at this shape: %.2f s (library heuristics for the new GEMM shapes)" % t_first)
    print("STG sample: %d node-trajectories (%d windows x 100 nodes x 100 samples) in %.2f s -> %.1f traj/s; finite %s; peak mem %.1f GB" % (n_traj, W, t, n_traj / t, bool(torch.isfinite(out).all()), torch.cuda.max_memory_allocated() / 2**30))
    if prof:
        from torch.profiler import profile, ProfilerActivity
        m.inference_diffusion_steps = 2
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            m.sample_windows(win, ei, 100, seed=1, window_base=0); torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

if which in ("nsx", "both"):
    from updgm_b200.nsdiff_spatial import NsDiff_model_spatial
    import networkx as nx
    g = np.load("tests/golden/nsx_yaml_evalstep.npz"); cfg = json.loads(str(g["cfg"]))
    cfg = dict(cfg, n_z_samples=100, parallel_sample=10)            # the biomass YAML's K / S on the spatial class
    torch.manual_seed(123)
    m = NsDiff_model_spatial(dict(cfg, device=DEV), "NsDiff_model").eval()
    with torch.no_grad():
        m.scaler_std.fill_(1.0)
    G = nx.barabasi_albert_graph(100, 12, seed=0)
    ei = torch.tensor(list(G.to_directed().edges)).t().contiguous()
    W = 4
    torch.manual_seed(0)
    win = 5.0 + (torch.randn(W, 100, 100, 1, device=DEV).cumsum(2) * 0.05)
    m.sample_windows(win[:1], ei, 100, seed=1, window_base=0)
    t_first, _ = timed(lambda: m.sample_windows(win, ei, 100, seed=1, window_base=0))
    t, out = timed(lambda: m.sample_windows(win, ei, 100, seed=1, window_base=0))
    n_traj = out.shape[0] * out.shape[1]
    print("NSX first call at this shape: %.2f s" % t_first)
    print("NSX sample: %d node-trajectories (%d windows x 100 nodes x 100 samples) in %.2f s -> %.1f traj/s; finite frac %.3f; peak mem %.1f GB" % (n_traj, W, t, n_traj / t, float(torch.isfinite(out).float().mean()), torch.cuda.max_memory_allocated() / 2**30))
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            m.sample_windows(win[:1], ei, 100, seed=1, window_base=0); torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
        print(p.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
