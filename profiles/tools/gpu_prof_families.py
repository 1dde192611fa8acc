"""Short DiffSTG / DiffusionTS / NsDiff_spatial steps for ncu: 2 DiffSTG denoise steps on 16384 replica rows, one DiffusionTS
loop iteration (t = 99 -> 98: x0 prediction, DDIM mean, 3 Langevin iterations, infill) on 1000 rows, then f(x) + g(x) and a
2-step NsDiff_spatial chain on 16000 replica rows.  Every family runs once un-profiled first (weight preparation, CUDA-graph
capture, library heuristics) and once between cudaProfilerStart / Stop: run ncu with --profile-from-start off to list the
steady state only."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from oracle import diffusionts_oracle as dto
DEV = "cuda:0"
from updgm_b200.diffstg import DiffSTG
from updgm_b200.diffusionts import DiffusionTS_model
import networkx as nx
g = np.load("tests/golden/stg_yaml_evalstep.npz"); cfg = json.loads(str(g["cfg"])); shapes = json.loads(str(g["keys"]))
cfg = dict(cfg, parallel_sampling=10, sequential_sampling=10, inference_diffusion_steps=2)
m = DiffSTG(dict(cfg, device=DEV)).eval()
sd = dto.synth_state_dict(shapes, int(g["seed"]))
for k in shapes:
    if ".net.0." in k: sd[k] = sd[k.replace(".net.0.", ".conv.")]
sd.update(scaler_mean=torch.zeros(1), scaler_std=torch.ones(1))
m.load_state_dict(sd, strict=True)
G = nx.barabasi_albert_graph(100, 12, seed=0)
ei = torch.tensor(list(G.to_directed().edges)).t().contiguous()
torch.manual_seed(0)
win = torch.randn(2, 100, 100, 1, device=DEV).cumsum(2) * 0.1
m.rows_per_launch = 16384
out = m.sample_windows(win, ei, 100, seed=1, window_base=0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = m.sample_windows(win, ei, 100, seed=1, window_base=0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("stg ok", tuple(out.shape), bool(torch.isfinite(out).all()))
g = np.load("tests/golden/dts_yaml_steps.npz"); cfg = json.loads(str(g["cfg"])); shapes = json.loads(str(g["keys"]))
d = DiffusionTS_model(dict(cfg, device=DEV)).eval()
d.load_state_dict(dto.synth_state_dict(shapes, int(g["seed"])), strict=False)
d.time_pairs = lambda: [(99, 98)]
tgt = torch.tanh(torch.randn(1000, 100, 1, device=DEV).cumsum(1) * 0.1)
gen = torch.Generator(device=DEV).manual_seed(0)
img = d._sample_rows(tgt, 1000, lambda i, shape: torch.randn(shape, device=DEV, generator=gen))
torch.cuda.synchronize()
torch.cuda.profiler.start()
img = d._sample_rows(tgt, 1000, lambda i, shape: torch.randn(shape, device=DEV, generator=gen))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("dts ok", tuple(img.shape), bool(torch.isfinite(img).all()))

from updgm_b200.nsdiff_spatial import NsDiff_model_spatial
g = np.load("tests/golden/nsx_yaml_evalstep.npz"); cfg = json.loads(str(g["cfg"]))
cfg = dict(cfg, n_z_samples=80, parallel_sample=10, diffusion_steps=2)
torch.manual_seed(123)
x = NsDiff_model_spatial(dict(cfg, device=DEV), "NsDiff_model").eval()
with torch.no_grad():
    x.scaler_std.fill_(1.0)
x.rows_per_launch = 16000
win = 5.0 + torch.randn(2, 100, 100, 1, device=DEV).cumsum(2) * 0.05
out = x.sample_windows(win, ei, 100, seed=1, window_base=0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = x.sample_windows(win, ei, 100, seed=1, window_base=0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("nsx ok", tuple(out.shape), bool(torch.isfinite(out).all()))
