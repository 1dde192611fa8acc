"""Throughput of the other BASELINE configs through the public sweep API (numbers for DESIGN.md section 5)."""
import json, os, sys, time
import numpy as np, torch, yaml
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import updgm_b200
from updgm_b200 import uncertainty as U
from oracle import diffusionts_oracle as dto
dev = torch.device("cuda:0")
G = os.path.join("tests", "golden")

def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()   # the previous result is still alive: every call pins a fresh cache
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, r

# ---- config 1: NsDiff SLBP, shipped checkpoint, W=981, K=100 ----
m, _ = U.load_model_from_dir(os.path.join(G, "ews_results", "NsDiff_machine", "wo_fx"), device=dev)
g = torch.Generator().manual_seed(0)
x = torch.zeros(10000, 2)
e = torch.randn(10000, 2, generator=g) * 0.1
for t in range(1, 10000): x[t] = 0.99 * x[t - 1] + e[t]
series = x * m.scaler_std.cpu() + m.scaler_mean.cpu()
wins = series.unfold(0, 200, 10).permute(0, 2, 1).unsqueeze(1).contiguous()      # [981, 1, 200, 2]
t, c = timed(lambda: U.sample_sweep(m, wins, device=dev))
print("config 1 (NsDiff SLBP wo_fx, W=%d, K=100): %.1f ms per sweep, %.3e traj/s (e2e, cache to host)" % (wins.shape[0], t * 1e3, wins.shape[0] * 100 / t))

# ---- config 3: TMDM neuronal ER-100, K=100, S=10, W=181 ----
from updgm_b200.tmdm import TMDM_model
cfg = yaml.safe_load(open(os.path.join(G, "ews_results", "model_compare", "TMDM", "neuronal", "model_trained.yaml")))
torch.manual_seed(123)
tm = TMDM_model(dict(cfg["net"], device=dev)).eval()
g = torch.Generator().manual_seed(0)
s = torch.sigmoid((torch.randn(100, 1000, 1, generator=g) * 0.1).cumsum(1))
wins = s.unfold(1, 100, 5).permute(1, 0, 3, 2).contiguous()                       # [181, 100, 100, 1]
t, c = timed(lambda: U.sample_sweep(tm, wins, device=dev))
print("config 3 (TMDM neuronal, W=%d, B=100, K=100): %.1f ms per sweep, %.3e traj/s (e2e)" % (wins.shape[0], t * 1e3, wins.shape[0] * 100 * 100 / t))

# ---- config 4: DiffusionTS, rows_per_launch sweep ----
from updgm_b200.diffusionts import DiffusionTS_model
gd = np.load(os.path.join(G, "dts_yaml_steps.npz")); cfg = json.loads(str(gd["cfg"])); shapes = json.loads(str(gd["keys"]))
d = DiffusionTS_model(dict(cfg, device=dev, n_z_samples=40, parallel_sample=10)).eval()
d.load_state_dict(dto.synth_state_dict(shapes, int(gd["seed"])), strict=False)
win = torch.tanh(torch.randn(1, 100, 100, 1, device=dev).cumsum(2) * 0.1)
for rpl in (2000, 4000):
    d.rows_per_launch = rpl
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter(); out = d.sample_windows(win, seed=1, window_base=0); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print("config 4 (DiffusionTS, 100 nodes x 40 samples, rows/launch %d): %.2f s -> %.1f traj/s, peak %.1f GB" % (rpl, t, out.shape[0] * out.shape[1] / t, torch.cuda.max_memory_allocated() / 2**30))
