"""Phases of one config-1 sweep (NsDiff SLBP, shipped wo_fx checkpoint, W = 981, K = 100) through sample_sweep."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200 import uncertainty as U, kernels
dev = torch.device("cuda:0")
m, _ = U.load_model_from_dir(os.path.join("tests", "golden", "ews_results", "NsDiff_machine", "wo_fx"), device=dev)
g = torch.Generator().manual_seed(0)
series = torch.randn(10000, 2, generator=g).cumsum(0) * 0.01 * m.scaler_std.cpu() + m.scaler_mean.cpu()
wins = series.unfold(0, 200, 10).permute(0, 2, 1).unsqueeze(1).contiguous().pin_memory()
for _ in range(2): U.sample_sweep(m, wins, device=dev)
def t(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = t(); x = U._scale_windows(m, wins, dev); t1 = t()
    traj = m.sample_windows(x, window_base=0); t2 = t()
    cache = torch.empty((981, 1, 100, 200, 2), dtype=torch.float32, pin_memory=True); t3 = t()
    cache.copy_(traj.view(981, 1, 100, 200, 2), non_blocking=True); t4 = t()
    r = kernels.mpv_reduce(traj, 981, 1, want_mean=True); r2 = kernels.mpv_reduce(traj, 981, 1, scale=U._scaler_table(m)); t5 = t()
    st = {k: v.cpu() for k, v in r.items()}; t6 = t()
    print("H2D+scale %.1f | sample_windows %.1f | pinned alloc %.1f | D2H %.1f | reduce x2 %.1f | stats %.1f ms" % tuple(1e3 * d for d in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)))
    del cache
    t0 = t(); c = U.sample_sweep(m, wins, device=dev); t1 = t(); print("sample_sweep total %.1f ms" % (1e3 * (t1 - t0))); del c
