import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from updgm_b200.nsdiff import NsDiff_model
from updgm_b200 import uncertainty as U, kernels
dev = torch.device("cuda:0")
cfg = bench.workload_config(); net = dict(cfg["net"], device=dev)
torch.manual_seed(123)
model = NsDiff_model(net, "NsDiff_model").eval(); model.scaler_std.fill_(1.0)
series = bench.make_series(0)
stacked = U.stacked_sliding_windows(series, 100, 5).contiguous().pin_memory()
for _ in range(2): U.sample_sweep(model, stacked, device=dev)
torch.cuda.synchronize()
def t(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = t(); cache = torch.empty((181, 100, 100, 100, 1), dtype=torch.float32, pin_memory=True); t1 = t()
    x = U._scale_windows(model, stacked, dev); t2 = t()
    traj = model.sample_windows(x, window_base=0); t3 = t()
    cache.copy_(traj.view(181, 100, 100, 100, 1), non_blocking=True); t4 = t()
    r = kernels.mpv_reduce(traj, 181, 100, want_mean=True); sc = U._scaler_table(model); r2 = kernels.mpv_reduce(traj, 181, 100, scale=sc); t5 = t()
    st = {k: v.cpu() for k, v in r.items()}; t6 = t()
    print("pinned alloc %.1f ms | H2D+scale %.1f | sample_windows %.1f | D2H copy %.1f | reduce x2 %.1f | stats cpu %.1f" % tuple(1e3 * d for d in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)))
    del cache
    t0 = t(); c = U.sample_sweep(model, stacked, device=dev); t1 = t(); print("sample_sweep total %.1f ms" % (1e3 * (t1 - t0))); del c
