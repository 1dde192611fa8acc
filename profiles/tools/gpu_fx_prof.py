import sys, torch, yaml, os
sys.path.insert(0, '/root/repo')
import bench
from updgm_b200.nsdiff import NsDiff_model
from updgm_b200 import uncertainty as U
dev = torch.device("cuda:0")
cfg = bench.workload_config(); net = dict(cfg["net"], device=dev)
torch.manual_seed(123)
model = NsDiff_model(net, "NsDiff_model").eval(); model.scaler_std.fill_(1.0)
series = bench.make_series(0)
stacked = U.stacked_sliding_windows(series, 100, 5).contiguous()
x = stacked.to(dev).view(-1, 100, 1)
with torch.no_grad():
    for _ in range(2): model.condition(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model.condition(x); e1.record(); torch.cuda.synchronize()
    print("condition ms", e0.elapsed_time(e1))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.condition(x); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
