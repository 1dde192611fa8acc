import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import updgm_b200
from updgm_b200 import _build
import os
if len(sys.argv) > 1: os.environ['UPD_LIB_PATH'] = sys.argv[1]
from updgm_b200 import kernels, schedules
from conftest import load_golden, load_wo_fx_checkpoint
dev = torch.device('cuda:0')
IMPLS = [int(a) for a in sys.argv[2:]] or [2, 4]
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tab = schedules.nsdiff_tables("linear", 20, 1e-4, 0.02)
g = load_golden("psample_loop_randF1.npz")
packed1 = kernels.pack_denoiser(g["sd"], 0, 1, 20, schedules.stack_rows(tab, schedules.NSDIFF_ROWS), dev)
n_win, B, K, O, F = 20, 100, 100, 100, 1
gx = torch.rand(n_win * B, O, F, device=dev) * 0.3 + 0.05
y0 = torch.randn(n_win * B, O, F, device=dev)
out = torch.empty(n_win * B, K, O, F, device=dev)
for impl in IMPLS:
    ms = timeit(lambda: kernels.nsdiff_sample(packed1, y0, gx, n_win, B, K, 10, O, F, 20, seed=1, out=out, impl=impl))
    print(sys.argv[1:] , "impl", impl, "cfg2 F=1: %.2f ms  %.3f G row-steps/s" % (ms, n_win*B*K*O*20/ms/1e6))
_, sd = load_wo_fx_checkpoint()
packed2 = kernels.pack_denoiser(sd, 0, 2, 20, schedules.stack_rows(tab, schedules.NSDIFF_ROWS), dev)
gx = torch.rand(148, 200, 2, device=dev) * 0.06 + 0.01
out = torch.empty(148, 100, 200, 2, device=dev)
for impl in IMPLS:
    ms = timeit(lambda: kernels.nsdiff_sample(packed2, None, gx, 148, 1, 100, 100, 200, 2, 20, seed=1, out=out, impl=impl))
    print(sys.argv[1:], "impl", impl, "cfg1 F=2: %.2f ms  %.3f G row-steps/s" % (ms, 148*100*200*20/ms/1e6))
# TMDM (config 3 shape: 150 positions per trajectory, F=1)
g = load_golden("tmdm_loop_randF1.npz")
tt = schedules.tmdm_tables("linear", 20, 1e-4, 0.02)
packed3 = kernels.pack_denoiser(g["sd"], 1, 1, 20, schedules.stack_rows(tt, schedules.TMDM_ROWS), dev)
n_win, B, K, Lr = 12, 100, 100, 150
y0 = torch.randn(n_win * B, Lr, 1, device=dev)
out = torch.empty(n_win * B, K, Lr, 1, device=dev)
for impl in IMPLS:
    ms = timeit(lambda: kernels.tmdm_sample(packed3, y0, n_win, B, K, 10, Lr, 1, 20, seed=1, out=out, impl=impl))
    print(sys.argv[1:], "impl", impl, "cfg3 TMDM F=1: %.2f ms  %.3f G row-steps/s" % (ms, n_win*B*K*Lr*20/ms/1e6))
