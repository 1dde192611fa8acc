"""Accuracy of an experimental sampler build against the release build: same Philox seed, config-2 shape (F = 1).
    python gpu_sampler_variant_diff.py save /tmp/a.pt [lib.so]   (run once per library)
    python gpu_sampler_variant_diff.py cmp /tmp/a.pt /tmp/b.pt"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
if sys.argv[1] == "cmp":
    a, b = torch.load(sys.argv[2]).double(), torch.load(sys.argv[3]).double()
    d = (a - b).abs()
    rms = b.pow(2).mean().sqrt()
    bound = 1e-3 * b.abs() + 1e-4 * rms
    mpv_a, mpv_b = a.var(dim=1, unbiased=True).mean(), b.var(dim=1, unbiased=True).mean()
    print("max|d|/rms %.3e  rms(d)/rms %.3e  worst |d|/bound %.3f  MPV rel %.3e" % (
        float(d.max() / rms), float(d.pow(2).mean().sqrt() / rms), float((d / bound).max()), float((mpv_a - mpv_b).abs() / mpv_b)))
    sys.exit(0)
if len(sys.argv) > 3:
    os.environ["UPD_LIB_PATH"] = sys.argv[3]
import updgm_b200
from updgm_b200 import kernels, schedules
from conftest import load_golden
dev = torch.device("cuda:0")
tab = schedules.nsdiff_tables("linear", 20, 1e-4, 0.02)
g = load_golden("psample_loop_randF1.npz")
packed = kernels.pack_denoiser(g["sd"], 0, 1, 20, schedules.stack_rows(tab, schedules.NSDIFF_ROWS), dev)
n_win, B, K, O, F = 2, 100, 100, 100, 1
torch.manual_seed(0)
gx = torch.rand(n_win * B, O, F, device=dev) * 0.3 + 0.05
y0 = torch.randn(n_win * B, O, F, device=dev)
out = torch.empty(n_win * B, K, O, F, device=dev)
kernels.nsdiff_sample(packed, y0, gx, n_win, B, K, 10, O, F, 20, seed=1, out=out, impl=4)
torch.cuda.synchronize()
torch.save(out.cpu(), sys.argv[2])
print("saved", sys.argv[2], float(out.abs().mean()))
