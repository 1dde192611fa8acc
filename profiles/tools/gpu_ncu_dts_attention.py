"""One forward + backward launch of the tcgen05 DiffusionTS attention kernels at the config-4 shape, for `ncu --set full`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200.diffusionts import FusedAttention
dev = torch.device("cuda:0")
H, hs, R, L = 4, 16, 2000, 200
d = H * hs
torch.manual_seed(0)
qb = torch.randn(R, L, 3 * d, device=dev, requires_grad=True)
w = torch.randn(R, L, d, device=dev) * 1e-6
for _ in range(2):
    out = FusedAttention.apply(qb, qb, 0, d, 2 * d, H, d)
    torch.autograd.grad(out, [qb], grad_outputs=w)
torch.cuda.synchronize()
