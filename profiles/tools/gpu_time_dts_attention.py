"""DiffusionTS attention (head size 16): the tcgen05 kernels (csrc/dts_attention_tc.cu) against the fp32 FFMA kernels they
replace (csrc/dts_attention.cu, forced with UPD_DTS_ATTN_FFMA=1), forward and backward, at the shapes of BASELINE
config 4 (seq 200, 4 heads, 1000 / 2000 rows).  Also prints the deviation of both from a float64 reference."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200.diffusionts import FusedAttention
dev = torch.device("cuda:0")
H, hs = 4, 16
d = H * hs
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())
for R, L in [(1000, 200), (2000, 200), (2000, 96)]:
    torch.manual_seed(R + L)
    qb = torch.randn(R, L, 3 * d, device=dev, requires_grad=True)
    w = torch.randn(R, L, d, device=dev) * 1e-6            # refinement-gradient sized cotangent
    res = {}
    for tag, env in (("tcgen05", "0"), ("ffma", "1")):
        os.environ["UPD_DTS_ATTN_FFMA"] = env
        out = FusedAttention.apply(qb, qb, 0, d, 2 * d, H, d)
        (g,) = torch.autograd.grad(out, [qb], grad_outputs=w)
        t_f = timeit(lambda: FusedAttention.apply(qb.detach(), qb.detach(), 0, d, 2 * d, H, d))
        def fb():
            o = FusedAttention.apply(qb, qb, 0, d, 2 * d, H, d)
            torch.autograd.grad(o, [qb], grad_outputs=w)
        t_fb = timeit(fb)
        res[tag] = (out.detach(), g, t_f, t_fb)
    n = min(R, 64)                                           # float64 reference on a slice
    q64 = qb.detach()[:n].double().requires_grad_(True)
    heads = lambda t: t.reshape(n, -1, H, hs).transpose(1, 2)
    att = torch.softmax(heads(q64[..., :d]) @ heads(q64[..., d:2 * d]).transpose(-1, -2) / math.sqrt(hs), -1)
    ref = (att @ heads(q64[..., 2 * d:])).transpose(1, 2).reshape(n, L, d)
    (gref,) = torch.autograd.grad(ref, [q64], grad_outputs=w[:n].double())
    for tag in ("tcgen05", "ffma"):
        out, g, t_f, t_fb = res[tag]
        print("R %5d L %4d %-8s fwd %.3f ms  fwd+bwd %.3f ms (bwd %.3f)   err fwd %.1e  grad %.1e" %
              (R, L, tag, t_f, t_fb, t_fb - t_f, rel(out[:n], ref.detach()), rel(g[:n], gref)))
