"""upd_gemm3 column-block choice (UPD_GEMM3_BN = 64 / 128 / 256) on the small-K shapes of DiffSTG / NsDiff_spatial / DiffusionTS,
with and without the fp32 addend (the graph blocks' shortcut)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200 import fx_encoder
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for M, N, K, add in [(100000, 320, 80, False), (100000, 1600, 160, False), (100000, 1600, 160, True), (163840, 320, 80, False),
                     (163840, 3200, 160, True), (163840, 160, 3200, False), (200000, 192, 64, False), (200000, 256, 64, False),
                     (200000, 64, 256, False), (409600, 512, 512, False)]:
    x = torch.randn(M, K, device=dev)
    lin = torch.nn.Linear(K, N).to(dev)
    w3 = fx_encoder._W3Cache().get([(lin.weight, lin.bias)])
    a3 = fx_encoder.a3_split(x)
    addend = torch.randn(M, N, device=dev) if add else None
    t_lib = timeit(lambda: torch.mm(a3, w3.t(), out_dtype=torch.float32))
    line = "M %7d N %5d K %4d add %d: library %.3f ms |" % (M, N, K, add, t_lib)
    for bn in ("64", "128", "256"):
        os.environ["UPD_GEMM3_BN"] = bn
        line += " BN%s %.3f" % (bn, timeit(lambda: fx_encoder.gemm3(a3, w3, N, addend=addend)))
    os.environ.pop("UPD_GEMM3_BN")
    print(line)
