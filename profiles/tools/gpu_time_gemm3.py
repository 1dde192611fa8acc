"""upd_gemm3 (csrc/gemm3.cu) against the library GEMM it replaces, on the shapes of the bench's condition encoder
(409 600 token rows, d_model 512, d_ff 256) and of the smaller families.  Prints ms, issued TFLOP/s, ratio."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import updgm_b200
from updgm_b200 import fx_encoder
dev = torch.device("cuda:0")
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for M, N, K in [(409600, 512, 512), (409600, 1536, 512), (409600, 256, 512), (409600, 512, 256), (614400, 1024, 512),
                (200000, 192, 64), (200000, 64, 64), (100000, 320, 80), (100000, 1600, 160)]:
    x = torch.randn(M, K, device=dev)
    lin = torch.nn.Linear(K, N).to(dev)
    w3 = fx_encoder._W3Cache().get([(lin.weight, lin.bias)])
    a3 = fx_encoder.a3_split(x)
    t_own = timeit(lambda: fx_encoder.gemm3(a3, w3, N))
    t_lib = timeit(lambda: torch.mm(a3, w3.t(), out_dtype=torch.float32))
    fl = 2.0 * M * N * a3.shape[1]
    print("M %7d N %5d K %4d (Kp %4d): upd_gemm3 %.3f ms (%.0f TFLOP/s issued)  library %.3f ms (%.0f)  ratio %.2f" %
          (M, N, K, a3.shape[1], t_own, fl / t_own / 1e9, t_lib, fl / t_lib / 1e9, t_lib / t_own))
