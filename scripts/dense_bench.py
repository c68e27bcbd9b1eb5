"""GPU timing of the dense internal-coordinate algebra at the C3 shape (1024 x 768 x 384)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200 import kernels as K
b, m, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 768, 384
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
A = torch.randn((b, m, n), dtype=torch.float64, device=dev, generator=g)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    z.record(); torch.cuda.synchronize()
    return a.elapsed_time(z) / reps, out
t, (Q, R) = timed(lambda: K.qr(A), 2)
print("qr            %8.2f ms  %6.2f TFLOP/s (2mn^2-2n^3/3 factor + same again for Q)" % (t, 2 * b * (2 * m * n * n - 2 * n ** 3 / 3) / t / 1e9))
t, (Rinv, st) = timed(lambda: K.trtri(R))
print("trtri         %8.2f ms" % t)
t, Binv = timed(lambda: K.gemm(Rinv, Q, transB=True))
print("gemm Binv     %8.2f ms  %6.2f TFLOP/s" % (t, 2.0 * b * n * n * m / t / 1e9))
D = torch.randn((b, n, n), dtype=torch.float64, device=dev, generator=g)
t, T = timed(lambda: K.gemm(D, Binv))
print("gemm D Binv   %8.2f ms  %6.2f TFLOP/s" % (t, 2.0 * b * n * n * m / t / 1e9))
t, Hc = timed(lambda: K.gemm(Binv, T, transA=True))
print("gemm Binv^T T %8.2f ms  %6.2f TFLOP/s" % (t, 2.0 * b * m * m * n / t / 1e9))
t, _ = timed(lambda: torch.matmul(Binv.transpose(1, 2), T))
print("torch.matmul  %8.2f ms  %6.2f TFLOP/s (cuBLAS, for scale)" % (t, 2.0 * b * m * m * n / t / 1e9))
