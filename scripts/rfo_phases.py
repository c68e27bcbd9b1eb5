"""GPU diagnostic: iteration counts of rfo_tr_kernel on the bench workload."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200 import _lib
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_batch_torch
dev = torch.device("cuda:0")
b, n = 1024, 384
A, xs, x0 = quadratic_batch_torch(b, n, dev, seed=1000)
eng = BatchedSella(QuadraticSurface(A, xs), x0, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8)
lib = _lib.get_lib()
out = (ctypes.c_ulonglong * 8)()
for t in range(11):
    torch.cuda.synchronize(); lib.sb_rfo_profile(out, 1)
    eng.step()
    torch.cuda.synchronize(); lib.sb_rfo_profile(out, 0)
    ns = max(1, out[4])
    print("step %2d systems %d  alpha evals/sys %.1f  root iters/root %.2f  roots/sys %.1f  kcycles/sys %.1f" %
          (t, out[4], out[0] / ns, out[1] / max(1, out[2]), out[2] / ns, out[3] / ns / 1e3), flush=True)
