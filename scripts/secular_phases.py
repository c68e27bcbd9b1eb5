"""GPU diagnostic: where do the cycles of secular_update_kernel go (bench workload)."""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200 import _lib
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_batch_torch
dev = torch.device("cuda:0")
b, n = 1024, 384
A, xs, x0 = quadratic_batch_torch(b, n, dev, seed=1000)
eng = BatchedSella(QuadraticSurface(A, xs), x0, method="qn", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8)
lib = _lib.get_lib()
out = (ctypes.c_ulonglong * 16)()
names = ["load+norm", "defl scan1", "householder", "defl scan2+givens", "secular+GE+Q", "row update", "pending z", "rank sort", "permute"]
for t in range(11):
    torch.cuda.synchronize(); lib.sb_secular_profile(out, 1)
    eng.step()
    torch.cuda.synchronize(); lib.sb_secular_profile(out, 0)
    tot = sum(out[:9]) or 1
    print("   moved rows per launch-CTA: %.1f  mean dist %.2f  dist==1: %.1f  equal-d moves: %.1f" % (out[9] / max(1, out[10]), out[11] / max(1, out[9]), out[12] / max(1, out[10]), out[13] / max(1, out[10])))
    print("step %2d total %.3f Mcyc/CTA  " % (t, tot / b / 1e6) + "  ".join("%s %.0f%%" % (nm, 100 * out[i] / tot) for i, nm in enumerate(names)), flush=True)
