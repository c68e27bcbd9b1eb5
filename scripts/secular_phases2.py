"""Where do the cycles of the per-term solve kernel (split mode, compact engine) go at the bench setting."""
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
out = (ctypes.c_ulonglong * 16)()
names = ["load+norm", "scan1", "householder", "scan2+givens", "secular+GE+Q", "rowupd/aux", "pending z", "rank sort", "permute"]
for t in range(26):
    if t >= 19:
        torch.cuda.synchronize(); lib.sb_secular_profile(out, 1)
    nd0 = eng.ndiag
    eng.step()
    if t >= 19:
        torch.cuda.synchronize(); lib.sb_secular_profile(out, 0)
        tot = sum(out[:9]) or 1
        print("step %2d diag %d rows %d..%d total %.3f Mcyc/system  " % (t, eng.ndiag - nd0, int(eng.mrows.min()), int(eng.mrows.max()), tot / b / 1e6)
              + "  ".join("%s %.0f%%" % (nm, 100 * out[i] / tot) for i, nm in enumerate(names))
              + "   secular iterations/root %.1f (roots/system %.1f) max %d, >=8: %d, >=20: %d" % (out[14] / max(1, out[15]), out[15] / b, out[13], out[12], out[11]), flush=True)
