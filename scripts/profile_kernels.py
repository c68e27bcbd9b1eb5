"""Minimal driver for `ncu --set full`: a few launches of the H.V kernel and one
batched eigh at the benchmark size (batch 1024, 3N = 384)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200 import kernels as K
from sella_b200.synthetic import quadratic_batch_torch
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 384
dev = torch.device("cuda:0")
A, xs, x0 = quadratic_batch_torch(b, n, dev)
x = x0.view(b, 1, n).contiguous()
for _ in range(4):
    y = K.hv(A, x)
    yt = K.hv(A, x, transposed=True)
w, Vt, st = K.eigh(A)
torch.cuda.synchronize()
print("done", float(y.abs().sum()), float(w.sum()))
