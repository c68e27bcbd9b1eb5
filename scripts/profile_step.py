"""Profiler target: `--warm` steps untimed, then `--steps` steps inside a cudaProfilerStart/Stop range.
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python scripts/profile_step.py"""
import argparse, sys
import torch
sys.path.insert(0, ".")
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_batch_torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024); ap.add_argument("--n", type=int, default=384)
ap.add_argument("--warm", type=int, default=19); ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--spectrum", default=None)
a = ap.parse_args()
dev = torch.device("cuda:0")
A, xs, x0 = quadratic_batch_torch(a.batch, a.n, dev, seed=1000)
kw = {} if a.spectrum is None else dict(spectrum=a.spectrum)
eng = BatchedSella(QuadraticSurface(A, xs), x0, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8, **kw)
for _ in range(a.warm):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("rows", int(eng.mrows.min()), int(eng.mrows.max()) if eng.compact else None)
