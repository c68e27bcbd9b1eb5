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
ap.add_argument("--workload", default="quadratic")
ap.add_argument("--kdiag", type=int, default=5)
ap.add_argument("--no-proj-rot", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
kw = {} if a.spectrum is None else dict(spectrum=a.spectrum)
if a.workload == "quadratic":
    A, xs, x0 = quadratic_batch_torch(a.batch, a.n, dev, seed=1000)
    surf, cons = QuadraticSurface(A, xs), None
else:
    import bench
    from sella_b200.emt import EMTSurface
    ns = argparse.Namespace(workload=a.workload, n=a.n)
    X0, C, cell, pbc = bench.emt_problem(ns, 0, a.batch)
    x0 = torch.from_numpy(X0).to(dev)
    surf = EMTSurface(a.batch, a.n // 3, dev, cell=cell, pbc=pbc)
    cons = (C, None)
    if a.workload == "emt-cluster" and not a.no_proj_rot:
        from sella_b200.internal import BatchedInternals
        cons = (C, None, BatchedInternals(a.n // 3, rotation_ref=X0.reshape(a.batch, a.n // 3, 3)), None)
eng = BatchedSella(surf, x0, method="prfo", rs="tr", diag_maxiter=a.kdiag, diag_every_n=3, kcap=max(8, a.kdiag + 1),
                   constraints=cons, **kw)
for _ in range(a.warm):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("rows", int(eng.mrows.min()), int(eng.mrows.max()) if eng.compact else None)
