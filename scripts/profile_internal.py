"""Profiler target for the internal-coordinate engine (C3 as named): `--warm` steps untimed, then `--steps`
steps inside a cudaProfilerStart/Stop range.
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X \
       python scripts/profile_internal.py"""
import argparse, sys
import torch
sys.path.insert(0, ".")
import bench
from sella_b200.batched_internal import BatchedInternalSella
from sella_b200.emt import EMTSurface

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256); ap.add_argument("--n", type=int, default=384)
ap.add_argument("--warm", type=int, default=3); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--inexact-geodesic", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
ns = argparse.Namespace(workload="emt-slab", n=a.n)
X0, cell, pbc, ints, rows, cs, h0 = bench.internal_problem(ns, 0, a.batch)
surf = EMTSurface(a.batch, a.n // 3, dev, cell=cell, pbc=pbc)
eng = BatchedInternalSella(surf, torch.from_numpy(X0).to(dev), ints.device_coordinates(), cons_rows=rows, h0=h0,
                           method="prfo", diag_maxiter=5, diag_every_n=3, kcap=8,
                           exact_geodesic=not a.inexact_geodesic)
for _ in range(a.warm):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ode steps", eng.ode_steps, "status", int(eng.status.max()))
