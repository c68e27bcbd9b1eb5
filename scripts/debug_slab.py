"""emt-slab diagnostics: explicit-row counts, per-step time and parity vs the oracle, compact vs dense."""
import sys, time, argparse
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from sella_b200.batched import BatchedSella
from sella_b200.emt import EMTSurface
from oracle.emt import emt_func
from oracle.pes import CartesianPES
from oracle.driver import SaddleSearch

dev = torch.device("cuda:0")
b, n = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 384
ns = argparse.Namespace(workload="emt-slab", n=n)
X0, C, cell, pbc = bench.emt_problem(ns, 0, b)
x0 = torch.from_numpy(X0).to(dev)
for spectrum in ("compact", "dense"):
    surf = EMTSurface(b, n // 3, dev, cell=cell, pbc=pbc)
    eng = BatchedSella(surf, x0, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8, constraints=(C, None),
                       spectrum=spectrum)
    orc = []
    for i in range(2):
        p = CartesianPES(emt_func(cell, pbc), X0[i], C, C @ X0[i])
        orc.append((p, SaddleSearch(p, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3)))
    print("=== %s compact=%s" % (spectrum, eng.compact))
    for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 14):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.step()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        x = eng.x.cpu().numpy()
        dx = []
        for i, (p, o) in enumerate(orc):
            o.step()
            dx.append(np.abs(x[i] - p.get_x()).max())
        extra = ""
        if eng.compact:
            mB, mP = eng.spB.mrows, eng.sp.mrows
            extra = "rowsB %d..%d rowsP %d..%d lam0 %.3f..%.3f" % (int(mB.min()), int(mB.max()), int(mP.min()), int(mP.max()),
                                                                  float(eng.lam0.min()), float(eng.lam0.max()))
        print("step %2d %.1f ms dx %.2e %.2e ndiag %d status %d %s" % (t, dt * 1e3, dx[0], dx[1], eng.ndiag,
                                                                  int(eng.status.max()), extra))
