"""GPU diagnostic: engine vs oracle at the larger BASELINE sizes (few systems), and invariants."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pes import CartesianPES
from oracle.driver import SaddleSearch
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_system, quadratic_func
dev = torch.device("cuda:0")
to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for n in [int(a) for a in sys.argv[1:]] or [768, 1536]:
    systems = [0, 1]
    data = [quadratic_system(b, n) for b in systems]
    A = np.stack([d[0] for d in data]); xs = np.stack([d[1] for d in data]); x0 = np.stack([d[2] for d in data])
    eng = BatchedSella(QuadraticSurface(to_dev(A), to_dev(xs)), to_dev(x0), method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8)
    orc = []
    for (Ai, xsi, x0i) in data:
        p = CartesianPES(quadratic_func(Ai, xsi), x0i)
        orc.append((p, SaddleSearch(p, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3)))
    out = []
    t0 = time.time()
    for t in range(6):
        eng.step()
        x = eng.x.cpu().numpy()
        errs = []
        for i, (p, o) in enumerate(orc):
            o.step(); errs.append(np.abs(x[i] - p.get_x()).max())
        out.append("%.1e" % max(errs))
    B = eng.B.cpu().numpy(); w = eng.evals.cpu().numpy(); Vt = eng.Vt.cpu().numpy()
    res = max(float(np.abs(B[i] @ Vt[i].T - Vt[i].T * w[i][None, :]).max()) for i in range(2))
    orth = max(float(np.abs(Vt[i] @ Vt[i].T - np.eye(n)).max()) for i in range(2))
    print("n=%d  max|dx| per step: %s  eig resid %.1e orth %.1e status %s  (%.0fs)" % (n, " ".join(out), res, orth, eng.status.cpu().numpy(), time.time() - t0), flush=True)
