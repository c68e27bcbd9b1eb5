"""Where does the compact spectrum drift away from the densely carried matrix?"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_system

dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
shift = int(sys.argv[3]) if len(sys.argv) > 3 else 30
data = [quadratic_system(b, n) for b in (0, 1, 2)]
eng = BatchedSella(QuadraticSurface(up(np.stack([d[0] for d in data])), up(np.stack([d[1] for d in data]))),
                   up(np.stack([d[2] for d in data])), method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8,
                   spectrum="compact", track_B=True)
print("split", eng.split_rotation, eng.split_min_rows)
prev = 0.0
for t in range(nsteps):
    nd0 = eng.ndiag
    eng.step()
    if shift and t % shift == shift - 1:
        eng.surface.xstar += 0.05
        eng._evaluated = False
        eng.surface.evaluate(eng.x, eng.f, eng.g)
    Bt = eng.tracked_B.cpu().numpy(); Bm = eng.B.cpu().numpy()
    err = np.abs(Bt - Bm).reshape(3, -1).max(axis=1)
    orth = []
    for i in range(3):
        th, VR, lam0, m = eng.explicit_pairs(i)
        orth.append(np.abs(VR @ VR.T - np.eye(m)).max() if m else 0.0)
    flag = " <== jump" if err.max() > 10 * max(prev, 1e-15) else ""
    if flag or t % 10 == 0:
        print("step %3d diag %d errB %s orth %s mrows %s skip %s smax %.2e%s" % (
            t, eng.ndiag - nd0, ["%.1e" % e for e in err], ["%.1e" % o for o in orth], eng.mrows.cpu().tolist(),
            eng.skip.cpu().tolist(), float(eng.s.abs().max()), flag))
    prev = err.max()
