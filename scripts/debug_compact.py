"""Step-by-step diagnostics of the compact representation against the oracle and against the dense
matrix carried independently (track_B)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_system, quadratic_func
from oracle.pes import CartesianPES
from oracle.driver import SaddleSearch

dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def run(n, method, rs, nsteps, nfix=0, **kw):
    systems = [0, 1, 2]
    data = [quadratic_system(b, n) for b in systems]
    C = np.eye(n)[:nfix] if nfix else None
    eng = BatchedSella(QuadraticSurface(up(np.stack([d[0] for d in data])), up(np.stack([d[1] for d in data]))),
                       up(np.stack([d[2] for d in data])), method=method, rs=rs, spectrum="compact", track_B=True,
                       constraints=None if C is None else (C, None), **kw)
    orc = []
    for (A, xs, x0) in data:
        p = CartesianPES(quadratic_func(A, xs), x0, C, None if C is None else C @ x0)
        orc.append((p, SaddleSearch(p, method=method, rs=rs, **{k: v for k, v in kw.items() if k in ("diag_maxiter", "diag_every_n")})))
    print("=== n=%d %s %s %s" % (n, method, rs, kw))
    for t in range(nsteps):
        try:
            eng.step()
        except Exception as exc:
            print("step %d raised %r" % (t, exc))
            break
        x = eng.x.cpu().numpy()
        dx = []
        for i, (p, o) in enumerate(orc):
            o.step()
            dx.append(np.abs(x[i] - p.get_x()).max())
        Bt = eng.tracked_B.cpu().numpy()
        Bm = eng.B.cpu().numpy()
        errB = np.abs(Bt - Bm).max()
        errO = max(np.abs(Bt[i] - orc[i][0].H.B).max() for i in range(3))
        orth = 0.0
        for i in range(3):
            th, VR, lam0, m = eng.explicit_pairs(i)
            if m:
                orth = max(orth, np.abs(VR @ VR.T - np.eye(m)).max())
        print("step %2d dx %.2e  |Btrack-Bspec| %.2e  |Btrack-Boracle| %.2e  orth %.2e  mrows %s rb %d status %s delta %s"
              % (t, max(dx), errB, errO, orth, (eng.mrows.cpu().tolist(), eng.sp.mrows.cpu().tolist()), eng._rb, eng.status.cpu().tolist(),
                 np.round(eng.delta.cpu().numpy(), 6).tolist()))
        if not np.isfinite(max(dx)) or max(dx) > 1e-3:
            print("diverged; stopping this case")
            break


if __name__ == "__main__":
    run(30, "qn", "tr", 14)
    run(30, "prfo", "tr", 14)
    run(30, "qn", "ras", 8)
    run(48, "prfo", "tr", 10, diag_maxiter=5, diag_every_n=3)
    run(48, "prfo", "ras", 6)
    run(30, "qn", "tr", 10, nfix=6)
    run(48, "prfo", "ras", 8, nfix=6)
    run(48, "prfo", "tr", 10, nfix=9, diag_maxiter=5, diag_every_n=3)
