"""Pinpoint the stage of the compact eigen-update that loses accuracy (see debug_drift.py)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_system

dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
n = 96
data = [quadratic_system(b, n) for b in (0, 1, 2)]
eng = BatchedSella(QuadraticSurface(up(np.stack([d[0] for d in data])), up(np.stack([d[1] for d in data]))),
                   up(np.stack([d[2] for d in data])), method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8,
                   spectrum="compact", track_B=True)
orig = eng._eigen_update
tag = {"t": 0}


def wrapped(sp, U, J, kc, kvec, nv, active):
    Bs0 = eng.B.cpu().numpy()
    m0 = sp.mrows.cpu().numpy().copy()
    dump = kc > 1 and tag["t"] == 24
    if dump:
        before = dict(VR0=sp.Vt.cpu().numpy(), ev0=sp.evals.cpu().numpy(), m0=m0, rb0=sp.rb, lam0=eng.lam0.cpu().numpy(),
                      U=U.cpu().numpy(), J=J.cpu().numpy(), C=eng.Cmat.cpu().numpy(), nv=nv, kc=kc)
    orig(sp, U, J, kc, kvec, nv, active)
    torch.cuda.synchronize()
    if dump:
        sec_ = eng.seck
        np.savez_compressed("gpurun_out/k_dump.npz", VR1=sp.Vt.cpu().numpy(), ev1=sp.evals.cpu().numpy(),
                            m1=sp.mrows.cpu().numpy(), ncand=eng.ncand.cpu().numpy(), nterm=eng.nterm.cpu().numpy(),
                            **{k: v.cpu().numpy() for k, v in sec_.items()}, **before)
    Bs1 = eng.B.cpu().numpy()
    Un, Jn = U.cpu().numpy(), J.cpu().numpy()
    C = eng.Cmat.view(-1, 32, 33).cpu().numpy()
    sec = eng.sec1 if kc == 1 else eng.seck
    P, sig, nt = sec["P"].cpu().numpy(), sec["sig"].cpu().numpy(), eng.nterm.cpu().numpy()
    kv = np.ones(3, dtype=int) if kvec is None else kvec.cpu().numpy()
    for i in range(3):
        k = kv[i]
        if eng.skip[i]:
            continue
        u, j = Un[i, :k], Jn[i, :k]
        c = 0.5 * (C[i, :k, :k] + C[i, :k, :k].T)
        D = u.T @ j + j.T @ u - u.T @ c @ u
        Dl = (P[i, :nt[i]].T * sig[i, :nt[i]]) @ P[i, :nt[i]]
        e_fac = np.abs(D - Dl).max()
        e_upd = np.abs((Bs1[i] - Bs0[i]) - D).max()
        th, VR, lam0, m = eng.explicit_pairs(i)
        res = P[i, :nt[i]] - (P[i, :nt[i]] @ VR.T) @ VR
        print("   upd t=%d sys %d k=%d |D| %.1e |U| %.1e |J| %.1e  factor err %.1e  spec-update err %.1e  span residual %.1e  m %d->%d sig %s"
              % (tag["t"], i, k, np.abs(D).max(), np.abs(u).max(), np.abs(j).max(), e_fac, e_upd,
                 np.abs(res).max(), m0[i], m, np.array2string(sig[i, :nt[i]], precision=2)))


eng._eigen_update = wrapped
for t in range(26):
    tag["t"] = t
    verbose = t >= 14
    if not verbose:
        eng._eigen_update = orig
    else:
        eng._eigen_update = wrapped
    eng.step()
    Bt = eng.tracked_B.cpu().numpy(); Bm = eng.B.cpu().numpy()
    print("step %d errB %s smax %.1e" % (t, ["%.1e" % e for e in np.abs(Bt - Bm).reshape(3, -1).max(axis=1)], float(eng.s.abs().max())))
