"""Side-by-side stages of the first internal-coordinate step: CUDA engine vs oracle (one slab)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_internal_pes import slab_problem, oracle_sets
from oracle.internal_pes import InternalPES
from oracle.driver import SaddleSearch
from oracle.restricted import MaxInternalStep
from sella_b200.batched_internal import BatchedInternalSella
from sella_b200.emt import EMTSurface

dev = torch.device("cuda:0")
at, cons, ints = slab_problem(40)
cs, csc, rows = oracle_sets(ints)
x0 = at.positions.ravel()[None]
surf = EMTSurface(1, ints.natoms, dev, cell=at.cell, pbc=tuple(at.pbc))
kw = dict(method="qn", diag_maxiter=6)
eng = BatchedInternalSella(surf, torch.from_numpy(x0.copy()).to(dev), ints.device_coordinates(), cons_rows=rows,
                           h0=np.diag(ints.guess_hessian()), **kw)
p = InternalPES(at.func, x0[0], cs, csc, integrator="rk")
o = SaddleSearch(p, rs="mis", **kw)
d = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max())
print("H0", d(eng.B[0].cpu().numpy(), p.H.B), "q0", d(eng.x[0].cpu().numpy(), p.get_x()))
eng.ensure_evaluated()
print("g", d(eng.g[0].cpu().numpy(), p.get_g()), "f", float(eng.f[0]) - p.get_f())
# one FD product along a fixed direction, both sides
rng = np.random.RandomState(0)
Uf = p.get_Ufree()
v = Uf @ rng.normal(size=Uf.shape[1]); v /= np.linalg.norm(v)
fo, go = p._calc_eg(p.get_x() + 1e-4 * v)
qt = eng.x + 1e-4 * torch.from_numpy(v[None]).to(dev)
fe, ge = torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, eng.n, dtype=torch.float64, device=dev)
eng.surface.evaluate(qt, fe, ge)
print("fd point: f", float(fe[0]) - fo, "g", d(ge[0].cpu().numpy(), go), "dg/eta", d((ge[0].cpu().numpy() - eng.g[0].cpu().numpy()) / 1e-4, (go - p.get_g()) / 1e-4))
# first diagonalisation
p.diag(gamma=0.1, maxiter=6)
eng._run_diag(None)
print("after diag: H", d(eng.B[0].cpu().numpy(), p.H.B), "relative", d(eng.B[0].cpu().numpy(), p.H.B) / np.abs(p.H.B).max())
print("Ritz", p.last_rr[0], eng.lams[0].cpu().numpy()[:6], "ksz", int(eng.ksz[0]), p.last_rr[1].shape)
print("evalsB vs eigvalsh(H)", d(np.sort(eng.evalsB[0].cpu().numpy()), np.linalg.eigvalsh(eng.B[0].cpu().numpy())))
# the step
o.initialized = True
o.nsteps_since_diag = -1
eng.initialized = True
eng.since_diag.fill_(-1)
p._update_basis()
rs = MaxInternalStep(p, 1, o.delta, method="qn")
s_o, smag_o = rs.get_s()
eng._model()
Ufr = p.get_Ufree()
print("model evals", d(eng.geo["evr"][0, :Ufr.shape[1]].cpu().numpy(), np.linalg.eigvalsh(Ufr.T @ (p.H.B - p.get_Hc()) @ Ufr)))
eng.step()
print("s", d(eng.s[0].cpu().numpy(), s_o), "smag", float(eng.smag[0]) - smag_o, "alpha", float(eng.alpha[0]), getattr(rs, "alpha", None))
o.step()
print("pos", d(eng.pos[0].cpu().numpy(), p.pos), "x", d(eng.x[0].cpu().numpy(), p.get_x()), "delta", float(eng.delta[0]) - o.delta,
      "rho", float(eng.rho[0]) - o.rho, "ode steps", eng.ode_steps, p.ode_nfev)
print("H after step", d(eng.B[0].cpu().numpy(), p.H.B))
print("status", int(eng.status[0]))
