"""GPU diagnostic: per-step max |x_engine - x_oracle| for engine variants."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pes import CartesianPES
from oracle.driver import SaddleSearch
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_system, quadratic_func
dev = torch.device("cuda:0")
to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
n, systems = 48, [0, 1, 2]
data = [quadratic_system(b, n) for b in systems]
A = np.stack([d[0] for d in data]); xs = np.stack([d[1] for d in data]); x0 = np.stack([d[2] for d in data])
for variant in sys.argv[1:] or ["base", "base_nodiag", "threepoint", "mjd0", "PSB"]:
    ekw, okw, pkw, upd = {}, {}, {}, None
    if ":" in variant:
        variant, mode = variant.split(":")
        ekw["eig_mode"] = mode
    common = dict(diag_every_n=3, diag_maxiter=5)
    if variant == "threepoint": ekw["threepoint"] = okw["threepoint"] = True
    elif variant in ("mjd0", "gd", "lanczos"): ekw["eigensolver"] = pkw["eigensolver"] = variant
    elif variant == "base_nodiag": common = {}
    elif variant != "base": ekw["update_method"] = upd = variant
    eng = BatchedSella(QuadraticSurface(to_dev(A), to_dev(xs)), to_dev(x0), method="qn", rs="tr", **common, **ekw)
    orc = []
    for (Ai, xsi, x0i) in data:
        p = CartesianPES(quadratic_func(Ai, xsi), x0i, **pkw)
        if upd: p.H.update_method = upd
        orc.append((p, SaddleSearch(p, method="qn", rs="tr", **common, **okw)))
    out = []
    for t in range(9):
        eng.step()
        x = eng.x.cpu().numpy()
        errs = []
        for i, (p, o) in enumerate(orc):
            o.step()
            errs.append(np.abs(x[i] - p.get_x()).max())
        out.append("%.1e" % max(errs))
        if t in (3, 4) and eng.eig_mode == "update":
            B = eng.B.cpu().numpy(); w = eng.evals.cpu().numpy(); Vt = eng.Vt.cpu().numpy()
            print("   step", t, "|B|", [float(np.abs(np.linalg.eigvalsh(B[i])).max()) for i in range(len(orc))],
                  "resid", [float(np.abs(B[i] @ Vt[i].T - Vt[i].T * w[i][None, :]).max()) for i in range(len(orc))],
                  "orth", [float(np.abs(Vt[i] @ Vt[i].T - np.eye(n)).max()) for i in range(len(orc))])
        if t == 5:
            B = eng.B.cpu().numpy()
            print("   step5 per-system err", errs, "nvec", eng.nvec.cpu().numpy(), "Berr",
                  [float(np.abs(B[i] - orc[i][0].H.B).max()) for i in range(len(orc))],
                  "neg evals", [(np.linalg.eigvalsh(orc[i][0].H.B) < 0).sum() for i in range(len(orc))])
    print(variant, " ".join(out), "status", eng.status.cpu().numpy(), "neval", eng.surface.neval, orc[0][0].neval, flush=True)
