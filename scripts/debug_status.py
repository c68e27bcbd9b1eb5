"""Debug helper (GPU): replay the bench workload, find systems whose status word is
set, dump their inputs and step state to gpurun_out/ for CPU analysis."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200.batched import BatchedSella, QuadraticSurface
from sella_b200.synthetic import quadratic_batch_torch
dev = torch.device("cuda:0")
b, n = 1024, 384
A, xs, x0 = quadratic_batch_torch(b, n, dev, seed=1000)
eng = BatchedSella(QuadraticSurface(A, xs), x0, method="qn", rs=sys.argv[1] if len(sys.argv) > 1 else "tr",
                   diag_maxiter=5, diag_every_n=3, kcap=8)
os.makedirs("gpurun_out", exist_ok=True)
for t in range(11):
    g_before = eng.g.clone(); delta_before = eng.delta.clone(); B_before = eng.B.clone()
    eng.step()
    st = eng.status.cpu().numpy()
    print("step", t, "bad", int((st != 0).sum()), "smag max", float(eng.smag.max()), "alpha max", float(eng.alpha.max()),
          "ndiag", eng.ndiag, flush=True)
    if st.any():
        i = int(np.nonzero(st)[0][0])
        np.savez("gpurun_out/debug_sys.npz", A=A[i].cpu().numpy(), xs=xs[i].cpu().numpy(), x0=x0[i].cpu().numpy(),
                 step=t, status=st[i], g=g_before[i].cpu().numpy(), delta=float(delta_before[i]),
                 B=B_before[i].cpu().numpy(), evals=eng.evals[i].cpu().numpy(), Vg=eng.Vg[i].cpu().numpy(),
                 alpha=float(eng.alpha[i]), smag=float(eng.smag[i]), idx=i)
        print("dumped system", i, "status", st[i])
        break
