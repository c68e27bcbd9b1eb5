"""GPU timing of the batched eigensolver on quasi-Newton Hessians (lam0 I + low rank) and on
random symmetric matrices."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200 import kernels as K
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    z.record(); torch.cuda.synchronize()
    return a.elapsed_time(z) / reps
for (b, n) in ((1024, 384), (256, 192)):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    U = torch.randn((b, n, 24), dtype=torch.float64, device="cuda", generator=g)
    A = 0.7 * torch.eye(n, dtype=torch.float64, device="cuda") + U @ torch.diag_embed(torch.randn((b, 24), dtype=torch.float64, device="cuda", generator=g)) @ U.transpose(1, 2) / n
    R = torch.randn((b, n, n), dtype=torch.float64, device="cuda", generator=g); R = R + R.transpose(1, 2)
    print(b, n, "quasi-Newton %.1f ms   random %.1f ms" % (timed(lambda: K.eigh(A)), timed(lambda: K.eigh(R))), flush=True)
