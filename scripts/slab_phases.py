"""GPU diagnostic: per-call times of the heavy C-ABI calls in the emt-slab workload."""
import os, sys, ctypes, collections, argparse, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from sella_b200 import batched, _lib, kernels
from sella_b200.emt import EMTSurface
dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "emt-slab"
args = argparse.Namespace(workload=wl, n=384 if wl != "emt-cluster" else 192, batch=int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
if wl == "quadratic":
    from sella_b200.synthetic import quadratic_batch_torch
    A, xs, x0 = quadratic_batch_torch(args.batch, args.n, dev, seed=1000)
    eng = batched.BatchedSella(batched.QuadraticSurface(A, xs), x0, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8)
else:
    X0, C, cell, pbc = bench.emt_problem(args, 1000, args.batch)
    surf = EMTSurface(args.batch, args.n // 3, dev, cell=cell, pbc=pbc)
    eng = batched.BatchedSella(surf, torch.from_numpy(X0).to(dev), method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8, constraints=(C, None))
times = collections.defaultdict(list)
orig_call = batched.call
def timed_call(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig_call(name, *a); e1.record()
    times[name].append((e0, e1))
batched.call = timed_call
orig_hv = kernels.hv_ld
def timed_hv(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig_hv(*a, **k); e1.record()
    times["hv_ld(nvec=%d%s)" % (a[3], ",T" if k.get("transposed") else "")].append((e0, e1))
    return r
batched.K.hv_ld = timed_hv
lib = _lib.get_lib()
t3 = (ctypes.c_float * 3)()
lib.sb_secular_timing(None, 1)
for t in range(8):
    times.clear()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.step(); z.record()
    torch.cuda.synchronize()
    lib.sb_secular_timing(t3, -1)
    tot = a.elapsed_time(z)
    rows = sorted(((sum(x.elapsed_time(y) for x, y in v), len(v), k) for k, v in times.items()), reverse=True)
    print("step %d total %.2f ms  nterm max %d mean %.1f  last secular call parts %s" % (t, tot, int(eng.nterm.max()), float(eng.nterm.float().mean()), ["%.3f" % v for v in t3]))
    print("    " + "  ".join("%s %.2f(%d)" % (k, ms, c) for ms, c, k in rows[:14]), flush=True)
