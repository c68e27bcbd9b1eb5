"""Profiler target: the Wilson-matrix factorisation of the internal-coordinate engine at the C3 shape
(1024 x [912 x 384]): Gram GEMM (gemm_kernel), sb_potrf (chol_panel_kernel + trailing GEMMs), sb_trtri.
   ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|chol_panel" -c 6 -o X python scripts/profile_wilson.py"""
import argparse, sys
import torch
sys.path.insert(0, ".")
import bench
from sella_b200 import kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024); ap.add_argument("--n", type=int, default=384)
a = ap.parse_args()
dev = torch.device("cuda:0")
ns = argparse.Namespace(workload="emt-slab", n=a.n)
X0, cell, pbc, ints, rows, cs, h0 = bench.internal_problem(ns, 0, a.batch)
Bw = ints.device_coordinates().jacobian(torch.from_numpy(X0).to(dev))
for _ in range(2):                      # first pass: module loading
    G = K.gemm(Bw, Bw, transA=True)
    G = (0.5 * (G + G.transpose(1, 2))).contiguous()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    G2 = K.gemm(Bw, Bw, transA=True)
    R, st = K.potrf(G)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("status", int(st.max()), "Bw", tuple(Bw.shape))
