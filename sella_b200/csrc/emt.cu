// EMT-form copper potential (energy + gradient) for a batch of configurations: the on-device
// surface evaluator for the EMT configurations of BASELINE.json (C1-C3), SURVEY.md 8(f1).
//
// The reference evaluates forces through an ASE calculator on the host (sella/peswrapper.py:
// 413-418, `atoms.get_forces()`), one call per optimiser step and per Davidson vector; ASE's
// EMT is a third-party dependency that is not part of the reference tree, so this file follows
// the published functional form (Jacobsen, Stoltze, Norskov, Surf. Sci. 366 (1996) 394) exactly
// as restated in oracle/emt.py -- "EMT-form", parity with ASE unpinned.
//
//   E = sum_i F(sigma1_i) + 1/2 sum_{i != j} phi(r_ij)
//   sigma1_i = sum_j rho(r_ij),  rho = w(r) exp(-eta2 (r - beta s0)) / gamma1
//   phi      = -V0 w(r) exp(-kappa (r/beta - s0)) / gamma2,   w = 1/(1 + exp(acut (r - rc)))
//   F(s)     = E0 ((1 + lam ds) exp(-lam ds) - 1) + 6 V0 exp(-kappa ds),  ds = -ln(s/12)/(beta eta2)
//   dE/dx_i  = sum_j [(F'_i + F'_j) rho'(r_ij) + phi'(r_ij)] (x_i - x_j)/r_ij
//
// One CTA per configuration; positions, sigma1 and F' live in shared memory.  A warp owns an
// atom i and its lanes sweep the partner atoms j (all periodic images inside the cutoff), so
// the O(N^2 * images) distance tests are spread over the whole CTA and the per-atom sums are
// warp reductions in a fixed order (deterministic).  Systems here have 64-512 atoms: the
// pair sweep is compute-bound and tiny next to the n^2 Hessian passes of a step.
#include "common.cuh"

namespace {

struct EmtPar {
    double E0, s0, V0, eta2, kappa, lam, beta, rc, acut, g1inv, g2inv, rlist2;
    int nimg[3];
};

constexpr int EMT_THREADS = 256;

__device__ __forceinline__ void emt_pair(const EmtPar& P, double r, double* rho, double* drho, double* phi,
                                         double* dphi) {
    const double ex = exp(P.acut * (r - P.rc));
    const double w = 1.0 / (1.0 + ex);
    const double dw = -P.acut * w * (1.0 - w);
    const double e1 = exp(-P.eta2 * (r - P.beta * P.s0)) * P.g1inv;
    const double e2 = exp(-P.kappa * (r / P.beta - P.s0)) * P.g2inv;
    *rho = w * e1;
    *drho = dw * e1 - P.eta2 * w * e1;
    *phi = -P.V0 * w * e2;
    *dphi = -P.V0 * (dw * e2 - P.kappa / P.beta * w * e2);
}

__global__ void __launch_bounds__(EMT_THREADS)
emt_kernel(const double* __restrict__ x_, int natoms, const double* __restrict__ cell_, long long cellstride,
           EmtPar P, double* __restrict__ f_, double* __restrict__ g_, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* pos = sm;                       // 3 N
    double* sig = pos + 3 * natoms;         // N
    double* dF = sig + natoms;              // N
    double* red = dF + natoms;              // SB_SCRATCH_DOUBLES
    __shared__ double cell[9];
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const int n = 3 * natoms;
    for (int i = tid; i < n; i += nt) pos[i] = x_[(size_t)b * n + i];
    if (tid < 9) cell[tid] = cell_ ? cell_[(size_t)b * cellstride + tid] : 0.0;
    __syncthreads();
    const int n0 = P.nimg[0], n1 = P.nimg[1], n2 = P.nimg[2];
    const int nimg = (2 * n0 + 1) * (2 * n1 + 1) * (2 * n2 + 1);
    // ---- pass 1: sigma1_i and the pair energy
    double epair = 0.0;
    for (int i = warp; i < natoms; i += nw) {
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        double s1 = 0.0, ep = 0.0;
        for (int im = 0; im < nimg; ++im) {
            const int a = im / ((2 * n1 + 1) * (2 * n2 + 1)) - n0;
            const int c = (im / (2 * n2 + 1)) % (2 * n1 + 1) - n1;
            const int d = im % (2 * n2 + 1) - n2;
            const double sx = a * cell[0] + c * cell[3] + d * cell[6];
            const double sy = a * cell[1] + c * cell[4] + d * cell[7];
            const double sz = a * cell[2] + c * cell[5] + d * cell[8];
            for (int j = lane; j < natoms; j += 32) {
                const double dx = xi - (pos[3 * j] + sx), dy = yi - (pos[3 * j + 1] + sy), dz = zi - (pos[3 * j + 2] + sz);
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < P.rlist2 && r2 > 1e-18) {
                    double rho, drho, phi, dphi;
                    emt_pair(P, sqrt(r2), &rho, &drho, &phi, &dphi);
                    s1 += rho;
                    ep += 0.5 * phi;
                }
            }
        }
        s1 = sb_warp_sum(s1);
        ep = sb_warp_sum(ep);
        if (lane == 0) {
            const double ds = -log(s1 / 12.0) / (P.beta * P.eta2);
            const double xl = P.lam * ds;
            const double exl = exp(-xl), ek = exp(-P.kappa * ds);
            sig[i] = s1;
            dF[i] = (P.E0 * P.lam * xl * exl + 6.0 * P.V0 * P.kappa * ek) / (P.beta * P.eta2 * s1);
            epair += ep + P.E0 * ((1.0 + xl) * exl - 1.0) + 6.0 * P.V0 * ek;
        }
    }
    const double etot = sb_block_sum(lane == 0 ? epair : 0.0, red);
    if (tid == 0) f_[b] = etot;
    __syncthreads();
    // ---- pass 2: gradient on atom i (every ordered pair is visited from i's side)
    for (int i = warp; i < natoms; i += nw) {
        const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
        const double dFi = dF[i];
        double gx = 0.0, gy = 0.0, gz = 0.0;
        for (int im = 0; im < nimg; ++im) {
            const int a = im / ((2 * n1 + 1) * (2 * n2 + 1)) - n0;
            const int c = (im / (2 * n2 + 1)) % (2 * n1 + 1) - n1;
            const int d = im % (2 * n2 + 1) - n2;
            const double sx = a * cell[0] + c * cell[3] + d * cell[6];
            const double sy = a * cell[1] + c * cell[4] + d * cell[7];
            const double sz = a * cell[2] + c * cell[5] + d * cell[8];
            for (int j = lane; j < natoms; j += 32) {
                const double dx = xi - (pos[3 * j] + sx), dy = yi - (pos[3 * j + 1] + sy), dz = zi - (pos[3 * j + 2] + sz);
                const double r2 = dx * dx + dy * dy + dz * dz;
                if (r2 < P.rlist2 && r2 > 1e-18) {
                    const double r = sqrt(r2);
                    double rho, drho, phi, dphi;
                    emt_pair(P, r, &rho, &drho, &phi, &dphi);
                    const double coef = ((dFi + dF[j]) * drho + dphi) / r;
                    gx = fma(coef, dx, gx); gy = fma(coef, dy, gy); gz = fma(coef, dz, gz);
                }
            }
        }
        gx = sb_warp_sum(gx); gy = sb_warp_sum(gy); gz = sb_warp_sum(gz);
        if (lane == 0) {
            double* g = g_ + (size_t)b * n + 3 * i;
            g[0] = gx; g[1] = gy; g[2] = gz;
        }
    }
}

}  // namespace

// par6 (host): E0, s0, V0, eta2, kappa, lambda in eV / Angstrom units; nimg3 (host): periodic images
// per lattice direction; cell (device): 9 doubles (rows = lattice vectors), shared (cellstride 0)
// or per configuration (cellstride 9); may be NULL for non-periodic clusters.
extern "C" int sb_emt_pes_impl(const double* x, int natoms, const double* cell, long long cellstride,
                               const int* nimg3, const double* par6, double* f, double* g, const int* active,
                               int batch, cudaStream_t st) {
    EmtPar P;
    P.E0 = par6[0]; P.s0 = par6[1]; P.V0 = par6[2]; P.eta2 = par6[3]; P.kappa = par6[4]; P.lam = par6[5];
    P.beta = 1.809;
    const double s3 = sqrt(3.0);
    P.rc = P.beta * P.s0 * 0.5 * (s3 + 2.0);
    const double rr = 4.0 * P.rc / (s3 + 2.0);
    P.acut = log(9999.0) / (rr - P.rc);
    double g1 = 0.0, g2 = 0.0;
    const int shell[3] = {12, 6, 24};
    for (int i = 0; i < 3; ++i) {
        const double r = P.s0 * P.beta * sqrt(i + 1.0);
        const double xx = shell[i] / (12.0 * (1.0 + exp(P.acut * (r - P.rc))));
        g1 += xx * exp(-P.eta2 * (r - P.beta * P.s0));
        g2 += xx * exp(-P.kappa / P.beta * (r - P.beta * P.s0));
    }
    P.g1inv = 1.0 / g1; P.g2inv = 1.0 / g2;
    const double rlist = P.rc + 0.5;
    P.rlist2 = rlist * rlist;
    for (int d = 0; d < 3; ++d) P.nimg[d] = (cell && nimg3) ? nimg3[d] : 0;
    const size_t smem = ((size_t)5 * natoms + SB_SCRATCH_DOUBLES) * sizeof(double);
    if (smem > 200 * 1024) return -2;
    cudaFuncSetAttribute(emt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    emt_kernel<<<batch, EMT_THREADS, smem, st>>>(x, natoms, cell, cellstride, P, f, g, active);
    return SB_LAUNCH_CHECK();
}
