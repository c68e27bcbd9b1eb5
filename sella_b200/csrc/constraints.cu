// Linear-constraint support kernels (per system or shared constraint data).
//
// Reference: sella/peswrapper.py
//   _split_cons_subspace / _calc_basis (:51-69, :395-407)  Ucons / Ufree (host, once)
//   get_scons (:429-438)            scons = -Ucons lstsq(drdx Ucons, res)
//   get_projected_forces (:558-562) -Ufree Ufree^T g
//   BaseRestrictedStep.__init__ (restricted_step.py:28-49): g + H scons, the
//   "constraint violation alone exceeds the radius" branch (NaiveStepper).
// For linear constraints C x = c the bases are geometry independent, so everything
// reduces to products with two thin matrices stored row-wise:
//   R[j,:] (nr x n)  ->  dots:  out[j] = R[j,:].x (- c[j]) ;  comb: out[:] (+)= sum_j coef[j] R[j,:]
// `rstride` = nr*n for per-system matrices, 0 when the whole batch shares one.
#include "small_dense.cuh"

namespace {

constexpr int CT = 256;

// out[b,j] = R[j,:].x[b,:] - (c ? c[b,j] : 0)
__global__ void __launch_bounds__(CT)
rect_dots_kernel(const double* __restrict__ R, size_t rstride, int nr, const double* __restrict__ x, size_t xstride,
                 const double* __restrict__ c, double* __restrict__ out, int n, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const double* Rb = R + (size_t)b * rstride;
    const double* xb = x + (size_t)b * xstride;
    for (int j = warp; j < nr; j += nw) {
        const double* row = Rb + (size_t)j * n;
        double acc = 0.0;
        for (int i = lane; i < n; i += 32) acc = fma(row[i], xb[i], acc);
        acc = sb_warp_sum(acc);
        if (lane == 0) out[(size_t)b * nr + j] = acc - (c ? c[(size_t)b * nr + j] : 0.0);
    }
}

// out[b,:] = beta*base[b,:] + scale * sum_j coef[b,j] R[j,:]      (base may alias out)
__global__ void __launch_bounds__(CT)
rect_comb_kernel(const double* __restrict__ R, size_t rstride, int nr, const double* __restrict__ coef,
                 double scale, const double* __restrict__ base, size_t bstride, double beta,
                 double* __restrict__ out, size_t ostride, int n, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double cs[];
    for (int j = threadIdx.x; j < nr; j += blockDim.x) cs[j] = coef[(size_t)b * nr + j];
    __syncthreads();
    const double* Rb = R + (size_t)b * rstride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
#pragma unroll 4
        for (int j = 0; j < nr; ++j) acc = fma(cs[j], Rb[(size_t)j * n + i], acc);
        const double b0 = base ? beta * base[(size_t)b * bstride + i] : 0.0;
        out[(size_t)b * ostride + i] = b0 + scale * acc;
    }
}

// After scons is known: its size under the step constraint, |scons|^2, and whether the
// violation alone exceeds the radius (restricted_step.py:44): naive[b] = 1 then.
// kind 0: |s| ; kind 1: max atomic displacement.
__global__ void __launch_bounds__(CT)
scons_measure_kernel(const double* __restrict__ scons, const double* __restrict__ delta, int kind, int n,
                     double* __restrict__ scons2, double* __restrict__ consval, int* __restrict__ naive,
                     int* __restrict__ regular) {
    const int b = blockIdx.x;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    __shared__ double red[CT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* s = scons + (size_t)b * n;
    double a = 0.0;
    for (int i = tid; i < n; i += blockDim.x) a = fma(s[i], s[i], a);
    a = sb_block_sum(a, scratch);
    double val = sqrt(a);
    if (kind == 1) {
        double best = 0.0;
        for (int at = tid; at < n / 3; at += blockDim.x) {
            const double x = s[3 * at], y = s[3 * at + 1], z = s[3 * at + 2];
            best = fmax(best, sqrt(x * x + y * y + z * z));
        }
        best = sb_warp_max(best);
        if (lane == 0) red[warp] = best;
        __syncthreads();
        if (tid == 0) { for (int w = 1; w < blockDim.x / 32; ++w) best = fmax(best, red[w]); red[0] = best; }
        __syncthreads();
        val = red[0];
    }
    if (tid == 0) {
        scons2[b] = a;
        consval[b] = val;
        const int nv = (val - delta[b] > 1e-8) ? 1 : 0;
        naive[b] = nv;
        regular[b] = 1 - nv;
    }
}

// stot = slift + scons for regular systems; the naive branch (NaiveStepper: s(alpha) = alpha*scons,
// alpha0 = 0.5) returns 0.5*scons when that is inside the radius, else the root alpha = delta/cons(scons).
__global__ void combine_step_kernel(const double* __restrict__ slift, const double* __restrict__ scons,
                                    const double* __restrict__ consval, const double* __restrict__ delta,
                                    const int* __restrict__ naive, double* __restrict__ stot, double* __restrict__ smag,
                                    int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)b * n + i;
    if (naive[b]) {
        // NaiveStepper (stepper.py:44-55) starts its search at alpha0 = 0.5 and the interior test of
        // restricted_step.py:81-84 comes first: for delta < cons(scons) < 2 delta the step is 0.5*scons
        const double half = 0.5 * consval[b];
        const bool interior = half < delta[b];
        stot[o] = scons[o] * (interior ? 0.5 : delta[b] / consval[b]);
        if (i == 0) smag[b] = interior ? half : delta[b];
    } else {
        stot[o] = slift[o] + scons[o];
    }
}

// PES.converged with constraints: fmax over atoms of |P_f g| (pg = projected gradient),
// cmax = |res|; conv = fmax < tol && cmax < 1e-5.
__global__ void __launch_bounds__(CT)
converged_cons_kernel(const double* __restrict__ pg, const double* __restrict__ res, int nr, int n, double fmax_tol,
                      double cmax_tol, double* __restrict__ fmax_out, double* __restrict__ cmax_out,
                      int* __restrict__ conv) {
    const int b = blockIdx.x;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    __shared__ double red[CT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double best = 0.0;
    for (int a = tid; a < n / 3; a += blockDim.x) {
        const double x = pg[(size_t)b * n + 3 * a], y = pg[(size_t)b * n + 3 * a + 1], z = pg[(size_t)b * n + 3 * a + 2];
        best = fmax(best, sqrt(x * x + y * y + z * z));
    }
    best = sb_warp_max(best);
    if (lane == 0) red[warp] = best;
    double r2 = 0.0;
    for (int j = tid; j < nr; j += blockDim.x) r2 = fma(res[(size_t)b * nr + j], res[(size_t)b * nr + j], r2);
    r2 = sb_block_sum(r2, scratch);
    if (tid == 0) {
        for (int w = 1; w < blockDim.x / 32; ++w) best = fmax(best, red[w]);
        const double cm = sqrt(r2);
        fmax_out[b] = best;
        cmax_out[b] = cm;
        conv[b] = (best < fmax_tol) && (cm < cmax_tol);
    }
}

// Position-dependent constraints (sella/peswrapper.py:395-407, 429-438, 467-481): with the rows
// of the constraint Jacobian orthonormalised, drdx = G Uc (G = drdx Uc^T, nc x nc),
//   scons = -Ucons lstsq(drdx Ucons, res)  = -sum_j res_j Mr[j,:],   Mr = G^-T Uc
//   L     = lstsq(drdx^T, g)               =  G^-T (Uc g)
// One CTA per system inverts G in shared memory (nc <= 32) and forms Mr and, if u = Uc g is
// given, the Lagrange multipliers.
struct ConsSolveShared {
    double G[SB_KMAT], Ginv[SB_KMAT];
    int ok;
};

__global__ void __launch_bounds__(CT)
cons_solve_kernel(const double* __restrict__ G_, const double* __restrict__ Uc_, const double* __restrict__ u_,
                  int nc, int n, double* __restrict__ Mr_, double* __restrict__ L_, int* __restrict__ status,
                  const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ unsigned char craw[];
    ConsSolveShared& S = *reinterpret_cast<ConsSolveShared*>(craw);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < nc * nc; i += nt) S.G[(i / nc) * SB_KLD + (i % nc)] = G_[(size_t)b * nc * nc + i];
    __syncthreads();
    if (tid == 0) {
        const bool ok = sbs_invert_serial(S.G, nc, S.Ginv);
        if (!ok && status) atomicOr(&status[b], SB_ST_SINGULAR);
    }
    __syncthreads();
    if (Mr_) {
        const double* Uc = Uc_ + (size_t)b * nc * n;
        double* Mr = Mr_ + (size_t)b * nc * n;
        for (int e = tid; e < n; e += nt) {
            double ucol[SB_KMAX];
            for (int k = 0; k < nc; ++k) ucol[k] = Uc[(size_t)k * n + e];
            for (int j = 0; j < nc; ++j) {
                double acc = 0.0;
                for (int k = 0; k < nc; ++k) acc = fma(S.Ginv[k * SB_KLD + j], ucol[k], acc);     // (G^-T)[j][k]
                Mr[(size_t)j * n + e] = acc;
            }
        }
    }
    if (L_ && u_) {
        for (int j = tid; j < nc; j += nt) {
            double acc = 0.0;
            for (int k = 0; k < nc; ++k) acc = fma(S.Ginv[k * SB_KLD + j], u_[(size_t)b * nc + k], acc);
            L_[(size_t)b * nc + j] = acc;
        }
    }
}

}  // namespace

extern "C" int sb_rect_dots_impl(const double* R, long long rstride, int nr, const double* x, long long xstride,
                                 const double* c, double* out, int n, const int* active, int batch, cudaStream_t st) {
    SB_COUNT(1);
    rect_dots_kernel<<<batch, CT, 0, st>>>(R, (size_t)rstride, nr, x, (size_t)xstride, c, out, n, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_rect_comb_impl(const double* R, long long rstride, int nr, const double* coef, double scale,
                                 const double* base, long long bstride, double beta, double* out, long long ostride,
                                 int n, const int* active, int batch, cudaStream_t st) {
    SB_COUNT(1);
    rect_comb_kernel<<<batch, CT, (size_t)nr * sizeof(double), st>>>(R, (size_t)rstride, nr, coef, scale, base,
                                                                     (size_t)bstride, beta, out, (size_t)ostride, n,
                                                                     active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_scons_measure_impl(const double* scons, const double* delta, int kind, int n, double* scons2,
                                     double* consval, int* naive, int* regular, int batch, cudaStream_t st) {
    SB_COUNT(1);
    scons_measure_kernel<<<batch, CT, 0, st>>>(scons, delta, kind, n, scons2, consval, naive, regular);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_combine_step_impl(const double* slift, const double* scons, const double* consval,
                                    const double* delta, const int* naive, double* stot, double* smag, int n,
                                    const int* active, int batch, cudaStream_t st) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    combine_step_kernel<<<grid, 256, 0, st>>>(slift, scons, consval, delta, naive, stot, smag, n, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_converged_cons_impl(const double* pg, const double* res, int nr, int n, double fmax_tol,
                                      double cmax_tol, double* fmax_out, double* cmax_out, int* conv, int batch,
                                      cudaStream_t st) {
    SB_COUNT(1);
    converged_cons_kernel<<<batch, CT, 0, st>>>(pg, res, nr, n, fmax_tol, cmax_tol, fmax_out, cmax_out, conv);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_cons_solve_impl(const double* G, const double* Uc, const double* u, int nc, int n, double* Mr,
                                  double* L, int* status, const int* active, int batch, cudaStream_t st) {
    const size_t smem = sizeof(ConsSolveShared);
    cudaFuncSetAttribute(cons_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    cons_solve_kernel<<<batch, CT, smem, st>>>(G, Uc, u, nc, n, Mr, L, status, active);
    return SB_LAUNCH_CHECK();
}
