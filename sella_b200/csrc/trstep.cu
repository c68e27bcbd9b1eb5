// Restricted-step ("trust region shift-and-solve") kernels and the per-step
// bookkeeping of the optimiser, one CTA per system.
//
// Restated from the reference (file:line relative to the reference tree):
//   QuasiNewton._stepper_init / get_s         sella/optimize/stepper.py:75-96
//   RationalFunctionOptimization.get_s        sella/optimize/stepper.py:128-157
//   PartitionedRFO                            sella/optimize/stepper.py:163-185
//   BaseRestrictedStep.eval / get_s           sella/optimize/restricted_step.py:72-121
//   TrustRegion.cons                          sella/optimize/restricted_step.py:136-142
//   RestrictedAtomicStep.cons                 sella/optimize/restricted_step.py:172-183
//   PES.kick (rho, secant pair)               sella/peswrapper.py:578-602
//   Sella.step (re-diagonalise test, radius)  sella/optimize/optimize.py:362-378, 412-434
//
// Everything is expressed in the eigenbasis of the (projected) Hessian:
// with Vg = V^T g, the quasi-Newton model step is s(alpha) = -V c,
// c_i = Vg_i / (L_i + alpha*sigma_i).  For the spherical trust region the
// constraint value only needs |c| (V is orthogonal), so the whole alpha search runs
// on length-n vectors in shared memory; the atomic-step variant needs s itself and
// streams V once per alpha evaluation inside the kernel.
#include "common.cuh"

namespace {

constexpr int TR_THREADS = 256;
constexpr int KIND_TR = 0;
constexpr int KIND_RAS = 1;

struct AlphaSearch {
    double alpha, lo, hi, err, val, dval;
    int iter, done, status;
};

// One bracketing/Newton update of alpha, restricted_step.py:90-117.  Returns true
// when another evaluation is required.
__device__ __forceinline__ bool alpha_next(AlphaSearch& s, double delta, double tol, double slope,
                                           bool newton_safe, int maxiter) {
    if (fabs(s.err) <= tol) return false;
    if (nextafter(s.lo, s.hi) >= s.hi) return false;
    if (s.iter >= maxiter) { s.status = SB_ST_TR_NOCONV; return false; }
    if (s.err * slope > 0.0) s.hi = s.alpha; else s.lo = s.alpha;
    const double a1 = s.alpha - s.err / s.dval;
    if (isnan(a1) || a1 <= s.lo || a1 >= s.hi || (s.iter > 4 && !newton_safe)) {
        const double a2 = (s.lo + s.hi) / 2.0;
        if (isinf(a2)) s.alpha = s.alpha + fmax(1.0, 0.5 * s.alpha) * (a2 > 0 ? 1.0 : -1.0);
        else s.alpha = a2;
    } else {
        s.alpha = a1;
    }
    ++s.iter;
    return true;
}

// Quasi-Newton model + spherical trust region, entirely in the eigenbasis.
// In : Vg[b,n] = V^T g, evals[b,n], delta[b], order.
// Out: coef[b,n] with s = V coef (i.e. coef = -c), smag[b], alpha_out[b].
__global__ void __launch_bounds__(TR_THREADS)
qn_tr_kernel(const double* __restrict__ Vg_, const double* __restrict__ evals_, const double* __restrict__ delta_,
             int order, int n, double* __restrict__ coef_, double* __restrict__ smag, double* __restrict__ alpha_out,
             int* __restrict__ status, const int* __restrict__ active, const double* __restrict__ extra2_) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* L = sm;            // signed |lambda|
    double* vg = sm + n;
    double* scratch = vg + n;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) {
        const double l = fabs(evals_[(size_t)b * n + i]);
        L[i] = i < order ? -l : l;
        vg[i] = Vg_[(size_t)b * n + i];
    }
    __syncthreads();
    const double delta = delta_[b];
    // |s_tot|^2 = |s_free|^2 + |scons|^2 (scons is orthogonal to the free space)
    const double extra2 = extra2_ ? extra2_[b] : 0.0;
    AlphaSearch S;
    S.alpha = 0.0; S.lo = 0.0; S.hi = INFINITY; S.iter = 0; S.status = 0;
    const double tol = 1e-10, slope = -1.0;
    bool interior = false;
    for (;;) {
        // eval(alpha): val = |s|, dval = (ds/dalpha . s)/max(val,1e-12)
        double a = 0.0, c = 0.0;
        for (int i = tid; i < n; i += nt) {
            const double sig = i < order ? -1.0 : 1.0;
            const double den = L[i] + S.alpha * sig;
            const double ci = vg[i] / den;
            a = fma(ci, ci, a);
            c = fma(ci, ci / den, c);
        }
        sb_block_sum2(a, c, scratch);
        S.val = sqrt(a + extra2);
        S.dval = -c / fmax(S.val, 1e-12);
        if (S.iter == 0 && S.alpha == 0.0 && S.val < delta) { interior = true; break; }
        S.err = S.val - delta;
        if (!alpha_next(S, delta, tol, slope, true, 1000)) break;
    }
    for (int i = tid; i < n; i += nt) {
        const double sig = i < order ? -1.0 : 1.0;
        coef_[(size_t)b * n + i] = -vg[i] / (L[i] + S.alpha * sig);
    }
    if (tid == 0) {
        smag[b] = interior ? S.val : delta;
        alpha_out[b] = S.alpha;
        if (S.status && status) atomicOr(&status[b], S.status);
    }
}

// Quasi-Newton model + restricted atomic step (max per-atom displacement), Cartesian
// coordinates with Ufree = I.  Streams Vt (rows = eigenvectors) once per alpha.
// Out: s[b,n] (the step itself), smag, alpha.
// The model is given as a list of np poles (np = n for the dense representation): pole i has
// eigenvalue evals[b*np + i], gradient coefficient Vg[b*np + i] and eigenvector row rowmap[b*np + i] of
// Vt (rowmap == NULL: row i); rowmap -1 = the unit vector gperp[b]/gam[b] (complement of the compact
// representation), -2 = padding.
__global__ void __launch_bounds__(TR_THREADS)
qn_ras_kernel(const double* __restrict__ Vg_, const double* __restrict__ evals_, const double* __restrict__ Vt_,
              const double* __restrict__ delta_, int order, int n, double* __restrict__ s_out,
              double* __restrict__ smag, double* __restrict__ alpha_out, int* __restrict__ status,
              const int* __restrict__ active, const double* __restrict__ sadd_, int np, const int* __restrict__ rowmap_,
              const double* __restrict__ gperp_, const double* __restrict__ gam_, long long vstride,
              const double* __restrict__ wmis) {
    // wmis != NULL: MaxInternalStep (restricted_step.py:206-216) -- the measure is max_j |s_j w_j| over the
    // n internal coordinates (weights wmis[n], shared by the batch) instead of the largest atomic displacement
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* L = sm;
    double* vg = L + np;
    double* c1 = vg + np;       // c_i
    double* c2 = c1 + np;       // c_i / den_i
    double* s = c2 + np;        // step
    double* ds = s + n;         // ds/dalpha
    double* scratch = ds + n;
    int* rm = reinterpret_cast<int*>(scratch + SB_SCRATCH_DOUBLES);
    __shared__ double best_val[TR_THREADS / 32];
    __shared__ int best_idx[TR_THREADS / 32];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const double* Vt = Vt_ + (size_t)b * vstride;
    const double* gp = gperp_ ? gperp_ + (size_t)b * n : nullptr;
    const double ginv = (gam_ && gam_[b] > 0.0) ? 1.0 / gam_[b] : 0.0;
    for (int i = tid; i < np; i += nt) {
        const double l = fabs(evals_[(size_t)b * np + i]);
        L[i] = i < order ? -l : l;
        vg[i] = Vg_[(size_t)b * np + i];
        rm[i] = rowmap_ ? rowmap_[(size_t)b * np + i] : i;
    }
    __syncthreads();
    const double delta = delta_[b];
    const int natoms = n / 3;
    AlphaSearch S;
    S.alpha = 0.0; S.lo = 0.0; S.hi = INFINITY; S.iter = 0; S.status = 0;
    bool interior = false;
    for (;;) {
        for (int i = tid; i < np; i += nt) {
            const double sig = i < order ? -1.0 : 1.0;
            const double den = L[i] + S.alpha * sig;
            const double ci = vg[i] / den;
            c1[i] = ci;
            c2[i] = ci / den;
        }
        __syncthreads();
        // s = -V c1, ds = V c2 : column combination of the rows of Vt
        for (int j = tid; j < n; j += nt) {
            double a = 0.0, d = 0.0;
            for (int i = 0; i < np; ++i) {
                const int r = rm[i];
                if (r < -1) continue;
                const double v = r >= 0 ? Vt[(size_t)r * n + j] : gp[j] * ginv;
                a = fma(v, c1[i], a);
                d = fma(v, c2[i], d);
            }
            s[j] = -a + (sadd_ ? sadd_[(size_t)b * n + j] : 0.0);      // + scons (constraint restoring part)
            ds[j] = d;
        }
        __syncthreads();
        // cons: largest atomic displacement (first maximum, as numpy argmax)
        double bv = -1.0; int bi = 0;
        if (wmis) {
            for (int j = tid; j < n; j += nt) {
                const double nr = fabs(s[j] * wmis[j]);
                if (nr > bv) { bv = nr; bi = j; }
            }
        } else
        for (int a = tid; a < natoms; a += nt) {
            const double x = s[3 * a], y = s[3 * a + 1], z = s[3 * a + 2];
            const double nr = sqrt(x * x + y * y + z * z);
            if (nr > bv) { bv = nr; bi = a; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { best_val[warp] = bv; best_idx[warp] = bi; }
        __syncthreads();
        bv = best_val[0]; bi = best_idx[0];
        for (int w = 1; w < nt / 32; ++w)
            if (best_val[w] > bv || (best_val[w] == bv && best_idx[w] < bi)) { bv = best_val[w]; bi = best_idx[w]; }
        S.val = bv;
        if (wmis) S.dval = (s[bi] < 0.0 ? -1.0 : (s[bi] > 0.0 ? 1.0 : 0.0)) * ds[bi] * wmis[bi];
        else
        S.dval = (ds[3 * bi] * s[3 * bi] + ds[3 * bi + 1] * s[3 * bi + 1] + ds[3 * bi + 2] * s[3 * bi + 2]) /
                 fmax(bv, 1e-12);
        __syncthreads();
        if (S.iter == 0 && S.alpha == 0.0 && S.val < delta) { interior = true; break; }
        S.err = S.val - delta;
        if (!alpha_next(S, delta, 1e-10, -1.0, true, 1000)) break;
    }
    for (int j = tid; j < n; j += nt) s_out[(size_t)b * n + j] = s[j];
    if (tid == 0) {
        smag[b] = interior ? S.val : delta;
        alpha_out[b] = S.alpha;
        if (S.status && status) atomicOr(&status[b], S.status);
    }
}

// out[b,0,:] = coef,  out[b,1,:] = |evals| * coef   (coefficients of s and |B| s in the
// eigenbasis: one transposed pass over Vt then yields both vectors)
__global__ void pack_coef_kernel(const double* __restrict__ coef, const double* __restrict__ evals,
                                 double* __restrict__ out, int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c = coef[(size_t)b * n + i];
    out[((size_t)b * 2) * n + i] = c;
    out[((size_t)b * 2 + 1) * n + i] = fabs(evals[(size_t)b * n + i]) * c;
}

// s = in[b,0,:], absBs = in[b,1,:]
__global__ void unpack2_kernel(const double* __restrict__ in, double* __restrict__ s, double* __restrict__ a,
                               int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    s[(size_t)b * n + i] = in[((size_t)b * 2) * n + i];
    a[(size_t)b * n + i] = in[((size_t)b * 2 + 1) * n + i];
}

// ------------------------------------------------------------------ step bookkeeping
// x_new = x + s
__global__ void axpy_kernel(const double* __restrict__ x, const double* __restrict__ s, double* __restrict__ out,
                            int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[(size_t)b * n + i] = x[(size_t)b * n + i] + s[(size_t)b * n + i];
}

struct StepParams {
    double rho_inc, rho_dec, sigma_inc, sigma_dec, delta_min;
    int order, eig, nsteps_per_diag, diag_every_n;   // diag_every_n < 0: never
};

// After the surface has been evaluated at x+s (PES.kick + Sella.step tail):
//   df_pred = g0.s + 1/2 s.Bs ; rho = (f1 - f0)/df_pred (None -> 1 when |df_pred| < 1e-14)
//   dg = g1 - g0  (secant pair (s, dg) for the Hessian update)
//   trust radius update from rho and smag
//   accept the new point: x <- x+s, f <- f1, g <- g1
// The re-diagonalisation decision uses the eigenvalues of the Hessian *before* the
// update (optimize.py:362-378) and is made by ev_decide_kernel below.
__global__ void __launch_bounds__(TR_THREADS)
kick_finish_kernel(double* __restrict__ x, double* __restrict__ f, double* __restrict__ g,
                   const double* __restrict__ xnew, const double* __restrict__ fnew, const double* __restrict__ gnew,
                   const double* __restrict__ s, const double* __restrict__ Bs, const double* __restrict__ smag,
                   double* __restrict__ dg, double* __restrict__ delta, double* __restrict__ rho_out,
                   int* __restrict__ nsteps, StepParams P, int n, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    const int tid = threadIdx.x, nt = blockDim.x;
    double gs = 0.0, sBs = 0.0;
    for (int i = tid; i < n; i += nt) {
        const size_t o = (size_t)b * n + i;
        gs = fma(g[o], s[o], gs);
        sBs = fma(s[o], Bs[o], sBs);
    }
    sb_block_sum2(gs, sBs, scratch);
    const double df_pred = gs + sBs / 2.0;
    const double df = fnew[b] - f[b];
    for (int i = tid; i < n; i += nt) {
        const size_t o = (size_t)b * n + i;
        dg[o] = gnew[o] - g[o];
        g[o] = gnew[o];
        x[o] = xnew[o];
    }
    if (tid == 0) {
        f[b] = fnew[b];
        double rho = 1.0;
        if (fabs(df_pred) >= 1e-14) {
            rho = df / df_pred;
            const double sm_ = smag[b];
            if (rho < 1.0 / P.rho_dec || rho > P.rho_dec) delta[b] = fmax(sm_ * P.sigma_dec, P.delta_min);
            else if (1.0 / P.rho_inc < rho && rho < P.rho_inc) delta[b] = fmax(P.sigma_inc * sm_, delta[b]);
        }
        rho_out[b] = rho;
        nsteps[b] += 1;
    }
}

// optimize.py:362-378.  evals: spectrum of the current (pre-update) projected
// Hessian; has_evals == 0 means "H.evals is None" (uninitialised Hessian).
__global__ void ev_decide_kernel(const double* __restrict__ evals, int n, int has_evals, int* __restrict__ since_diag,
                                 int* __restrict__ ev, StepParams P, int batch, const int* __restrict__ active) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    if (active && !active[b]) { ev[b] = 0; return; }
    const int c = since_diag[b];
    int e = 0;
    if (P.diag_every_n >= 0 && c >= P.diag_every_n) e = 1;
    else if (P.eig && c >= P.nsteps_per_diag) {
        if (!has_evals) e = 1;
        else
            for (int i = 0; i < P.order && i < n; ++i) e |= (evals[(size_t)b * n + i] > 0.0);
    }
    ev[b] = e;
    since_diag[b] = e ? 0 : c + 1;
}

// PES.converged (peswrapper.py:558-568) for Ufree = I: fmax = largest atomic
// |gradient|; conv[b] = fmax < fmax_tol.  Also refreshes the active mask.
__global__ void __launch_bounds__(TR_THREADS)
converged_kernel(const double* __restrict__ g, int n, double fmax_tol, double* __restrict__ fmax_out,
                 int* __restrict__ conv) {
    const int b = blockIdx.x;
    __shared__ double red[TR_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double best = 0.0;
    for (int a = tid; a < n / 3; a += blockDim.x) {
        const double x = g[(size_t)b * n + 3 * a], y = g[(size_t)b * n + 3 * a + 1], z = g[(size_t)b * n + 3 * a + 2];
        best = fmax(best, sqrt(x * x + y * y + z * z));
    }
    best = sb_warp_max(best);
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < blockDim.x / 32; ++w) best = fmax(best, red[w]);
        fmax_out[b] = best;
        conv[b] = best < fmax_tol;
    }
}

}  // namespace

extern "C" int sb_qn_tr_impl(const double* Vg, const double* evals, const double* delta, int order, int n,
                             double* coef, double* smag, double* alpha, int* status, const int* active,
                             const double* extra2, int batch, cudaStream_t st) {
    const size_t smem = (size_t)(2 * n + SB_SCRATCH_DOUBLES) * sizeof(double);
    cudaFuncSetAttribute(qn_tr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    qn_tr_kernel<<<batch, TR_THREADS, smem, st>>>(Vg, evals, delta, order, n, coef, smag, alpha, status, active,
                                                  extra2);
    return SB_LAUNCH_CHECK();
}

static int launch_qn_ras(const double* Vg, const double* evals, const double* Vt, const double* delta,
                         int order, int n, double* s, double* smag, double* alpha, int* status,
                         const int* active, const double* sadd, int np, const int* rowmap, const double* gperp,
                         const double* gam, long long vstride, const double* wmis, int batch, cudaStream_t st) {
    const size_t smem = (size_t)(4 * np + 2 * n + SB_SCRATCH_DOUBLES) * sizeof(double) + (size_t)(np + 2) * sizeof(int);
    if (smem > 200 * 1024) return -2;
    cudaFuncSetAttribute(qn_ras_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    qn_ras_kernel<<<batch, TR_THREADS, smem, st>>>(Vg, evals, Vt, delta, order, n, s, smag, alpha, status, active,
                                                   sadd, np, rowmap, gperp, gam, vstride, wmis);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_qn_ras_c_impl(const double* Vg, const double* evals, const double* Vt, const double* delta,
                                int order, int n, double* s, double* smag, double* alpha, int* status,
                                const int* active, const double* sadd, int np, const int* rowmap, const double* gperp,
                                const double* gam, long long vstride, int batch, cudaStream_t st) {
    return launch_qn_ras(Vg, evals, Vt, delta, order, n, s, smag, alpha, status, active, sadd, np, rowmap, gperp, gam,
                         vstride, nullptr, batch, st);
}

// MaxInternalStep: npole poles whose eigenvector rows (length n = number of internal coordinates) are the
// first npole rows of Wt[b] (vstride doubles per system); weights w[n]
extern "C" int sb_qn_mis_impl(const double* Vg, const double* evals, const double* Wt, const double* delta,
                              int order, int n, double* s, double* smag, double* alpha, int* status,
                              const int* active, const double* sadd, int np, long long vstride, const double* w,
                              int batch, cudaStream_t st) {
    if (!w) return -1;
    return launch_qn_ras(Vg, evals, Wt, delta, order, n, s, smag, alpha, status, active, sadd, np, nullptr, nullptr,
                         nullptr, vstride, w, batch, st);
}

extern "C" int sb_qn_ras_impl(const double* Vg, const double* evals, const double* Vt, const double* delta,
                              int order, int n, double* s, double* smag, double* alpha, int* status,
                              const int* active, const double* sadd, int batch, cudaStream_t st) {
    return sb_qn_ras_c_impl(Vg, evals, Vt, delta, order, n, s, smag, alpha, status, active, sadd, n, nullptr, nullptr,
                            nullptr, (long long)n * n, batch, st);
}

extern "C" int sb_pack_coef_impl(const double* coef, const double* evals, double* out, int n, const int* active,
                                 int batch, cudaStream_t st) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    pack_coef_kernel<<<grid, 256, 0, st>>>(coef, evals, out, n, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_unpack2_impl(const double* in, double* s, double* a, int n, const int* active, int batch,
                               cudaStream_t st) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    unpack2_kernel<<<grid, 256, 0, st>>>(in, s, a, n, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_axpy_impl(const double* x, const double* s, double* out, int n, const int* active, int batch,
                            cudaStream_t st) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    axpy_kernel<<<grid, 256, 0, st>>>(x, s, out, n, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_kick_finish_impl(double* x, double* f, double* g, const double* xnew, const double* fnew,
                                   const double* gnew, const double* s, const double* Bs, const double* smag,
                                   double* dg, double* delta, double* rho, int* nsteps, const double* dpar,
                                   const int* ipar, int n, const int* active, int batch, cudaStream_t st) {
    StepParams P;
    P.rho_inc = dpar[0]; P.rho_dec = dpar[1]; P.sigma_inc = dpar[2]; P.sigma_dec = dpar[3]; P.delta_min = dpar[4];
    P.order = ipar[0]; P.eig = ipar[1]; P.nsteps_per_diag = ipar[2]; P.diag_every_n = ipar[3];
    SB_COUNT(1);
    kick_finish_kernel<<<batch, TR_THREADS, 0, st>>>(x, f, g, xnew, fnew, gnew, s, Bs, smag, dg, delta, rho,
                                                     nsteps, P, n, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_ev_decide_impl(const double* evals, int n, int has_evals, int* since_diag, int* ev,
                                 const double* dpar, const int* ipar, const int* active, int batch,
                                 cudaStream_t st) {
    StepParams P;
    P.rho_inc = dpar[0]; P.rho_dec = dpar[1]; P.sigma_inc = dpar[2]; P.sigma_dec = dpar[3]; P.delta_min = dpar[4];
    P.order = ipar[0]; P.eig = ipar[1]; P.nsteps_per_diag = ipar[2]; P.diag_every_n = ipar[3];
    SB_COUNT(1);
    ev_decide_kernel<<<(batch + 127) / 128, 128, 0, st>>>(evals, n, has_evals, since_diag, ev, P, batch, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_converged_impl(const double* g, int n, double fmax_tol, double* fmax_out, int* conv, int batch,
                                 cudaStream_t st) {
    SB_COUNT(1);
    converged_kernel<<<batch, TR_THREADS, 0, st>>>(g, n, fmax_tol, fmax_out, conv);
    return SB_LAUNCH_CHECK();
}
