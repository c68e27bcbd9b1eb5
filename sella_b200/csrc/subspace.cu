// Davidson / Rayleigh-Ritz subspace kernels (one CTA per system, fp64).
//
// Restated from the reference (file:line relative to the reference tree):
//   mgs / modified_gram_schmidt      sella/utilities/math.pyx:74-140, 143-159
//   symmetrize_Y2                    sella/hessian_update.py:12-24
//   rayleigh_ritz  (one iteration)   sella/eigensolvers.py:56-112
//   expand (all six methods)         sella/eigensolvers.py:115-153
//   NumericalHessian._matvec         sella/linalg.py:39-95 (sign rule :59-73)
//   PES.diag tail (re-Ritz of the operator history)   sella/peswrapper.py:541-551
//
// Vector blocks are vector-major: V[b, kcap, n] (vector j of system b contiguous
// at V + (b*kcap + j)*n).  Subspace sizes live in int32 arrays on the device, so a
// whole batch iterates without host synchronisation; systems that are finished are
// masked by their own state words.
#include "small_dense.cuh"

namespace {

constexpr int SS_THREADS = 256;

// Davidson state words (dav_state[b])
constexpr int DAV_EXPAND = 0;   // a target was chosen, correction vector requested
constexpr int DAV_DONE = 1;     // converged / maxiter / stalled
constexpr int DAV_IDLE = 2;     // not taking part in this diagonalisation

__device__ __forceinline__ double dot_block(const double* __restrict__ a, const double* __restrict__ b, int n,
                                            double* scratch) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(a[i], b[i], acc);
    return sb_block_sum(acc, scratch);
}

// ------------------------------------------------------------------ MGS
// One column: x (shared memory, length n) against `ny` vectors Y and `m` vectors X
// (global memory).  Semantics of math.pyx:97-133.  Returns 1 accepted, 0 dropped,
// -2 iteration limit.
__device__ int mgs_column(double* xs, int n, const double* __restrict__ Y, int ny,
                          const double* __restrict__ Xacc, int m, double eps1, double eps2, int maxiter,
                          double* scratch) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double acc = 0.0;
    for (int i = tid; i < n; i += nt) acc = fma(xs[i], xs[i], acc);
    double nrm = sqrt(sb_block_sum(acc, scratch));
    for (int i = tid; i < n; i += nt) xs[i] /= nrm;
    for (int it = 0; it < maxiter; ++it) {
        double normtot = 1.0;
        bool dropped = false;
        for (int pass = 0; pass < 2 && !dropped; ++pass) {
            const double* basis = pass == 0 ? Y : Xacc;
            const int cnt = pass == 0 ? ny : m;
            for (int j = 0; j < cnt; ++j) {
                const double* y = basis + (size_t)j * n;
                acc = 0.0;
                for (int i = tid; i < n; i += nt) acc = fma(y[i], xs[i], acc);
                const double d = sb_block_sum(acc, scratch);
                acc = 0.0;
                for (int i = tid; i < n; i += nt) {
                    const double v = fma(-d, y[i], xs[i]);
                    xs[i] = v;
                    acc = fma(v, v, acc);
                }
                nrm = sqrt(sb_block_sum(acc, scratch));
                normtot *= nrm;
                if (normtot < eps2) { dropped = true; break; }
                for (int i = tid; i < n; i += nt) xs[i] /= nrm;
            }
            if (normtot < eps2) dropped = true;     // tested after each sweep, even an empty one
        }
        if (dropped) return 0;
        const double gap = 1.0 - normtot;
        if (gap >= 0.0 && gap <= eps1) return 1;
    }
    return -2;
}

// In-place MGS of the nx columns of X (global) against Y (ny, assumed
// orthonormal) and themselves; leftover columns zeroed.  Returns #kept or -2.
__device__ int mgs_block(double* X, int nx, int n, const double* Y, int ny, double eps1, double eps2,
                         int maxiter, double* xs, double* scratch) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int kept = 0;
    for (int c = 0; c < nx; ++c) {
        for (int i = tid; i < n; i += nt) xs[i] = X[(size_t)c * n + i];
        const int r = mgs_column(xs, n, Y, ny, X, kept, eps1, eps2, maxiter, scratch);
        if (r < 0) return r;
        if (r == 1) {
            for (int i = tid; i < n; i += nt) X[(size_t)kept * n + i] = xs[i];
            ++kept;
        }
        __syncthreads();
    }
    for (int c = kept; c < nx; ++c)
        for (int i = tid; i < n; i += nt) X[(size_t)c * n + i] = 0.0;
    __syncthreads();
    return kept;
}

// modified_gram_schmidt(X, Y): Ywork receives the orthonormalised copy of Y.
__global__ void __launch_bounds__(SS_THREADS)
mgs_kernel(double* __restrict__ X, int nx, const double* __restrict__ Y, double* __restrict__ Ywork, int ny,
           int n, double eps1, double eps2, int maxiter, int* __restrict__ nkept, int* __restrict__ status,
           const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* xs = sm;
    double* scratch = sm + n;
    int nyk = 0;
    double* Yb = nullptr;
    if (ny > 0) {
        Yb = Ywork + (size_t)b * ny * n;
        for (int i = threadIdx.x; i < ny * n; i += blockDim.x) Yb[i] = Y[(size_t)b * ny * n + i];
        __syncthreads();
        nyk = mgs_block(Yb, ny, n, nullptr, 0, eps1, eps2, maxiter, xs, scratch);
        if (nyk < 0) nyk = 0;       // the reference slices Yout[:, :ny] without checking
    }
    const int r = mgs_block(X + (size_t)b * nx * n, nx, n, Yb, nyk, eps1, eps2, maxiter, xs, scratch);
    if (threadIdx.x == 0) {
        nkept[b] = r;
        if (r < 0 && status) atomicOr(&status[b], SB_ST_MGS_MAXITER);
    }
}

// ------------------------------------------------------------------ Gram blocks
// G[i][j] = P_i . Q_j for i < ki, j < kj (k x k blocks in shared memory).
__device__ void gram_block(const double* __restrict__ P, const double* __restrict__ Q, int ki, int kj, int n,
                           double* G, bool symmetric) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int npairs = ki * kj;
    for (int pr = warp; pr < npairs; pr += nw) {
        const int i = pr / kj, j = pr % kj;
        if (symmetric && j < i) continue;
        const double* a = P + (size_t)i * n;
        const double* c = Q + (size_t)j * n;
        double acc = 0.0;
        for (int e = lane; e < n; e += 32) acc = fma(a[e], c[e], acc);
        acc = sb_warp_sum(acc);
        if (lane == 0) {
            G[i * SB_KLD + j] = acc;
            if (symmetric) G[j * SB_KLD + i] = acc;
        }
    }
}

// Coefficients of symmetrize_Y2 (serial, thread 0).  STS = S^T S, YTS[i][j] = Y_i.S_j.
// coef[i][j] (j<i): Ytilde_i = Y_i - sum_j coef[i][j] S_j.   dYTS[i][a] = (dY_i).S_a.
// `upto`: only columns i < upto are needed.  T1, T2: k x k scratch.
__device__ bool symmetrize_coeffs_serial(const double* STS, const double* YTS, int k, int upto, double* coef,
                                         double* dYTS, double* T1, double* T2) {
    bool ok = true;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) { coef[i * SB_KLD + j] = 0.0; dYTS[i * SB_KLD + j] = 0.0; }
    for (int i = 1; i < upto; ++i) {
        for (int r = 0; r < i; ++r) {
            for (int c = 0; c < i; ++c) T1[r * SB_KLD + c] = STS[r * SB_KLD + c];
            T2[r * SB_KLD + 0] = YTS[i * SB_KLD + r] - YTS[r * SB_KLD + i] - dYTS[r * SB_KLD + i];
        }
        ok = sbs_solve_serial(T1, i, T2, 1) && ok;
        // dY_i = -S[:, :i] rhs  ->  coef[i][j] = rhs[j]   (Ytilde_i = Y_i - sum rhs_j S_j)
        for (int j = 0; j < i; ++j) coef[i * SB_KLD + j] = T2[j * SB_KLD + 0];
        for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int j = 0; j < i; ++j) acc += STS[a * SB_KLD + j] * T2[j * SB_KLD + 0];
            dYTS[i * SB_KLD + a] = -acc;
        }
    }
    return ok;
}

// In-place rotation of a vector block: X'_a = sum_j R[j][a] X_j  (X' = X R in the
// reference's column convention).
__device__ void rotate_block(double* X, int k, int n, const double* R) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        double col[SB_KMAX];
        for (int j = 0; j < k; ++j) col[j] = X[(size_t)j * n + e];
        for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int j = 0; j < k; ++j) acc = fma(col[j], R[j * SB_KLD + a], acc);
            X[(size_t)a * n + e] = acc;
        }
    }
}

// T = R^T G R  (serial; all k x k)
__device__ void congruence_serial(const double* G, const double* R, int k, double* T, double* tmp) {
    for (int i = 0; i < k; ++i)
        for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int j = 0; j < k; ++j) acc += G[i * SB_KLD + j] * R[j * SB_KLD + a];
            tmp[i * SB_KLD + a] = acc;
        }
    for (int c = 0; c < k; ++c)
        for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int i = 0; i < k; ++i) acc += R[i * SB_KLD + c] * tmp[i * SB_KLD + a];
            T[c * SB_KLD + a] = acc;
        }
}

struct RRShared {
    double STS[SB_KMAT], YTS[SB_KMAT], dYTS[SB_KMAT], coef[SB_KMAT];
    double Asub[SB_KMAT], R[SB_KMAT], T1[SB_KMAT], T2[SB_KMAT];
    double w[SB_KMAX];
    int perm[SB_KMAX];
    double scratch[SB_SCRATCH_DOUBLES];
    int flag, target, nneg;
};

// One Rayleigh-Ritz step of rayleigh_ritz (eigensolvers.py:56-89): Ritz values,
// rotation of (V, AV) to Ritz vectors, residuals, choice of the pair to improve.
// Outputs (for systems that continue): rv[b,0,:] = r_target, rv[b,1,:] = v_target,
// theta[b], dav_state[b] = DAV_EXPAND; otherwise dav_state[b] = DAV_DONE.
__global__ void __launch_bounds__(SS_THREADS)
rr_kernel(double* __restrict__ V_, double* __restrict__ AV_, int kcap, const int* __restrict__ ksz, int n,
          double gamma, int maxiter_eff, double* __restrict__ lams_, double* __restrict__ rv_,
          double* __restrict__ theta_, int* __restrict__ dav_state, int* __restrict__ status) {
    const int b = blockIdx.x;
    if (dav_state[b] != DAV_EXPAND) return;
    extern __shared__ unsigned char rr_raw[];
    RRShared& S = *reinterpret_cast<RRShared*>(rr_raw);
    const int k = ksz[b];
    double* V = V_ + (size_t)b * kcap * n;
    double* AV = AV_ + (size_t)b * kcap * n;
    const int tid = threadIdx.x, warp = tid >> 5;

    gram_block(V, V, k, k, n, S.STS, true);
    gram_block(AV, V, k, k, n, S.YTS, false);
    __syncthreads();
    if (tid == 0) {
        bool ok = symmetrize_coeffs_serial(S.STS, S.YTS, k, k, S.coef, S.dYTS, S.T1, S.T2);
        // Atilde[a][i] = V_a . Ytilde_i = YTS[i][a] + dYTS[i][a]
        for (int a = 0; a < k; ++a)
            for (int i = 0; i < k; ++i) S.Asub[a * SB_KLD + i] = S.YTS[i * SB_KLD + a] + S.dYTS[i * SB_KLD + a];
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) S.T2[i * SB_KLD + j] = S.STS[i * SB_KLD + j];   // metric copy
        S.flag = ok ? 0 : SB_ST_SINGULAR;
    }
    __syncthreads();
    if (warp == 0) {
        const bool ok = sbs_gen_eigh_warp(S.Asub, S.T2, k, S.R, S.w, S.T1, S.perm);
        if (!ok && (tid & 31) == 0) S.flag |= SB_ST_SINGULAR;
    }
    __syncthreads();
    rotate_block(V, k, n, S.R);
    rotate_block(AV, k, n, S.R);
    if (tid == 0) {
        int nneg = 0;
        for (int i = 0; i < k; ++i) nneg += (S.w[i] < 0.0);
        S.nneg = nneg < 1 ? 1 : nneg;
        // Gram blocks of the rotated basis (R^T G R) for the second symmetrisation
        congruence_serial(S.STS, S.R, k, S.Asub, S.T1);
        congruence_serial(S.YTS, S.R, k, S.T2, S.T1);
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) { S.STS[i * SB_KLD + j] = S.Asub[i * SB_KLD + j]; S.YTS[i * SB_KLD + j] = S.T2[i * SB_KLD + j]; }
        symmetrize_coeffs_serial(S.STS, S.YTS, k, S.nneg, S.coef, S.dYTS, S.T1, S.T2);
    }
    for (int i = tid; i < k; i += blockDim.x) lams_[(size_t)b * kcap + i] = S.w[i];
    __syncthreads();
    if (tid == 0 && S.flag && status) atomicOr(&status[b], S.flag);
    if (k >= maxiter_eff) {
        if (tid == 0) dav_state[b] = DAV_DONE;
        return;
    }
    // residuals of the lowest nneg pairs, in order, until one is not converged
    double* r = rv_ + (size_t)b * 2 * n;
    double* vt = r + n;
    const int nneg = S.nneg;
    int target = -1;
    for (int i = 0; i < nneg; ++i) {
        const double th = S.w[i];
        double acc = 0.0;
        for (int e = tid; e < n; e += blockDim.x) {
            double y = AV[(size_t)i * n + e];
            for (int j = 0; j < i; ++j) y = fma(-S.coef[i * SB_KLD + j], V[(size_t)j * n + e], y);
            const double res = y - V[(size_t)i * n + e] * th;
            r[e] = res;
            acc = fma(res, res, acc);
        }
        const double rnorm = sqrt(sb_block_sum(acc, S.scratch));
        if (k == 1 || rnorm >= gamma * fabs(th)) { target = i; break; }
    }
    if (target < 0) {
        if (tid == 0) dav_state[b] = DAV_DONE;
        return;
    }
    for (int e = tid; e < n; e += blockDim.x) vt[e] = V[(size_t)target * n + e];
    if (tid == 0) theta_[b] = S.w[target];
}

// Correction vector in the eigenbasis of the preconditioner P = Q diag(pl) Q^T:
// rvhat[b,0,:] = Q^T r, rvhat[b,1,:] = Q^T v  ->  that[b,:] (then t = Q that).
// method 0: jd0 / jd0_alt   that = -a + eps b,  a = rhat/(pl-theta), b = vhat/(pl-theta),
//                           eps = (vhat.a)/(vhat.b)       (eigensolvers.py:123-139)
// method 1: gd              that = a                       (eigensolvers.py:121-122)
__global__ void __launch_bounds__(SS_THREADS)
jd_coeff_kernel(const double* __restrict__ rvhat, const double* __restrict__ pl, const double* __restrict__ theta,
                double* __restrict__ that, int n, int method, const int* __restrict__ dav_state) {
    const int b = blockIdx.x;
    if (dav_state[b] != DAV_EXPAND) return;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    const double* rh = rvhat + (size_t)b * 2 * n;
    const double* vh = rh + n;
    const double* lam = pl + (size_t)b * n;
    const double th = theta[b];
    double va = 0.0, vb = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = lam[i] - th;
        va = fma(vh[i], rh[i] / d, va);
        vb = fma(vh[i], vh[i] / d, vb);
    }
    sb_block_sum2(va, vb, scratch);
    double eps = va / vb;
    if (method == 0 && fabs(vb) < 1e-12) eps = 0.0;   // jd0_alt's guard; never reached by jd0 proper
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = lam[i] - th;
        const double a = rh[i] / d, bb = vh[i] / d;
        that[(size_t)b * n + i] = method == 1 ? a : (eps * bb - a);
    }
}

struct MjdShared {
    double G[SB_KMAT], rhs[SB_KMAT];
    int ok;
};

// 'mjd0' / 'mjd0_alt' (eigensolvers.py:140-151): the correction is kept orthogonal to ALL k
// Ritz vectors.  In the eigenbasis of P:  that = (-rhat + Vhat eps)/(pl - theta) with
// (Vhat^T D Vhat) eps = Vhat^T D rhat,  D = diag(1/(pl - theta)),  Vhat[b,j,:] = Q^T V_j.
__global__ void __launch_bounds__(SS_THREADS)
mjd_coeff_kernel(const double* __restrict__ Vhat_, int kcap, const int* __restrict__ ksz,
                 const double* __restrict__ rvhat, const double* __restrict__ pl, const double* __restrict__ theta,
                 double* __restrict__ that, int n, const int* __restrict__ dav_state, int* __restrict__ status) {
    const int b = blockIdx.x;
    if (dav_state[b] != DAV_EXPAND) return;
    extern __shared__ unsigned char mjd_raw[];
    MjdShared& M = *reinterpret_cast<MjdShared*>(mjd_raw);
    const int k = ksz[b];
    const double* Vh = Vhat_ + (size_t)b * kcap * n;
    const double* rh = rvhat + (size_t)b * 2 * n;
    const double* lam = pl + (size_t)b * n;
    const double th = theta[b];
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int pr = warp; pr < k * (k + 1); pr += nw) {
        const int i = pr / (k + 1), j = pr % (k + 1);
        if (j < i) continue;
        const double* a = Vh + (size_t)i * n;
        const double* c = j < k ? Vh + (size_t)j * n : rh;
        double acc = 0.0;
        for (int e = lane; e < n; e += 32) acc = fma(a[e], c[e] / (lam[e] - th), acc);
        acc = sb_warp_sum(acc);
        if (lane == 0) {
            if (j < k) { M.G[i * SB_KLD + j] = acc; M.G[j * SB_KLD + i] = acc; }
            else M.rhs[i * SB_KLD] = acc;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const bool ok = sbs_solve_serial(M.G, k, M.rhs, 1);
        if (!ok && status) atomicOr(&status[b], SB_ST_SINGULAR);
    }
    __syncthreads();
    for (int e = tid; e < n; e += nt) {
        double acc = -rh[e];
        for (int i = 0; i < k; ++i) acc = fma(M.rhs[i * SB_KLD], Vh[(size_t)i * n + e], acc);
        that[(size_t)b * n + e] = acc / (lam[e] - th);
    }
}

struct ExpShared {
    double scratch[SB_SCRATCH_DOUBLES];
    double tmp[SB_KMAX];
};

// Finish one expansion (eigensolvers.py:90-112) and prepare the finite-difference
// displacement of the new direction (linalg.py:39-81):
//   t <- t/|t| ; Lanczos fallback if |t - V V^T t| < 1e-2 ; t <- mgs(t, V) ;
//   V[k] <- t ; xdisp = x0 + eta * t / (sign*|t|), signnorm[b] = sign*|t|.
// p_identity != 0: the preconditioner is the identity (uninitialised Hessian) and
// the jd0 correction has the closed form  t = -(r - v (v.r)/(v.v)) / (1 - theta).
// `tin` [b,n] holds the correction vector otherwise.  Ywork: [b,kcap,n] scratch.
__global__ void __launch_bounds__(SS_THREADS)
expand_finish_kernel(const double* __restrict__ tin, const double* __restrict__ rv_, const double* __restrict__ theta_,
                     double* __restrict__ V_, double* __restrict__ Ywork_, int kcap, const int* __restrict__ ksz,
                     int n, int p_identity, int lanczos, double* __restrict__ vnew_, int* __restrict__ dav_state,
                     int* __restrict__ status) {
    const int b = blockIdx.x;
    if (dav_state[b] != DAV_EXPAND) return;
    extern __shared__ double sm[];
    double* xs = sm;                                  // n
    ExpShared& S = *reinterpret_cast<ExpShared*>(sm + n);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int k = ksz[b];
    double* V = V_ + (size_t)b * kcap * n;
    double* Yw = Ywork_ + (size_t)b * kcap * n;
    const double* r = rv_ + (size_t)b * 2 * n;
    const double* v = r + n;

    if (k >= kcap) {                                  // compiled capacity reached
        if (tid == 0) { dav_state[b] = DAV_DONE; if (status) atomicOr(&status[b], SB_ST_DAVIDSON_CAP); }
        return;
    }
    if (lanczos) {
        for (int i = tid; i < n; i += nt) xs[i] = r[i];
    } else if (p_identity) {
        double a = 0.0, c = 0.0;
        for (int i = tid; i < n; i += nt) { a = fma(v[i], r[i], a); c = fma(v[i], v[i], c); }
        sb_block_sum2(a, c, S.scratch);
        const double eps = a / c, den = 1.0 - theta_[b];
        for (int i = tid; i < n; i += nt) xs[i] = -(r[i] - eps * v[i]) / den;
    } else {
        for (int i = tid; i < n; i += nt) xs[i] = tin[(size_t)b * n + i];
    }
    // t /= |t|
    double acc = 0.0;
    for (int i = tid; i < n; i += nt) acc = fma(xs[i], xs[i], acc);
    double nrm = sqrt(sb_block_sum(acc, S.scratch));
    for (int i = tid; i < n; i += nt) xs[i] /= nrm;
    // |t - V (V^T t)| < 1e-2  ->  Lanczos direction instead
    {
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        __syncthreads();
        for (int j = warp; j < k; j += nw) {
            double d = 0.0;
            for (int e = lane; e < n; e += 32) d = fma(V[(size_t)j * n + e], xs[e], d);
            d = sb_warp_sum(d);
            if (lane == 0) S.tmp[j] = d;
        }
        __syncthreads();
        acc = 0.0;
        for (int i = tid; i < n; i += nt) {
            double p = 0.0;
            for (int j = 0; j < k; ++j) p = fma(V[(size_t)j * n + i], S.tmp[j], p);
            const double q = xs[i] - p;
            acc = fma(q, q, acc);
        }
        const double out = sqrt(sb_block_sum(acc, S.scratch));
        if (out < 1e-2) {
            acc = 0.0;
            for (int i = tid; i < n; i += nt) acc = fma(r[i], r[i], acc);
            nrm = sqrt(sb_block_sum(acc, S.scratch));
            for (int i = tid; i < n; i += nt) xs[i] = r[i] / nrm;
        }
    }
    // modified_gram_schmidt(t, V): the wrapper first re-orthonormalises a copy of V
    for (int i = tid; i < k * n; i += nt) Yw[i] = V[i];
    __syncthreads();
    double* col = Yw + (size_t)k * n;                 // spare row of the scratch block (k < kcap)
    int nyk = mgs_block(Yw, k, n, nullptr, 0, 1e-15, 1e-6, 100, col, S.scratch);
    if (nyk < 0) nyk = 0;
    int res = mgs_column(xs, n, Yw, nyk, nullptr, 0, 1e-15, 1e-6, 100, S.scratch);
    if (res == 0) {
        // Davidson failed to find a new direction: try the residual (eigensolvers.py:99-104)
        __syncthreads();
        for (int i = tid; i < n; i += nt) xs[i] = r[i];
        __syncthreads();
        res = mgs_column(xs, n, Yw, nyk, nullptr, 0, 1e-15, 1e-6, 100, S.scratch);
    }
    if (res != 1) {
        if (tid == 0) {
            dav_state[b] = DAV_DONE;
            if (status) atomicOr(&status[b], res < 0 ? SB_ST_MGS_MAXITER : SB_ST_DAVIDSON_STALL);
        }
        return;
    }
    for (int i = tid; i < n; i += nt) {
        V[(size_t)k * n + i] = xs[i];
        vnew_[(size_t)b * n + i] = xs[i];
    }
}

// Displacement for the finite-difference Hessian-vector product of direction
// vfull (full space): canonical sign rule linalg.py:59-73, then
// xdisp = x0 + eta * v/(sign*|v|); signnorm = sign*|v| (0 marks a null vector).
__global__ void __launch_bounds__(SS_THREADS)
hvp_prepare_kernel(const double* __restrict__ vfull_, size_t vstride, const double* __restrict__ x0_,
                   const double* __restrict__ g0_, double eta, double* __restrict__ xdisp_,
                   double* __restrict__ signnorm, int n, const int* __restrict__ mask, int maskval) {
    const int b = blockIdx.x;
    if (mask && mask[b] != maskval) return;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    __shared__ int first_idx;
    const double* v = vfull_ + (size_t)b * vstride;
    const double* x0 = x0_ + (size_t)b * n;
    const double* g0 = g0_ + (size_t)b * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    double vg = 0.0, vx = 0.0, vv = 0.0;
    for (int i = tid; i < n; i += nt) {
        vg = fma(v[i], g0[i], vg);
        vx = fma(v[i], x0[i], vx);
        vv = fma(v[i], v[i], vv);
    }
    sb_block_sum2(vg, vx, scratch);
    vv = sb_block_sum(vv, scratch);
    if (tid == 0) first_idx = n;
    __syncthreads();
    double sign = 1.0;
    if (fabs(vg) > 1e-4) sign = vg < 0.0 ? 1.0 : -1.0;
    else if (fabs(vx) > 1e-4) sign = vx < 0.0 ? 1.0 : -1.0;
    else {
        int mine = n;
        for (int i = tid; i < n; i += nt)
            if (fabs(v[i]) > 1e-4) { mine = i; break; }
        atomicMin(&first_idx, mine);
        __syncthreads();
        const int fi = first_idx;
        if (fi < n) sign = v[fi] > 0.0 ? 1.0 : -1.0;
    }
    const double vnorm = sqrt(vv);
    double* xd = xdisp_ + (size_t)b * n;
    if (vnorm < 1e-12) {
        for (int i = tid; i < n; i += nt) xd[i] = x0[i];
        if (tid == 0) signnorm[b] = 0.0;
        return;
    }
    const double sn = vnorm * sign;
    for (int i = tid; i < n; i += nt) xd[i] = x0[i] + eta * v[i] / sn;
    if (tid == 0) signnorm[b] = sn;
}

// Av = signnorm * (gplus - g0)/eta ; record (v, Av) in the operator history and as
// the new column of AV; bump the counters.
__global__ void __launch_bounds__(SS_THREADS)
hvp_finish_kernel(const double* __restrict__ vfull_, size_t vstride, const double* __restrict__ gplus_, const double* __restrict__ g0_,
                  const double* __restrict__ signnorm, double eta, double* __restrict__ AV_,
                  double* __restrict__ Vs_, double* __restrict__ AVs_, int kcap, int* __restrict__ ksz,
                  int* __restrict__ nhist, int n, const int* __restrict__ mask, int maskval) {
    const int b = blockIdx.x;
    if (mask && mask[b] != maskval) return;
    const int k = ksz[b], h = nhist[b];
    const double sn = signnorm[b];
    const double* gp = gplus_ + (size_t)b * n;
    const double* g0 = g0_ + (size_t)b * n;
    const double* v = vfull_ + (size_t)b * vstride;
    double* av = AV_ + ((size_t)b * kcap + k) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = sn == 0.0 ? 0.0 : sn * (gp[i] - g0[i]) / eta;
        av[i] = a;
        if (sn != 0.0 && h < kcap) {
            Vs_[((size_t)b * kcap + h) * n + i] = v[i];
            AVs_[((size_t)b * kcap + h) * n + i] = a;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ksz[b] = k + 1;
        if (sn != 0.0 && h < kcap) nhist[b] = h + 1;
    }
}

// Start of a diagonalisation (eigensolvers.py:44-52, peswrapper.py:522-529).
//   mode 0: V[0] = mgs(v0)            (first diagonalisation: v0 = Ufree^T g)
//   mode 1: V[:nneg] = mgs(rows of Pvt with negative eigenvalue, at least one)
// Sets ksz = 0 (columns are appended by the HVP loop), ninit = #start vectors,
// nhist = 0, dav_state = EXPAND for participating systems (part[b] != 0), IDLE else.
__global__ void __launch_bounds__(SS_THREADS)
davidson_init_kernel(const double* __restrict__ v0_, const double* __restrict__ pl, const double* __restrict__ Pvt,
                     int mode, double* __restrict__ V_, int kcap, int n, int* __restrict__ ksz,
                     int* __restrict__ ninit, int* __restrict__ nhist, int* __restrict__ dav_state,
                     int* __restrict__ status, const int* __restrict__ part, const int* __restrict__ mrows,
                     const double* __restrict__ lam0, const double* __restrict__ gperp, long long estride,
                     long long vstride) {
    // mrows != NULL: compact representation -- only the first mrows[b] entries of pl / rows of Pvt are
    // explicit eigenpairs, every other eigenvalue of the preconditioner equals lam0[b]
    const int b = blockIdx.x;
    if (part && !part[b]) {
        if (threadIdx.x == 0) dav_state[b] = DAV_IDLE;
        return;
    }
    extern __shared__ double sm[];
    double* xs = sm;
    double* scratch = sm + n;
    double* V = V_ + (size_t)b * kcap * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    int nstart = 1;
    if (mode == 0) {
        for (int i = tid; i < n; i += nt) V[i] = v0_[(size_t)b * n + i];
    } else {
        const int m = mrows ? mrows[b] : n;
        // rows with a negative eigenvalue, most negative first (at most kcap); none: the lowest one.
        // Dense representation: pl ascending, so these are the leading rows; compact: any storage order
        __shared__ int sel[32];
        __shared__ int nsel_s, lowest_s;
        if (tid == 0) {
            int cnt = 0, lowest = 0;
            if (!mrows) {
                for (int i = 0; i < m && i < kcap; ++i) if (pl[(size_t)b * estride + i] < 0.0) sel[cnt++] = i;
            } else {
                for (int i = 0; i < m; ++i) {
                    const double v = pl[(size_t)b * estride + i];
                    if (v < pl[(size_t)b * estride + lowest]) lowest = i;
                    if (v < 0.0) {
                        int p = cnt < kcap ? cnt : kcap - 1;            // insertion, ascending, capacity kcap
                        if (cnt == kcap && !(v < pl[(size_t)b * estride + sel[kcap - 1]])) continue;
                        while (p > 0 && pl[(size_t)b * estride + sel[p - 1]] > v) { sel[p] = sel[p - 1]; --p; }
                        sel[p] = i;
                        if (cnt < kcap) ++cnt;
                    }
                }
                if (m < n && lam0[b] < 0.0) cnt = 0;    // (never: lam0 is a geometric mean of |Ritz values|)
            }
            if (cnt == 0) sel[0] = lowest;
            nsel_s = cnt; lowest_s = lowest;
        }
        __syncthreads();
        const int nneg = nsel_s;
        nstart = nneg < 1 ? 1 : nneg;
        // no negative eigenvalue: the lowest eigenvector.  In the compact representation that is an explicit
        // row unless the complement's lam0 lies below every explicit eigenvalue; any unit vector of the
        // complement is then "the" lowest eigenvector (eigh of a degenerate matrix returns an arbitrary
        // one): the gradient's component in it
        const bool from_complement = mrows && nneg == 0 && m < n &&
                                     (m == 0 || pl[(size_t)b * estride + lowest_s] > lam0[b]);
        if (from_complement) {
            for (int i = tid; i < n; i += nt) V[i] = gperp[(size_t)b * n + i];
        } else {
            for (int i = tid; i < nstart * n; i += nt) V[i] = Pvt[(size_t)b * vstride + (size_t)sel[i / n] * n + (i % n)];
        }
    }
    __syncthreads();
    const int kept = mgs_block(V, nstart, n, nullptr, 0, 1e-15, 1e-6, 100, xs, scratch);
    if (tid == 0) {
        ksz[b] = 0;
        nhist[b] = 0;
        ninit[b] = kept < 0 ? 0 : kept;
        dav_state[b] = (kept > 0) ? DAV_EXPAND : DAV_DONE;
        if (kept < 0 && status) atomicOr(&status[b], SB_ST_MGS_MAXITER);
    }
}

// PES.diag tail (peswrapper.py:541-551): Ritz-rotate the operator history
//   Atilde = Vs^T symmetrize_Y(Vs, AVs, 2);  X = eigvecs(Atilde);  S = Vs X, Y = AVs X
// in place.  nvec_out[b] = number of history vectors (0 for idle systems).
__global__ void __launch_bounds__(SS_THREADS)
history_ritz_kernel(double* __restrict__ Vs_, double* __restrict__ AVs_, int kcap, const int* __restrict__ nhist,
                    int n, int* __restrict__ nvec_out, const int* __restrict__ dav_state,
                    int* __restrict__ status, const double* __restrict__ HcVs_) {
    const int b = blockIdx.x;
    if (dav_state[b] == DAV_IDLE) {
        if (threadIdx.x == 0) nvec_out[b] = 0;
        return;
    }
    extern __shared__ unsigned char rr_raw[];
    RRShared& S = *reinterpret_cast<RRShared*>(rr_raw);
    const int k = nhist[b];
    if (threadIdx.x == 0) nvec_out[b] = k;
    if (k == 0) return;
    double* Vs = Vs_ + (size_t)b * kcap * n;
    double* AVs = AVs_ + (size_t)b * kcap * n;
    const int tid = threadIdx.x, warp = tid >> 5;
    gram_block(Vs, Vs, k, k, n, S.STS, true);
    gram_block(AVs, Vs, k, k, n, S.YTS, false);
    // constraint-Hessian term of peswrapper.py:546: Atilde -= Vs^T Hc Vs  (R[i][a] = (Hc Vs_i).Vs_a)
    if (HcVs_) gram_block(HcVs_ + (size_t)b * kcap * n, Vs, k, k, n, S.R, false);
    __syncthreads();
    if (tid == 0) {
        bool ok = symmetrize_coeffs_serial(S.STS, S.YTS, k, k, S.coef, S.dYTS, S.T1, S.T2);
        for (int a = 0; a < k; ++a)
            for (int i = 0; i < k; ++i)
                S.Asub[a * SB_KLD + i] = S.YTS[i * SB_KLD + a] + S.dYTS[i * SB_KLD + a] - (HcVs_ ? S.R[i * SB_KLD + a] : 0.0);
        // scipy eigh reads the lower triangle
        for (int i = 0; i < k; ++i)
            for (int j = i + 1; j < k; ++j) S.Asub[i * SB_KLD + j] = S.Asub[j * SB_KLD + i];
        if (!ok && status) atomicOr(&status[b], SB_ST_SINGULAR);
    }
    __syncthreads();
    if (warp == 0) sbs_jacobi_warp(S.Asub, k, S.R, S.w, S.perm);
    __syncthreads();
    rotate_block(Vs, k, n, S.R);
    rotate_block(AVs, k, n, S.R);
}

}  // namespace

// ------------------------------------------------------------------ launchers
extern "C" int sb_mgs_impl(double* X, int nx, const double* Y, double* Ywork, int ny, int n, double eps1,
                           double eps2, int maxiter, int* nkept, int* status, const int* active, int batch,
                           cudaStream_t st) {
    const size_t smem = (size_t)(n + SB_SCRATCH_DOUBLES) * sizeof(double);
    cudaFuncSetAttribute(mgs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    mgs_kernel<<<batch, SS_THREADS, smem, st>>>(X, nx, Y, Ywork, ny, n, eps1, eps2, maxiter, nkept, status,
                                                active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_davidson_init_c_impl(const double* v0, const double* pl, const double* Pvt, int mode, double* V,
                                       int kcap, int n, int* ksz, int* ninit, int* nhist, int* dav_state,
                                       int* status, const int* part, const int* mrows, const double* lam0,
                                       const double* gperp, long long estride, long long vstride, int batch,
                                       cudaStream_t st) {
    const size_t smem = (size_t)(n + SB_SCRATCH_DOUBLES) * sizeof(double);
    cudaFuncSetAttribute(davidson_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    davidson_init_kernel<<<batch, SS_THREADS, smem, st>>>(v0, pl, Pvt, mode, V, kcap, n, ksz, ninit, nhist,
                                                          dav_state, status, part, mrows, lam0, gperp, estride,
                                                          vstride);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_davidson_init_impl(const double* v0, const double* pl, const double* Pvt, int mode, double* V,
                                     int kcap, int n, int* ksz, int* ninit, int* nhist, int* dav_state,
                                     int* status, const int* part, int batch, cudaStream_t st) {
    return sb_davidson_init_c_impl(v0, pl, Pvt, mode, V, kcap, n, ksz, ninit, nhist, dav_state, status, part, nullptr,
                                   nullptr, nullptr, (long long)n, (long long)n * n, batch, st);
}

extern "C" int sb_davidson_rr_impl(double* V, double* AV, int kcap, const int* ksz, int n, double gamma,
                                   int maxiter_eff, double* lams, double* rv, double* theta, int* dav_state,
                                   int* status, int batch, cudaStream_t st) {
    const size_t smem = sizeof(RRShared);
    cudaFuncSetAttribute(rr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    rr_kernel<<<batch, SS_THREADS, smem, st>>>(V, AV, kcap, ksz, n, gamma, maxiter_eff, lams, rv, theta,
                                               dav_state, status);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_davidson_jd_coeff_impl(const double* rvhat, const double* pl, const double* theta, double* that,
                                         int n, int method, const int* dav_state, int batch, cudaStream_t st) {
    SB_COUNT(1);
    jd_coeff_kernel<<<batch, SS_THREADS, 0, st>>>(rvhat, pl, theta, that, n, method, dav_state);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_davidson_mjd_coeff_impl(const double* Vhat, int kcap, const int* ksz, const double* rvhat,
                                          const double* pl, const double* theta, double* that, int n,
                                          const int* dav_state, int* status, int batch, cudaStream_t st) {
    const size_t smem = sizeof(MjdShared);
    cudaFuncSetAttribute(mjd_coeff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    mjd_coeff_kernel<<<batch, SS_THREADS, smem, st>>>(Vhat, kcap, ksz, rvhat, pl, theta, that, n, dav_state, status);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_davidson_expand_impl(const double* tin, const double* rv, const double* theta, double* V,
                                       double* Ywork, int kcap, const int* ksz, int n, int p_identity,
                                       int lanczos, double* vnew, int* dav_state, int* status, int batch,
                                       cudaStream_t st) {
    const size_t smem = (size_t)n * sizeof(double) + sizeof(ExpShared);
    cudaFuncSetAttribute(expand_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    expand_finish_kernel<<<batch, SS_THREADS, smem, st>>>(tin, rv, theta, V, Ywork, kcap, ksz, n, p_identity,
                                                          lanczos, vnew, dav_state, status);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_hvp_prepare_impl(const double* vfull, long long vstride, const double* x0, const double* g0,
                                   double eta, double* xdisp, double* signnorm, int n, const int* mask,
                                   int maskval, int batch, cudaStream_t st) {
    SB_COUNT(1);
    hvp_prepare_kernel<<<batch, SS_THREADS, 0, st>>>(vfull, (size_t)vstride, x0, g0, eta, xdisp, signnorm, n, mask,
                                                     maskval);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_hvp_finish_impl(const double* vfull, long long vstride, const double* gplus, const double* g0,
                                  const double* signnorm, double eta, double* AV, double* Vs, double* AVs,
                                  int kcap, int* ksz, int* nhist, int n, const int* mask, int maskval, int batch,
                                  cudaStream_t st) {
    SB_COUNT(1);
    hvp_finish_kernel<<<batch, SS_THREADS, 0, st>>>(vfull, (size_t)vstride, gplus, g0, signnorm, eta, AV, Vs, AVs,
                                                    kcap, ksz, nhist, n, mask, maskval);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_history_ritz_impl(double* Vs, double* AVs, int kcap, const int* nhist, int n, int* nvec_out,
                                    const int* dav_state, int* status, const double* HcVs, int batch,
                                    cudaStream_t st) {
    const size_t smem = sizeof(RRShared);
    cudaFuncSetAttribute(history_ritz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    history_ritz_kernel<<<batch, SS_THREADS, smem, st>>>(Vs, AVs, kcap, nhist, n, nvec_out, dav_state, status, HcVs);
    return SB_LAUNCH_CHECK();
}
