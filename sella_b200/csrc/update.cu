// Symmetric multi-secant quasi-Newton updates of the dense approximate Hessian.
//
// Restated from the reference (file:line relative to the reference tree):
//   update_H dispatcher, first-update scaling, final (B + B^T)/2   sella/hessian_update.py:40-111
//   TS-BFGS  :118-125,  PSB  :128-132,  Greenstadt  :147-152
//   symmetrize_Y (symm=2)                                           sella/hessian_update.py:12-37
//   ApproximateHessian.update (first update on the Cartesian block) sella/linalg.py:274-304
//
// Per system, with S, Y of k columns (k = 1 for a step update, k = #history for the
// update after a diagonalisation):
//   prep  : Ytilde = symmetrize_Y(S, Y, 2);  first update: lam0 = geometric mean of
//           |eig(S^T Ytilde)|  ->  B = lam0 * I                     (kernel 1 + fill)
//   (hv)  : BS = B S,  VtS = Vt S,  absBS = Vt^T (|lam| * VtS)       (hv.cu)
//   mid   : J = Ytilde - BS,  U per method,  W = U (J^T S)          (kernel 2)
//   apply : B += sum_a (U_a J_a^T + J_a U_a^T) - 1/2 (W_a U_a^T + U_a W_a^T)   (kernel 3)
// The last line is (Delta + Delta^T)/2 with Delta = U J^T + J U^T - U (J^T S) U^T,
// i.e. the reference's `Bplus = B + Delta; Bplus = (Bplus + Bplus.T)/2` for a
// symmetric B; it is evaluated so that the stored B stays bitwise symmetric.
// The apply kernel is a pure streaming read-modify-write of B (16 bytes / element).
#include "small_dense.cuh"

namespace {

constexpr int UP_THREADS = 256;

__device__ void gram_block_u(const double* __restrict__ P, const double* __restrict__ Q, int ki, int kj, int n,
                             double* G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int pr = warp; pr < ki * kj; pr += nw) {
        const int i = pr / kj, j = pr % kj;
        const double* a = P + (size_t)i * n;
        const double* c = Q + (size_t)j * n;
        double acc = 0.0;
        for (int e = lane; e < n; e += 32) acc = fma(a[e], c[e], acc);
        acc = sb_warp_sum(acc);
        if (lane == 0) G[i * SB_KLD + j] = acc;
    }
}

struct PrepShared {
    double STS[SB_KMAT], YTS[SB_KMAT], dYTS[SB_KMAT], coef[SB_KMAT], T1[SB_KMAT], T2[SB_KMAT];
    double w[SB_KMAX];
    int perm[SB_KMAX];
    int ok;
};

// Ytil = symmetrize_Y(S, Y, 2); skip[b] = 1 when the update is a no-op
// (k == 0, or a single step shorter than 1e-8: hessian_update.py:49-52);
// first != 0: lam0[b] = exp(mean(log(max(|eig(S^T Ytil)|, 1e-12)))).
__global__ void __launch_bounds__(UP_THREADS)
update_prep_kernel(const double* __restrict__ S_, const double* __restrict__ Y_, double* __restrict__ Ytil_,
                   int kcap, const int* __restrict__ kvec, int n, int ncart, int first, int symm,
                   double* __restrict__ lam0,
                   int* __restrict__ skip, int* __restrict__ status, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) { if (threadIdx.x == 0) skip[b] = 1; return; }
    extern __shared__ unsigned char raw[];
    PrepShared& P = *reinterpret_cast<PrepShared*>(raw);
    const int k = kvec ? kvec[b] : 1;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5;
    const double* S = S_ + (size_t)b * kcap * n;
    const double* Y = Y_ + (size_t)b * kcap * n;
    double* Yt = Ytil_ + (size_t)b * kcap * n;
    if (k == 0) { if (tid == 0) skip[b] = 1; return; }
    // the first update only sees the Cartesian block (linalg.py:282-287)
    const int nn = first ? ncart : n;
    gram_block_u(S, S, k, k, nn, P.STS);
    gram_block_u(Y, S, k, k, nn, P.YTS);
    __syncthreads();
    if (k == 1 && sqrt(P.STS[0]) < 1e-8) { if (tid == 0) skip[b] = 1; return; }
    if (tid == 0) {
        skip[b] = 0;
        // same routine as in subspace.cu, restated inline on the k x k blocks
        bool ok = true;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) { P.coef[i * SB_KLD + j] = 0.0; P.dYTS[i * SB_KLD + j] = 0.0; }
        if (symm == 2) {
            for (int i = 1; i < k; ++i) {
                for (int r = 0; r < i; ++r) {
                    for (int c = 0; c < i; ++c) P.T1[r * SB_KLD + c] = P.STS[r * SB_KLD + c];
                    P.T2[r * SB_KLD] = P.YTS[i * SB_KLD + r] - P.YTS[r * SB_KLD + i] - P.dYTS[r * SB_KLD + i];
                }
                ok = sbs_solve_serial(P.T1, i, P.T2, 1) && ok;
                for (int j = 0; j < i; ++j) P.coef[i * SB_KLD + j] = P.T2[j * SB_KLD];
                for (int a = 0; a < k; ++a) {
                    double acc = 0.0;
                    for (int j = 0; j < i; ++j) acc += P.STS[a * SB_KLD + j] * P.T2[j * SB_KLD];
                    P.dYTS[i * SB_KLD + a] = -acc;
                }
            }
        } else if (k > 1) {
            // symm 0: Ytil = Y + S X, X = (S^T S)^-1 tril(S^T Y - Y^T S, -1)^T
            // symm 1: Ytil = Y + Y X, X = (S^T Y)^-1 (same right-hand side)   hessian_update.py:30-33
            for (int a = 0; a < k; ++a)
                for (int c = 0; c < k; ++c) {
                    P.T1[a * SB_KLD + c] = symm == 0 ? P.STS[a * SB_KLD + c] : P.YTS[c * SB_KLD + a];
                    P.T2[a * SB_KLD + c] = c > a ? P.YTS[a * SB_KLD + c] - P.YTS[c * SB_KLD + a] : 0.0;
                }
            ok = sbs_solve_serial(P.T1, k, P.T2, k);                 // T2 <- X
            const double* base = symm == 0 ? P.STS : P.YTS;           // (basis vector j) . S_a
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) P.coef[i * SB_KLD + j] = -P.T2[j * SB_KLD + i];
            for (int i = 0; i < k; ++i)
                for (int a = 0; a < k; ++a) {
                    double acc = 0.0;
                    for (int j = 0; j < k; ++j) acc += P.T2[j * SB_KLD + i] * base[j * SB_KLD + a];
                    P.dYTS[i * SB_KLD + a] = acc;
                }
        }
        P.ok = ok ? 1 : 0;
        if (!ok && status) atomicOr(&status[b], SB_ST_SINGULAR);
    }
    __syncthreads();
    for (int i = 0; i < k; ++i)
        for (int e = tid; e < n; e += nt) {
            double y = Y[(size_t)i * n + e];
            if (e < nn) {
                const int jmax = symm == 2 ? i : k;
                const double* basis = symm == 1 ? Y : S;
                for (int j = 0; j < jmax; ++j) y = fma(-P.coef[i * SB_KLD + j], basis[(size_t)j * n + e], y);
            }
            Yt[(size_t)i * n + e] = y;
        }
    if (first) {
        // T1 = S^T Ytilde (k x k), symmetric up to round-off; eigh reads the lower triangle
        if (tid == 0) {
            for (int a = 0; a < k; ++a)
                for (int i = 0; i < k; ++i) P.T1[a * SB_KLD + i] = P.YTS[i * SB_KLD + a] + P.dYTS[i * SB_KLD + a];
            for (int i = 0; i < k; ++i)
                for (int j = i + 1; j < k; ++j) P.T1[i * SB_KLD + j] = P.T1[j * SB_KLD + i];
        }
        __syncthreads();
        if (warp == 0) sbs_jacobi_warp(P.T1, k, P.T2, P.w, P.perm);
        __syncthreads();
        if (tid == 0) {
            double acc = 0.0;
            for (int i = 0; i < k; ++i) acc += log(fmax(fabs(P.w[i]), 1e-12));
            lam0[b] = exp(acc / k);
        }
    }
}

// B[:ncart,:ncart] = lam0 * I, rest 0; evals = lam0 (ncart) / 0, Vt = I.
__global__ void fill_scaled_identity_kernel(double* __restrict__ B_, double* __restrict__ evals_,
                                            double* __restrict__ Vt_, const double* __restrict__ lam0, int n,
                                            int ncart, const int* __restrict__ skip) {
    const int b = blockIdx.y;
    if (skip && skip[b]) return;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    const double l = lam0[b];
    B_[(size_t)b * n * n + idx] = (i == j && i < ncart) ? l : 0.0;
    if (Vt_) Vt_[(size_t)b * n * n + idx] = (i == j) ? 1.0 : 0.0;
    if (evals_ && j == 0) evals_[(size_t)b * n + i] = (i < ncart) ? l : 0.0;
}

// absBS coefficients: C[b,a,i] = |lam_i| * VtS[b,a,i]   (|B| S = V (|lam| * V^T S))
__global__ void abs_scale_kernel(const double* __restrict__ VtS, const double* __restrict__ evals,
                                 double* __restrict__ out, int kcap, int n, const int* __restrict__ skip) {
    const int b = blockIdx.y;
    if (skip && skip[b]) return;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)kcap * n) return;
    const int i = (int)(idx % n);
    out[(size_t)b * kcap * n + idx] = fabs(evals[(size_t)b * n + i]) * VtS[(size_t)b * kcap * n + idx];
}

struct MidShared {
    double G1[SB_KMAT], G2[SB_KMAT], XS[SB_KMAT], Minv[SB_KMAT], C[SB_KMAT];
    double w[SB_KMAX];
    int perm[SB_KMAX];
    int ok, meth;
};

// method: 0 TS-BFGS, 1 PSB, 2 Greenstadt, 3 DFP, 4 BFGS, 5 SR1, 6 BFGS_auto
// (hessian_update.py:77-101, 114-152).  In: S, Ytil, BS, absBS (TS-BFGS / BFGS_auto),
// evals (BFGS_auto: spectrum of B, ascending).  Out: kout[b] secant "pairs" (U_a, J_a, W_a),
// each [b,kcap,n], with  B+ = B + sum_a (U_a J_a^T + J_a U_a^T) - 1/2 (W_a U_a^T + U_a W_a^T).
// Methods 0-3 are the reference's U J^T + J U^T - U (J^T S) U^T family (k pairs); BFGS and SR1
// are sums of symmetric terms  X A X^T = sum_a X_a (1/2 sum_c A[a][c] X_c)^T + transpose,
// written as pairs with W = 0 (BFGS needs 2k <= kcap pairs).  Xw: [b,kcap,n] scratch.
__global__ void __launch_bounds__(UP_THREADS)
update_mid_kernel(const double* __restrict__ S_, const double* __restrict__ Yt_, const double* __restrict__ BS_,
                  const double* __restrict__ aBS_, double* __restrict__ U_, double* __restrict__ J_,
                  double* __restrict__ W_, double* __restrict__ Xw_, int kcap, const int* __restrict__ kvec, int n,
                  int method, const int* __restrict__ skip, int* __restrict__ status, double* __restrict__ Cout,
                  const double* __restrict__ evals, int* __restrict__ kout) {
    const int b = blockIdx.x;
    if (skip[b]) { if (kout && threadIdx.x == 0) kout[b] = 0; return; }
    extern __shared__ unsigned char raw[];
    MidShared& M = *reinterpret_cast<MidShared*>(raw);
    const int k = kvec ? kvec[b] : 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t off = (size_t)b * kcap * n;
    const double* S = S_ + off;
    const double* Yt = Yt_ + off;
    const double* BS = BS_ + off;
    const double* aBS = aBS_ ? aBS_ + off : nullptr;
    double* U = U_ + off;
    double* J = J_ + off;
    double* W = W_ + off;
    double* Xw = Xw_ + off;

    int meth = method;
    if (method == 6) {
        // BFGS only if B and the pencil (S^T Ytil, S^T S) are both positive definite (:80-87)
        gram_block_u(S, Yt, k, k, n, M.G1);
        gram_block_u(S, S, k, k, n, M.G2);
        __syncthreads();
        if (tid < 32) {
            bool pd = evals != nullptr && evals[(size_t)b * n] > 0.0;
            if (pd) {
                pd = sbs_gen_eigh_warp(M.G1, M.G2, k, M.XS, M.w, M.Minv, M.perm);
                if (pd) pd = M.w[0] > 0.0;
            }
            if (tid == 0) M.meth = pd ? 4 : 0;
        }
        __syncthreads();
        meth = M.meth;
        __syncthreads();
    }
    if (meth == 4 && 2 * k > kcap) {
        if (tid == 0) { if (status) atomicOr(&status[b], SB_ST_CAPACITY); if (kout) kout[b] = 0; }
        return;
    }

    for (int i = tid; i < k * n; i += nt) J[i] = Yt[i] - BS[i];
    if (meth == 4 || meth == 5) {
        // symmetric-term family
        const int nblk = meth == 4 ? 2 : 1;
        __syncthreads();
        for (int blk = 0; blk < nblk; ++blk) {
            const double* X = meth == 5 ? J : (blk == 0 ? Yt : BS);
            const double sgn = blk == 0 ? 0.5 : -0.5;
            if (meth == 5) gram_block_u(J, S, k, k, n, M.XS);          // J^T S
            else if (blk == 0) gram_block_u(Yt, S, k, k, n, M.XS);     // Ytil^T S
            else gram_block_u(S, BS, k, k, n, M.XS);                   // S^T B S
            __syncthreads();
            if (tid == 0) {
                const bool ok = sbs_invert_serial(M.XS, k, M.Minv);
                if (!ok && status) atomicOr(&status[b], SB_ST_SINGULAR);
            }
            __syncthreads();
            // pair a of this block: U = X_a, J' = sgn * sum_c A[a][c] X_c   (J' staged in Xw / W)
            double* Jst = blk == 0 ? Xw : W;
            for (int e = tid; e < n; e += nt) {
                double x[SB_KMAX];
                for (int c = 0; c < k; ++c) x[c] = X[(size_t)c * n + e];
                for (int a = 0; a < k; ++a) {
                    double acc = 0.0;
                    for (int c = 0; c < k; ++c) acc = fma(M.Minv[a * SB_KLD + c], x[c], acc);
                    U[(size_t)(blk * k + a) * n + e] = x[a];
                    Jst[(size_t)a * n + e] = sgn * acc;
                }
            }
            __syncthreads();
        }
        const int kp = nblk * k;
        for (int i = tid; i < k * n; i += nt) {
            J[i] = Xw[i];
            if (nblk == 2) J[(size_t)k * n + i] = W[i];
        }
        __syncthreads();
        for (int i = tid; i < kp * n; i += nt) W[i] = 0.0;
        if (Cout)
            for (int i = tid; i < kp * kp; i += nt) Cout[(size_t)b * SB_KMAT + (i / kp) * SB_KLD + (i % kp)] = 0.0;
        if (kout && tid == 0) kout[b] = kp;
        return;
    }
    if (meth == 0) {
        gram_block_u(S, Yt, k, k, n, M.G1);       // S^T Ytilde
        gram_block_u(S, aBS, k, k, n, M.G2);      // S^T |B|S
        __syncthreads();
        // X_a = sum_b G1[a][b] Ytilde_b + G2[a][b] absBS_b      (rows of X1 + X2)
        for (int e = tid; e < n; e += nt) {
            double y[SB_KMAX], z[SB_KMAX];
            for (int c = 0; c < k; ++c) { y[c] = Yt[(size_t)c * n + e]; z[c] = aBS[(size_t)c * n + e]; }
            for (int a = 0; a < k; ++a) {
                double x1 = 0.0, x2 = 0.0;
                for (int c = 0; c < k; ++c) { x1 = fma(M.G1[a * SB_KLD + c], y[c], x1); x2 = fma(M.G2[a * SB_KLD + c], z[c], x2); }
                Xw[(size_t)a * n + e] = x1 + x2;
            }
        }
        __syncthreads();
        gram_block_u(Xw, S, k, k, n, M.XS);       // (X1+X2) S
    } else if (meth == 1) {
        for (int i = tid; i < k * n; i += nt) Xw[i] = S[i];
        gram_block_u(S, S, k, k, n, M.XS);
    } else if (meth == 2) {
        for (int i = tid; i < k * n; i += nt) Xw[i] = BS[i];
        gram_block_u(S, BS, k, k, n, M.XS);
    } else {
        for (int i = tid; i < k * n; i += nt) Xw[i] = Yt[i];
        gram_block_u(S, Yt, k, k, n, M.XS);       // DFP: U = Ytil (S^T Ytil)^-T
    }
    __syncthreads();
    if (tid == 0) {
        const bool ok = sbs_invert_serial(M.XS, k, M.Minv);
        M.ok = ok;
        if (!ok && status) atomicOr(&status[b], SB_ST_SINGULAR);
    }
    __syncthreads();
    // U_a = sum_c Minv[a][c] X_c
    for (int e = tid; e < n; e += nt) {
        double x[SB_KMAX];
        for (int c = 0; c < k; ++c) x[c] = Xw[(size_t)c * n + e];
        for (int a = 0; a < k; ++a) {
            double acc = 0.0;
            for (int c = 0; c < k; ++c) acc = fma(M.Minv[a * SB_KLD + c], x[c], acc);
            U[(size_t)a * n + e] = acc;
        }
    }
    __syncthreads();
    gram_block_u(J, S, k, k, n, M.C);             // C = J^T S
    __syncthreads();
    if (Cout)
        for (int i = tid; i < k * k; i += nt)
            Cout[(size_t)b * SB_KMAT + (i / k) * SB_KLD + (i % k)] = M.C[(i / k) * SB_KLD + (i % k)];
    // W[:, c] = sum_a U[:, a] C[a][c]
    for (int e = tid; e < n; e += nt) {
        double u[SB_KMAX];
        for (int a = 0; a < k; ++a) u[a] = U[(size_t)a * n + e];
        for (int c = 0; c < k; ++c) {
            double acc = 0.0;
            for (int a = 0; a < k; ++a) acc = fma(u[a], M.C[a * SB_KLD + c], acc);
            W[(size_t)c * n + e] = acc;
        }
    }
    if (kout && tid == 0) kout[b] = k;
}

// B += sum_a (U_a J_a^T + J_a U_a^T) - 1/2 sum_a (W_a U_a^T + U_a W_a^T), streaming.
// Grid: (row chunks, batch).  Column-side vectors of a chunk of KC secant pairs
// are staged in shared memory; products are formed without FMA contraction so that
// element (i,j) and (j,i) receive bit-identical increments.
constexpr int KC = 4;
constexpr int ROWS_PER_CTA = 32;

__global__ void __launch_bounds__(UP_THREADS)
update_apply_kernel(double* __restrict__ B_, const double* __restrict__ U_, const double* __restrict__ J_,
                    const double* __restrict__ W_, int kcap, const int* __restrict__ kvec, int n,
                    const int* __restrict__ skip, int kcs) {
    const int b = blockIdx.y;
    if (skip[b]) return;
    extern __shared__ double sm[];
    const int k = kvec ? kvec[b] : 1;
    const size_t off = (size_t)b * kcap * n;
    const double* U = U_ + off;
    const double* J = J_ + off;
    const double* W = W_ + off;
    double* B = B_ + (size_t)b * n * n;
    const int r0 = blockIdx.x * ROWS_PER_CTA;
    const int r1 = min(n, r0 + ROWS_PER_CTA);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int a0 = 0; a0 < k; a0 += kcs) {
        const int kc = min(kcs, k - a0);
        double* cu = sm;                       // [kc][n]; kcs = staged pairs per pass (1 when every system has k = 1)
        double* cj = cu + (size_t)kcs * n;
        double* cw = cj + (size_t)kcs * n;
        __syncthreads();
        for (int i = tid; i < kc * n; i += nt) {
            cu[i] = U[(size_t)a0 * n + i];
            cj[i] = J[(size_t)a0 * n + i];
            cw[i] = W[(size_t)a0 * n + i];
        }
        __syncthreads();
        // warps own rows, lanes own columns (8 per lane and chunk, all loads issued
        // before the arithmetic): coalesced 2 KB bursts, 8 independent loads per lane
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        constexpr int LPC = 8;
        for (int i = r0 + warp; i < r1; i += nw) {
            double ui[KC], ji[KC], wi[KC];
#pragma unroll
            for (int a = 0; a < KC; ++a) {
                ui[a] = a < kc ? cu[(size_t)a * n + i] : 0.0;
                ji[a] = a < kc ? cj[(size_t)a * n + i] : 0.0;
                wi[a] = a < kc ? cw[(size_t)a * n + i] : 0.0;
            }
            double* row = B + (size_t)i * n;
            for (int c0 = 0; c0 < n; c0 += 32 * LPC) {
                double bv[LPC];
#pragma unroll
                for (int q = 0; q < LPC; ++q) {
                    const int j = c0 + lane + 32 * q;
                    bv[q] = j < n ? row[j] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < LPC; ++q) {
                    const int j = c0 + lane + 32 * q;
                    if (j >= n) continue;
                    double inc = 0.0;
#pragma unroll
                    for (int a = 0; a < KC; ++a) {
                        if (a < kc) {
                            const double uj = cu[(size_t)a * n + j], jj = cj[(size_t)a * n + j], wj = cw[(size_t)a * n + j];
                            const double p1 = __dmul_rn(ui[a], jj), p2 = __dmul_rn(ji[a], uj);
                            const double q1 = __dmul_rn(wi[a], uj), q2 = __dmul_rn(ui[a], wj);
                            const double t = __dadd_rn(__dadd_rn(p1, p2), __dmul_rn(-0.5, __dadd_rn(q1, q2)));
                            inc = __dadd_rn(inc, t);
                        }
                    }
                    row[j] = __dadd_rn(bv[q], inc);
                }
            }
        }
    }
}

}  // namespace

extern "C" int sb_update_prep_impl(const double* S, const double* Y, double* Ytil, int kcap, const int* kvec,
                                   int n, int ncart, int first, int symm, double* lam0, int* skip, int* status,
                                   const int* active, int batch, cudaStream_t st) {
    const size_t smem = sizeof(PrepShared);
    cudaFuncSetAttribute(update_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    if (symm < 0 || symm > 2) return -1;
    update_prep_kernel<<<batch, UP_THREADS, smem, st>>>(S, Y, Ytil, kcap, kvec, n, ncart, first, symm, lam0, skip,
                                                        status, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_fill_scaled_identity_impl(double* B, double* evals, double* Vt, const double* lam0, int n,
                                            int ncart, const int* skip, int batch, cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)n * n + 255) / 256), batch);
    SB_COUNT(1);
    fill_scaled_identity_kernel<<<grid, 256, 0, st>>>(B, evals, Vt, lam0, n, ncart, skip);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_abs_scale_impl(const double* VtS, const double* evals, double* out, int kcap, int n,
                                 const int* skip, int batch, cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)kcap * n + 255) / 256), batch);
    SB_COUNT(1);
    abs_scale_kernel<<<grid, 256, 0, st>>>(VtS, evals, out, kcap, n, skip);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_update_mid_impl(const double* S, const double* Ytil, const double* BS, const double* aBS,
                                  double* U, double* J, double* W, double* Xw, int kcap, const int* kvec, int n,
                                  int method, const int* skip, int* status, double* Cout, const double* evals,
                                  int* kout, int batch, cudaStream_t st) {
    if (method < 0 || method > 6) return -1;
    if ((method == 0 || method == 6) && aBS == nullptr) return -1;
    if ((method == 4 || method == 6) && kout == nullptr) return -1;
    const size_t smem = sizeof(MidShared);
    cudaFuncSetAttribute(update_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    update_mid_kernel<<<batch, UP_THREADS, smem, st>>>(S, Ytil, BS, aBS, U, J, W, Xw, kcap, kvec, n, method,
                                                       skip, status, Cout, evals, kout);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_update_apply_impl(double* B, const double* U, const double* J, const double* W, int kcap,
                                    const int* kvec, int n, const int* skip, int batch, cudaStream_t st) {
    const int kcs = kvec ? (kcap < KC ? kcap : KC) : 1;
    const size_t smem = (size_t)3 * kcs * n * sizeof(double);
    cudaFuncSetAttribute(update_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((n + ROWS_PER_CTA - 1) / ROWS_PER_CTA, batch);
    SB_COUNT(1);
    update_apply_kernel<<<grid, UP_THREADS, smem, st>>>(B, U, J, W, kcap, kvec, n, skip, kcs);
    return SB_LAUNCH_CHECK();
}
