// Eigendecomposition UPDATE of the approximate Hessian after a low-rank secant
// update:  B+ = B + Delta,  Delta = U J^T + J U^T - U C_sym U^T  (rank <= 2k).
//
// The reference recomputes scipy.linalg.eigh(B) from scratch after every update
// (sella/linalg.py:293 -> 174-195; 10 n^3 flops, its single biggest per-step cost).
// Given B = V diag(d) V^T this file instead applies the classical
// "diagonal plus rank-one" machinery (Bunch-Nielsen-Sorensen 1978; the merge step of
// LAPACK's divide & conquer, dlaed2/dlaed4/dlaed3, with the Gu-Eisenstat recomputed
// z-vector for orthogonality):
//
//   1. lowrank_factor : Delta = sum_t sigma_t p_t p_t^T with orthonormal p_t
//                       (Gram matrix of [U J], small symmetric eigenproblems);
//   2. (hv.cu)        : z_t = V^T p_t for all terms in one pass over Vt;
//   3. secular_update : for each term: deflation (negligible z_i, or nearly equal
//                       d_i, d_j combined by a plane rotation of two eigenvectors),
//                       roots of 1 + rho sum z_i^2/(d_i - lambda) = 0 for the rest,
//                       z recomputed from the roots, new eigenvectors
//                       V_nd <- V_nd Qhat; finally eigenvalues sorted ascending and the
//                       rows of Vt permuted in place.
//
// Quasi-Newton Hessians are "scaled identity + low rank": almost everything deflates
// and an update costs a few passes over Vt instead of a full eigensolve.  The result
// is the eigendecomposition of the same matrix, to rounding.
#include <cstdlib>
#include "small_dense.cuh"

namespace {

constexpr int SEC_THREADS = 256;
constexpr double SEC_EPS = 2.220446049250313e-16;

struct FactorShared {
    double R[SB_KMAT], K[SB_KMAT], Y[SB_KMAT], Mc[SB_KMAT], A[SB_KMAT];
    double sg[SB_KMAX], c[SB_KMAX];
    double scratch[SB_SCRATCH_DOUBLES];
    int perm[SB_KMAX];
};

// F = [U_0..U_{k-1}, J_0..J_{k-1}] (m = 2k vectors, m <= 32).  P[b,t,:] = p_t, sig[b,t],
// nterm[b].  Cmat: k x k (ld SB_KLD) = J^T S as written by update_mid.
//   F = Q R   (classical Gram-Schmidt with one re-orthogonalisation pass, "CGS2": errors of
//              order eps*cond(F), where a Gram-matrix route would square the condition number --
//              DFP/SR1-type updates put vectors of very different length into F)
//   Delta = F Mc F^T = Q (R Mc R^T) Q^T,  Mc = [[-C_sym, I], [I, 0]];  R Mc R^T = Y diag(sig) Y^T
//   p_t = Q y_t.   Q is built in the output buffer P and rotated in place.
__global__ void __launch_bounds__(SEC_THREADS)
lowrank_factor_kernel(const double* __restrict__ U_, const double* __restrict__ J_, const double* __restrict__ Cmat_,
                      int kcap, const int* __restrict__ kvec, int n, double* __restrict__ P_, double* __restrict__ sig_,
                      int* __restrict__ nterm, const int* __restrict__ skip) {
    const int b = blockIdx.x;
    if (skip[b]) { if (threadIdx.x == 0) nterm[b] = 0; return; }
    extern __shared__ unsigned char raw[];
    FactorShared& S = *reinterpret_cast<FactorShared*>(raw);
    double* xs = reinterpret_cast<double*>(raw + ((sizeof(FactorShared) + 15) / 16) * 16);   // n doubles
    const int k = kvec ? kvec[b] : 1;
    const int m = 2 * k;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const double* U = U_ + (size_t)b * kcap * n;
    const double* J = J_ + (size_t)b * kcap * n;
    const double* C = Cmat_ + (size_t)b * SB_KMAT;
    double* P = P_ + (size_t)b * 2 * kcap * n;
    auto F = [&](int a) { return a < k ? U + (size_t)a * n : J + (size_t)(a - k) * n; };
    if (m == 0) { if (tid == 0) nterm[b] = 0; return; }
    // Mc = [[-C_sym, I], [I, 0]]
    for (int idx = tid; idx < m * m; idx += nt) {
        const int i = idx / m, j = idx % m;
        double v = 0.0;
        if (i < k && j < k) v = -0.5 * (C[i * SB_KLD + j] + C[j * SB_KLD + i]);
        else if (i < k && j == i + k) v = 1.0;
        else if (j < k && i == j + k) v = 1.0;
        S.Mc[i * SB_KLD + j] = v;
        S.R[i * SB_KLD + j] = 0.0;
    }
    __syncthreads();
    int r = 0;
    for (int a = 0; a < m; ++a) {
        const double* f = F(a);
        double acc = 0.0;
        for (int e = tid; e < n; e += nt) { const double v = f[e]; xs[e] = v; acc = fma(v, v, acc); }
        const double nrm0 = sqrt(sb_block_sum(acc, S.scratch));
        if (!(nrm0 > 0.0)) continue;                          // zero column: R[:, a] = 0
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = warp; i < r; i += nw) {
                const double* q = P + (size_t)i * n;
                double d = 0.0;
                for (int e = lane; e < n; e += 32) d = fma(q[e], xs[e], d);
                d = sb_warp_sum(d);
                if (lane == 0) { S.c[i] = d; S.R[i * SB_KLD + a] += d; }
            }
            __syncthreads();
            for (int e = tid; e < n; e += nt) {
                double v = xs[e];
                for (int i = 0; i < r; ++i) v = fma(-S.c[i], P[(size_t)i * n + e], v);
                xs[e] = v;
            }
            __syncthreads();
        }
        acc = 0.0;
        for (int e = tid; e < n; e += nt) acc = fma(xs[e], xs[e], acc);
        const double nrm = sqrt(sb_block_sum(acc, S.scratch));
        if (!(nrm > 4.0 * SEC_EPS * nrm0)) continue;          // numerically dependent column
        if (tid == 0) S.R[r * SB_KLD + a] = nrm;
        const double inv = 1.0 / nrm;
        for (int e = tid; e < n; e += nt) P[(size_t)r * n + e] = xs[e] * inv;
        ++r;
        __syncthreads();
    }
    if (r == 0) { if (tid == 0) nterm[b] = 0; return; }
    // K = R Mc R^T  (r x r), symmetrised
    for (int idx = tid; idx < r * m; idx += nt) {
        const int i = idx / m, bq = idx % m;
        double acc = 0.0;
        for (int a = 0; a < m; ++a) acc = fma(S.R[i * SB_KLD + a], S.Mc[a * SB_KLD + bq], acc);
        S.A[i * SB_KLD + bq] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += nt) {
        const int i = idx / r, j = idx % r;
        double acc = 0.0;
        for (int bq = 0; bq < m; ++bq) acc = fma(S.A[i * SB_KLD + bq], S.R[j * SB_KLD + bq], acc);
        S.K[i * SB_KLD + j] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < r * r; idx += nt) {
        const int i = idx / r, j = idx % r;
        if (j > i) {
            const double v = 0.5 * (S.K[i * SB_KLD + j] + S.K[j * SB_KLD + i]);
            S.K[i * SB_KLD + j] = v; S.K[j * SB_KLD + i] = v;
        }
    }
    __syncthreads();
    if (warp == 0) sbs_jacobi_warp(S.K, r, S.Y, S.sg, S.perm);    // K = Y diag(sg) Y^T
    __syncthreads();
    // p_t = sum_i Y[i][t] q_i, in place (each thread owns its elements)
    for (int e = tid; e < n; e += nt) {
        double q[SB_KMAX];
        for (int i = 0; i < r; ++i) q[i] = P[(size_t)i * n + e];
        for (int t = 0; t < r; ++t) {
            double acc = 0.0;
            for (int i = 0; i < r; ++i) acc = fma(q[i], S.Y[i * SB_KLD + t], acc);
            P[(size_t)t * n + e] = acc;
        }
    }
    if (tid < r) sig_[(size_t)b * 2 * kcap + tid] = S.sg[tid];
    if (tid == 0) nterm[b] = r;
}

// optional in-kernel phase profile (cycles, summed over CTAs by thread 0); [14] secular iterations, [15] roots
__device__ unsigned long long sec_prof[16];

// ------------------------------------------------------------------------------
// One secular root of  f(lam) = 1 + rho * sum_i y_i^2 / (e_i - lam),  rho > 0,
// e ascending, in (e_j, e_{j+1})  (j = r-1: (e_{r-1}, e_{r-1} + rho*|y|^2)).
// Returns the pole index `org` and the offset mu with lam = e[org] + mu.
__device__ void secular_root(const double* __restrict__ e, const double* __restrict__ y2, int r, double rho, int j,
                             double ysum, int* org_out, double* mu_out) {
    // executed by one full warp: lanes split the r terms of the secular function
    const int lane = threadIdx.x & 31;
    const double left = e[j];
    const double gap = (j + 1 < r) ? (e[j + 1] - e[j]) : rho * ysum;
    int org = j;
    if (j + 1 < r) {
        const double mid = 0.5 * gap;
        double fm = 0.0;
        for (int i = lane; i < r; i += 32) fm += rho * y2[i] / ((e[i] - left) - mid);
        fm = 1.0 + sb_warp_sum(fm);
        if (fm < 0.0) org = j + 1;         // root in the right half: measure from e_{j+1}
    }
    const double eo = e[org];
    double lo, hi;
    if (org == j) { lo = 0.0; hi = (j + 1 < r) ? 0.5 * gap : gap; }
    else { lo = -0.5 * gap; hi = 0.0; }
    // f = p + q with p = -w/mu the pole the offset is measured from (w = rho y_org^2) and q the
    // smooth remainder; each step solves  q(mu_k) + q'(mu_k)(t - mu_k) - w/t = 0  (Newton on q with
    // the pole kept exact: a few steps even when the root hugs the pole), inside the bracket
    // with (geometric) bisection as the safeguard
    const double w = rho * y2[org];
    double mu = (org == j) ? ((j + 1 < r) ? 0.25 * gap : 0.5 * gap) : -0.25 * gap;
    {
        // first guess from the pole-exact model expanded AT the pole: q(0) + q'(0) t - w/t = 0.  A pole
        // with a tiny weight (a direction the update barely touches: most explicit pairs of a compact
        // spectrum) has its root at a distance ~ w/q(0) from it, many orders of magnitude inside the
        // bracket; started from the middle of the bracket the safeguarded iteration needs dozens of
        // (geometric) bisection steps to get there, started here two or three
        double q0 = 0.0, dq0 = 0.0;
        for (int i = lane; i < r; i += 32) {
            if (i == org) continue;
            const double den = e[i] - eo;
            const double t = y2[i] / den;
            q0 += rho * t;
            dq0 += rho * t / den;
        }
        q0 = 1.0 + sb_warp_sum(q0);
        dq0 = sb_warp_sum(dq0);
        const double sq0 = sqrt(fma(q0, q0, 4.0 * dq0 * w));
        double cand;
        if (dq0 > 0.0) {
            if (org == j) cand = q0 <= 0.0 ? (sq0 - q0) / (2.0 * dq0) : 2.0 * w / (q0 + sq0);
            else cand = q0 >= 0.0 ? -(q0 + sq0) / (2.0 * dq0) : -2.0 * w / (sq0 - q0);
        } else {
            cand = (q0 != 0.0) ? w / q0 : mu;
        }
        if (cand > lo && cand < hi) mu = cand;
    }
    int nit_ = 0;
    for (int it = 0; it < 200; ++it) {
        ++nit_;
        double q = 0.0, dq = 0.0, qa = 0.0;
        for (int i = lane; i < r; i += 32) {
            if (i == org) continue;
            const double den = (e[i] - eo) - mu;
            const double t = y2[i] / den;
            q += rho * t;
            qa += fabs(rho * t);
            dq += rho * t / den;
        }
        q = 1.0 + sb_warp_sum(q);
        dq = sb_warp_sum(dq);
        qa = sb_warp_sum(qa);
        const double f = q - w / mu;
        // f cannot be evaluated more accurately than eps * (sum of |terms|): stop there (LAPACK
        // dlaed4's criterion); the Gu-Eisenstat z-vector keeps the eigenvectors orthogonal for
        // whatever roots were computed
        if (fabs(f) <= 4.0 * SEC_EPS * (1.0 + qa + fabs(w / mu))) break;
        if (f < 0.0) lo = mu; else hi = mu;
        const double Bq = q - dq * mu;
        const double sq = sqrt(fma(Bq, Bq, 4.0 * dq * w));
        double next;
        if (dq > 0.0) {
            if (org == j) next = Bq <= 0.0 ? (sq - Bq) / (2.0 * dq) : 2.0 * w / (Bq + sq);
            else next = Bq >= 0.0 ? -(Bq + sq) / (2.0 * dq) : -2.0 * w / (sq - Bq);
        } else {
            next = (Bq != 0.0) ? w / Bq : 0.5 * (lo + hi);   // no other pole: q is constant, w/t = q
        }
        if (!(next > lo && next < hi)) {
            if (org == j) {
                if (lo > 0.0 && hi > 4.0 * lo) next = sqrt(lo) * sqrt(hi);
                else if (lo == 0.0) next = (hi > 1e-290) ? hi * 0.0625 : 0.5 * hi;
                else next = 0.5 * (lo + hi);
            } else {
                if (hi < 0.0 && lo < 4.0 * hi) next = -sqrt(-lo) * sqrt(-hi);
                else if (hi == 0.0) next = (lo < -1e-290) ? lo * 0.0625 : 0.5 * lo;
                else next = 0.5 * (lo + hi);
            }
        }
        if (fabs(next - mu) <= 2.0 * SEC_EPS * fabs(next) || next == mu) { mu = next; break; }
        mu = next;
        if (hi - lo <= 2.0 * SEC_EPS * fmax(fabs(lo), fabs(hi))) break;
    }
    *org_out = org;
    *mu_out = mu;
#ifdef SB_SEC_COUNT_ITERS
    if (lane == 0) { atomicAdd(&sec_prof[14], (unsigned long long)nit_); atomicAdd(&sec_prof[15], 1ull); }
#endif
}

// Reciprocal for the secular sums: hardware approximation (2^-23) + two Newton steps, i.e. full fp64
// accuracy up to the last bit at a fifth of the instructions of an IEEE division.  The root search does
// O(r^2) of these per rank-one term: with one thread per root they are what the solve kernel spends its
// time on.
__device__ __forceinline__ double sec_rcp(double a) {
    if (!(fabs(a) > 1e-290 && fabs(a) < 1e290)) return 1.0 / a;       // 0, denormals, huge, NaN: the IEEE path
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    return x;
}

// One secular root with ONE THREAD per root (split-mode solve kernel: r-way parallel over the roots, plain
// loops over the r poles in shared memory -- all threads of a warp read the same e[i], y2[i], a broadcast).
// Iteration: the "middle way" of LAPACK's dlaed4 (Li 1994): f = 1 + psi + phi with psi the poles up to j and
// phi the poles above; psi is modelled as s + a/(delta_j - eta), phi as S + b/(delta_{j+1} - eta) (value and
// slope matched), so BOTH neighbouring poles are exact and each step solves a quadratic.  A Taylor model of
// everything but the nearest pole (secular_root above) needs 10-35 steps when the other neighbour is close
// (gaps of 1e-8 are common in a compact spectrum) and the slowest root sets the time of the whole CTA; this
// one needs at most ~7.  First guess: both neighbouring poles exact, the rest frozen at the origin pole.
__device__ void secular_root_serial(const double* __restrict__ e, const double* __restrict__ y2, int r, double rho,
                                    int j, double ysum, int* org_out, double* mu_out) {
    const bool last = j + 1 >= r;
    const double left = e[j];
    const double gap = last ? rho * ysum : (e[j + 1] - e[j]);
    int org = j;
    if (!last) {
        const double mid = 0.5 * gap;
        double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0;
        int i = 0;
        for (; i + 3 < r; i += 4) {
            f0 = fma(y2[i], sec_rcp((e[i] - left) - mid), f0);
            f1 = fma(y2[i + 1], sec_rcp((e[i + 1] - left) - mid), f1);
            f2 = fma(y2[i + 2], sec_rcp((e[i + 2] - left) - mid), f2);
            f3 = fma(y2[i + 3], sec_rcp((e[i + 3] - left) - mid), f3);
        }
        for (; i < r; ++i) f0 = fma(y2[i], sec_rcp((e[i] - left) - mid), f0);
        if (1.0 + rho * ((f0 + f1) + (f2 + f3)) < 0.0) org = j + 1;
    }
    const double eo = e[org];
    double lo, hi;
    if (org == j) { lo = 0.0; hi = last ? gap : 0.5 * gap; }
    else { lo = -0.5 * gap; hi = 0.0; }
    const double dj = e[j] - eo;                            // 0, or -gap when the origin is pole j+1
    const double dj1 = last ? 0.0 : e[j + 1] - eo;
    const double w = rho * y2[org];
    double mu = (org == j) ? (last ? 0.5 * gap : 0.25 * gap) : -0.25 * gap;
    {
        // first guess: R0 - w/t + wK/(D - t) = 0 with R0 = 1 + rho sum over the OTHER poles of y2/(e - eo)
        const int other = last ? -1 : (org == j ? j + 1 : j);
        double r0 = 0.0, r1 = 0.0;
        int i = 0;
        auto term = [&](int k) {
            const bool skip = (k == org) || (k == other);
            return (skip ? 0.0 : y2[k]) * sec_rcp(skip ? 1.0 : e[k] - eo);
        };
        for (; i + 1 < r; i += 2) { r0 += term(i); r1 += term(i + 1); }
        for (; i < r; ++i) r0 += term(i);
        const double R0 = 1.0 + rho * (r0 + r1);
        double cand = mu;
        if (other < 0) {
            if (R0 > 0.0) cand = w / R0;
        } else {
            const double D = e[other] - eo, wK = rho * y2[other];
            // -R0 t^2 + (R0 D + w + wK) t - w D = 0
            const double qa_ = -R0, qb_ = fma(R0, D, w + wK), qc_ = -w * D;
            const double disc = fma(qb_, qb_, -4.0 * qa_ * qc_);
            if (disc >= 0.0) {
                const double sq = sqrt(disc);
                const double qq = -0.5 * (qb_ + (qb_ >= 0.0 ? sq : -sq));
                const double x1 = qq != 0.0 ? qc_ / qq : mu, x2 = qa_ != 0.0 ? qq / qa_ : mu;
                cand = (x1 > lo && x1 < hi) ? x1 : x2;
            }
        }
        if (cand > lo && cand < hi) mu = cand;
    }
    for (int it = 0; it < 200; ++it) {
        // psi, psi' over the poles <= j and phi, phi' over the poles > j at the current iterate
        double ps0 = 0.0, ps1 = 0.0, dp0 = 0.0, dp1 = 0.0, ph0 = 0.0, ph1 = 0.0, dh0 = 0.0, dh1 = 0.0, qa0 = 0.0, qa1 = 0.0;
        auto term = [&](int k, double& ps, double& dp, double& ph, double& dh, double& qa) {
            const double rc = sec_rcp((e[k] - eo) - mu);
            const double t = y2[k] * rc, t2 = t * rc;
            const bool lower = k <= j;
            ps += lower ? t : 0.0;  dp += lower ? t2 : 0.0;
            ph += lower ? 0.0 : t;  dh += lower ? 0.0 : t2;
            qa += fabs(t);
        };
        int i = 0;
        for (; i + 1 < r; i += 2) { term(i, ps0, dp0, ph0, dh0, qa0); term(i + 1, ps1, dp1, ph1, dh1, qa1); }
        for (; i < r; ++i) term(i, ps0, dp0, ph0, dh0, qa0);
        const double psi = rho * (ps0 + ps1), dpsi = rho * (dp0 + dp1), phi = rho * (ph0 + ph1), dphi = rho * (dh0 + dh1);
        const double qa = rho * (qa0 + qa1);
        const double f = 1.0 + psi + phi;
        if (fabs(f) <= 4.0 * SEC_EPS * (1.0 + qa)) break;
        if (f < 0.0) lo = mu; else hi = mu;
        const double Dj = dj - mu;
        const double a = dpsi * Dj * Dj, sA = psi - dpsi * Dj;
        double next = 0.0;
        bool ok = false;
        if (last) {
            const double c = 1.0 + sA;
            if (c != 0.0) { next = mu + Dj + a / c; ok = next > lo && next < hi; }
        } else {
            const double Dj1 = dj1 - mu;
            const double bB = dphi * Dj1 * Dj1, sB = phi - dphi * Dj1;
            const double c = 1.0 + sA + sB;
            const double Bc = c * (Dj + Dj1) + a + bB;
            const double C0 = c * Dj * Dj1 + a * Dj1 + bB * Dj;
            double disc = fma(Bc, Bc, -4.0 * c * C0);
            if (disc < 0.0) disc = 0.0;
            const double sq = sqrt(disc);
            const double qq = 0.5 * (Bc + (Bc >= 0.0 ? sq : -sq));
            if (qq != 0.0) { next = mu + C0 / qq; ok = next > lo && next < hi; }
            if (!ok && c != 0.0) { next = mu + qq / c; ok = next > lo && next < hi; }
        }
        if (!ok) {
            if (lo >= 0.0) {
                if (lo > 0.0 && hi > 4.0 * lo) next = sqrt(lo) * sqrt(hi);
                else if (lo == 0.0) next = (hi > 1e-290) ? hi * 0.0625 : 0.5 * hi;
                else next = 0.5 * (lo + hi);
            } else {
                if (hi < 0.0 && lo < 4.0 * hi) next = -sqrt(-lo) * sqrt(-hi);
                else if (hi == 0.0) next = (lo < -1e-290) ? lo * 0.0625 : 0.5 * lo;
                else next = 0.5 * (lo + hi);
            }
        }
        if (fabs(next - mu) <= 2.0 * SEC_EPS * fabs(next) || next == mu) { mu = next; break; }
        mu = next;
        if (hi - lo <= 2.0 * SEC_EPS * fmax(fabs(lo), fabs(hi))) break;
    }
    *org_out = org;
    *mu_out = mu;
}

// ------------------------------------------------------------------------------
// Pre-phase: ONE block reflector for the big degenerate cluster.
//
// A quasi-Newton Hessian is lam0*I + low rank: most of its eigenvalues are the same number
// and the matching rows of Vt are an arbitrary orthonormal basis of that eigenspace.  Every
// rank-one term has weight on all of them, and deflating term by term reflects the whole
// cluster once per term (2 reads + 1 write of ~n rows each).  Instead, with Z_C the m x T
// block of all T pending z vectors on the m cluster rows, one Householder QR  Z_C = Q R  is
// applied as  Rows_C <- Q^T Rows_C  (compact WY: I - W Tm^T W^T) in a single streaming pass
// pair; afterwards term t only touches the first t+1 cluster rows (R is upper triangular)
// and the per-term code below finds nothing left to reflect.
//   cluster_qr_kernel      (1 CTA / system): finds the cluster, factors Z_C in place (Z gets R),
//                          writes W [T][m] to `work`, Tm [16x16] and {c0, m, T} to `qwork`.
//   cluster_reflect_kernel (column chunks x systems): the bandwidth-bound application.
constexpr int CQ_TMAX = 16;
constexpr int CR_THREADS = 128;
constexpr int CR_ROWS = 64;

__global__ void __launch_bounds__(SEC_THREADS)
cluster_qr_kernel(const double* __restrict__ evals_, double* __restrict__ Z_, int zcap, const int* __restrict__ nterm,
                  int n, double* __restrict__ work_, double* __restrict__ qwork_, const int* __restrict__ skip) {
    const int b = blockIdx.x;
    int* meta = reinterpret_cast<int*>(qwork_ + (size_t)b * n * n);
    double* TmG = qwork_ + (size_t)b * n * n + 8;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const int T = skip[b] ? 0 : min(nterm[b], CQ_TMAX);
    if (T == 0 || nterm[b] > CQ_TMAX) { if (tid == 0) { meta[0] = 0; meta[1] = 0; meta[2] = 0; } return; }
    extern __shared__ double sm[];
    double* d = sm;                          // n
    double* scratch = d + n;                 // SB_SCRATCH_DOUBLES
    double* G = scratch + SB_SCRATCH_DOUBLES;   // 16 x 16
    double* Tm = G + CQ_TMAX * CQ_TMAX;      // 16 x 16
    double* beta = Tm + CQ_TMAX * CQ_TMAX;   // 16
    double* dots = beta + CQ_TMAX;           // 16
    int* ibuf = reinterpret_cast<int*>(dots + CQ_TMAX);   // 16 ints
    double dmax = 0.0;
    for (int i = tid; i < n; i += nt) { const double v = evals_[(size_t)b * n + i]; d[i] = v; dmax = fmax(dmax, fabs(v)); }
    dmax = sb_warp_max(dmax);
    if (lane == 0) scratch[warp] = dmax;
    __syncthreads();
    dmax = 0.0;
    for (int w = 0; w < nw; ++w) dmax = fmax(dmax, scratch[w]);
    const double tolc = 8.0 * SEC_EPS * dmax;
    // Window [lo, i] of the ascending spectrum with d[i] - d[lo] <= tolc that holds the most rows
    // on which some pending z is non-negligible (a cluster the update does not touch -- e.g. the
    // constraint block of a projected Hessian -- needs no reflection however large it is).
    double* Zb = Z_ + (size_t)b * zcap * n;
    for (int t = warp; t < T; t += nw) {                    // tolz_t = 8 eps |z_t|
        double acc = 0.0;
        for (int e = lane; e < n; e += 32) { const double v = Zb[(size_t)t * n + e]; acc = fma(v, v, acc); }
        acc = sb_warp_sum(acc);
        if (lane == 0) dots[t] = 8.0 * SEC_EPS * sqrt(acc);
    }
    __syncthreads();
    int* pre = reinterpret_cast<int*>(ibuf + 16);           // n + 1 ints: prefix counts of "live" rows
    {
        const int per = (n + nt - 1) / nt;
        const int i0 = tid * per, i1 = min(n, i0 + per);
        int cnt = 0;
        for (int i = i0; i < i1; ++i) {
            bool live = false;
            for (int t = 0; t < T; ++t) live = live || fabs(Zb[(size_t)t * n + i]) > dots[t];
            pre[i + 1] = live ? 1 : 0;
            cnt += live ? 1 : 0;
        }
        // exclusive scan of the per-thread counts
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) ibuf[warp] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += ibuf[w];
        int run = base + incl - cnt;
        if (tid == 0) pre[0] = 0;
        for (int i = i0; i < i1; ++i) { run += pre[i + 1]; pre[i + 1] = run; }
        __syncthreads();
    }
    int best = 0;
    for (int i = tid; i < n; i += nt) {
        int a = 0, e = i;                                   // smallest j in [0, i] with d[i] - d[j] <= tolc
        const double di = d[i];
        while (a < e) { const int mid = (a + e) >> 1; if (di - d[mid] <= tolc) e = mid; else a = mid + 1; }
        const int live = pre[i + 1] - pre[a];
        // key: live rows first, then window length; the window start is recovered below
        const int key = (live << 12) | i;                   // n <= 4095
        best = max(best, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    __syncthreads();
    if (lane == 0) ibuf[warp] = best;
    __syncthreads();
    best = 0;
    for (int w = 0; w < nw; ++w) best = max(best, ibuf[w]);
    const int iend = best & 4095, nlive = best >> 12;
    int c0;
    {
        int a = 0, e = iend;
        const double di = d[iend];
        while (a < e) { const int mid = (a + e) >> 1; if (di - d[mid] <= tolc) e = mid; else a = mid + 1; }
        c0 = a;
    }
    const int m = iend - c0 + 1;
    if (nlive < T + 2) { if (tid == 0) { meta[0] = 0; meta[1] = 0; meta[2] = 0; } return; }
    if (m < T + 2 || m < 8) { if (tid == 0) { meta[0] = 0; meta[1] = 0; meta[2] = 0; } return; }
    double* Z = Z_ + (size_t)b * zcap * n;
    double* W = work_ + (size_t)b * n * n;                  // [T][m]
    auto row = [&](int j) { return c0 + m - 1 - j; };
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        double* zt = Z + (size_t)t * n;
        double acc = 0.0;
        for (int j = t + 1 + tid; j < m; j += nt) { const double v = zt[row(j)]; acc = fma(v, v, acc); }
        const double tail2 = sb_block_sum(acc, scratch);
        const double x0 = zt[row(t)];
        if (!(tail2 > 0.0)) {                               // nothing below the diagonal
            for (int j = tid; j < m; j += nt) W[(size_t)t * m + j] = 0.0;
            if (tid == 0) beta[t] = 0.0;
            __syncthreads();
            continue;
        }
        const double nrm = sqrt(tail2 + x0 * x0);
        const double alpha = x0 > 0.0 ? -nrm : nrm;
        const double wt = x0 - alpha;
        const double bt = 2.0 / (tail2 + wt * wt);
        for (int j = tid; j < m; j += nt) W[(size_t)t * m + j] = j < t ? 0.0 : (j == t ? wt : zt[row(j)]);
        __syncthreads();
        for (int s2 = t + 1 + warp; s2 < T; s2 += nw) {
            const double* zs = Z + (size_t)s2 * n;
            double dd = 0.0;
            for (int j = t + lane; j < m; j += 32) dd = fma(W[(size_t)t * m + j], zs[row(j)], dd);
            dd = sb_warp_sum(dd);
            if (lane == 0) dots[s2] = dd;
        }
        __syncthreads();
        for (int idx = tid; idx < (T - t - 1) * (m - t); idx += nt) {
            const int s2 = t + 1 + idx / (m - t), j = t + idx % (m - t);
            double* zs = Z + (size_t)s2 * n;
            zs[row(j)] = fma(-bt * dots[s2], W[(size_t)t * m + j], zs[row(j)]);
        }
        for (int j = t + 1 + tid; j < m; j += nt) zt[row(j)] = 0.0;
        if (tid == 0) { zt[row(t)] = alpha; beta[t] = bt; }
        __syncthreads();
    }
    // compact WY: Q = H_0 ... H_{T-1} = I - W Tm W^T
    for (int pr = warp; pr < T * T; pr += nw) {
        const int i = pr / T, j = pr % T;
        if (j <= i) continue;
        double dd = 0.0;
        for (int e = lane; e < m; e += 32) dd = fma(W[(size_t)i * m + e], W[(size_t)j * m + e], dd);
        dd = sb_warp_sum(dd);
        if (lane == 0) G[i * CQ_TMAX + j] = dd;
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < T; ++t) {
            for (int i = 0; i < t; ++i) {
                double acc = 0.0;
                for (int l = i; l < t; ++l) acc += Tm[i * CQ_TMAX + l] * G[l * CQ_TMAX + t];
                Tm[i * CQ_TMAX + t] = -beta[t] * acc;
            }
            Tm[t * CQ_TMAX + t] = beta[t];
            for (int i = t + 1; i < T; ++i) Tm[i * CQ_TMAX + t] = 0.0;
        }
        meta[0] = c0; meta[1] = m; meta[2] = T;
    }
    __syncthreads();
    for (int i = tid; i < T * T; i += nt) TmG[(i / T) * CQ_TMAX + (i % T)] = Tm[(i / T) * CQ_TMAX + (i % T)];
}

template <int TT>
__global__ void __launch_bounds__(CR_THREADS)
cluster_reflect_kernel(double* __restrict__ Vt_, const double* __restrict__ work_, const double* __restrict__ qwork_, int n) {
    const int b = blockIdx.y;
    const int* meta = reinterpret_cast<const int*>(qwork_ + (size_t)b * n * n);
    const int c0 = meta[0], m = meta[1], T = meta[2];
    if (m == 0 || T > TT || (TT > 2 && T <= TT / 2)) return;       // exactly one instantiation serves a system
    __shared__ double Wb[TT][CR_ROWS];
    __shared__ double Tm[TT * TT];
    const double* TmG = qwork_ + (size_t)b * n * n + 8;
    const double* W = work_ + (size_t)b * n * n;
    const int tid = threadIdx.x;
    const int col = blockIdx.x * CR_THREADS + tid;
    const bool live = col < n;
    double* X = Vt_ + (size_t)b * n * n + (live ? col : 0);
    for (int i = tid; i < TT * TT; i += CR_THREADS) {
        const int r = i / TT, c = i % TT;
        Tm[i] = (r < T && c < T) ? TmG[r * CQ_TMAX + c] : 0.0;
    }
    double y[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) y[t] = 0.0;
    for (int j0 = 0; j0 < m; j0 += CR_ROWS) {
        const int jn = min(CR_ROWS, m - j0);
        __syncthreads();
        for (int i = tid; i < TT * CR_ROWS; i += CR_THREADS) {
            const int t = i / CR_ROWS, jj = i % CR_ROWS;
            Wb[t][jj] = (t < T && jj < jn) ? W[(size_t)t * m + j0 + jj] : 0.0;
        }
        __syncthreads();
        if (live) {
            for (int jj = 0; jj < jn; jj += 8) {
                double x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = (jj + u < jn) ? X[(size_t)(c0 + m - 1 - (j0 + jj + u)) * n] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int t = 0; t < TT; ++t) y[t] = fma(Wb[t][jj + u], x[u], y[t]);
            }
        }
    }
    double yt[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < TT; ++i) acc = fma(Tm[i * TT + t], y[i], acc);      // (Tm^T y)_t, Tm upper triangular
        yt[t] = acc;
    }
    for (int j0 = 0; j0 < m; j0 += CR_ROWS) {
        const int jn = min(CR_ROWS, m - j0);
        __syncthreads();
        for (int i = tid; i < TT * CR_ROWS; i += CR_THREADS) {
            const int t = i / CR_ROWS, jj = i % CR_ROWS;
            Wb[t][jj] = (t < T && jj < jn) ? W[(size_t)t * m + j0 + jj] : 0.0;
        }
        __syncthreads();
        if (live) {
            for (int jj = 0; jj < jn; jj += 8) {
                double x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = (jj + u < jn) ? X[(size_t)(c0 + m - 1 - (j0 + jj + u)) * n] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    double acc = x[u];
#pragma unroll
                    for (int t = 0; t < TT; ++t) acc = fma(-Wb[t][jj + u], yt[t], acc);
                    if (jj + u < jn) X[(size_t)(c0 + m - 1 - (j0 + jj + u)) * n] = acc;
                }
            }
        }
    }
}

#define SEC_MARK(ph)                                                        \
    do {                                                                    \
        if (tid == 0) {                                                     \
            const long long now_ = clock64();                               \
            atomicAdd(&sec_prof[ph], (unsigned long long)(now_ - tmark_)); \
            tmark_ = now_;                                                  \
        }                                                                   \
    } while (0)

// the per-term kernel is latency-bound (short serial phases, few rows touched): small CTAs,
// several per SM, hide that better than one wide CTA
constexpr int SECK_THREADS = 128;
constexpr int SEC_QS_MAX = 54;          // r <= this: eigenvector block Qh lives in shared memory

struct SecShared {
    double scratch[SB_SCRATCH_DOUBLES];
    int r, nrot, flag;
    int wcnt[SECK_THREADS / 32];
    double rho;
};

// Sequential application of `nrot` plane rotations (rows ia[q], ib[q] of Vt):
//   x' = c x + s y ;  y' = c y - s x        (LAPACK drot on the two eigenvectors)
template <int CPT>
__device__ void apply_rotations(double* __restrict__ Vt, int n, const int* __restrict__ ia, const int* __restrict__ ib,
                                const double* __restrict__ rc, const double* __restrict__ rs, int nrot) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double carry[CPT];
    int carried = -1;
    for (int q = 0; q < nrot; ++q) {
        const int a = ia[q], bb = ib[q];
        const double c = rc[q], s = rs[q];
        if (carried != a) {
            if (carried >= 0) {
#pragma unroll
                for (int u = 0; u < CPT; ++u) { const int col = tid + u * nt; if (col < n) Vt[(size_t)carried * n + col] = carry[u]; }
            }
#pragma unroll
            for (int u = 0; u < CPT; ++u) { const int col = tid + u * nt; carry[u] = col < n ? Vt[(size_t)a * n + col] : 0.0; }
        }
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            const int col = tid + u * nt;
            if (col < n) {
                const double x = carry[u], y = Vt[(size_t)bb * n + col];
                Vt[(size_t)a * n + col] = c * x + s * y;
                carry[u] = c * y - s * x;
            }
        }
        carried = bb;
    }
    if (carried >= 0) {
#pragma unroll
        for (int u = 0; u < CPT; ++u) { const int col = tid + u * nt; if (col < n) Vt[(size_t)carried * n + col] = carry[u]; }
    }
}

// Applies nterm[b] rank-one updates (sig, Z = Vt P^T) to (evals, Vt).
// Zs: [b, zcap, n] (row t = V^T p_t in the CURRENT row order of Vt), work: [b, n, n],
// qwork: [b, n, n].  On exit evals ascending, Vt rows permuted accordingly.
template <int CPT, bool SPLIT>
__global__ void __launch_bounds__(SECK_THREADS, SPLIT ? 7 : 4)
secular_update_kernel(double* __restrict__ evals_, double* __restrict__ Vt_, double* __restrict__ Z_, int zcap,
                      const double* __restrict__ sig_, const int* __restrict__ nterm, int n,
                      double* __restrict__ work_, double* __restrict__ qwork_, int* __restrict__ status,
                      const int* __restrict__ skip, int tile_doubles, const int* __restrict__ mrows, int mc,
                      long long estride, long long vstride, int t_only, int* __restrict__ aux_, int auxs) {
    // t_only == -1: the whole update in this launch (all terms, rotation of the rows included, final sort).
    // Split mode (compact representation): t_only >= 0 processes term t_only only -- deflation, secular roots,
    // Qh into qwork, the row map and r into aux -- and leaves the rotation of the rows (a batched GEMM) to
    // secular_apply_kernel / secular_copyback_kernel; the eigenvalues stay in ROW order in evals between
    // launches.  t_only == -2: only the final sort + row permutation.
    // m = number of (explicit) eigenpairs = rows of Vt that take part; mc >= m sizes the shared arrays;
    // n = length of a row.  Dense representation: m = mc = n, estride = n, vstride = n*n.
    const int b = blockIdx.x;
    if (skip[b]) return;
    const int nterms = nterm[b];
    if (nterms == 0) return;
    const int m = mrows ? mrows[b] : n;
    if (m == 0) return;
    extern __shared__ double sm[];
    SecShared& S = *reinterpret_cast<SecShared*>(sm);
    double* d = sm + (sizeof(SecShared) + 7) / 8;   // current eigenvalues, row order
    double* z = d + mc;              // current z, row order
    double* dd = z + mc;             // non-deflated poles (ascending; mirrored when rho<0)
    double* y2 = dd + mc;            // squared weights of the non-deflated
    double* zh = y2 + mc;            // recomputed z (Gu-Eisenstat)
    double* mu = zh + mc;            // root offsets
    double* rc = mu + mc;            // rotation cosines
    double* rs = rc + mc;            // rotation sines
    int* ord = reinterpret_cast<int*>(rs + mc);   // sorted position -> row
    int* nd = ord + mc;              // non-deflated rows, ascending d
    int* org = nd + mc;              // origin pole index of each root
    int* ia = org + mc;              // rotation row a
    int* ib = ia + mc;               // rotation row b
    double* tile = reinterpret_cast<double*>(ib + mc + (mc & 1));   // staging tile for the row update
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Vt = Vt_ + (size_t)b * vstride;
    double* Z = Z_ + (size_t)b * zcap * n;
    double* work = work_ + (size_t)b * vstride;
    double* Qh = qwork_ + (size_t)b * vstride;

    long long tmark_ = clock64();
    int* aux = aux_ ? aux_ + (size_t)b * auxs : nullptr;
    if (t_only >= 0 && tid == 0) aux[0] = 0;
    for (int i = tid; i < m; i += nt) { d[i] = evals_[(size_t)b * estride + i]; ord[i] = i; }   // dense: sorted on entry
    __syncthreads();
    if (mrows) {
        // compact representation: freshly appended rows (eigenvalue lam0) sit at the end, out of order
        for (int i = tid; i < m; i += nt) {
            const double di = d[i];
            int rank = 0;
            for (int j = 0; j < m; ++j) { const double dj = d[j]; rank += (dj < di) || (dj == di && j < i); }
            ord[rank] = i;
        }
        __syncthreads();
    }

    const int t_begin = t_only >= 0 ? t_only : 0;
    const int t_end = t_only >= 0 ? min(t_only + 1, nterms) : (t_only == -2 ? 0 : nterms);
    for (int t = t_begin; t < t_end; ++t) {
        double* zt = Z + (size_t)t * n;
        double acc = 0.0;
        for (int i = tid; i < m; i += nt) { const double v = zt[i]; z[i] = v; acc = fma(v, v, acc); }
        const double znorm2 = sb_block_sum(acc, S.scratch);
        double rho = sig_[(size_t)b * zcap + t] * znorm2;
        if (znorm2 == 0.0 || rho == 0.0) continue;
        const double zinv = 1.0 / sqrt(znorm2);
        double dmax = 0.0, zmax = 0.0;
        for (int i = tid; i < m; i += nt) { z[i] *= zinv; dmax = fmax(dmax, fabs(d[i])); zmax = fmax(zmax, fabs(z[i])); }
        dmax = sb_warp_max(dmax); zmax = sb_warp_max(zmax);
        __syncthreads();
        if ((tid & 31) == 0) { S.scratch[40 + (tid >> 5)] = dmax; S.scratch[52 + (tid >> 5)] = zmax; }
        __syncthreads();
        dmax = 0.0; zmax = 0.0;
        for (int w = 0; w < nt / 32; ++w) { dmax = fmax(dmax, S.scratch[40 + w]); zmax = fmax(zmax, S.scratch[52 + w]); }
        const double tol = 8.0 * SEC_EPS * fmax(dmax, fabs(rho));
        if (fabs(rho) * zmax <= tol) continue;          // the whole term is negligible
        SEC_MARK(0);
        // ---------------- deflation, stage 1: rows whose z is negligible keep their
        // eigenpair; runs of (numerically) equal eigenvalues among the rest are an exact
        // eigenspace, so ONE Householder reflection of those eigenvectors can move all
        // of z's weight onto a single row (instead of a chain of plane rotations).
        // candidates (non-negligible z) in ascending-d order: ballot compaction
        if (tid == 0) S.flag = 0;
        __syncthreads();
        for (int p0 = 0; p0 < m; p0 += nt) {
            const int p = p0 + tid;
            const int i = p < m ? ord[p] : 0;
            const bool f = p < m && fabs(rho * z[i]) > tol;
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if ((tid & 31) == 0) S.wcnt[tid >> 5] = __popc(bal);
            __syncthreads();
            int off = S.flag;
            for (int w = 0; w < (tid >> 5); ++w) off += S.wcnt[w];
            if (f) nd[off + __popc(bal & ((1u << (tid & 31)) - 1u))] = i;
            __syncthreads();
            if (tid == 0) { int tot = 0; for (int w = 0; w < nt / 32; ++w) tot += S.wcnt[w]; S.flag += tot; }
            __syncthreads();
        }
        if (tid == 0) {
            const int na = S.flag;
            int nc = 0, ncand = 0;
            const double tolc = 8.0 * SEC_EPS * dmax;
            int a = 0;
            while (a < na) {
                int e = a;
                while (e + 1 < na && d[nd[e + 1]] - d[nd[a]] <= tolc) ++e;
                if (e > a) { ia[nc] = a; ib[nc] = e - a + 1; ++nc; }   // cluster [a, e] in nd
                // the survivor's eigenvalue moves up for rho > 0 and down for rho < 0: take the
                // row at that end of the run so that it does not have to cross the cluster
                org[ncand++] = rho > 0.0 ? nd[e] : nd[a];
                a = e + 1;
            }
            S.r = ncand; S.nrot = nc; S.flag = na;
        }
        __syncthreads();
        SEC_MARK(1);
        {
            const int nclus = S.nrot;
            for (int c = 0; c < nclus; ++c) {
                const int a0 = ia[c], m = ib[c];
                const int piv = rho > 0.0 ? m - 1 : 0;
                const int last = nd[a0 + piv];
                // w = z_c - alpha e_piv, alpha = -sign(z_piv) |z_c|
                double part = 0.0;
                for (int i = tid; i < m; i += nt) { const double v = z[nd[a0 + i]]; part = fma(v, v, part); }
                const double beta = sqrt(sb_block_sum(part, S.scratch));
                const double zl = z[last];
                const double alpha = zl > 0.0 ? -beta : beta;
                const double wl = zl - alpha;
                const double wn2 = beta * beta - zl * zl + wl * wl;
                const double coef = 2.0 / wn2;
                for (int i = tid; i < m; i += nt) zh[i] = (i == piv) ? wl : z[nd[a0 + i]];   // w in cluster order
                __syncthreads();
                // y = coef * sum_i w_i row_i ; row_i -= w_i y.   Warps own rows, lanes own
                // columns (8 per lane and chunk): every row access is one coalesced burst
                // and each lane keeps 8 independent loads in flight.
                {
                    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
                    constexpr int LPC = 8;                       // columns per lane per chunk
                    const int chunk = 32 * LPC;
                    double* ypart = tile;                        // [nw][chunk]
                    double* ysum_ = tile + (size_t)nw * chunk;   // [chunk]
                    for (int c0 = 0; c0 < n; c0 += chunk) {
                        double yp[LPC];
#pragma unroll
                        for (int q = 0; q < LPC; ++q) yp[q] = 0.0;
                        {
                            constexpr int RU = 2;                // rows in flight per warp
                            for (int i0 = warp; i0 < m; i0 += nw * RU) {
                                double v[RU][LPC];
                                double wv[RU];
#pragma unroll
                                for (int u = 0; u < RU; ++u) {
                                    const int i = i0 + u * nw;
                                    const bool live = i < m;
                                    wv[u] = live ? zh[i] : 0.0;
                                    const double* row = Vt + (size_t)nd[a0 + (live ? i : 0)] * n + c0;
#pragma unroll
                                    for (int q = 0; q < LPC; ++q) {
                                        const int cc = lane + 32 * q;
                                        v[u][q] = (live && c0 + cc < n) ? row[cc] : 0.0;
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < RU; ++u)
#pragma unroll
                                    for (int q = 0; q < LPC; ++q) yp[q] = fma(wv[u], v[u][q], yp[q]);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < LPC; ++q) ypart[warp * chunk + lane + 32 * q] = yp[q];
                        __syncthreads();
                        for (int cc = tid; cc < chunk; cc += nt) {
                            double a = 0.0;
                            for (int w2 = 0; w2 < nw; ++w2) a += ypart[w2 * chunk + cc];
                            ysum_[cc] = a * coef;
                        }
                        __syncthreads();
#pragma unroll
                        for (int q = 0; q < LPC; ++q) yp[q] = ysum_[lane + 32 * q];
                        {
                            constexpr int RU = 2;
                            for (int i0 = warp; i0 < m; i0 += nw * RU) {
                                double v[RU][LPC];
#pragma unroll
                                for (int u = 0; u < RU; ++u) {
                                    const int i = i0 + u * nw;
                                    const bool live = i < m;
                                    const double* row = Vt + (size_t)nd[a0 + (live ? i : 0)] * n + c0;
#pragma unroll
                                    for (int q = 0; q < LPC; ++q) {
                                        const int cc = lane + 32 * q;
                                        v[u][q] = (live && c0 + cc < n) ? row[cc] : 0.0;
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < RU; ++u) {
                                    const int i = i0 + u * nw;
                                    if (i >= m) continue;
                                    const double wi = zh[i];
                                    double* row = Vt + (size_t)nd[a0 + i] * n + c0;
#pragma unroll
                                    for (int q = 0; q < LPC; ++q) {
                                        const int cc = lane + 32 * q;
                                        if (c0 + cc < n) row[cc] = fma(-wi, yp[q], v[u][q]);
                                    }
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
                // pending z vectors of later terms
                for (int s2 = t + 1 + tid; s2 < nterms; s2 += nt) {
                    double* zp = Z + (size_t)s2 * n;
                    double dot = 0.0;
                    for (int i = 0; i < m; ++i) dot = fma(zh[i], zp[nd[a0 + i]], dot);
                    dot *= coef;
                    for (int i = 0; i < m; ++i) zp[nd[a0 + i]] = fma(-zh[i], dot, zp[nd[a0 + i]]);
                }
                __syncthreads();
                for (int i = tid; i < m; i += nt) z[nd[a0 + i]] = (i == piv) ? alpha : 0.0;
                __syncthreads();
            }
        }
        SEC_MARK(2);
        // every thread must have read S.nrot (the cluster count of stage 1) before thread 0 reuses it below:
        // with no cluster there is no barrier inside the loop above (compute-sanitizer racecheck)
        __syncthreads();
        // ---------------- deflation, stage 2 (LAPACK dlaed2): neighbouring survivors whose
        // eigenvalues are close enough are combined by a plane rotation
        if (tid == 0) {
            const int ncand = S.r;
            int r = 0, nrot = 0, prev = -1;
            for (int p = 0; p < ncand; ++p) {
                const int i = org[p];
                if (fabs(rho * z[i]) <= tol) continue;
                if (prev < 0) { prev = i; continue; }
                double s = z[prev], c = z[i];
                const double tt = d[i] - d[prev];
                // |tt c s| <= tol with c = z_i/tau, s = -z_prev/tau, tau^2 = z_i^2 + z_prev^2: test without the
                // square root and the divisions first (almost no pair merges)
                if (!(fabs(tt * c * s) <= 2.0 * tol * (c * c + s * s))) { nd[r++] = prev; prev = i; continue; }
                const double tau = hypot(c, s);
                c /= tau; s = -s / tau;
                if (fabs(tt * c * s) <= tol) {
                    z[i] = tau; z[prev] = 0.0;
                    ia[nrot] = prev; ib[nrot] = i; rc[nrot] = c; rs[nrot] = s; ++nrot;
                    const double tnew = d[prev] * c * c + d[i] * s * s;
                    d[i] = d[prev] * s * s + d[i] * c * c;
                    d[prev] = tnew;
                    prev = i;
                } else {
                    nd[r++] = prev;
                    prev = i;
                }
            }
            if (prev >= 0) nd[r++] = prev;
            S.r = r; S.nrot = nrot; S.rho = rho;
        }
        __syncthreads();
        const int r = S.r, nrot = S.nrot;
        if (nrot > 0) {
            apply_rotations<CPT>(Vt, n, ia, ib, rc, rs, nrot);
            // pending z vectors of later terms see the same rotations
            for (int s2 = t + 1 + tid; s2 < nterms; s2 += nt) {
                double* zp = Z + (size_t)s2 * n;
                for (int q = 0; q < nrot; ++q) {
                    const double x = zp[ia[q]], y = zp[ib[q]];
                    zp[ia[q]] = rc[q] * x + rs[q] * y;
                    zp[ib[q]] = rc[q] * y - rs[q] * x;
                }
            }
        }
        __syncthreads();
        if (r == 0) continue;
        // non-deflated d may be slightly out of order after rotations changed d: insertion sort (r small shifts)
        if (tid == 0) {
            for (int a = 1; a < r; ++a) {
                const int row = nd[a];
                int p = a - 1;
                while (p >= 0 && d[nd[p]] > d[row]) { nd[p + 1] = nd[p]; --p; }
                nd[p + 1] = row;
            }
        }
        __syncthreads();
        SEC_MARK(3);
        // ---------------- secular equation on the r non-deflated poles
        const bool neg = rho < 0.0;
        const double arho = fabs(rho);
        acc = 0.0;
        for (int j = tid; j < r; j += nt) {
            const int src = neg ? nd[r - 1 - j] : nd[j];
            dd[j] = neg ? -d[src] : d[src];
            const double v = z[src];
            y2[j] = v * v;
            acc += v * v;
        }
        const double ysum = sb_block_sum(acc, S.scratch);
        if constexpr (SPLIT) {
            for (int j = tid; j < r; j += nt) {                // one thread per root
                int o; double m_;
                if (r == 1) { o = 0; m_ = arho * y2[0]; }
                else secular_root_serial(dd, y2, r, arho, j, ysum, &o, &m_);
                org[j] = o; mu[j] = m_;
            }
        } else {
            for (int j = tid >> 5; j < r; j += nt >> 5) {      // one warp per root
                int o; double m_;
                if (r == 1) { o = 0; m_ = arho * y2[0]; }
                else secular_root(dd, y2, r, arho, j, ysum, &o, &m_);
                if ((tid & 31) == 0) { org[j] = o; mu[j] = m_; }
            }
        }
        __syncthreads();
        // Gu-Eisenstat: zh_i^2 = (lam_i - dd_i)/rho * prod_{j != i} (lam_j - dd_i)/(dd_j - dd_i)
        for (int i = tid; i < r; i += nt) {
            double prod = ((dd[org[i]] - dd[i]) + mu[i]) / arho;
            {
                const double di = dd[i];
                double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
                auto fac = [&](int j) {
                    const bool self = j == i;
                    return self ? 1.0 : ((dd[org[j]] - di) + mu[j]) * sec_rcp(dd[j] - di);
                };
                int j = 0;
                for (; j + 3 < r; j += 4) { p0 *= fac(j); p1 *= fac(j + 1); p2 *= fac(j + 2); p3 *= fac(j + 3); }
                for (; j < r; ++j) p0 *= fac(j);
                prod *= (p0 * p1) * (p2 * p3);
            }
            const int src = neg ? nd[r - 1 - i] : nd[i];
            zh[i] = copysign(sqrt(fabs(prod)), z[src]);
        }
        __syncthreads();
        // eigenvector matrix (mirrored index space): Qh[i*r + j] = zh_i / (dd_i - lam_j), columns normalised
        const bool qsmem = !SPLIT && r <= SEC_QS_MAX && (size_t)r * r <= (size_t)tile_doubles;
        if (qsmem) Qh = tile;
        else Qh = qwork_ + (size_t)b * vstride;
        // column norms straight from (zh, dd, roots) in shared memory, then Qh written once, normalised
        for (int j = tid; j < r; j += nt) {
            const double base = dd[org[j]], m_ = mu[j];
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
            int i = 0;
            for (; i + 3 < r; i += 4) {
                const double q0 = zh[i] * sec_rcp((dd[i] - base) - m_), q1 = zh[i + 1] * sec_rcp((dd[i + 1] - base) - m_);
                const double q2 = zh[i + 2] * sec_rcp((dd[i + 2] - base) - m_), q3 = zh[i + 3] * sec_rcp((dd[i + 3] - base) - m_);
                n0 = fma(q0, q0, n0); n1 = fma(q1, q1, n1); n2 = fma(q2, q2, n2); n3 = fma(q3, q3, n3);
            }
            for (; i < r; ++i) { const double q = zh[i] * sec_rcp((dd[i] - base) - m_); n0 = fma(q, q, n0); }
            y2[j] = 1.0 / sqrt((n0 + n1) + (n2 + n3));     // y2 is free again (weights consumed by the roots / zh)
        }
        __syncthreads();
        for (int idx = tid; idx < r * r; idx += nt) {
            const int i = idx / r, j = idx % r;
            Qh[idx] = zh[i] * sec_rcp((dd[i] - dd[org[j]]) - mu[j]) * y2[j];
        }
        __syncthreads();
        SEC_MARK(4);
        // ---------------- new eigenvectors: row_new(j) = sum_i Qh[i][j] row_old(i)   (mirrored index i,j)
        {
            auto rowof = [&](int i) { return neg ? nd[r - 1 - i] : nd[i]; };
            if constexpr (SPLIT) {
                // split mode: hand (r, row map, Qh) to the GEMM kernels
                for (int i = tid; i < r; i += nt) aux[4 + i] = rowof(i);
                if (tid == 0) aux[0] = r;
            } else {
            // out[j][col] = sum_i Qh[i][j] old[rowof(i)][col], column chunks staged in smem
            {
                // warps own blocks of 8 new rows, lanes own 4 columns of a 128-column chunk; old rows
                // stream from global/L1 (each is read once per row block), Qh from shared memory
                // (r <= SEC_QS_MAX) or from L1/L2 (8 consecutive doubles per old row, broadcast);
                // results go through `work` and are copied back below
                constexpr int JB = 8, CC = 4, IU = 2;
                const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
                const int nblk = (r + JB - 1) / JB;
                for (int c0 = 0; c0 < n; c0 += 32 * CC) {
                    for (int blk0 = 0; blk0 < nblk; blk0 += nw) {
                        // all row blocks of this chunk must be computed before any is stored: when
                        // nblk > nw the extra rounds go through `work`
                        const int blk = blk0 + warp;
                        const int jb0 = blk * JB;
                        if (blk < nblk) {
                            double acc[JB][CC];
#pragma unroll
                            for (int jj = 0; jj < JB; ++jj)
#pragma unroll
                                for (int q = 0; q < CC; ++q) acc[jj][q] = 0.0;
                            for (int i0 = 0; i0 < r; i0 += IU) {
                                double x[IU][CC];
#pragma unroll
                                for (int u = 0; u < IU; ++u) {
                                    const int i = i0 + u;
                                    const double* src = Vt + (size_t)rowof(i < r ? i : 0) * n + c0;
#pragma unroll
                                    for (int q = 0; q < CC; ++q) {
                                        const int col = lane + 32 * q;
                                        x[u][q] = (i < r && c0 + col < n) ? src[col] : 0.0;
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < IU; ++u) {
                                    const int i = i0 + u;
                                    if (i >= r) break;
                                    const double* qrow = Qh + (size_t)i * r + jb0;
#pragma unroll
                                    for (int jj = 0; jj < JB; ++jj) {
                                        const double qv = (jb0 + jj < r) ? qrow[jj] : 0.0;
#pragma unroll
                                        for (int q = 0; q < CC; ++q) acc[jj][q] = fma(qv, x[u][q], acc[jj][q]);
                                    }
                                }
                            }
#pragma unroll
                            for (int jj = 0; jj < JB; ++jj) {
                                if (jb0 + jj >= r) break;
#pragma unroll
                                for (int q = 0; q < CC; ++q) {
                                    const int col = lane + 32 * q;
                                    if (c0 + col < n) work[(size_t)(jb0 + jj) * n + c0 + col] = acc[jj][q];
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            for (int idx = tid; idx < r * n; idx += nt) {
                const int j = idx / n, col = idx % n;
                Vt[(size_t)rowof(j) * n + col] = work[(size_t)j * n + col];
            }
            }
            SEC_MARK(5);
            // pending z vectors: z_s[row(j)] <- sum_i Qh[i][j] z_s[row(i)]
            for (int s2 = t + 1; s2 < nterms; ++s2) {
                double* zp = Z + (size_t)s2 * n;
                __syncthreads();
                for (int j = tid; j < r; j += nt) {
                    double a = 0.0;
                    for (int i = 0; i < r; ++i) a = fma(Qh[(size_t)i * r + j], zp[rowof(i)], a);
                    zh[j] = a;                // reuse zh as temp (roots already consumed)
                }
                __syncthreads();
                for (int j = tid; j < r; j += nt) zp[rowof(j)] = zh[j];
            }
            __syncthreads();
            // new eigenvalues (undo the mirror)
            for (int j = tid; j < r; j += nt) {
                const double lam = dd[org[j]] + mu[j];
                d[rowof(j)] = neg ? -lam : lam;
            }
        }
        __syncthreads();
        SEC_MARK(6);
        // ---------------- ascending order of the rows for the next term
        for (int i = tid; i < m; i += nt) {
            const double di = d[i];
            int rank = 0;
            for (int j = 0; j < m; ++j) { const double dj = d[j]; rank += (dj < di) || (dj == di && j < i); }
            ord[rank] = i;
        }
        __syncthreads();
    }
    SEC_MARK(7);
    if (t_only >= 0) {
        for (int i = tid; i < m; i += nt) evals_[(size_t)b * estride + i] = d[i];      // row order, until the final sort
        return;
    }
    // ---------------- write back: evals ascending, rows of Vt permuted (new[p] = old[ord[p]])
    // Rows with equal eigenvalues are interchangeable: a row that already sits in a slot
    // whose sorted value equals its own eigenvalue stays put, only the others are matched
    // (in ascending order) to the remaining slots.  This keeps a degenerate cluster in
    // place when one row crosses it, instead of shifting every member by one.
    for (int p = tid; p < m; p += nt) dd[p] = d[ord[p]];            // sorted values
    __syncthreads();
    for (int p = tid; p < m; p += nt) {
        evals_[(size_t)b * estride + p] = dd[p];
        org[p] = (d[p] == dd[p]) ? 1 : 0;                             // fixed point
    }
    __syncthreads();
    for (int p = tid; p < m; p += nt) {
        if (org[p]) continue;
        int kslot = 0;
        for (int q = 0; q < p; ++q) kslot += !org[q];
        ia[kslot] = p;                                                // kslot-th free slot
    }
    __syncthreads();
    for (int i = tid; i < m; i += nt) {
        if (org[i]) { ib[i] = i; continue; }
        const double di = d[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            if (org[j]) continue;
            const double dj = d[j];
            rank += (dj < di) || (dj == di && j < i);
        }
        ib[ia[rank]] = i;                                             // slot -> row
    }
    __syncthreads();
    for (int p = tid; p < m; p += nt) ord[p] = ib[p];
    if (tid == 0) S.flag = 0;
    __syncthreads();
    for (int p2 = tid; p2 < m; p2 += nt)
        if (ord[p2] != p2) {
            nd[atomicAdd(&S.flag, 1)] = p2;          // destinations that change
            const int dist = abs(ord[p2] - p2);
            atomicAdd(&sec_prof[11], (unsigned long long)dist);
            if (dist == 1) atomicAdd(&sec_prof[12], 1ull);
            if (d[ord[p2]] == d[p2]) atomicAdd(&sec_prof[13], 1ull);   // moved although eigenvalue equal to the old occupant
        }
    __syncthreads();
    {
        const int nm = S.flag;
        if (tid == 0) { atomicAdd(&sec_prof[9], (unsigned long long)nm); atomicAdd(&sec_prof[10], 1ull); }
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        for (int e = warp; e < nm; e += nw) {
            const double* src = Vt + (size_t)ord[nd[e]] * n;
            double* dst = work + (size_t)e * n;
            for (int col = lane; col < n; col += 32) dst[col] = src[col];
        }
        __syncthreads();
        for (int e = warp; e < nm; e += nw) {
            const double* src = work + (size_t)e * n;
            double* dst = Vt + (size_t)nd[e] * n;
            for (int col = lane; col < n; col += 32) dst[col] = src[col];
        }
    }
    __syncthreads();
    SEC_MARK(8);
    (void)status;
}

// ------------------------------------------------------------------------------
// Rotation of the eigenvector rows of one rank-one term (split mode), a batched GEMM on the fp64
// tensor-core path (DMMA m8n8k4):
//     work[j][c] = sum_i Qh[i][j] * Vt[rowmap[i]][c],   j < r, c < n        (r, rowmap, Qh per system)
// 64 x 64 output tiles, K blocks of 16 staged through a 3-deep cp.async ring (8-byte copies: Qh rows
// start at arbitrary offsets), 8 warps of 16 x 32, padded shared tiles (row stride 68 doubles: the
// fragment reads of a half warp fall into 16 different bank pairs).
constexpr int AP_BM = 64, AP_BN = 64, AP_BK = 16, AP_THREADS = 256, AP_LD = 68, AP_STAGES = 3;

__device__ __forceinline__ void sec_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void sec_cp8(double* dst, const double* src, bool valid) {
    const int sz = valid ? 8 : 0;                        // src-size 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sb_smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(AP_THREADS)
secular_apply_kernel(const double* __restrict__ Vt_, const double* __restrict__ qwork_, double* __restrict__ work_,
                     const int* __restrict__ aux_, int auxs, int n, long long vstride, const int* __restrict__ skip) {
    const int b = blockIdx.z;
    if (skip && skip[b]) return;
    const int* aux = aux_ + (size_t)b * auxs;
    const int r = aux[0];
    const int m0 = blockIdx.y * AP_BM, n0 = blockIdx.x * AP_BN;
    if (m0 >= r) return;
    extern __shared__ double sm[];
    double* As = sm;                                              // [STAGES][BK][LD]   As[k][j]
    double* Bs = sm + (size_t)AP_STAGES * AP_BK * AP_LD;          // [STAGES][BK][LD]   Bs[k][c]
    const double* Vt = Vt_ + (size_t)b * vstride;
    const double* Qh = qwork_ + (size_t)b * vstride;
    double* work = work_ + (size_t)b * vstride;
    const int* rowmap = aux + 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
    const int gid = lane >> 2, tig = lane & 3;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const int nk = (r + AP_BK - 1) / AP_BK;
    auto stage = [&](int kb, int slot) {
        const int k0 = kb * AP_BK;
        double* as = As + (size_t)slot * AP_BK * AP_LD;
        double* bs = Bs + (size_t)slot * AP_BK * AP_LD;
#pragma unroll
        for (int e = tid; e < AP_BK * AP_BM; e += AP_THREADS) {
            const int kk = e / AP_BM, mm = e % AP_BM;
            const int gk = k0 + kk, gm = m0 + mm;
            const bool ok = gk < r && gm < r;
            sec_cp8(as + kk * AP_LD + mm, Qh + (ok ? (size_t)gk * r + gm : 0), ok);
        }
#pragma unroll
        for (int e = tid; e < AP_BK * AP_BN; e += AP_THREADS) {
            const int kk = e / AP_BN, nn = e % AP_BN;
            const int gk = k0 + kk, gn = n0 + nn;
            const bool ok = gk < r && gn < n;
            sec_cp8(bs + kk * AP_LD + nn, Vt + (ok ? (size_t)rowmap[gk] * n + gn : 0), ok);
        }
    };
#pragma unroll
    for (int p = 0; p < AP_STAGES - 1; ++p) {
        if (p < nk) stage(p, p);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < nk; ++kb) {
        asm volatile("cp.async.wait_group %0;" ::"n"(AP_STAGES - 2) : "memory");
        __syncthreads();
        {
            const int nxt = kb + AP_STAGES - 1;
            if (nxt < nk) stage(nxt, nxt % AP_STAGES);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const double* as = As + (size_t)(kb % AP_STAGES) * AP_BK * AP_LD;
        const double* bs = Bs + (size_t)(kb % AP_STAGES) * AP_BK * AP_LD;
#pragma unroll
        for (int ks = 0; ks < AP_BK; ks += 4) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = as[(ks + tig) * AP_LD + wm + 8 * i + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = bs[(ks + tig) * AP_LD + wn + 8 * j + gid];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sec_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + wm + 8 * i + gid, gn = n0 + wn + 8 * j + 2 * tig;
            if (gm < r) {
                if (gn + 1 < n && !(n & 1)) *reinterpret_cast<double2*>(work + (size_t)gm * n + gn) = make_double2(acc[i][j][0], acc[i][j][1]);
                else {
                    if (gn < n) work[(size_t)gm * n + gn] = acc[i][j][0];
                    if (gn + 1 < n) work[(size_t)gm * n + gn + 1] = acc[i][j][1];
                }
            }
        }
}

// Vt[rowmap[j]][:] = work[j][:] for j < r  (warp per row)
__global__ void secular_copyback_kernel(double* __restrict__ Vt_, const double* __restrict__ work_,
                                        const int* __restrict__ aux_, int auxs, int n, long long vstride,
                                        const int* __restrict__ skip) {
    const int b = blockIdx.y;
    if (skip && skip[b]) return;
    const int* aux = aux_ + (size_t)b * auxs;
    const int r = aux[0];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int j = blockIdx.x * nw + warp;
    if (j >= r) return;
    const double* src = work_ + (size_t)b * vstride + (size_t)j * n;
    double* dst = Vt_ + (size_t)b * vstride + (size_t)aux[4 + j] * n;
    if (!(n & 1)) {
        for (int c = lane; c < (n >> 1); c += 32) reinterpret_cast<double2*>(dst)[c] = reinterpret_cast<const double2*>(src)[c];
    } else {
        for (int c = lane; c < n; c += 32) dst[c] = src[c];
    }
}

__global__ void apply_bench_fill_kernel(int* __restrict__ aux_, int auxs, int r) {
    int* aux = aux_ + (size_t)blockIdx.x * auxs;
    if (threadIdx.x == 0) aux[0] = r;
    for (int i = threadIdx.x; i < r; i += blockDim.x) aux[4 + i] = i;
}

}  // namespace

// diagnostic (bench.py roofline of the rotation GEMM): `reps` launches of secular_apply_kernel with r rows for
// every system (identity row map; Vt, qwork hold whatever they hold), average milliseconds per launch.
extern "C" int sb_secular_apply_bench_impl(const double* Vt, const double* qwork, double* work, int* aux, int r,
                                           int n, long long vstride, int batch, int reps, float* ms_out,
                                           cudaStream_t st) {
    const int auxs = r + 4;
    apply_bench_fill_kernel<<<batch, 128, 0, st>>>(aux, auxs, r);
    const size_t apsm = (size_t)2 * AP_STAGES * AP_BK * AP_LD * sizeof(double);
    cudaFuncSetAttribute(secular_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apsm);
    dim3 ga((n + AP_BN - 1) / AP_BN, (r + AP_BM - 1) / AP_BM, batch);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    secular_apply_kernel<<<ga, AP_THREADS, apsm, st>>>(Vt, qwork, work, aux, auxs, n, vstride, nullptr);
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps; ++i)
        secular_apply_kernel<<<ga, AP_THREADS, apsm, st>>>(Vt, qwork, work, aux, auxs, n, vstride, nullptr);
    cudaEventRecord(e1, st);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) return (int)err;
    *ms_out = ms / reps;
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_secular_profile_impl(unsigned long long* out16, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, sec_prof, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return (int)e;
    if (reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(sec_prof, z, sizeof(z));
    }
    return (int)e;
}

extern "C" int sb_lowrank_factor_impl(const double* U, const double* J, const double* Cmat, int kcap, const int* kvec,
                                      int n, double* P, double* sig, int* nterm, const int* skip, int batch,
                                      cudaStream_t st) {
    const size_t smem = ((sizeof(FactorShared) + 15) / 16) * 16 + (size_t)n * sizeof(double);
    cudaFuncSetAttribute(lowrank_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    lowrank_factor_kernel<<<batch, SEC_THREADS, smem, st>>>(U, J, Cmat, kcap, kvec, n, P, sig, nterm, skip);
    return SB_LAUNCH_CHECK();
}

// optional per-kernel timing of the three kernels of one eigen-update (bench.py roofline)
static int sec_timing_on = 0;
static cudaEvent_t sec_ev[4];
extern "C" int sb_secular_timing_impl(float* out3, int enable) {
    if (enable > 0 && !sec_timing_on) {
        for (int i = 0; i < 4; ++i) cudaEventCreate(&sec_ev[i]);
        sec_timing_on = 1;
        return 0;
    }
    if (!sec_timing_on) return -1;
    cudaError_t e = cudaEventSynchronize(sec_ev[3]);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&out3[i], sec_ev[i], sec_ev[i + 1]);
    if (enable == 0) sec_timing_on = 0;
    return 0;
}

// Compact representation: only the first mrows[b] rows of Vt (explicit eigenpairs) take part; mcap is the
// host-side bound on mrows that sizes the shared arrays.  mrows == NULL: dense (m = n).
extern "C" int sb_secular_update_c_impl(double* evals, double* Vt, double* Z, int zcap, const double* sig,
                                        const int* nterm, int n, double* work, double* qwork, int* status,
                                        const int* skip, const int* mrows, int mcap, long long estride,
                                        long long vstride, int batch, cudaStream_t st, int nterm_max, int* aux) {
    const int mc = mrows ? (mcap < 1 ? 1 : (mcap > n ? n : mcap)) : n;
    const size_t base = (size_t)mc * (8 * sizeof(double) + 5 * sizeof(int)) + sizeof(SecShared) + 64;
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // 23 KB of staging: Qh of up to 54 x 54 in shared memory, and 4 CTAs per SM at n = 384
    size_t tile_bytes = 23 * 1024;
    if (base + tile_bytes > (size_t)optin) tile_bytes = base < (size_t)optin ? ((size_t)optin - base) / 1024 * 1024 : 0;
    const int tile_doubles = (int)(tile_bytes / sizeof(double));
    if (tile_doubles < (SECK_THREADS / 32 + 1) * 256) return -2;   // n too large for this build
    const size_t smem = base + tile_bytes;
    const int cpt = (n + SECK_THREADS - 1) / SECK_THREADS;
    if (sec_timing_on) cudaEventRecord(sec_ev[0], st);
    if (!mrows && n >= 32 && n < 4096 && !getenv("SB_NO_CLUSTER_QR")) {
        // pre-phase: one block reflector for the degenerate cluster (all terms at once)
        const size_t qsm = ((size_t)n + SB_SCRATCH_DOUBLES + 2 * CQ_TMAX * CQ_TMAX + 2 * CQ_TMAX) * sizeof(double) + 64 +
                           ((size_t)n + 20) * sizeof(int);
        cudaFuncSetAttribute(cluster_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsm);
        SB_COUNT(1);
        cluster_qr_kernel<<<batch, SEC_THREADS, qsm, st>>>(evals, Z, zcap, nterm, n, work, qwork, skip);
        if (sec_timing_on) cudaEventRecord(sec_ev[1], st);
        dim3 grid((n + CR_THREADS - 1) / CR_THREADS, batch);
        const int tmax = zcap < CQ_TMAX ? zcap : CQ_TMAX;
        SB_COUNT(1);
        cluster_reflect_kernel<2><<<grid, CR_THREADS, 0, st>>>(Vt, work, qwork, n);
        if (tmax > 2) { SB_COUNT(1); cluster_reflect_kernel<4><<<grid, CR_THREADS, 0, st>>>(Vt, work, qwork, n); }
        if (tmax > 4) { SB_COUNT(1); cluster_reflect_kernel<8><<<grid, CR_THREADS, 0, st>>>(Vt, work, qwork, n); }
        if (tmax > 8) { SB_COUNT(1); cluster_reflect_kernel<16><<<grid, CR_THREADS, 0, st>>>(Vt, work, qwork, n); }
    } else if (sec_timing_on) {
        cudaEventRecord(sec_ev[1], st);
    }
    if (sec_timing_on) cudaEventRecord(sec_ev[2], st);
    SB_COUNT(1);
    const long long es = mrows ? estride : (long long)n;
    const long long vs = mrows ? vstride : (long long)n * n;
    const int auxs = mc + 4;
#define SB_SEC_LAUNCH(C, TONLY)                                                                                   \
    if ((TONLY) == -1) {                                                                                          \
        cudaFuncSetAttribute(secular_update_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        secular_update_kernel<C, false><<<batch, SECK_THREADS, smem, st>>>(evals, Vt, Z, zcap, sig, nterm, n, work,   \
                                                                          qwork, status, skip, tile_doubles, mrows,   \
                                                                          mc, es, vs, TONLY, aux, auxs);              \
    } else {                                                                                                      \
        cudaFuncSetAttribute(secular_update_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        secular_update_kernel<C, true><<<batch, SECK_THREADS, smem, st>>>(evals, Vt, Z, zcap, sig, nterm, n, work,    \
                                                                         qwork, status, skip, tile_doubles, mrows,    \
                                                                         mc, es, vs, TONLY, aux, auxs);               \
    }
#define SB_SEC_DISPATCH(TONLY)                          \
    if (cpt <= 1) { SB_SEC_LAUNCH(1, TONLY); }          \
    else if (cpt <= 2) { SB_SEC_LAUNCH(2, TONLY); }     \
    else if (cpt <= 4) { SB_SEC_LAUNCH(4, TONLY); }     \
    else if (cpt <= 8) { SB_SEC_LAUNCH(8, TONLY); }     \
    else if (cpt <= 16) { SB_SEC_LAUNCH(16, TONLY); }   \
    else return -2;
    if (mrows && aux && nterm_max > 0) {
        // split mode: per term one solve launch (CTA per system), the rotation of the rows as a batched
        // DMMA GEMM over the whole GPU and the copy back; the final sort + row permutation last
        const size_t apsm = (size_t)2 * AP_STAGES * AP_BK * AP_LD * sizeof(double);
        cudaFuncSetAttribute(secular_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)apsm);
        const int tmax = nterm_max < zcap ? nterm_max : zcap;
        for (int t = 0; t < tmax; ++t) {
            SB_COUNT(3);
            SB_SEC_DISPATCH(t)
            dim3 ga((n + AP_BN - 1) / AP_BN, (mc + AP_BM - 1) / AP_BM, batch);
            secular_apply_kernel<<<ga, AP_THREADS, apsm, st>>>(Vt, qwork, work, aux, auxs, n, vs, skip);
            dim3 gc((mc + 7) / 8, batch);
            secular_copyback_kernel<<<gc, 256, 0, st>>>(Vt, work, aux, auxs, n, vs, skip);
        }
        // no final sort: in the compact representation the explicit pairs may sit in any order (evals[i]
        // belongs to row i); every consumer ranks them on the fly
    } else {
        SB_SEC_DISPATCH(-1)
    }
#undef SB_SEC_DISPATCH
#undef SB_SEC_LAUNCH
    if (sec_timing_on) cudaEventRecord(sec_ev[3], st);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_secular_update_impl(double* evals, double* Vt, double* Z, int zcap, const double* sig,
                                      const int* nterm, int n, double* work, double* qwork, int* status,
                                      const int* skip, int batch, cudaStream_t st) {
    return sb_secular_update_c_impl(evals, Vt, Z, zcap, sig, nterm, n, work, qwork, status, skip, nullptr, n, n,
                                    (long long)n * n, batch, st, 0, nullptr);
}
