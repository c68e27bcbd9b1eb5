// Rational-function step models in the eigenbasis of the (projected) Hessian:
//   RationalFunctionOptimization   sella/optimize/stepper.py:114-157
//   PartitionedRFO (Sella's default for saddles)   sella/optimize/stepper.py:160-185
// combined with the restricted-step search of sella/optimize/restricted_step.py:78-121
// (newton_safe = False: Newton for the first iterations, then bisection down to a
// collapsed bracket, tol = 1e-15) and the spherical trust region (:136-142).
//
// The reference diagonalises the (m+1)x(m+1) bordered matrix [[a^2 H, a g],[a g^T, 0]]
// with a dense eigh for EVERY alpha (its slowest piece: 53 ms at 3N=384 per
// evaluation, dozens of evaluations per step).  In the eigenbasis of H the bordered
// matrix is an arrow-head matrix [[D, z],[z^T, 0]], D = a^2 diag(lam), z = a ghat, whose
// eigenvalues are the roots of  F(mu) = mu + sum_i z_i^2/(d_i - mu) = 0  (one root per
// interval between consecutive d_i) and whose eigenvector gives
//   s_i = -a^2 ghat_i / (a^2 lam_i - mu),
// with ds/da from implicit differentiation of F.  One root = O(n) work; no n^2 pass is
// needed during the alpha search for the spherical trust region (|V s| = |s|).
//
// One WARP per system (shuffle reductions, no block barriers).
#include "common.cuh"

namespace {

constexpr int RFO_WARPS = 4;
constexpr double RFO_EPS = 2.220446049250313e-16;

// Root number `idx` (0..m) of F(mu) = mu + sum_{i<m} z2_i/(d_i - mu), d ascending:
// idx = 0: (-inf, d_0); idx = m: (d_{m-1}, inf); else (d_{idx-1}, d_idx).
// Returns the pole index `org` (or -1 when m == 0) and tt with mu = d[org] + tt.
__device__ void arrow_root(const double* __restrict__ d, const double* __restrict__ z2, int m, int idx, double znorm,
                           int* org_out, double* tt_out) {
    const int lane = threadIdx.x & 31;
    if (m == 0) { *org_out = -1; *tt_out = 0.0; return; }
    int org;
    double lo, hi;           // bracket of tt (offset from d[org])
    if (idx == 0) {
        org = 0; hi = 0.0; lo = -(fabs(d[0]) + znorm);
    } else if (idx == m) {
        org = m - 1; lo = 0.0; hi = fabs(d[m - 1]) + znorm;
    } else {
        const double gap = d[idx] - d[idx - 1];
        double fm = 0.0;
        const double mid = d[idx - 1] + 0.5 * gap;
        for (int i = lane; i < m; i += 32) fm += z2[i] / (d[i] - mid);
        fm = mid + sb_warp_sum(fm);
        if (fm > 0.0) { org = idx - 1; lo = 0.0; hi = 0.5 * gap; }     // root in the left half
        else { org = idx; lo = -0.5 * gap; hi = 0.0; }
    }
    const double dorg = d[org];
    double tt = 0.5 * (lo + hi);
    for (int it = 0; it < 200; ++it) {
        double f = 0.0, df = 0.0;
        for (int i = lane; i < m; i += 32) {
            const double den = (d[i] - dorg) - tt;
            const double q = z2[i] / den;
            f += q;
            df += q / den;
        }
        f = (dorg + tt) + sb_warp_sum(f);
        df = 1.0 + sb_warp_sum(df);
        if (f == 0.0) break;
        if (f < 0.0) lo = tt; else hi = tt;
        double next = tt - f / df;
        if (!(next > lo && next < hi)) {
            // bisection, geometric next to the pole the offset is measured from
            if (lo >= 0.0) {
                if (lo > 0.0 && hi > 4.0 * lo) next = sqrt(lo) * sqrt(hi);
                else if (lo == 0.0) next = (hi > 1e-290) ? hi * 0.0625 : 0.5 * hi;
                else next = 0.5 * (lo + hi);
            } else if (hi <= 0.0) {
                if (hi < 0.0 && lo < 4.0 * hi) next = -sqrt(-lo) * sqrt(-hi);
                else if (hi == 0.0) next = (lo < -1e-290) ? lo * 0.0625 : 0.5 * lo;
                else next = 0.5 * (lo + hi);
            } else {
                next = 0.5 * (lo + hi);
            }
        }
        if (fabs(next - tt) <= 2.0 * RFO_EPS * fabs(next) || next == tt) { tt = next; break; }
        tt = next;
        if (hi - lo <= 2.0 * RFO_EPS * fmax(fabs(lo), fabs(hi))) break;
    }
    *org_out = org;
    *tt_out = tt;
}

// One RFO block on entries [i0, i0+m): writes s (the block of the step in the
// eigenbasis) and ds/dalpha; returns sum s^2 and sum s*ds via pointers.
__device__ void rfo_block(const double* __restrict__ lam, const double* __restrict__ gh, int i0, int m, int idx,
                          double alpha, double* __restrict__ dwork, double* __restrict__ zwork,
                          double* __restrict__ s, double* __restrict__ ds, double* ss_out, double* sds_out) {
    const int lane = threadIdx.x & 31;
    double zn = 0.0;
    for (int i = lane; i < m; i += 32) {
        const double di = alpha * alpha * lam[i0 + i];
        const double zi = alpha * gh[i0 + i];
        dwork[i] = di;
        zwork[i] = zi * zi;
        zn += zi * zi;
    }
    zn = sqrt(sb_warp_sum(zn));
    __syncwarp();
    int org; double tt;
    arrow_root(dwork, zwork, m, idx, zn, &org, &tt);
    // Dn_i = a^2 lam_i - mu = (d_i - d_org) - tt
    double q2 = 0.0;
    for (int i = lane; i < m; i += 32) {
        const double Dn = (dwork[i] - dwork[org]) - tt;
        const double g = gh[i0 + i];
        q2 += g * g / (Dn * Dn);
    }
    q2 = sb_warp_sum(q2);
    const double mu = (m > 0) ? dwork[org] + tt : 0.0;
    const double dmu = 2.0 * alpha * mu * q2 / (1.0 + alpha * alpha * q2);
    double ss = 0.0, sds = 0.0;
    for (int i = lane; i < m; i += 32) {
        const double Dn = (dwork[i] - dwork[org]) - tt;
        const double g = gh[i0 + i];
        const double si = -alpha * alpha * g / Dn;
        const double dsi = (-2.0 * alpha * g * Dn + alpha * alpha * g * (2.0 * alpha * lam[i0 + i] - dmu)) / (Dn * Dn);
        s[i0 + i] = si;
        ds[i0 + i] = dsi;
        ss += si * si;
        sds += si * dsi;
    }
    *ss_out = sb_warp_sum(ss);
    *sds_out = sb_warp_sum(sds);
    __syncwarp();
}

// mode 0: rfo (one arrow-head over all n entries, eigenvector index `order`);
// mode 1: prfo (maximise along the lowest `order` modes, minimise along the rest).
// Spherical trust region.  coef[b,:] = step in the eigenbasis (s = V coef).
__global__ void __launch_bounds__(RFO_WARPS * 32)
rfo_tr_kernel(const double* __restrict__ Vg_, const double* __restrict__ evals_, const double* __restrict__ delta_,
              int order, int n, int mode, double* __restrict__ coef_, double* __restrict__ smag, double* __restrict__ alpha_out,
              int* __restrict__ status, const int* __restrict__ active, int batch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int b = blockIdx.x * wpc + warp;
    if (b >= batch) return;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* lam = sm + (size_t)warp * 6 * n;
    double* gh = lam + n;
    double* dwork = gh + n;
    double* zwork = dwork + n;
    double* s = zwork + n;
    double* ds = s + n;
    for (int i = lane; i < n; i += 32) { lam[i] = evals_[(size_t)b * n + i]; gh[i] = Vg_[(size_t)b * n + i]; }
    __syncwarp();
    const double delta = delta_[b];
    const int mo = order < n ? order : n;
    auto eval = [&](double alpha, double* val, double* dval) {
        double ss, sds;
        if (mode == 0) {
            rfo_block(lam, gh, 0, n, mo, alpha, dwork, zwork, s, ds, &ss, &sds);
        } else {
            double s1, d1, s2, d2;
            rfo_block(lam, gh, 0, mo, mo, alpha, dwork, zwork, s, ds, &s1, &d1);
            rfo_block(lam, gh, mo, n - mo, 0, alpha, dwork, zwork, s, ds, &s2, &d2);
            ss = s1 + s2; sds = d1 + d2;
        }
        *val = sqrt(ss);
        *dval = sds / fmax(*val, 1e-12);
    };
    // restricted_step.py:78-121 with alpha0 = 1, [0, 1], slope = +1, newton_safe = False, tol = 1e-15
    double alpha = 1.0, val, dval;
    eval(alpha, &val, &dval);
    bool interior = val < delta;
    int st = 0;
    if (!interior) {
        double err = val - delta, lo = 0.0, hi = 1.0;
        int it = 0;
        for (;; ++it) {
            if (fabs(err) <= 1e-15) break;
            if (nextafter(lo, hi) >= hi) break;
            if (it >= 1000) { st = SB_ST_TR_NOCONV; break; }
            if (err > 0.0) hi = alpha; else lo = alpha;
            const double a1 = alpha - err / dval;
            if (isnan(a1) || a1 <= lo || a1 >= hi || it > 4) alpha = 0.5 * (lo + hi);
            else alpha = a1;
            eval(alpha, &val, &dval);
            err = val - delta;
        }
    }
    for (int i = lane; i < n; i += 32) coef_[(size_t)b * n + i] = s[i];
    if (lane == 0) {
        smag[b] = interior ? val : delta;
        alpha_out[b] = alpha;
        if (st && status) atomicOr(&status[b], st);
    }
}

}  // namespace

extern "C" int sb_rfo_tr_impl(const double* Vg, const double* evals, const double* delta, int order, int n, int mode,
                              double* coef, double* smag, double* alpha, int* status, const int* active, int batch,
                              cudaStream_t st) {
    int wpc = RFO_WARPS;
    while (wpc > 1 && (size_t)wpc * 6 * n * sizeof(double) > 200 * 1024) wpc >>= 1;
    const size_t smem = (size_t)wpc * 6 * n * sizeof(double);
    if (smem > 220 * 1024) return -2;
    cudaFuncSetAttribute(rfo_tr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    rfo_tr_kernel<<<(batch + wpc - 1) / wpc, wpc * 32, smem, st>>>(
        Vg, evals, delta, order, n, mode, coef, smag, alpha, status, active, batch);
    return SB_LAUNCH_CHECK();
}
