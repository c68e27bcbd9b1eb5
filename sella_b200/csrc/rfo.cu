// Rational-function step models in the eigenbasis of the (projected) Hessian:
//   RationalFunctionOptimization   sella/optimize/stepper.py:114-157
//   PartitionedRFO (Sella's default for saddles)   sella/optimize/stepper.py:160-185
// combined with the restricted-step search of sella/optimize/restricted_step.py:78-121
// and the spherical trust region (:136-142).
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
// The alpha search: the reference (newton_safe = False, tol = 1e-15) falls back to pure
// bisection after five Newton steps and stops when the bracket has collapsed, i.e. at
// the root of |s(alpha)| = delta to the last bit.  Here the same root is found by a
// bracketed Newton iteration with the exact derivative (a handful of evaluations instead
// of ~55); the step agrees to rounding.
//
// One WARP per system; eigenvalues and gradient components live in registers.
#include "common.cuh"

namespace {

constexpr int RFO_WARPS = 8;
constexpr double RFO_EPS = 2.220446049250313e-16;

// diagnostic counters (summed over warps by lane 0): 0 alpha evaluations, 1 arrow-head root
// iterations, 2 arrow-head roots, 3 cycles, 4 systems
__device__ unsigned long long rfo_prof[8];

template <int NPL>
struct Sys {
    double lam[NPL], g2[NPL], g[NPL];
    int n;
};

// Root of F(mu) = mu + a2 * sum_{i in [i0,i1)} g2_i / (a2*lam_i - mu) in the interval
// selected by `which`: 0 = below the lowest pole of the block, 1 = above the highest,
// 2 = between poles (idx-1, idx) (absolute indices).  mu = a2*lam[org] + tt.
// tt_guess (0 = none) warm-starts Newton.
template <int NPL>
__device__ void arrow_root(const Sys<NPL>& S, int i0, int i1, int which, int idx, double a2, double gnorm2,
                           double tt_guess, int* org_out, double* tt_out, int* iters = nullptr) {
    const int lane = threadIdx.x & 31;
    auto lam_at = [&](int i) {            // broadcast lam[i] from the owning lane
        const int q = i >> 5, src = i & 31;
        double v = 0.0;
#pragma unroll
        for (int u = 0; u < NPL; ++u) if (u == q) v = S.lam[u];
        return __shfl_sync(0xffffffffu, v, src);
    };
    int org;
    double lo, hi;
    const double znorm = sqrt(a2 * gnorm2);
    if (which == 0) {
        org = i0;
        const double d0 = a2 * lam_at(i0);
        hi = 0.0; lo = -(fabs(d0) + znorm);
    } else if (which == 1) {
        org = i1 - 1;
        const double d1 = a2 * lam_at(i1 - 1);
        lo = 0.0; hi = fabs(d1) + znorm;
    } else {
        const double la = lam_at(idx - 1), lb = lam_at(idx);
        const double gap = a2 * (lb - la);
        const double half = 0.5 * gap;
        double fm = 0.0;
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            const int i = lane + 32 * u;
            if (i >= i0 && i < i1) fm += a2 * S.g2[u] / (a2 * (S.lam[u] - la) - half);
        }
        fm = (a2 * la + half) + sb_warp_sum(fm);
        if (fm > 0.0) { org = idx - 1; lo = 0.0; hi = half; }
        else { org = idx; lo = -half; hi = 0.0; }
    }
    const double lorg = lam_at(org);
    // weight of the pole the offset is measured from: every entry whose eigenvalue is identical
    // to lam[org] sits on that pole (the degenerate cluster of a quasi-Newton Hessian does)
    double w = 0.0;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int i = lane + 32 * u;
        if (i >= i0 && i < i1 && S.lam[u] == lorg) w += S.g2[u];
    }
    w = a2 * sb_warp_sum(w);
    double tt = (tt_guess > lo && tt_guess < hi) ? tt_guess : 0.5 * (lo + hi);
    // F(tt) = p(tt) + r(tt): p = -w/tt is the nearest pole, r the smooth remainder.  Each step
    // solves  r(tt_k) + r'(tt_k)(t - tt_k) - w/t = 0  (a quadratic; Newton on r with the pole kept
    // exact), which converges in a few steps even when the root hugs the pole; bracket and
    // bisection safeguards as before.
    for (int it = 0; it < 200; ++it) {
        if (iters) ++*iters;
        double r = 0.0, dr = 0.0, ra = 0.0;
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            const int i = lane + 32 * u;
            if (i >= i0 && i < i1 && S.lam[u] != lorg) {
                const double den = a2 * (S.lam[u] - lorg) - tt;
                const double rc = 1.0 / den;
                const double q = a2 * S.g2[u] * rc;
                r += q;
                ra += fabs(q);
                dr = fma(q, rc, dr);
            }
        }
        r = (a2 * lorg + tt) + sb_warp_sum(r);
        dr = 1.0 + sb_warp_sum(dr);
        ra = sb_warp_sum(ra);
        const double f = r - w / tt;
        // F cannot be evaluated more accurately than eps * (sum of |terms|): stop there
        // (LAPACK dlaed4's criterion) instead of chasing the last bit of tt through noise
        if (fabs(f) <= 4.0 * RFO_EPS * (fabs(a2 * lorg) + fabs(tt) + ra + fabs(w / tt))) break;
        if (f < 0.0) lo = tt; else hi = tt;
        const double Bq = r - dr * tt;
        const double sq = sqrt(fma(Bq, Bq, 4.0 * dr * w));
        double next;
        if (lo >= 0.0) next = Bq <= 0.0 ? (sq - Bq) / (2.0 * dr) : 2.0 * w / (Bq + sq);
        else next = Bq >= 0.0 ? -(Bq + sq) / (2.0 * dr) : -2.0 * w / (sq - Bq);
        if (!(next > lo && next < hi)) {
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

// One RFO block on entries [i0, i1).  Accumulates sum s^2 and sum s*ds; optionally
// stores s into sreg.  `guess` in/out: tt / alpha^2 of this block's root (0 = none).
template <int NPL>
__device__ void rfo_block(const Sys<NPL>& S, int i0, int i1, int which, int idx, double alpha, double* guess,
                          double* ss_out, double* sds_out, double* sreg, int* iters = nullptr) {
    const int lane = threadIdx.x & 31;
    if (i1 <= i0) { *ss_out = 0.0; *sds_out = 0.0; return; }
    const double a2 = alpha * alpha;
    double gn = 0.0;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int i = lane + 32 * u;
        if (i >= i0 && i < i1) gn += S.g2[u];
    }
    gn = sb_warp_sum(gn);
    int org; double tt;
    arrow_root<NPL>(S, i0, i1, which, idx, a2, gn, (*guess) * a2, &org, &tt, iters);
    if (which != 2) *guess = (a2 > 0.0) ? tt / a2 : 0.0;
    double lorg;
    {
        const int q = org >> 5, src = org & 31;
        double v = 0.0;
#pragma unroll
        for (int u = 0; u < NPL; ++u) if (u == q) v = S.lam[u];
        lorg = __shfl_sync(0xffffffffu, v, src);
    }
    // Dn_i = a^2 lam_i - mu
    double q2 = 0.0;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int i = lane + 32 * u;
        if (i >= i0 && i < i1) {
            const double Dn = a2 * (S.lam[u] - lorg) - tt;
            q2 += S.g2[u] / (Dn * Dn);
        }
    }
    q2 = sb_warp_sum(q2);
    const double mu = a2 * lorg + tt;
    const double dmu = 2.0 * alpha * mu * q2 / (1.0 + a2 * q2);
    double ss = 0.0, sds = 0.0;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int i = lane + 32 * u;
        if (i >= i0 && i < i1) {
            const double Dn = a2 * (S.lam[u] - lorg) - tt;
            const double g = S.g[u];
            const double si = -a2 * g / Dn;
            const double dsi = (-2.0 * alpha * g * Dn + a2 * g * (2.0 * alpha * S.lam[u] - dmu)) / (Dn * Dn);
            if (sreg) sreg[u] = si;
            ss += si * si;
            sds += si * dsi;
        }
    }
    *ss_out = sb_warp_sum(ss);
    *sds_out = sb_warp_sum(sds);
}

// mode 0: rfo (one arrow-head over all n entries, eigenvector index `order`);
// mode 1: prfo (maximise along the lowest `order` modes, minimise along the rest).
template <int NPL>
__global__ void __launch_bounds__(RFO_WARPS * 32)
rfo_tr_kernel(const double* __restrict__ Vg_, const double* __restrict__ evals_, const double* __restrict__ delta_,
              int order, int n, int mode, double* __restrict__ coef_, double* __restrict__ smag,
              double* __restrict__ alpha_out, int* __restrict__ status, const int* __restrict__ active, int batch,
              const double* __restrict__ extra2_) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * RFO_WARPS + warp;
    if (b >= batch) return;
    if (active && !active[b]) return;
    Sys<NPL> S;
    S.n = n;
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int i = lane + 32 * u;
        S.lam[u] = i < n ? evals_[(size_t)b * n + i] : 0.0;
        S.g[u] = i < n ? Vg_[(size_t)b * n + i] : 0.0;
        S.g2[u] = S.g[u] * S.g[u];
    }
    const double delta = delta_[b];
    const double extra2 = extra2_ ? extra2_[b] : 0.0;
    const int mo = order < n ? order : n;
    double guess_a = 0.0, guess_b = 0.0;
    double sreg[NPL];
    int n_eval = 0, n_iter = 0, n_root = 0;
    const long long t_begin = clock64();
    auto eval = [&](double alpha, double* val, double* dval, bool store) {
        double ss, sds;
        ++n_eval;
        n_root += mode == 0 ? 1 : 2;
        if (mode == 0) {
            const int which = mo == 0 ? 0 : (mo == n ? 1 : 2);
            rfo_block<NPL>(S, 0, n, which, mo, alpha, &guess_a, &ss, &sds, store ? sreg : nullptr, &n_iter);
        } else {
            double s1, d1, s2, d2;
            rfo_block<NPL>(S, 0, mo, 1, mo, alpha, &guess_a, &s1, &d1, store ? sreg : nullptr, &n_iter);
            rfo_block<NPL>(S, mo, n, 0, mo, alpha, &guess_b, &s2, &d2, store ? sreg : nullptr, &n_iter);
            ss = s1 + s2; sds = d1 + d2;
        }
        *val = sqrt(ss + extra2);
        *dval = sds / fmax(*val, 1e-12);
    };
    // restricted_step.py:78-121 with alpha0 = 1 on [0, 1], slope = +1
    double alpha = 1.0, val, dval;
    eval(alpha, &val, &dval, false);
    const bool interior = val < delta;
    int st = 0;
    if (!interior) {
        double err = val - delta, lo = 0.0, hi = 1.0;
        for (int it = 0;; ++it) {
            if (fabs(err) <= 1e-15) break;
            if (nextafter(lo, hi) >= hi) break;
            if (it >= 200) { st = SB_ST_TR_NOCONV; break; }
            if (err > 0.0) hi = alpha; else lo = alpha;
            double a1 = alpha - err / dval;
            if (isnan(a1) || a1 <= lo || a1 >= hi) a1 = 0.5 * (lo + hi);
            if (a1 == alpha) break;
            alpha = a1;
            eval(alpha, &val, &dval, false);
            err = val - delta;
        }
    }
    eval(alpha, &val, &dval, true);
#pragma unroll
    for (int u = 0; u < NPL; ++u) {
        const int i = lane + 32 * u;
        if (i < n) coef_[(size_t)b * n + i] = sreg[u];
    }
    if (lane == 0) {
        smag[b] = interior ? val : delta;
        alpha_out[b] = alpha;
        if (st && status) atomicOr(&status[b], st);
        atomicAdd(&rfo_prof[0], (unsigned long long)n_eval);
        atomicAdd(&rfo_prof[1], (unsigned long long)n_iter);
        atomicAdd(&rfo_prof[2], (unsigned long long)n_root);
        atomicAdd(&rfo_prof[3], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&rfo_prof[4], 1ull);
    }
}

// Restricted atomic step (max per-atom displacement, restricted_step.py:172-183) with
// the rfo / prfo models: warp 0 solves the arrow-head problems for the current alpha
// (coefficients s_hat, ds_hat in the eigenbasis), then the whole CTA streams Vt once to
// form the Cartesian step and its derivative (s = V s_hat), as qn_ras_kernel does.
template <int NPL>
__global__ void __launch_bounds__(256)
rfo_ras_kernel(const double* __restrict__ Vg_, const double* __restrict__ evals_, const double* __restrict__ Vt_,
               const double* __restrict__ delta_, int order, int n, int mode, double* __restrict__ s_out,
               double* __restrict__ smag, double* __restrict__ alpha_out, int* __restrict__ status,
               const int* __restrict__ active, const double* __restrict__ sadd_, int np,
               const int* __restrict__ rowmap_, const double* __restrict__ gperp_, const double* __restrict__ gam_,
               long long vstride, const double* __restrict__ wmis) {
    // wmis != NULL: MaxInternalStep measure max_j |s_j w_j| over the n internal coordinates (restricted_step.py:206-216)
    // np poles (np = n for the dense representation); pole i -> eigenvector row rowmap[i] of Vt
    // (NULL: row i; -1: the unit vector gperp/gam; -2: padding), as in qn_ras_kernel
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* c1 = sm;            // s_hat
    double* c2 = c1 + np;       // ds_hat
    double* s = c2 + np;
    double* ds = s + n;
    int* rm = reinterpret_cast<int*>(ds + n);
    __shared__ double best_val[8];
    __shared__ int best_idx[8];
    __shared__ double sh_alpha, sh_val, sh_dval;
    __shared__ int sh_go;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const double* Vt = Vt_ + (size_t)b * vstride;
    const double* gp = gperp_ ? gperp_ + (size_t)b * n : nullptr;
    const double ginv = (gam_ && gam_[b] > 0.0) ? 1.0 / gam_[b] : 0.0;
    for (int i = tid; i < np; i += nt) rm[i] = rowmap_ ? rowmap_[(size_t)b * np + i] : i;
    const double delta = delta_[b];
    const int natoms = n / 3;
    const int mo = order < np ? order : np;
    Sys<NPL> S;
    S.n = np;
    double guess_a = 0.0, guess_b = 0.0;
    if (warp == 0) {
#pragma unroll
        for (int u = 0; u < NPL; ++u) {
            const int i = lane + 32 * u;
            S.lam[u] = i < np ? evals_[(size_t)b * np + i] : 0.0;
            S.g[u] = i < np ? Vg_[(size_t)b * np + i] : 0.0;
            S.g2[u] = S.g[u] * S.g[u];
        }
    }
    double alpha = 1.0, lo = 0.0, hi = 1.0, val = 0.0, dval = 0.0;
    bool interior = false, first = true;
    int st = 0;
    for (int it = 0;; ++it) {
        // ---- coefficients at the current alpha (warp 0)
        if (warp == 0) {
            double sreg[NPL], dreg[NPL];
            const double a2 = alpha * alpha;
            auto block = [&](int i0, int i1, int which, int idx, double* guess) {
                if (i1 <= i0) return;
                double gn = 0.0;
#pragma unroll
                for (int u = 0; u < NPL; ++u) { const int i = lane + 32 * u; if (i >= i0 && i < i1) gn += S.g2[u]; }
                gn = sb_warp_sum(gn);
                int org; double tt;
                arrow_root<NPL>(S, i0, i1, which, idx, a2, gn, (*guess) * a2, &org, &tt);
                if (which != 2) *guess = (a2 > 0.0) ? tt / a2 : 0.0;
                double lorg;
                {
                    const int q = org >> 5, src = org & 31;
                    double v = 0.0;
#pragma unroll
                    for (int u = 0; u < NPL; ++u) if (u == q) v = S.lam[u];
                    lorg = __shfl_sync(0xffffffffu, v, src);
                }
                double q2 = 0.0;
#pragma unroll
                for (int u = 0; u < NPL; ++u) {
                    const int i = lane + 32 * u;
                    if (i >= i0 && i < i1) { const double Dn = a2 * (S.lam[u] - lorg) - tt; q2 += S.g2[u] / (Dn * Dn); }
                }
                q2 = sb_warp_sum(q2);
                const double mu = a2 * lorg + tt;
                const double dmu = 2.0 * alpha * mu * q2 / (1.0 + a2 * q2);
#pragma unroll
                for (int u = 0; u < NPL; ++u) {
                    const int i = lane + 32 * u;
                    if (i >= i0 && i < i1) {
                        const double Dn = a2 * (S.lam[u] - lorg) - tt;
                        const double g = S.g[u];
                        sreg[u] = -a2 * g / Dn;
                        dreg[u] = (-2.0 * alpha * g * Dn + a2 * g * (2.0 * alpha * S.lam[u] - dmu)) / (Dn * Dn);
                    }
                }
            };
            if (mode == 0) block(0, np, mo == 0 ? 0 : (mo == np ? 1 : 2), mo, &guess_a);
            else { block(0, mo, 1, mo, &guess_a); block(mo, np, 0, mo, &guess_b); }
#pragma unroll
            for (int u = 0; u < NPL; ++u) {
                const int i = lane + 32 * u;
                if (i < np) { c1[i] = sreg[u]; c2[i] = dreg[u]; }
            }
        }
        __syncthreads();
        // ---- s = V c1, ds = V c2
        for (int j = tid; j < n; j += nt) {
            double a = 0.0, d = 0.0;
#pragma unroll 4
            for (int i = 0; i < np; ++i) {
                const int r = rm[i];
                if (r < -1) continue;
                const double v = r >= 0 ? Vt[(size_t)r * n + j] : gp[j] * ginv;
                a = fma(v, c1[i], a);
                d = fma(v, c2[i], d);
            }
            s[j] = a + (sadd_ ? sadd_[(size_t)b * n + j] : 0.0);
            ds[j] = d;
        }
        __syncthreads();
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
        if (tid == 0) {
            bv = best_val[0]; bi = best_idx[0];
            for (int w = 1; w < nt / 32; ++w)
                if (best_val[w] > bv || (best_val[w] == bv && best_idx[w] < bi)) { bv = best_val[w]; bi = best_idx[w]; }
            val = bv;
            if (wmis) dval = (s[bi] < 0.0 ? -1.0 : (s[bi] > 0.0 ? 1.0 : 0.0)) * ds[bi] * wmis[bi];
            else
            dval = (ds[3 * bi] * s[3 * bi] + ds[3 * bi + 1] * s[3 * bi + 1] + ds[3 * bi + 2] * s[3 * bi + 2]) /
                   fmax(bv, 1e-12);
            int go = 1;
            if (first && val < delta) { interior = true; go = 0; }
            else {
                const double err = val - delta;
                if (fabs(err) <= 1e-15 || nextafter(lo, hi) >= hi) go = 0;
                else if (it >= 200) { st = SB_ST_TR_NOCONV; go = 0; }
                else {
                    if (err > 0.0) hi = alpha; else lo = alpha;
                    double a1 = alpha - err / dval;
                    if (isnan(a1) || a1 <= lo || a1 >= hi) a1 = 0.5 * (lo + hi);
                    if (a1 == alpha) go = 0;
                    alpha = a1;
                }
            }
            sh_alpha = alpha; sh_val = val; sh_dval = dval; sh_go = go;
        }
        first = false;
        __syncthreads();
        alpha = sh_alpha;
        if (!sh_go) break;
    }
    for (int j = tid; j < n; j += nt) s_out[(size_t)b * n + j] = s[j];
    if (tid == 0) {
        smag[b] = interior ? sh_val : delta;
        alpha_out[b] = alpha;
        if (st && status) atomicOr(&status[b], st);
    }
}

}  // namespace

extern "C" int sb_rfo_profile_impl(unsigned long long* out8, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, rfo_prof, sizeof(unsigned long long) * 8);
    if (e != cudaSuccess) return (int)e;
    if (reset) {
        unsigned long long z[8] = {0};
        e = cudaMemcpyToSymbol(rfo_prof, z, sizeof(z));
    }
    return (int)e;
}

extern "C" int sb_rfo_tr_impl(const double* Vg, const double* evals, const double* delta, int order, int n, int mode,
                              double* coef, double* smag, double* alpha, int* status, const int* active,
                              const double* extra2, int batch, cudaStream_t st) {
    const int grid = (batch + RFO_WARPS - 1) / RFO_WARPS;
    const int npl = (n + 31) / 32;
    SB_COUNT(1);
#define SB_RFO(N)                                                                                                    \
    rfo_tr_kernel<N><<<grid, RFO_WARPS * 32, 0, st>>>(Vg, evals, delta, order, n, mode, coef, smag, alpha, status, \
                                                      active, batch, extra2)
    if (npl <= 4) SB_RFO(4);
    else if (npl <= 8) SB_RFO(8);
    else if (npl <= 12) SB_RFO(12);
    else if (npl <= 16) SB_RFO(16);
    else if (npl <= 24) SB_RFO(24);
    else if (npl <= 32) SB_RFO(32);
    else if (npl <= 48) SB_RFO(48);
    else return -2;      // n > 1536 not supported by this build
#undef SB_RFO
    return SB_LAUNCH_CHECK();
}

static int launch_rfo_ras(const double* Vg, const double* evals, const double* Vt, const double* delta,
                          int order, int n, int mode, double* s, double* smag, double* alpha, int* status,
                          const int* active, const double* sadd, int np, const int* rowmap, const double* gperp,
                          const double* gam, long long vstride, const double* wmis, int batch, cudaStream_t st) {
    const int npl = (np + 31) / 32;
    const size_t smem = (size_t)(2 * np + 2 * n) * sizeof(double) + (size_t)(np + 2) * sizeof(int);
    if (smem > 200 * 1024) return -2;
    SB_COUNT(1);
#define SB_RFOR(N)                                                                                             \
    cudaFuncSetAttribute(rfo_ras_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    rfo_ras_kernel<N><<<batch, 256, smem, st>>>(Vg, evals, Vt, delta, order, n, mode, s, smag, alpha, status, \
                                                active, sadd, np, rowmap, gperp, gam, vstride, wmis)
    if (npl <= 4) { SB_RFOR(4); }
    else if (npl <= 8) { SB_RFOR(8); }
    else if (npl <= 12) { SB_RFOR(12); }
    else if (npl <= 16) { SB_RFOR(16); }
    else if (npl <= 24) { SB_RFOR(24); }
    else if (npl <= 32) { SB_RFOR(32); }
    else if (npl <= 48) { SB_RFOR(48); }
    else return -2;
#undef SB_RFOR
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_rfo_ras_c_impl(const double* Vg, const double* evals, const double* Vt, const double* delta,
                                 int order, int n, int mode, double* s, double* smag, double* alpha, int* status,
                                 const int* active, const double* sadd, int np, const int* rowmap, const double* gperp,
                                 const double* gam, long long vstride, int batch, cudaStream_t st) {
    return launch_rfo_ras(Vg, evals, Vt, delta, order, n, mode, s, smag, alpha, status, active, sadd, np, rowmap, gperp,
                          gam, vstride, nullptr, batch, st);
}

// MaxInternalStep with the rfo / prfo models: see sb_qn_mis_impl
extern "C" int sb_rfo_mis_impl(const double* Vg, const double* evals, const double* Wt, const double* delta, int order,
                               int n, int mode, double* s, double* smag, double* alpha, int* status,
                               const int* active, const double* sadd, int np, long long vstride, const double* w,
                               int batch, cudaStream_t st) {
    if (!w) return -1;
    return launch_rfo_ras(Vg, evals, Wt, delta, order, n, mode, s, smag, alpha, status, active, sadd, np, nullptr,
                          nullptr, nullptr, vstride, w, batch, st);
}

extern "C" int sb_rfo_ras_impl(const double* Vg, const double* evals, const double* Vt, const double* delta, int order,
                               int n, int mode, double* s, double* smag, double* alpha, int* status,
                               const int* active, const double* sadd, int batch, cudaStream_t st) {
    return sb_rfo_ras_c_impl(Vg, evals, Vt, delta, order, n, mode, s, smag, alpha, status, active, sadd, n, nullptr,
                             nullptr, nullptr, (long long)n * n, batch, st);
}
