// Compact representation of the approximate Hessian and of its spectrum:
//
//     B = lam0 * I + VR^T diag(theta - lam0) VR,      VR [m, n] orthonormal rows, theta ascending,
//
// m explicit eigenpairs (theta_i, row i of VR) plus the eigenvalue lam0 on the whole orthogonal
// complement (multiplicity n - m), whose eigenvectors are never stored.  A quasi-Newton Hessian that
// starts as lam0*I (sella/hessian_update.py:58-67) and receives low-rank secant updates
// (hessian_update.py:77-111) has exactly this form with m = number of rank-one terms applied so far,
// so everything the reference gets from a fresh scipy.linalg.eigh(B) (sella/linalg.py:174-195 for
// |B|S, sella/optimize/stepper.py:75-185 for the step models, sella/eigensolvers.py:115-139 for the
// Davidson preconditioner, sella/optimize/optimize.py:369-371 for the "lowest mode went positive"
// test) follows from
//
//     f(B) x = f(lam0) x + VR^T [ (f(theta) - f(lam0)) * (VR x) ]
//
// at O(n m) per pass instead of O(n^2); once m reaches n the representation IS the dense
// eigendecomposition (no complement left) and nothing changes in the algebra.
//
// Kernels here (one CTA per system unless noted); the heavy passes over VR are the rectangular H.V
// kernels of hv.cu (sb_hv_rect) and the eigenvector rotation is secular.cu (sb_secular_update_c):
//   append_a / append_b : new orthonormal directions for the part of an update that lies outside
//                         span(VR) -- they leave the complement with eigenvalue lam0 and join VR
//   prepare / finish    : merged pole list (explicit eigenvalues + the complement as ONE pole with
//                         weight |g_perp|) for the restricted-step kernels, and back
//   finish2             : s, |B| s and B s from one transposed pass
//   scale / axpy        : spectral functions of B applied to vector blocks (|B| S, B S)
//   jd_coeff / jd_finish: Jacobi-Davidson correction in the eigenbasis of the preconditioner
// numpy blueprint with the same operation order: tests/compact_proto.py.
#include "small_dense.cuh"
#include "../../include/sella_b200.h"

namespace {

constexpr int CP_THREADS = 256;
constexpr double CP_EPS = 2.220446049250313e-16;
constexpr int CP_TMAX = 32;            // rank-one terms per update (2 * kcap, kcap <= 16)
constexpr double CP_DROP = 4.0e-13;    // smallest out-of-span remainder of a unit update vector that is kept

// dots[j] = basis_j . xs for j < cnt (warp per basis vector), then xs -= sum_j dots[j] basis_j
__device__ void cgs_sweep(double* xs, int n, const double* __restrict__ basis, size_t bstride, int cnt, double* dots) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    __syncthreads();
    for (int j = warp; j < cnt; j += nw) {
        const double* q = basis + (size_t)j * bstride;
        double d = 0.0;
        for (int e = lane; e < n; e += 32) d = fma(q[e], xs[e], d);
        d = sb_warp_sum(d);
        if (lane == 0) dots[j] = d;
    }
    __syncthreads();
    for (int e = tid; e < n; e += nt) {
        double v = xs[e];
        for (int j = 0; j < cnt; ++j) v = fma(-dots[j], basis[(size_t)j * bstride + e], v);
        xs[e] = v;
    }
    __syncthreads();
}

// Candidates for new explicit directions: p_perp_t = P_t - W1_t (W1 = VR^T VR P_t), orthonormalised
// among themselves (classical Gram-Schmidt, two sweeps), negligible ones dropped.
__global__ void __launch_bounds__(CP_THREADS)
append_a_kernel(const double* __restrict__ P_, const double* __restrict__ W1_, int zcap, const int* __restrict__ nterm,
                const int* __restrict__ mrows, int n, double* __restrict__ Qc_, int* __restrict__ ncand,
                const int* __restrict__ skip) {
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    extern __shared__ double sm[];
    double* xs = sm;                       // n
    double* scratch = xs + n;              // SB_SCRATCH_DOUBLES
    double* dots = scratch + SB_SCRATCH_DOUBLES;   // CP_TMAX
    const int T = (skip && skip[b]) ? 0 : min(nterm[b], min(zcap, CP_TMAX));
    const int m = mrows[b];
    double* Qc = Qc_ + (size_t)b * zcap * n;
    int cnt = 0;
    if (T > 0 && m < n) {
        const double* P = P_ + (size_t)b * zcap * n;
        const double* W1 = W1_ + (size_t)b * zcap * n;
        for (int t = 0; t < T; ++t) {
            for (int e = tid; e < n; e += nt) xs[e] = P[(size_t)t * n + e] - W1[(size_t)t * n + e];
            cgs_sweep(xs, n, Qc, n, cnt, dots);
            cgs_sweep(xs, n, Qc, n, cnt, dots);
            double acc = 0.0;
            for (int e = tid; e < n; e += nt) acc = fma(xs[e], xs[e], acc);
            const double nrm = sqrt(sb_block_sum(acc, scratch));
            // p_t has unit length and p_perp carries an absolute error of a few eps (plus the orthogonality
            // defect of VR): a remainder below ~2000 eps is noise.  Kept, it would become a unit vector with
            // O(1) components inside span(VR) and pollute every later candidate through its projection;
            // dropped, at most |sigma| * 4e-13 of the update is lost.
            if (!(nrm > CP_DROP) || cnt >= n - m) continue;
            const double inv = 1.0 / nrm;
            for (int e = tid; e < n; e += nt) Qc[(size_t)cnt * n + e] = xs[e] * inv;
            ++cnt;
            __syncthreads();
        }
    }
    for (int i = cnt * n + tid; i < zcap * n; i += nt) Qc[i] = 0.0;       // unused slots stay finite
    if (tid == 0) ncand[b] = cnt;
}

// q_j = Qc_j - W2_j (W2 = VR^T VR Qc_j: second projection of the UNIT candidates, orthogonal to VR to
// rounding), re-orthonormalised among themselves, appended as rows m, m+1, ... of VR with eigenvalue
// lam0; Z[t, m + j] = q_j . p_t (coefficients of the update vectors on the new rows).
__global__ void __launch_bounds__(CP_THREADS)
append_b_kernel(const double* __restrict__ P_, const double* __restrict__ Qc_, const double* __restrict__ W2_, int zcap,
                const int* __restrict__ nterm, const int* __restrict__ ncand, int n, double* __restrict__ evals_,
                long long estride, double* __restrict__ VR_, long long vstride, int* __restrict__ mrows,
                const double* __restrict__ lam0, double* __restrict__ Z_, const int* __restrict__ skip) {
    const int b = blockIdx.x;
    if (skip && skip[b]) return;
    const int C = ncand[b];
    if (C == 0) return;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    extern __shared__ double sm[];
    double* xs = sm;
    double* scratch = xs + n;
    double* dots = scratch + SB_SCRATCH_DOUBLES;
    const int T = min(nterm[b], min(zcap, CP_TMAX));
    const int m = mrows[b];
    const double* P = P_ + (size_t)b * zcap * n;
    const double* Qc = Qc_ + (size_t)b * zcap * n;
    const double* W2 = W2_ + (size_t)b * zcap * n;
    double* VR = VR_ + (size_t)b * vstride;
    double* Z = Z_ + (size_t)b * zcap * n;
    double* rows = VR + (size_t)m * n;
    int cnt = 0;
    for (int j = 0; j < C; ++j) {
        for (int e = tid; e < n; e += nt) xs[e] = Qc[(size_t)j * n + e] - W2[(size_t)j * n + e];
        cgs_sweep(xs, n, rows, n, cnt, dots);
        cgs_sweep(xs, n, rows, n, cnt, dots);
        double acc = 0.0;
        for (int e = tid; e < n; e += nt) acc = fma(xs[e], xs[e], acc);
        const double nrm = sqrt(sb_block_sum(acc, scratch));
        // a unit candidate that loses most of its length here was round-off to begin with
        if (!(nrm > 0.1) || m + cnt >= n) continue;
        const double inv = 1.0 / nrm;
        for (int e = tid; e < n; e += nt) rows[(size_t)cnt * n + e] = xs[e] * inv;
        ++cnt;
        __syncthreads();
    }
    const double l0 = lam0[b];
    for (int j = tid; j < cnt; j += nt) evals_[(size_t)b * estride + m + j] = l0;
    for (int pr = warp; pr < T * cnt; pr += nw) {
        const int t = pr / cnt, j = pr % cnt;
        const double* q = rows + (size_t)j * n;
        const double* p = P + (size_t)t * n;
        double d = 0.0;
        for (int e = lane; e < n; e += 32) d = fma(q[e], p[e], d);
        d = sb_warp_sum(d);
        if (lane == 0) Z[(size_t)t * n + m + j] = d;
    }
    if (tid == 0) mrows[b] = m + cnt;
}

// Merged ascending pole list for the restricted-step kernels (width entries per system, stride
// width): explicit eigenvalues with coefficients Vg = VR g; if m < n the complement as ONE pole lam0
// with coefficient |g_perp| at its sorted position, followed by zero-weight copies of lam0 as padding.
// rowmap: explicit row index, -1 = complement pole, -2 = padding.  The explicit pairs may be stored in
// any order (evals[i] belongs to row i): they are ranked here.
__global__ void __launch_bounds__(CP_THREADS)
prepare_kernel(const double* __restrict__ g_, const double* __restrict__ Vg_, const double* __restrict__ Wg_,
               const double* __restrict__ evals_, long long estride, const int* __restrict__ mrows,
               const double* __restrict__ lam0, int n, int width, double* __restrict__ gperp_, double* __restrict__ gam,
               double* __restrict__ cev, double* __restrict__ cvg, int* __restrict__ rowmap,
               const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double evs[];                 // width doubles
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    __shared__ int spos;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int m = min(mrows[b], width);
    const double l0 = lam0[b];
    const bool cluster = m < n;
    double acc = 0.0;
    for (int e = tid; e < n; e += nt) {
        const double v = cluster ? g_[(size_t)b * n + e] - Wg_[(size_t)b * n + e] : 0.0;
        gperp_[(size_t)b * n + e] = v;
        acc = fma(v, v, acc);
    }
    const double gm = sqrt(sb_block_sum(acc, scratch));
    if (tid == 0) { spos = 0; gam[b] = gm; }
    for (int i = tid; i < m; i += nt) evs[i] = evals_[(size_t)b * estride + i];
    __syncthreads();
    int cntl = 0;
    for (int i = tid; i < m; i += nt) cntl += (evs[i] < l0) ? 1 : 0;
    if (cntl) atomicAdd(&spos, cntl);
    __syncthreads();
    const int pos = spos;
    const int nfill = width - m;           // complement pole (if any) + padding
    for (int i = tid; i < m; i += nt) {
        const double ei = evs[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) { const double ej = evs[j]; rank += (ej < ei) || (ej == ei && j < i); }
        const int idx = rank < pos ? rank : rank + nfill;
        cev[(size_t)b * width + idx] = ei;
        cvg[(size_t)b * width + idx] = Vg_[(size_t)b * n + i];
        rowmap[(size_t)b * width + idx] = i;
    }
    for (int f = tid; f < nfill; f += nt) {
        const bool first = (f == 0) && cluster;
        cev[(size_t)b * width + pos + f] = l0;
        cvg[(size_t)b * width + pos + f] = first ? gm : 0.0;
        rowmap[(size_t)b * width + pos + f] = first ? -1 : -2;
    }
}

// Pole-list coefficients back to explicit rows: C4[b,0,:] = c, C4[b,1,:] = |theta| c, C4[b,2,:] = theta c
// (zero beyond the explicit rows), kappa[b] = c_complement / |g_perp|.
__global__ void __launch_bounds__(CP_THREADS)
finish_kernel(const double* __restrict__ ccoef, const int* __restrict__ rowmap, int width,
              const double* __restrict__ evals_, long long estride, const double* __restrict__ gam, int n, int mb,
              double* __restrict__ C4, double* __restrict__ kappa, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* c4 = C4 + (size_t)b * 4 * n;
    for (int i = tid; i < 3 * n; i += nt) if ((i % n) < mb) c4[i] = 0.0;
    if (tid == 0) kappa[b] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < width; idx += nt) {
        const int r = rowmap[(size_t)b * width + idx];
        const double c = ccoef[(size_t)b * width + idx];
        if (r >= 0) {
            const double th = evals_[(size_t)b * estride + r];
            c4[r] = c;
            c4[n + r] = fabs(th) * c;
            c4[2 * n + r] = th * c;
        } else if (r == -1) {
            const double gm = gam[b];
            kappa[b] = gm > 0.0 ? c / gm : 0.0;
        }
    }
}

// s = T4[0] + kappa g_perp, |B| s = T4[1] + |lam0| kappa g_perp, B s = T4[2] + lam0 kappa g_perp, xnew = x + s
__global__ void finish2_kernel(const double* __restrict__ T4, const double* __restrict__ gperp,
                               const double* __restrict__ kappa, const double* __restrict__ lam0,
                               const double* __restrict__ x, int n, double* __restrict__ s, double* __restrict__ absBs,
                               double* __restrict__ Bs, double* __restrict__ xnew, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t o = (size_t)b * n + i;
    const double kg = kappa[b] * gperp[o], l0 = lam0[b];
    const double* t4 = T4 + (size_t)b * 4 * n;
    const double sv = t4[i] + kg;
    s[o] = sv;
    if (absBs) absBs[o] = t4[n + i] + fabs(l0) * kg;
    if (Bs) Bs[o] = t4[2 * n + i] + l0 * kg;
    if (xnew) xnew[o] = x[o] + sv;
}

// out[b,a,i] = (f(theta_i) - f(lam0)) * VtS[b,a,i] for i < m, 0 for m <= i < mb;  mode 0: f = |.|, 1: f = id
__global__ void scale_kernel(const double* __restrict__ VtS, const double* __restrict__ evals_, long long estride,
                             const int* __restrict__ mrows, const double* __restrict__ lam0, int kcap, int nv, int n,
                             int mb, int mode, double* __restrict__ out, const int* __restrict__ skip) {
    const int b = blockIdx.y;
    if (skip && skip[b]) return;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nv * n) return;
    const int i = (int)(idx % n);
    if (i >= mb) return;
    const int m = mrows[b];
    double v = 0.0;
    if (i < m) {
        const double th = evals_[(size_t)b * estride + i], l0 = lam0[b];
        const double f = mode == 0 ? fabs(th) - fabs(l0) : th - l0;
        v = f * VtS[(size_t)b * kcap * n + idx];
    }
    out[(size_t)b * kcap * n + idx] = v;
}

// out = f(lam0) * S + T   (T = VR^T scaled coefficients)
__global__ void axpy_kernel(const double* __restrict__ S, const double* __restrict__ T, const double* __restrict__ lam0,
                            int kcap, int nv, int n, int mode, double* __restrict__ out, const int* __restrict__ skip) {
    const int b = blockIdx.y;
    if (skip && skip[b]) return;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nv * n) return;
    const double l0 = lam0[b];
    const size_t o = (size_t)b * kcap * n + idx;
    out[o] = (mode == 0 ? fabs(l0) : l0) * S[o] + T[o];
}

// Jacobi-Davidson correction in the eigenbasis of P = B (eigensolvers.py:115-139):
//   a = (P - theta)^-1 r, b = (P - theta)^-1 v, eps = (v.a)/(v.b), t = -a + eps b   (jd0; gd: t = a)
// with (P - theta)^-1 x = x/d0 + VR^T [(1/d_i - 1/d0) (VR x)_i], d_i = theta_i - theta, d0 = lam0 - theta.
// rvhat[b,0,:] = VR r, rvhat[b,1,:] = VR v; rv[b,0,:] = r, rv[b,1,:] = v.
// Out: that[b, :mb] (coefficients for the transposed pass), ed[b,0] = eps, ed[b,1] = 1/d0 (0 if m == n).
__global__ void __launch_bounds__(CP_THREADS)
jd_coeff_kernel(const double* __restrict__ rvhat, const double* __restrict__ rv, const double* __restrict__ evals_,
                long long estride, const int* __restrict__ mrows, const double* __restrict__ lam0,
                const double* __restrict__ theta, int n, int mb, int method, double* __restrict__ that,
                double* __restrict__ ed, const int* __restrict__ dav_state) {
    const int b = blockIdx.x;
    if (dav_state[b] != 0) return;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int m = min(mrows[b], mb);
    const double* rh = rvhat + (size_t)b * 2 * n;
    const double* vh = rh + n;
    const double* r = rv + (size_t)b * 2 * n;
    const double* v = r + n;
    const double* lam = evals_ + (size_t)b * estride;
    const double th = theta[b];
    const bool cluster = m < n;
    const double inv0 = cluster ? 1.0 / (lam0[b] - th) : 0.0;
    double va = 0.0, vb = 0.0, hr = 0.0, hv = 0.0;
    for (int i = tid; i < m; i += nt) {
        const double d = lam[i] - th;
        va = fma(vh[i], rh[i] / d, va);
        vb = fma(vh[i], vh[i] / d, vb);
        hr = fma(vh[i], rh[i], hr);
        hv = fma(vh[i], vh[i], hv);
    }
    sb_block_sum2(va, vb, scratch);
    sb_block_sum2(hr, hv, scratch);
    if (cluster) {
        double vr = 0.0, vv = 0.0;
        for (int e = tid; e < n; e += nt) { vr = fma(v[e], r[e], vr); vv = fma(v[e], v[e], vv); }
        sb_block_sum2(vr, vv, scratch);
        va += (vr - hr) * inv0;
        vb += (vv - hv) * inv0;
    }
    double eps = va / vb;
    if (method == 0 && fabs(vb) < 1e-12) eps = 0.0;
    for (int i = tid; i < mb; i += nt) {
        double o = 0.0;
        if (i < m) {
            const double w = 1.0 / (lam[i] - th) - inv0;
            o = method == 1 ? w * rh[i] : w * (eps * vh[i] - rh[i]);
        }
        that[(size_t)b * n + i] = o;
    }
    if (tid == 0) { ed[2 * b] = eps; ed[2 * b + 1] = inv0; }
}

// t += (eps v - r)/d0   (gd: t += r/d0)
__global__ void jd_finish_kernel(double* __restrict__ t, const double* __restrict__ rv, const double* __restrict__ ed,
                                 int n, int method, const int* __restrict__ dav_state) {
    const int b = blockIdx.y;
    if (dav_state[b] != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double eps = ed[2 * b], inv0 = ed[2 * b + 1];
    const double r = rv[(size_t)b * 2 * n + i], v = rv[(size_t)b * 2 * n + n + i];
    t[(size_t)b * n + i] += method == 1 ? r * inv0 : (eps * v - r) * inv0;
}

// the k lowest eigenvalues of B per system (ascending): the explicit theta (any storage order) merged with
// lam0 (multiplicity n - m); k <= 8
__global__ void lowest_kernel(const double* __restrict__ evals_, long long estride, const int* __restrict__ mrows,
                              const double* __restrict__ lam0, int n, int k, double* __restrict__ out, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int m = mrows[b];
    double best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = INFINITY;
    auto insert = [&](double v) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < k && v < best[j]) { const double t = best[j]; best[j] = v; v = t; }
        }
    };
    for (int i = 0; i < m; ++i) insert(evals_[(size_t)b * estride + i]);
    const int copies = min(k, n - m);
    for (int c = 0; c < copies; ++c) insert(lam0[b]);
    for (int j = 0; j < k; ++j) out[(size_t)b * k + j] = isinf(best[j]) ? lam0[b] : best[j];
}

// out[b,v,:] = mask[:] * X[b,v,:]  for v < nv  (projection onto the free Cartesian coordinates)
__global__ void mask_kernel(const double* __restrict__ X, const double* __restrict__ mask, double* __restrict__ out,
                            int ld, int nv, int n) {
    const int b = blockIdx.y;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nv * n) return;
    const size_t o = (size_t)b * ld * n + idx;
    out[o] = X[o] * mask[idx % n];
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" {

int sb_compact_append_a(const double* P, const double* W1, int zcap, const int32_t* nterm, const int32_t* mrows,
                        int n, double* Qc, int32_t* ncand, const int32_t* skip, int batch, void* stream) {
    if (zcap < 1 || zcap > CP_TMAX || n < 1 || batch < 1) return -1;
    const size_t smem = ((size_t)n + SB_SCRATCH_DOUBLES + CP_TMAX) * sizeof(double);
    cudaFuncSetAttribute(append_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    append_a_kernel<<<batch, CP_THREADS, smem, ST>>>(P, W1, zcap, nterm, mrows, n, Qc, ncand, skip);
    return SB_LAUNCH_CHECK();
}

int sb_compact_append_b(const double* P, const double* Qc, const double* W2, int zcap, const int32_t* nterm,
                        const int32_t* ncand, int n, double* evals, long long estride, double* VR, long long vstride,
                        int32_t* mrows, const double* lam0, double* Z, const int32_t* skip, int batch, void* stream) {
    if (zcap < 1 || zcap > CP_TMAX || n < 1 || batch < 1) return -1;
    const size_t smem = ((size_t)n + SB_SCRATCH_DOUBLES + CP_TMAX) * sizeof(double);
    cudaFuncSetAttribute(append_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    append_b_kernel<<<batch, CP_THREADS, smem, ST>>>(P, Qc, W2, zcap, nterm, ncand, n, evals, estride, VR, vstride,
                                                     mrows, lam0, Z, skip);
    return SB_LAUNCH_CHECK();
}

int sb_compact_prepare(const double* g, const double* Vg, const double* Wg, const double* evals, long long estride,
                       const int32_t* mrows, const double* lam0, int n, int width, double* gperp, double* gam,
                       double* cev, double* cvg, int32_t* rowmap, const int32_t* active, int batch, void* stream) {
    if (width < 1 || width > n || batch < 1) return -1;
    SB_COUNT(1);
    const size_t smem = (size_t)width * sizeof(double);
    cudaFuncSetAttribute(prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    prepare_kernel<<<batch, CP_THREADS, smem, ST>>>(g, Vg, Wg, evals, estride, mrows, lam0, n, width, gperp, gam, cev,
                                                    cvg, rowmap, active);
    return SB_LAUNCH_CHECK();
}

int sb_compact_finish(const double* ccoef, const int32_t* rowmap, int width, const double* evals, long long estride,
                      const double* gam, int n, int mbound, double* C4, double* kappa, const int32_t* active,
                      int batch, void* stream) {
    if (width < 1 || batch < 1) return -1;
    SB_COUNT(1);
    finish_kernel<<<batch, CP_THREADS, 0, ST>>>(ccoef, rowmap, width, evals, estride, gam, n, mbound, C4, kappa,
                                                active);
    return SB_LAUNCH_CHECK();
}

int sb_compact_finish2(const double* T4, const double* gperp, const double* kappa, const double* lam0,
                       const double* x, int n, double* s, double* absBs, double* Bs, double* xnew,
                       const int32_t* active, int batch, void* stream) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    finish2_kernel<<<grid, 256, 0, ST>>>(T4, gperp, kappa, lam0, x, n, s, absBs, Bs, xnew, active);
    return SB_LAUNCH_CHECK();
}

int sb_compact_scale(const double* VtS, const double* evals, long long estride, const int32_t* mrows,
                     const double* lam0, int kcap, int nvec, int n, int mbound, int mode, double* out,
                     const int32_t* skip, int batch, void* stream) {
    if (nvec < 1 || nvec > kcap) return -1;
    dim3 grid((unsigned)(((size_t)nvec * n + 255) / 256), batch);
    SB_COUNT(1);
    scale_kernel<<<grid, 256, 0, ST>>>(VtS, evals, estride, mrows, lam0, kcap, nvec, n, mbound, mode, out, skip);
    return SB_LAUNCH_CHECK();
}

int sb_compact_axpy(const double* S, const double* T, const double* lam0, int kcap, int nvec, int n, int mode,
                    double* out, const int32_t* skip, int batch, void* stream) {
    if (nvec < 1 || nvec > kcap) return -1;
    dim3 grid((unsigned)(((size_t)nvec * n + 255) / 256), batch);
    SB_COUNT(1);
    axpy_kernel<<<grid, 256, 0, ST>>>(S, T, lam0, kcap, nvec, n, mode, out, skip);
    return SB_LAUNCH_CHECK();
}

int sb_compact_jd_coeff(const double* rvhat, const double* rv, const double* evals, long long estride,
                        const int32_t* mrows, const double* lam0, const double* theta, int n, int mbound, int method,
                        double* that, double* ed, const int32_t* dav_state, int batch, void* stream) {
    SB_COUNT(1);
    jd_coeff_kernel<<<batch, CP_THREADS, 0, ST>>>(rvhat, rv, evals, estride, mrows, lam0, theta, n, mbound, method,
                                                  that, ed, dav_state);
    return SB_LAUNCH_CHECK();
}

int sb_compact_jd_finish(double* t, const double* rv, const double* ed, int n, int method, const int32_t* dav_state,
                         int batch, void* stream) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    jd_finish_kernel<<<grid, 256, 0, ST>>>(t, rv, ed, n, method, dav_state);
    return SB_LAUNCH_CHECK();
}

int sb_compact_lowest(const double* evals, long long estride, const int32_t* mrows, const double* lam0, int n,
                      int k, double* out, int batch, void* stream) {
    if (k < 1 || k > 8) return -1;
    SB_COUNT(1);
    lowest_kernel<<<(batch + 127) / 128, 128, 0, ST>>>(evals, estride, mrows, lam0, n, k, out, batch);
    return SB_LAUNCH_CHECK();
}

int sb_mask_vec(const double* X, const double* mask, double* out, int ldv, int nvec, int n, int batch, void* stream) {
    if (nvec < 1 || nvec > ldv || n < 1 || batch < 1) return -1;
    dim3 grid((unsigned)(((size_t)nvec * n + 255) / 256), batch);
    SB_COUNT(1);
    mask_kernel<<<grid, 256, 0, ST>>>(X, mask, out, ldv, nvec, n);
    return SB_LAUNCH_CHECK();
}

}  // extern "C"
