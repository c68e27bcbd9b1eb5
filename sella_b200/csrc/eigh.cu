// Batched dense symmetric eigensolver (fp64), one CTA per system.
//
// The reference leans on LAPACK ?syevd through scipy.linalg.eigh for
//   * |B| S in the TS-BFGS update          (sella/linalg.py:293 -> 174-195,
//                                            sella/hessian_update.py:121)
//   * the quasi-Newton / P-RFO step models  (sella/optimize/stepper.py:75-89,163-165)
//   * the "is the lowest mode still negative" test (sella/optimize/optimize.py:369-371)
//   * the Davidson start vectors            (sella/eigensolvers.py:46-50)
// LAPACK is third-party to the reference; this file restates the classical
// published algorithm it uses for the tridiagonal path (Householder reduction,
// then the implicit-shift QL iteration of EISPACK imtql2 / LAPACK dsteqr with
// accumulated plane rotations) as four kernels:
//
//   1. tridiag : A -> Q T Q^T, reflectors kept in the strict upper part of the
//                work copy; the rank-2 trailing update of step k is fused with the
//                symmetric matrix-vector product of step k+1 (one read + one
//                write of the trailing block per step, 16*n^3/3 bytes in total).
//   2. form_qt : Mt[j,:] = (Q e_j)^T, one warp per column, vector kept on chip.
//   3. ql      : implicit QL on (d,e); every plane rotation mixes two *rows* of
//                Mt (contiguous, coalesced).
//   4. sort    : ascending eigenvalues, rows of Mt permuted in place.
//
// Output convention: evals[b,:] ascending, Vt[b,i,:] = i-th eigenvector (i.e.
// the transpose of scipy's column convention), so V^T g is a row-dot pass and
// V c a row-combination pass for hv / hv_t.
#include "common.cuh"

namespace {

constexpr int EIG_THREADS = 256;

// ---------------------------------------------------------------- 1. tridiag
// Only the upper triangle of the work copy is read or written (the matrix is symmetric): in the
// pass over the trailing block a warp owns row i and sweeps the columns j >= i; element (i, j)
// contributes to p_i (row part, warp reduction) and, for j > i, to p_j (column part, kept in
// registers per lane -- lane l owns the columns j = l mod 32 -- and summed over the warps through
// shared memory at the end of the pass).  Half the traffic of a full-square sweep.
template <int NPL>
__global__ void __launch_bounds__(EIG_THREADS)
tridiag_kernel(double* __restrict__ Wk, double* __restrict__ dout, double* __restrict__ eout,
               double* __restrict__ tauout, const int* __restrict__ active, int n) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* v = sm;            // pending reflector (absolute indexing 0..n-1)
    double* w = v + n;         // pending companion vector
    double* vn = w + n;        // new reflector
    double* p = vn + n;        // tau * T v
    double* rowbuf = p + n;    // updated row k
    double* scratch = rowbuf + n;
    double* colpart = scratch + SB_SCRATCH_DOUBLES;     // [nw][n] column parts of p
    double* W = Wk + (size_t)b * n * n;
    double* d = dout + (size_t)b * n;
    double* e = eout + (size_t)b * n;
    double* tau = tauout + (size_t)b * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;

    bool pending = false;
    for (int k = 0; k < n; ++k) {
        // (a) bring row k (columns k..n-1) up to date
        const double vk = pending ? v[k] : 0.0, wk = pending ? w[k] : 0.0;
        for (int j = k + tid; j < n; j += nt) {
            double val = W[(size_t)k * n + j];
            if (pending) val -= vk * w[j] + wk * v[j];
            rowbuf[j] = val;
        }
        __syncthreads();
        if (tid == 0) d[k] = rowbuf[k];
        if (k == n - 1) {
            if (tid == 0) { e[k] = 0.0; tau[k] = 0.0; }
            break;
        }
        if (k == n - 2) {
            if (tid == 0) { e[k] = rowbuf[k + 1]; tau[k] = 0.0; }
            // last diagonal entry still needs the pending update
            __syncthreads();
            if (tid == 0) {
                double val = W[(size_t)(n - 1) * n + (n - 1)];
                if (pending) val -= 2.0 * v[n - 1] * w[n - 1];
                d[n - 1] = val;
                e[n - 1] = 0.0;
                tau[n - 1] = 0.0;
            }
            break;
        }
        // (b) Householder vector annihilating rowbuf[k+2..]
        double part = 0.0;
        for (int j = k + 2 + tid; j < n; j += nt) part += rowbuf[j] * rowbuf[j];
        const double xnorm2 = sb_block_sum(part, scratch);
        const double alpha = rowbuf[k + 1];
        double tk, beta, scale;
        if (xnorm2 == 0.0) {
            tk = 0.0; beta = alpha; scale = 0.0;
        } else {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tk = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        for (int j = k + 1 + tid; j < n; j += nt) {
            const double val = (j == k + 1) ? 1.0 : rowbuf[j] * scale;
            vn[j] = val;
            W[(size_t)k * n + j] = val;          // reflector storage (row k, right of diagonal)
        }
        if (tid == 0) { e[k] = beta; tau[k] = tk; }
        __syncthreads();
        // (c) one pass over the upper triangle of the trailing block: apply the pending rank-2
        //     update, write it back, and form p = tau * T vn on the fly
        double cacc[NPL];
#pragma unroll
        for (int q = 0; q < NPL; ++q) cacc[q] = 0.0;
        for (int i = k + 1 + warp; i < n; i += nw) {
            double* row = W + (size_t)i * n;
            const double vi = pending ? v[i] : 0.0, wi = pending ? w[i] : 0.0;
            const double vni = vn[i];
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < NPL; ++q) {
                const int j = lane + 32 * q;
                if (j >= i && j < n) {
                    double val = row[j];
                    if (pending) {
                        val -= vi * w[j] + wi * v[j];
                        row[j] = val;
                    }
                    acc = fma(val, vn[j], acc);
                    if (j > i) cacc[q] = fma(val, vni, cacc[q]);
                }
            }
            acc = sb_warp_sum(acc);
            if (lane == 0) p[i] = acc;
        }
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
            const int j = lane + 32 * q;
            if (j < n) colpart[(size_t)warp * n + j] = cacc[q];
        }
        __syncthreads();
        for (int j = k + 1 + tid; j < n; j += nt) {
            double acc = p[j];
            for (int w2 = 0; w2 < nw; ++w2) acc += colpart[(size_t)w2 * n + j];
            p[j] = tk * acc;
        }
        __syncthreads();
        // (d) w = p - (tau/2)(p.v) v
        part = 0.0;
        for (int i = k + 1 + tid; i < n; i += nt) part += p[i] * vn[i];
        const double pv = sb_block_sum(part, scratch);
        const double coef = 0.5 * tk * pv;
        for (int i = k + 1 + tid; i < n; i += nt) {
            const double vi = vn[i];
            w[i] = p[i] - coef * vi;
            v[i] = vi;
        }
        pending = true;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 2. form Q^T
__global__ void __launch_bounds__(EIG_THREADS)
formqt_kernel(const double* __restrict__ Wk, const double* __restrict__ tauin, double* __restrict__ Mt_,
              const int* __restrict__ active, int n) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    double* y = sm + (size_t)warp * n;
    const double* W = Wk + (size_t)b * n * n;
    const double* tau = tauin + (size_t)b * n;
    double* Mt = Mt_ + (size_t)b * n * n;
    for (int j = warp; j < n; j += nw) {
        for (int i = lane; i < n; i += 32) y[i] = (i == j) ? 1.0 : 0.0;
        __syncwarp();
        int kstart = j - 1;
        if (kstart > n - 3) kstart = n - 3;
        for (int k = kstart; k >= 0; --k) {
            const double tk = tau[k];
            if (tk == 0.0) continue;
            const double* vk = W + (size_t)k * n;
            double acc = 0.0;
            for (int i = k + 1 + lane; i < n; i += 32) acc = fma(vk[i], y[i], acc);
            acc = sb_warp_sum(acc) * tk;
            for (int i = k + 1 + lane; i < n; i += 32) y[i] = fma(-acc, vk[i], y[i]);
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) Mt[(size_t)j * n + i] = y[i];
        __syncwarp();
    }
}

// ---------------------------------------------------------------- 2b. form Q^T with GEMMs
// Blocked accumulation (LAPACK dorgtr/dorgqr scheme, compact WY): the reflectors of a panel of
// FQ_NB columns, B_p = H_k0 ... H_k1 = I - V T V^T, are applied to the trailing block only,
//   Mt22 <- Mt22 (I - V T^T V^T) = Mt22 - (Mt22 V22^T) T^T V22,     Mt22 = Mt[k0+1:, k0+1:],
// going from the last panel to the first, as three fp64 tensor-core GEMMs per panel.  This kernel
// prepares a panel: the explicit reflector rows V22 (zeros left of the unit entry) and T (dlarft).
constexpr int FQ_NB = 16;

__global__ void __launch_bounds__(EIG_THREADS)
formqt_panel_kernel(const double* __restrict__ Wk, const double* __restrict__ tauin, double* __restrict__ Vp_,
                    double* __restrict__ T_, const int* __restrict__ active, int n, int batch) {
    const int b = blockIdx.y, pnl = blockIdx.x;
    if (active && !active[b]) return;
    __shared__ double Tm[FQ_NB * FQ_NB];
    __shared__ double G[FQ_NB * FQ_NB];
    const int k0 = pnl * FQ_NB;
    const int nref = n - 2;                                   // reflectors 0 .. n-3
    const int nb = min(FQ_NB, nref - k0);
    if (nb <= 0) return;
    const double* W = Wk + (size_t)b * n * n;
    const double* tau = tauin + (size_t)b * n;
    double* Vp = Vp_ + (size_t)b * n * n;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int idx = tid; idx < nb * n; idx += nt) {
        const int r = idx / n, j = idx % n, k = k0 + r;
        Vp[(size_t)k * n + j] = j >= k + 1 ? W[(size_t)k * n + j] : 0.0;
    }
    __syncthreads();
    for (int pr = warp; pr < nb * nb; pr += nw) {
        const int i = pr / nb, j = pr % nb;
        if (i >= j) continue;
        const double* vi = Vp + (size_t)(k0 + i) * n;
        const double* vj = Vp + (size_t)(k0 + j) * n;
        double d = 0.0;
        for (int e = k0 + j + 1 + lane; e < n; e += 32) d = fma(vi[e], vj[e], d);
        d = sb_warp_sum(d);
        if (lane == 0) G[i * FQ_NB + j] = d;
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < FQ_NB * FQ_NB; ++i) Tm[i] = 0.0;
        for (int j = 0; j < nb; ++j) {
            const double tj = tau[k0 + j];
            for (int i = 0; i < j; ++i) {
                double acc = 0.0;
                for (int l = i; l < j; ++l) acc += Tm[i * FQ_NB + l] * G[l * FQ_NB + j];
                Tm[i * FQ_NB + j] = -tj * acc;
            }
            Tm[j * FQ_NB + j] = tj;
        }
    }
    __syncthreads();
    for (int i = tid; i < FQ_NB * FQ_NB; i += nt) T_[((size_t)pnl * batch + b) * FQ_NB * FQ_NB + i] = Tm[i];
}

__global__ void identity_kernel(double* __restrict__ M_, int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n * n) M_[(size_t)b * n * n + idx] = (idx / n == idx % n) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------- 3. implicit QL
template <int CPT>
__device__ __forceinline__ void apply_sweep(double* __restrict__ Mt, const double* __restrict__ cs, int n,
                                            int l, int mtop, int tid, int nt) {
    // rotations i = mtop-1 .. l, each mixing rows i and i+1 of Mt
    double carry[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int col = tid + q * nt;
        carry[q] = (col < n) ? Mt[(size_t)mtop * n + col] : 0.0;
    }
#pragma unroll 4
    for (int i = mtop - 1; i >= l; --i) {
        const double c = cs[2 * i], s = cs[2 * i + 1];
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
            const int col = tid + q * nt;
            if (col < n) {
                const double zi = Mt[(size_t)i * n + col];
                const double f = carry[q];
                Mt[(size_t)(i + 1) * n + col] = s * zi + c * f;
                carry[q] = c * zi - s * f;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int col = tid + q * nt;
        if (col < n) Mt[(size_t)l * n + col] = carry[q];
    }
}

template <int CPT>
__global__ void __launch_bounds__(EIG_THREADS)
ql_kernel(double* __restrict__ dio, double* __restrict__ eio, double* __restrict__ Mt_,
          int* __restrict__ status, const int* __restrict__ active, int n) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* d = sm;
    double* e = d + n;
    double* cs = e + n;          // (c,s) pairs of the current sweep
    __shared__ int sh_m, sh_l_lo, sh_flag;
    double* Mt = Mt_ + (size_t)b * n * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) {
        d[i] = dio[(size_t)b * n + i];
        e[i] = eio[(size_t)b * n + i];
    }
    __syncthreads();
    const double eps = 2.220446049250313e-16;
    // An off-diagonal element is negligible relative to its diagonal neighbours OR relative to the whole
    // matrix (EISPACK tql2 / LAPACK dsteqr): without the second test a rank-deficient matrix -- the model
    // Hessian P diag(h0) P of an internal-coordinate search has an exact (nint - ncart)-fold zero eigenvalue --
    // leaves d ~ e ~ 1e-16 |A| in the zero cluster and the purely relative test never fires.
    __shared__ double sh_anorm;
    if (tid == 0) {
        double a = 0.0;
        for (int i = 0; i < n; ++i) a = fmax(a, fabs(d[i]) + (i < n - 1 ? fabs(e[i]) : 0.0));
        sh_anorm = a;
    }
    __syncthreads();
    const double tiny = eps * sh_anorm;
    bool failed = false;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            if (tid == 0) {
                int m = l;
                for (; m < n - 1; ++m) {
                    const double dd = fabs(d[m]) + fabs(d[m + 1]);
                    if (fabs(e[m]) <= eps * dd || fabs(e[m]) <= tiny) break;
                }
                sh_m = m;
                sh_flag = 0;
                if (m != l) {
                    if (iter >= 60) {
                        sh_flag = 2;            // no convergence
                    } else {
                        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                        double r = hypot(g, 1.0);
                        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                        double s = 1.0, c = 1.0, p = 0.0;
                        int i = m - 1;
                        bool underflow = false;
                        for (; i >= l; --i) {
                            double f = s * e[i];
                            const double bb = c * e[i];
                            r = hypot(f, g);
                            e[i + 1] = r;
                            if (r == 0.0) {
                                d[i + 1] -= p;
                                e[m] = 0.0;
                                underflow = true;
                                break;
                            }
                            s = f / r;
                            c = g / r;
                            g = d[i + 1] - p;
                            r = (d[i] - g) * s + 2.0 * c * bb;
                            p = s * r;
                            d[i + 1] = g + p;
                            g = c * r - bb;
                            cs[2 * i] = c;
                            cs[2 * i + 1] = s;
                        }
                        if (!underflow) {
                            d[l] -= p;
                            e[l] = g;
                            e[m] = 0.0;
                            sh_l_lo = l;
                        } else {
                            sh_l_lo = i + 1;    // rotations i+1 .. m-1 were generated
                        }
                        sh_flag = 1;
                    }
                }
            }
            __syncthreads();
            const int flag = sh_flag, m = sh_m, lo = sh_l_lo;
            if (flag != 1) {
                // leaving the sweep loop: thread 0 goes straight on to the next eigenvalue and rewrites the flags --
                // every thread must have read them first (compute-sanitizer racecheck, profiles/sanitizer_r2_*)
                __syncthreads();
                if (flag == 2) failed = true;
                break;
            }
            if (Mt_ && lo <= m - 1) apply_sweep<CPT>(Mt, cs, n, lo, m, tid, nt);
            ++iter;
            __syncthreads();
        }
        if (failed) break;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) dio[(size_t)b * n + i] = d[i];
    if (tid == 0 && failed && status) atomicOr(&status[b], SB_ST_EIGH_NOCONV);
}

// ---------------------------------------------------------------- 4. sort
__global__ void __launch_bounds__(EIG_THREADS)
sort_kernel(const double* __restrict__ din, double* __restrict__ evals, double* __restrict__ Mt_,
            const int* __restrict__ active, int n) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* d = sm;
    int* src = reinterpret_cast<int*>(d + n);     // src[r] = unsorted index that lands at rank r
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) d[i] = din[(size_t)b * n + i];
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const double di = d[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double dj = d[j];
            rank += (dj < di) || (dj == di && j < i);
        }
        src[rank] = i;
        evals[(size_t)b * n + rank] = di;
    }
    __syncthreads();
    if (!Mt_) return;                              // eigenvalues only
    // rows: new[r] = old[src[r]]; cycle-following, independently per column.
    // A cycle is walked from its smallest member ("leader"), found once per start.
    int* leader = src + n;
    for (int start = tid; start < n; start += nt) {
        int j = src[start];
        int isl = (j != start);
        while (isl && j != start) {
            if (j < start) isl = 0;
            j = src[j];
        }
        leader[start] = isl;
    }
    __syncthreads();
    double* Mt = Mt_ + (size_t)b * n * n;
    for (int col = tid; col < n; col += nt) {
        for (int start = 0; start < n; ++start) {
            if (!leader[start]) continue;
            const double tmp = Mt[(size_t)start * n + col];
            int cur = start;
            int nxt = src[cur];
            while (nxt != start) {
                Mt[(size_t)cur * n + col] = Mt[(size_t)nxt * n + col];
                cur = nxt;
                nxt = src[cur];
            }
            Mt[(size_t)cur * n + col] = tmp;
        }
    }
}

}  // namespace

// A (read only) -> evals ascending, Vt rows = eigenvectors.  work: n*n doubles per
// system (reflectors), small: 3*n doubles per system (d, e, tau).
extern "C" int sb_gemm_impl(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                            long long sA, const double* B, int ldb, long long sB, double beta, double* C, int ldc,
                            long long sC, const int* active, int batch, cudaStream_t st);

// work2 (may be NULL): batch * (n*n + 32*n + 256*ceil(n/16)) doubles; when given, Q^T is formed by the
// blocked GEMM scheme instead of the warp-per-column kernel.
extern "C" int sb_eigh_impl(const double* A, double* evals, double* Vt, double* work, double* small,
                            double* work2, int* status, const int* active, int batch, int n, cudaStream_t st) {
    if (n < 1) return -1;
    cudaError_t err = cudaMemcpyAsync(work, A, (size_t)batch * n * n * sizeof(double),
                                      cudaMemcpyDeviceToDevice, st);
    if (err != cudaSuccess) return (int)err;
    double* d = small;
    double* e = small + (size_t)batch * n;
    double* tau = small + (size_t)2 * batch * n;
    {
        const size_t smem = (size_t)(5 * n + SB_SCRATCH_DOUBLES + (EIG_THREADS / 32) * n) * sizeof(double);
        const int npl = (n + 31) / 32;
        SB_COUNT(1);
#define SB_TRIDIAG(N)                                                                                         \
    cudaFuncSetAttribute(tridiag_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    tridiag_kernel<N><<<batch, EIG_THREADS, smem, st>>>(work, d, e, tau, active, n)
        if (npl <= 2) { SB_TRIDIAG(2); }
        else if (npl <= 4) { SB_TRIDIAG(4); }
        else if (npl <= 6) { SB_TRIDIAG(6); }
        else if (npl <= 8) { SB_TRIDIAG(8); }
        else if (npl <= 12) { SB_TRIDIAG(12); }
        else if (npl <= 16) { SB_TRIDIAG(16); }
        else if (npl <= 24) { SB_TRIDIAG(24); }
        else if (npl <= 32) { SB_TRIDIAG(32); }
        else if (npl <= 48) { SB_TRIDIAG(48); }
        else if (npl <= 64) { SB_TRIDIAG(64); }
        else return -2;
#undef SB_TRIDIAG
    }
    if (!Vt) {
        // eigenvalues only (the "did the lowest modes turn positive" test, optimize.py:362-371): no Q^T, and
        // the QL iteration runs on (d, e) alone -- the rotation sweeps over Vt are most of a full eigensolve
    } else if (work2 && n >= 4 * FQ_NB) {
        double* Vp = work2;
        double* W1 = Vp + (size_t)batch * n * n;
        double* W2 = W1 + (size_t)batch * FQ_NB * n;
        double* T = W2 + (size_t)batch * FQ_NB * n;
        const int nref = n - 2, npan = (nref + FQ_NB - 1) / FQ_NB;
        const long long sN = (long long)n * n, sW = (long long)FQ_NB * n, sT = FQ_NB * FQ_NB;
        dim3 gp(npan, batch), gi((n * n + 255) / 256, batch);
        SB_COUNT(2);
        formqt_panel_kernel<<<gp, EIG_THREADS, 0, st>>>(work, tau, Vp, T, active, n, batch);
        identity_kernel<<<gi, 256, 0, st>>>(Vt, n, active);
        int rc;
        for (int p = npan - 1; p >= 0; --p) {
            const int k0 = p * FQ_NB, nb = (nref - k0 < FQ_NB) ? nref - k0 : FQ_NB, np = n - k0 - 1;
            const double* V22 = Vp + (size_t)k0 * n + k0 + 1;             // [nb, np], ld n
            double* M22 = Vt + (size_t)(k0 + 1) * n + k0 + 1;             // [np, np], ld n
            const double* Tp = T + (size_t)p * batch * sT;
            // W1 = M22 V22^T [np, nb];  W2 = W1 T^T;  M22 -= W2 V22
            if ((rc = sb_gemm_impl(0, 1, np, nb, np, 1.0, M22, n, sN, V22, n, sN, 0.0, W1, nb, sW, active, batch, st))) return rc;
            if ((rc = sb_gemm_impl(0, 1, np, nb, nb, 1.0, W1, nb, sW, Tp, FQ_NB, sT, 0.0, W2, nb, sW, active, batch, st))) return rc;
            if ((rc = sb_gemm_impl(0, 0, np, np, nb, -1.0, W2, nb, sW, V22, n, sN, 1.0, M22, n, sN, active, batch, st))) return rc;
        }
    } else {
        int threads = EIG_THREADS;
        size_t smem = (size_t)(threads / 32) * n * sizeof(double);
        cudaFuncSetAttribute(formqt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        SB_COUNT(1);
        formqt_kernel<<<batch, threads, smem, st>>>(work, tau, Vt, active, n);
    }
    {
        // threads: multiple of 32 that divides the columns as evenly as possible
        int cpt = (n + EIG_THREADS - 1) / EIG_THREADS;
        int threads = ((n + cpt - 1) / cpt + 31) / 32 * 32;
        if (threads > EIG_THREADS) threads = EIG_THREADS;
        if (threads < 32) threads = 32;
        const size_t smem = (size_t)4 * n * sizeof(double);
        switch (cpt) {
            case 1: SB_COUNT(1); cudaFuncSetAttribute(ql_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ql_kernel<1><<<batch, threads, smem, st>>>(d, e, Vt, status, active, n); break;
            case 2: SB_COUNT(1); cudaFuncSetAttribute(ql_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ql_kernel<2><<<batch, threads, smem, st>>>(d, e, Vt, status, active, n); break;
            case 3: SB_COUNT(1); cudaFuncSetAttribute(ql_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ql_kernel<3><<<batch, threads, smem, st>>>(d, e, Vt, status, active, n); break;
            case 4: SB_COUNT(1); cudaFuncSetAttribute(ql_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ql_kernel<4><<<batch, threads, smem, st>>>(d, e, Vt, status, active, n); break;
            case 5: case 6: SB_COUNT(1); cudaFuncSetAttribute(ql_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ql_kernel<6><<<batch, threads, smem, st>>>(d, e, Vt, status, active, n); break;
            case 7: case 8: SB_COUNT(1); cudaFuncSetAttribute(ql_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ql_kernel<8><<<batch, threads, smem, st>>>(d, e, Vt, status, active, n); break;
            default: return -2;  // n > 2048 not supported by this build
        }
    }
    {
        const size_t smem = (size_t)n * (sizeof(double) + 2 * sizeof(int));
        SB_COUNT(1);
        sort_kernel<<<batch, EIG_THREADS, smem, st>>>(d, evals, Vt, active, n);
    }
    return SB_LAUNCH_CHECK();
}
