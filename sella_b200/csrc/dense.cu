// Batched dense fp64 linear algebra for the internal-coordinate path:
//   sb_gemm   C = alpha op(A) op(B) + beta C         DMMA (mma.sync m8n8k4 f64) tensor-core tiles
//   sb_qr     economy Householder QR  A = Q R  (blocked, compact WY; trailing updates are GEMMs)
//                                                     sella/_gpu.py:100-111 (gpu_qr),
//                                                     sella/peswrapper.py:674-709 (_get_jacobian_qr)
//   sb_trtri  inverse of the upper-triangular R        sella/peswrapper.py:711-736 (_get_Binv:
//                                                     Binv = R^-1 Q^T = sb_gemm(trtri(R), Q^T))
// These serve the Wilson-matrix algebra of InternalPES (B = QR, B+, g_int = B+^T g_cart,
// Hc = B+^T (D_c - D_q) B+, U^T H U): rectangular nint x ncart matrices, a few hundred rows,
// one matrix per system.  fp64 has no tcgen05 kind; the dense contractions use the fp64
// tensor-core path that exists on sm_100a (DMMA), the factorisations are CTA-per-system.
#include "common.cuh"

namespace {

constexpr int GM_BM = 64, GM_BN = 64, GM_BK = 16, GM_THREADS = 256, GM_PAD = 8;

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// op(A): M x K, op(B): K x N, C: M x N; all row-major with leading dimensions lda/ldb/ldc.
// transA: A is stored K x M; transB: B is stored N x K.  Strides in doubles (0 = shared).
__global__ void __launch_bounds__(GM_THREADS)
gemm_kernel(int transA, int transB, int M, int N, int K, double alpha, const double* __restrict__ A_, int lda,
            long long sA, const double* __restrict__ B_, int ldb, long long sB, double beta, double* __restrict__ C_,
            int ldc, long long sC, const int* __restrict__ active) {
    const int b = blockIdx.z;
    if (active && !active[b]) return;
    __shared__ double As[GM_BK][GM_BM + GM_PAD];     // As[k][m]
    __shared__ double Bs[GM_BK][GM_BN + GM_PAD];     // Bs[k][n]
    const double* A = A_ + (size_t)b * sA;
    const double* B = B_ + (size_t)b * sB;
    double* C = C_ + (size_t)b * sC;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;     // warp tile 16 x 32
    const int gid = lane >> 2, tig = lane & 3;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int k0 = 0; k0 < K; k0 += GM_BK) {
        // stage the tiles (4 elements of each per thread); index order chosen so that the
        // contiguous global dimension runs over consecutive threads
        for (int e = tid; e < GM_BM * GM_BK; e += GM_THREADS) {
            int mm, kk;
            if (transA) { mm = e % GM_BM; kk = e / GM_BM; } else { kk = e % GM_BK; mm = e / GM_BK; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < M && gk < K) v = transA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[kk][mm] = v;
        }
        for (int e = tid; e < GM_BN * GM_BK; e += GM_THREADS) {
            int nn, kk;
            if (transB) { kk = e % GM_BK; nn = e / GM_BK; } else { nn = e % GM_BN; kk = e / GM_BN; }
            const int gn = n0 + nn, gk = k0 + kk;
            double v = 0.0;
            if (gn < N && gk < K) v = transB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GM_BK; ks += 4) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = As[ks + tig][wm + 8 * i + gid];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[ks + tig][wn + 8 * j + gid];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gm = m0 + wm + 8 * i + gid, gn = n0 + wn + 8 * j + 2 * tig + q;
                if (gm < M && gn < N) {
                    double* c = C + (size_t)gm * ldc + gn;
                    *c = beta == 0.0 ? alpha * acc[i][j][q] : fma(alpha, acc[i][j][q], beta * *c);
                }
            }
}

// fp64 peak micro-benchmarks (bench.py: the denominator of the tensor/FMA rooflines; MEASURED_PEAKS.json
// carries no fp64 figure).  kind 0: dependent-chain-free DFMA from registers (8 accumulators per thread);
// kind 1: DMMA m8n8k4 from registers (4 accumulator tiles per warp).  Each thread writes one value so
// that nothing is optimised away.
__global__ void __launch_bounds__(256)
fp64_peak_kernel(int kind, int iters, double* __restrict__ out) {
    const double seed = 1.0 + 1e-9 * (threadIdx.x + blockIdx.x);
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i;
    const double x = 0.999999, y = 1e-7;
    if (kind == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma_8x8x4(a[2 * i], a[2 * i + 1], x, y);
        }
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}

constexpr int QR_THREADS = 256;
constexpr int QR_NB = 16;          // panel width

// Blocked Householder QR (LAPACK dgeqrf / dorgqr scheme, compact WY):
//   for each panel of QR_NB columns: qr_panel_kernel factors it in shared memory (dgeqr2 on an
//   m' x 16 block), writes R's rows, the explicit reflector block V (unit lower trapezoid, zeros
//   above) and the triangular factor T (dlarft); the trailing block then gets
//   A22 <- (I - V T^T V^T) A22 as three DMMA GEMMs (W1 = V^T A22, W2 = T^T W1, A22 -= V W2);
//   Q = H_0 ... H_{n-1} [I; 0] is accumulated backwards, panel by panel, the same way
//   (Q22 <- (I - V T V^T) Q22).
// Conventions as dgeqr2: v_j[j] = 1, H_j = I - tau_j v_j v_j^T, beta_j = -sign(alpha)|x|.
//
// Panel kernel: one CTA per system.  Ap = A + p0*n + p0 (leading dimension n), mp = m - p0 rows,
// nb columns.  Vp [b, m, n] receives the explicit V of this panel at rows p0.., columns p0..p0+nb.
__global__ void __launch_bounds__(QR_THREADS)
qr_panel_kernel(double* __restrict__ A_, int m, int n, int p0, int nb, double* __restrict__ Vp_,
                double* __restrict__ T_, double* __restrict__ R_, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    const int mp = m - p0;
    double* P = sm;                              // [nb][mp] column-major panel
    double* tau = P + (size_t)nb * mp;           // nb
    double* Tm = tau + QR_NB;                    // QR_NB x QR_NB
    double* dots = Tm + QR_NB * QR_NB;           // QR_NB
    double* scratch = dots + QR_NB;              // SB_SCRATCH_DOUBLES
    double* A = A_ + (size_t)b * m * n;
    double* Vp = Vp_ + (size_t)b * m * n;
    double* R = R_ + (size_t)b * n * n;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    for (int idx = tid; idx < mp * nb; idx += nt) {
        const int i = idx / nb, c = idx % nb;
        P[(size_t)c * mp + i] = A[(size_t)(p0 + i) * n + p0 + c];
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        double* col = P + (size_t)j * mp;
        double acc = 0.0;
        for (int i = j + 1 + tid; i < mp; i += nt) acc = fma(col[i], col[i], acc);
        const double sigma = sb_block_sum(acc, scratch);
        const double alpha = col[j];
        double tj = 0.0;
        if (sigma > 0.0) {
            const double nrm = sqrt(alpha * alpha + sigma);
            const double beta = alpha >= 0.0 ? -nrm : nrm;
            tj = (beta - alpha) / beta;
            const double scale = 1.0 / (alpha - beta);
            __syncthreads();
            for (int i = j + 1 + tid; i < mp; i += nt) col[i] *= scale;
            if (tid == 0) col[j] = beta;
        }
        if (tid == 0) tau[j] = tj;
        __syncthreads();
        if (tj != 0.0) {
            // w_c = v^T P[:, c] (v_j = 1 implicit), then P[:, c] -= tau w_c v   for the later columns
            for (int c = j + 1 + warp; c < nb; c += nw) {
                double* pc = P + (size_t)c * mp;
                const double pj = pc[j];
                double d = 0.0;
                for (int i = j + 1 + lane; i < mp; i += 32) d = fma(col[i], pc[i], d);
                d = sb_warp_sum(d) + pj;
                const double w = tj * d;
                __syncwarp();
                if (lane == 0) pc[j] = pj - w;
                for (int i = j + 1 + lane; i < mp; i += 32) pc[i] = fma(-w, col[i], pc[i]);
            }
        }
        __syncthreads();
    }
    // T (dlarft, forward, columnwise): T[j][j] = tau_j, T[0:j, j] = -tau_j T[0:j,0:j] (V[:,0:j]^T v_j)
    for (int j = 0; j < nb; ++j) {
        for (int i = warp; i < j; i += nw) {          // dots[i] = v_i . v_j  (unit diagonals implicit)
            const double* vi = P + (size_t)i * mp;
            const double* vj = P + (size_t)j * mp;
            double d = 0.0;
            for (int r = j + 1 + lane; r < mp; r += 32) d = fma(vi[r], vj[r], d);
            d = sb_warp_sum(d) + vi[j];               // row j: v_j[j] = 1, v_i[j] stored
            if (lane == 0) dots[i] = d;
        }
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < j; ++i) {
                double acc = 0.0;
                for (int l = i; l < j; ++l) acc += Tm[i * QR_NB + l] * dots[l];
                Tm[i * QR_NB + j] = -tau[j] * acc;
            }
            Tm[j * QR_NB + j] = tau[j];
            for (int i = j + 1; i < QR_NB; ++i) Tm[i * QR_NB + j] = 0.0;
        }
        __syncthreads();
    }
    // write back: R rows of this panel (columns p0..p0+nb), explicit V, reflectors into A, T
    for (int idx = tid; idx < mp * nb; idx += nt) {
        const int i = idx / nb, c = idx % nb;
        const double val = P[(size_t)c * mp + i];
        A[(size_t)(p0 + i) * n + p0 + c] = val;
        Vp[(size_t)(p0 + i) * n + p0 + c] = i > c ? val : (i == c ? 1.0 : 0.0);
        if (i <= c) R[(size_t)(p0 + i) * n + p0 + c] = val;
    }
    for (int i = tid; i < QR_NB * QR_NB; i += nt) {
        const int r = i / QR_NB, c = i % QR_NB;
        T_[(size_t)b * QR_NB * QR_NB + i] = (r < nb && c < nb) ? Tm[i] : 0.0;
    }
}

// R[p0.., p0+nb..] rows of the panel right of it (copied out after the trailing update), zero lower part
__global__ void qr_copy_r_kernel(const double* __restrict__ A_, int m, int n, double* __restrict__ R_,
                                 const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int i = idx / n, c = idx % n;
    R_[(size_t)b * n * n + idx] = c >= i ? A_[(size_t)b * m * n + (size_t)i * n + c] : 0.0;
}

__global__ void qr_init_q_kernel(double* __restrict__ Q_, int m, int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * n) return;
    Q_[(size_t)b * m * n + idx] = (idx / n == idx % n) ? 1.0 : 0.0;
}

// Inverse of an upper-triangular diagonal block R[r0:r0+sz, r0:r0+sz] (sz <= 48) of a row-major
// [n, n] matrix into the same block of X; thread j owns column j of the block (back substitution).
// Status bit SB_ST_SINGULAR on a zero diagonal.  Larger matrices are assembled from such blocks by
// the block formula  inv([[R11, R12], [0, R22]]) = [[X11, -X11 R12 X22], [0, X22]]  with DMMA GEMMs.
constexpr int TRI_BS = 48;          // 2 x 48 x 49 doubles of static shared memory
__global__ void __launch_bounds__(64)
trtri_block_kernel(const double* __restrict__ R_, double* __restrict__ X_, int n, int r0, int sz,
                   int* __restrict__ status, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    __shared__ double Rs[TRI_BS][TRI_BS + 1];
    __shared__ double Xs[TRI_BS][TRI_BS + 1];
    const double* R = R_ + (size_t)b * n * n;
    double* X = X_ + (size_t)b * n * n;
    const int j = threadIdx.x;
    for (int idx = threadIdx.x; idx < sz * sz; idx += blockDim.x) {
        const int i = idx / sz, c = idx % sz;
        Rs[i][c] = R[(size_t)(r0 + i) * n + r0 + c];
        Xs[i][c] = 0.0;
    }
    __syncthreads();
    if (j < sz) {
        for (int i = j; i >= 0; --i) {
            double acc = (i == j) ? 1.0 : 0.0;
            for (int k = i + 1; k <= j; ++k) acc = fma(-Rs[i][k], Xs[k][j], acc);
            const double dgn = Rs[i][i];
            if (dgn == 0.0) { if (status) atomicOr(&status[b], SB_ST_SINGULAR); Xs[i][j] = 0.0; }
            else Xs[i][j] = acc / dgn;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < sz * sz; idx += blockDim.x) {
        const int i = idx / sz, c = idx % sz;
        X[(size_t)(r0 + i) * n + r0 + c] = Xs[i][c];
    }
}

// ---------------------------------------------------------------------------------------------------
// Blocked Cholesky G = R^T R (R upper) of symmetric positive definite matrices, in place, one CTA per system
// and panel: the diagonal block is factored in shared memory, the block row to its right is solved against it
// (R12 = R11^-T G12, one thread per column, coalesced along the row), the trailing block is updated by
// ONE GEMM per panel (G22 -= R12^T R12) launched by the host loop.  With G = Bw^T Bw this yields the R factor
// of the Wilson matrix from one big GEMM + n/32 small panels instead of a Householder QR (the geodesic stages
// only need R: B+ w = R^-1 R^-T Bw^T w).
constexpr int CH_NB = 32, CH_THREADS = 256;

__global__ void __launch_bounds__(CH_THREADS)
chol_panel_kernel(double* __restrict__ A_, int n, int p0, int nb, int* __restrict__ status,
                  const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    __shared__ double D[CH_NB][CH_NB + 1];
    __shared__ int bad;
    double* A = A_ + (size_t)b * n * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) bad = 0;
    for (int idx = tid; idx < nb * nb; idx += nt) {
        const int i = idx / nb, j = idx % nb;
        D[i][j] = A[(size_t)(p0 + i) * n + p0 + j];
    }
    __syncthreads();
    // unblocked upper Cholesky of the diagonal block: row k of R, then the rank-one downdate of the rest
    for (int k = 0; k < nb; ++k) {
        const double akk = D[k][k];
        __syncthreads();
        if (!(akk > 0.0)) {
            if (tid == 0) { bad = 1; }
            // keep going with a unit pivot so that every thread follows the same control flow
        }
        const double d = (akk > 0.0) ? sqrt(akk) : 1.0;
        for (int j = k + tid; j < nb; j += nt) D[k][j] = (j == k) ? d : D[k][j] / d;
        __syncthreads();
        const int rem = nb - k - 1;
        for (int idx = tid; idx < rem * rem; idx += nt) {
            const int i = k + 1 + idx / rem, j = k + 1 + idx % rem;
            if (j >= i) D[i][j] = fma(-D[k][i], D[k][j], D[i][j]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nb * nb; idx += nt) {
        const int i = idx / nb, j = idx % nb;
        A[(size_t)(p0 + i) * n + p0 + j] = (j >= i) ? D[i][j] : 0.0;
    }
    // block row to the right: solve R11^T x = g for every column (forward substitution)
    const int nr = n - p0 - nb;
    for (int c = tid; c < nr; c += nt) {
        double x[CH_NB];                       // fully unrolled below: stays in registers
        double* col = A + (size_t)p0 * n + p0 + nb + c;
#pragma unroll
        for (int i = 0; i < CH_NB; ++i) x[i] = (i < nb) ? col[(size_t)i * n] : 0.0;
#pragma unroll
        for (int k = 0; k < CH_NB; ++k) {
            // (rows beyond nb: D is not loaded there; their x stays 0 and is never stored)
            const double xk = (k < nb) ? x[k] / D[k][k] : 0.0;
            x[k] = xk;
#pragma unroll
            for (int i = k + 1; i < CH_NB; ++i) x[i] = (i < nb) ? fma(-D[k][i], xk, x[i]) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < CH_NB; ++i)
            if (i < nb) col[(size_t)i * n] = x[i];
    }
    if (tid == 0 && bad && status) atomicOr(&status[b], SB_ST_SINGULAR);
}

__global__ void zero_lower_kernel(double* __restrict__ X_, int n, const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    if (idx / n > idx % n) X_[(size_t)b * n * n + idx] = 0.0;
}

}  // namespace

extern "C" int sb_gemm_impl(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                            long long sA, const double* B, int ldb, long long sB, double beta, double* C, int ldc,
                            long long sC, const int* active, int batch, cudaStream_t st) {
    dim3 grid((N + GM_BN - 1) / GM_BN, (M + GM_BM - 1) / GM_BM, batch);
    SB_COUNT(1);
    gemm_kernel<<<grid, GM_THREADS, 0, st>>>(transA, transB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC,
                                             active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_gemm_impl(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
                            long long sA, const double* B, int ldb, long long sB, double beta, double* C, int ldc,
                            long long sC, const int* active, int batch, cudaStream_t st);

// work: batch * (m*n + 2*QR_NB*n + QR_NB*QR_NB) doubles (explicit reflector blocks, W1, W2, T)
extern "C" int sb_qr_impl(double* A, int m, int n, double* Q, double* R, double* work, const int* active, int batch,
                          cudaStream_t st) {
    double* Vp = work;
    double* W1 = Vp + (size_t)batch * m * n;
    double* W2 = W1 + (size_t)batch * QR_NB * n;
    double* T = W2 + (size_t)batch * QR_NB * n;
    const long long sA = (long long)m * n, sW = (long long)QR_NB * n, sT = QR_NB * QR_NB;
    const int npan = (n + QR_NB - 1) / QR_NB;
    // T of every panel is needed again for Q: keep them after W2 (npan blocks)
    // (work therefore holds batch * npan * QR_NB^2 doubles of T; see the size formula in the header)
    int rc = 0;
    for (int p = 0; p < npan; ++p) {
        const int p0 = p * QR_NB, nb = (n - p0 < QR_NB) ? n - p0 : QR_NB, mp = m - p0, nr = n - p0 - nb;
        double* Tp = T + (size_t)p * batch * sT;
        const size_t smem = ((size_t)nb * mp + 2 * QR_NB + QR_NB * QR_NB + SB_SCRATCH_DOUBLES) * sizeof(double);
        if (smem > 200 * 1024) return -2;
        cudaFuncSetAttribute(qr_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        SB_COUNT(1);
        qr_panel_kernel<<<batch, QR_THREADS, smem, st>>>(A, m, n, p0, nb, Vp, Tp, R, active);
        if (nr > 0) {
            const double* V = Vp + (size_t)p0 * n + p0;            // [mp, nb], ld n
            double* A22 = A + (size_t)p0 * n + p0 + nb;            // [mp, nr], ld n
            // W1 = V^T A22 [nb, nr];  W2 = T^T W1;  A22 -= V W2
            if ((rc = sb_gemm_impl(1, 0, nb, nr, mp, 1.0, V, n, sA, A22, n, sA, 0.0, W1, nr, sW, active, batch, st))) return rc;
            if ((rc = sb_gemm_impl(1, 0, nb, nr, nb, 1.0, Tp, QR_NB, sT, W1, nr, sW, 0.0, W2, nr, sW, active, batch, st))) return rc;
            if ((rc = sb_gemm_impl(0, 0, mp, nr, nb, -1.0, V, n, sA, W2, nr, sW, 1.0, A22, n, sA, active, batch, st))) return rc;
        }
    }
    {
        dim3 g1((n * n + 255) / 256, batch), g2((m * n + 255) / 256, batch);
        SB_COUNT(1);
        qr_copy_r_kernel<<<g1, 256, 0, st>>>(A, m, n, R, active);
        // Q == NULL: R only (the caller works with the semi-normal equations B+ w = R^-1 R^-T B^T w, e.g.
        // every stage of the geodesic integrator) -- the accumulation of Q is half of the work
        if (!Q) return SB_LAUNCH_CHECK();
        SB_COUNT(1);
        qr_init_q_kernel<<<g2, 256, 0, st>>>(Q, m, n, active);
    }
    for (int p = npan - 1; p >= 0; --p) {
        const int p0 = p * QR_NB, nb = (n - p0 < QR_NB) ? n - p0 : QR_NB, mp = m - p0, nq = n - p0;
        const double* Tp = T + (size_t)p * batch * sT;
        const double* V = Vp + (size_t)p0 * n + p0;
        double* Q22 = Q + (size_t)p0 * n + p0;                     // [mp, nq], ld n
        // Q22 <- (I - V T V^T) Q22
        if ((rc = sb_gemm_impl(1, 0, nb, nq, mp, 1.0, V, n, sA, Q22, n, sA, 0.0, W1, nq, sW, active, batch, st))) return rc;
        if ((rc = sb_gemm_impl(0, 0, nb, nq, nb, 1.0, Tp, QR_NB, sT, W1, nq, sW, 0.0, W2, nq, sW, active, batch, st))) return rc;
        if ((rc = sb_gemm_impl(0, 0, mp, nq, nb, -1.0, V, n, sA, W2, nq, sW, 1.0, Q22, n, sA, active, batch, st))) return rc;
    }
    return SB_LAUNCH_CHECK();
}

static int trtri_rec(const double* R, double* X, double* work, int n, int r0, int sz, int* status,
                     const int* active, int batch, cudaStream_t st) {
    if (sz <= TRI_BS) {
        SB_COUNT(1);
        trtri_block_kernel<<<batch, 64, 0, st>>>(R, X, n, r0, sz, status, active);
        return 0;
    }
    int h = ((sz / 2 + 15) / 16) * 16;
    if (h >= sz) h = sz / 2;
    int rc;
    if ((rc = trtri_rec(R, X, work, n, r0, h, status, active, batch, st))) return rc;
    if ((rc = trtri_rec(R, X, work, n, r0 + h, sz - h, status, active, batch, st))) return rc;
    const long long sN = (long long)n * n;
    const double* R12 = R + (size_t)r0 * n + r0 + h;          // [h, sz-h]
    const double* X11 = X + (size_t)r0 * n + r0;
    const double* X22 = X + (size_t)(r0 + h) * n + r0 + h;
    double* X12 = X + (size_t)r0 * n + r0 + h;
    // work <- R12 X22 ;  X12 <- -X11 work
    if ((rc = sb_gemm_impl(0, 0, h, sz - h, sz - h, 1.0, R12, n, sN, X22, n, sN, 0.0, work, sz - h, sN, active, batch, st))) return rc;
    return sb_gemm_impl(0, 0, h, sz - h, h, -1.0, X11, n, sN, work, sz - h, sN, 0.0, X12, n, sN, active, batch, st);
}

// work: batch * n * n doubles
extern "C" int sb_trtri_impl(const double* R, double* X, double* work, int n, int* status, const int* active,
                             int batch, cudaStream_t st) {
    // the strictly lower part must be zero BEFORE the merges: X22 of a level is read as a full block
    dim3 grid((n * n + 255) / 256, batch);
    SB_COUNT(1);
    zero_lower_kernel<<<grid, 256, 0, st>>>(X, n, active);
    const int rc = trtri_rec(R, X, work, n, 0, n, status, active, batch, st);
    if (rc) return rc;
    return SB_LAUNCH_CHECK();
}

// In place: A [b, n, n] symmetric positive definite (full storage) -> upper Cholesky factor R (A = R^T R),
// strictly lower part zeroed.  SB_ST_SINGULAR: a non-positive pivot (rank-deficient / indefinite input).
extern "C" int sb_potrf_impl(double* A, int n, int* status, const int* active, int batch, cudaStream_t st) {
    const long long sN = (long long)n * n;
    for (int p0 = 0; p0 < n; p0 += CH_NB) {
        const int nb = (n - p0 < CH_NB) ? n - p0 : CH_NB, nr = n - p0 - nb;
        SB_COUNT(1);
        chol_panel_kernel<<<batch, CH_THREADS, 0, st>>>(A, n, p0, nb, status, active);
        if (nr > 0) {
            const double* R12 = A + (size_t)p0 * n + p0 + nb;               // [nb, nr], ld n
            double* A22 = A + (size_t)(p0 + nb) * n + p0 + nb;             // [nr, nr], ld n
            const int rc = sb_gemm_impl(1, 0, nr, nr, nb, -1.0, R12, n, sN, R12, n, sN, 1.0, A22, n, sN, active, batch, st);
            if (rc) return rc;
        }
    }
    dim3 grid((n * n + 255) / 256, batch);
    SB_COUNT(1);
    zero_lower_kernel<<<grid, 256, 0, st>>>(A, n, active);
    return SB_LAUNCH_CHECK();
}

// TFLOP/s of the fp64 FMA pipe (kind 0) or of the fp64 tensor-core path (kind 1) on the current device:
// `ctas_per_sm` CTAs of 256 threads per SM, `iters` iterations; scratch: sms*ctas_per_sm*256 doubles.
// Synchronises the device (diagnostic, not on any product path).
extern "C" int sb_fp64_peak_impl(int kind, int iters, int ctas_per_sm, double* scratch, double* tflops) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_peak_kernel<<<grid, 256>>>(kind, iters / 8 + 1, scratch);       // warm-up
    cudaEventRecord(e0);
    fp64_peak_kernel<<<grid, 256>>>(kind, iters, scratch);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) return (int)err;
    // kind 0: 8 FMA = 16 flops per thread and iteration; kind 1: 4 DMMA of 8x8x4 = 4 * 512 flops per warp
    const double flops = kind == 0 ? (double)grid * 256 * 16.0 * iters : (double)grid * 8 * 4.0 * 512.0 * iters;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return 0;
}
