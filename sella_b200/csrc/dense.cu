// Batched dense fp64 linear algebra for the internal-coordinate path:
//   sb_gemm   C = alpha op(A) op(B) + beta C         DMMA (mma.sync m8n8k4 f64) tensor-core tiles
//   sb_qr     economy Householder QR  A = Q R         sella/_gpu.py:100-111 (gpu_qr),
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

constexpr int QR_THREADS = 256;

// In-place unblocked Householder QR of A [m, n] (row-major, m >= n), LAPACK dgeqr2 conventions
// (v_j[j] = 1 implicit, H_j = I - tau_j v_j v_j^T, R = upper triangle, beta_j = -sign(alpha)|x|),
// followed by the explicit economy Q [m, n] (dorg2r) and R [n, n].  One CTA per system; threads
// own columns, so every access to a matrix row is coalesced and the two passes over the trailing
// block (w = v^T A, A -= tau v w^T) need no barrier between them.
__global__ void __launch_bounds__(QR_THREADS)
qr_kernel(double* __restrict__ A_, int m, int n, double* __restrict__ Q_, double* __restrict__ R_,
          const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    extern __shared__ double sm[];
    double* v = sm;                         // m
    double* tau = v + m;                    // n
    double* scratch = tau + n;              // SB_SCRATCH_DOUBLES
    double* A = A_ + (size_t)b * m * n;
    double* Q = Q_ + (size_t)b * m * n;
    double* R = R_ + (size_t)b * n * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int i = j + 1 + tid; i < m; i += nt) { const double x = A[(size_t)i * n + j]; v[i] = x; acc = fma(x, x, acc); }
        const double sigma = sb_block_sum(acc, scratch);
        const double alpha = A[(size_t)j * n + j];
        double tj = 0.0;
        if (sigma > 0.0) {
            const double nrm = sqrt(alpha * alpha + sigma);
            const double beta = alpha >= 0.0 ? -nrm : nrm;
            tj = (beta - alpha) / beta;
            const double scale = 1.0 / (alpha - beta);
            for (int i = j + 1 + tid; i < m; i += nt) { const double x = v[i] * scale; v[i] = x; A[(size_t)i * n + j] = x; }
            if (tid == 0) A[(size_t)j * n + j] = beta;
        }
        if (tid == 0) { tau[j] = tj; v[j] = 1.0; }
        __syncthreads();
        if (tj != 0.0) {
            for (int c = j + 1 + tid; c < n; c += nt) {
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
                int i = j;
                for (; i + 3 < m; i += 4) {
                    w0 = fma(v[i], A[(size_t)i * n + c], w0);
                    w1 = fma(v[i + 1], A[(size_t)(i + 1) * n + c], w1);
                    w2 = fma(v[i + 2], A[(size_t)(i + 2) * n + c], w2);
                    w3 = fma(v[i + 3], A[(size_t)(i + 3) * n + c], w3);
                }
                for (; i < m; ++i) w0 = fma(v[i], A[(size_t)i * n + c], w0);
                const double w = tj * ((w0 + w1) + (w2 + w3));
                for (i = j; i < m; ++i) A[(size_t)i * n + c] = fma(-w, v[i], A[(size_t)i * n + c]);
            }
        }
        __syncthreads();
    }
    // R and Q = H_0 ... H_{n-1} [I; 0]
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx / n, c = idx % n;
        R[idx] = c >= i ? A[(size_t)i * n + c] : 0.0;
    }
    for (int idx = tid; idx < m * n; idx += nt) Q[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
    __syncthreads();
    for (int j = n - 1; j >= 0; --j) {
        for (int i = j + 1 + tid; i < m; i += nt) v[i] = A[(size_t)i * n + j];
        if (tid == 0) v[j] = 1.0;
        __syncthreads();
        const double tj = tau[j];
        if (tj != 0.0) {
            for (int c = j + tid; c < n; c += nt) {
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
                int i = j;
                for (; i + 3 < m; i += 4) {
                    w0 = fma(v[i], Q[(size_t)i * n + c], w0);
                    w1 = fma(v[i + 1], Q[(size_t)(i + 1) * n + c], w1);
                    w2 = fma(v[i + 2], Q[(size_t)(i + 2) * n + c], w2);
                    w3 = fma(v[i + 3], Q[(size_t)(i + 3) * n + c], w3);
                }
                for (; i < m; ++i) w0 = fma(v[i], Q[(size_t)i * n + c], w0);
                const double w = tj * ((w0 + w1) + (w2 + w3));
                for (i = j; i < m; ++i) Q[(size_t)i * n + c] = fma(-w, v[i], Q[(size_t)i * n + c]);
            }
        }
        __syncthreads();
    }
}

// Rinv = R^-1 for upper-triangular R [n, n] (row-major).  Thread j owns column j of the
// inverse (back substitution); status bit SB_ST_SINGULAR on a zero diagonal.
__global__ void __launch_bounds__(256)
trtri_kernel(const double* __restrict__ R_, double* __restrict__ X_, int n, int* __restrict__ status,
             const int* __restrict__ active) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const double* R = R_ + (size_t)b * n * n;
    double* X = X_ + (size_t)b * n * n;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = n - 1; i > j; --i) X[(size_t)i * n + j] = 0.0;
    for (int i = j; i >= 0; --i) {
        double acc = (i == j) ? 1.0 : 0.0;
        for (int k = i + 1; k <= j; ++k) acc = fma(-R[(size_t)i * n + k], X[(size_t)k * n + j], acc);
        const double dgn = R[(size_t)i * n + i];
        if (dgn == 0.0) { if (status) atomicOr(&status[b], SB_ST_SINGULAR); X[(size_t)i * n + j] = 0.0; }
        else X[(size_t)i * n + j] = acc / dgn;
    }
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

extern "C" int sb_qr_impl(double* A, int m, int n, double* Q, double* R, const int* active, int batch,
                          cudaStream_t st) {
    const size_t smem = ((size_t)m + n + SB_SCRATCH_DOUBLES) * sizeof(double);
    cudaFuncSetAttribute(qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SB_COUNT(1);
    qr_kernel<<<batch, QR_THREADS, smem, st>>>(A, m, n, Q, R, active);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_trtri_impl(const double* R, double* X, int n, int* status, const int* active, int batch,
                             cudaStream_t st) {
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    trtri_kernel<<<grid, 256, 0, st>>>(R, X, n, status, active);
    return SB_LAUNCH_CHECK();
}
