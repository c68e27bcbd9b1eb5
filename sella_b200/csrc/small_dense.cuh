// Tiny dense linear algebra on k x k matrices (k <= SB_KMAX) held in shared
// memory with leading dimension SB_KLD.  These stand in for the LAPACK drivers the
// reference calls on subspace-sized problems: numpy.linalg.lstsq / scipy solve on
// the Gram blocks (sella/hessian_update.py:18-21,124,130), and scipy.linalg.eigh
// on the Rayleigh-Ritz matrices, plain and generalised
// (sella/eigensolvers.py:58, sella/peswrapper.py:547, sella/hessian_update.py:61).
//
// "serial" routines are meant to be executed by ONE thread; "warp" routines by one
// full warp (all 32 lanes must call).
#pragma once
#include "common.cuh"

#define SB_KMAX 32
#define SB_KLD 33
#define SB_KMAT (SB_KMAX * SB_KLD)

// Solve A X = B in place (Gaussian elimination, partial pivoting).  A: k x k
// (destroyed), B: k x nrhs with the same leading dimension.  Returns false on an
// exactly zero pivot (B is then left with the minimum-effort partial result).
__device__ inline bool sbs_solve_serial(double* A, int k, double* B, int nrhs) {
    bool ok = true;
    for (int c = 0; c < k; ++c) {
        int piv = c;
        double best = fabs(A[c * SB_KLD + c]);
        for (int r = c + 1; r < k; ++r) {
            const double a = fabs(A[r * SB_KLD + c]);
            if (a > best) { best = a; piv = r; }
        }
        if (best == 0.0) { ok = false; continue; }
        if (piv != c) {
            for (int j = 0; j < k; ++j) {
                const double t = A[c * SB_KLD + j]; A[c * SB_KLD + j] = A[piv * SB_KLD + j]; A[piv * SB_KLD + j] = t;
            }
            for (int j = 0; j < nrhs; ++j) {
                const double t = B[c * SB_KLD + j]; B[c * SB_KLD + j] = B[piv * SB_KLD + j]; B[piv * SB_KLD + j] = t;
            }
        }
        const double inv = 1.0 / A[c * SB_KLD + c];
        for (int r = c + 1; r < k; ++r) {
            const double f = A[r * SB_KLD + c] * inv;
            if (f == 0.0) continue;
            for (int j = c + 1; j < k; ++j) A[r * SB_KLD + j] -= f * A[c * SB_KLD + j];
            for (int j = 0; j < nrhs; ++j) B[r * SB_KLD + j] -= f * B[c * SB_KLD + j];
        }
    }
    for (int c = k - 1; c >= 0; --c) {
        const double dgn = A[c * SB_KLD + c];
        if (dgn == 0.0) continue;
        for (int j = 0; j < nrhs; ++j) {
            double acc = B[c * SB_KLD + j];
            for (int r = c + 1; r < k; ++r) acc -= A[c * SB_KLD + r] * B[r * SB_KLD + j];
            B[c * SB_KLD + j] = acc / dgn;
        }
    }
    return ok;
}

// Inverse of a k x k matrix: Ainv <- A^-1 (A destroyed).
__device__ inline bool sbs_invert_serial(double* A, int k, double* Ainv) {
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) Ainv[i * SB_KLD + j] = (i == j) ? 1.0 : 0.0;
    return sbs_solve_serial(A, k, Ainv, k);
}

// In-place lower Cholesky factor of the (lower triangle of the) SPD matrix M.
__device__ inline bool sbs_cholesky_serial(double* M, int k) {
    for (int j = 0; j < k; ++j) {
        double dj = M[j * SB_KLD + j];
        for (int p = 0; p < j; ++p) dj -= M[j * SB_KLD + p] * M[j * SB_KLD + p];
        if (!(dj > 0.0)) return false;
        dj = sqrt(dj);
        M[j * SB_KLD + j] = dj;
        for (int i = j + 1; i < k; ++i) {
            double v = M[i * SB_KLD + j];
            for (int p = 0; p < j; ++p) v -= M[i * SB_KLD + p] * M[j * SB_KLD + p];
            M[i * SB_KLD + j] = v / dj;
        }
    }
    return true;
}

// Cyclic two-sided Jacobi eigensolver executed by one warp.  C: symmetric k x k
// (both triangles valid on entry; destroyed).  On exit w[0..k) ascending and the
// columns of Z the matching orthonormal eigenvectors.  `perm` is k ints of scratch.
__device__ inline void sbs_jacobi_warp(double* C, int k, double* Z, double* w, int* perm) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < k * k; i += 32) Z[(i / k) * SB_KLD + (i % k)] = ((i / k) == (i % k)) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        // off-diagonal and diagonal magnitudes
        double off = 0.0, dg = 0.0;
        for (int i = lane; i < k * k; i += 32) {
            const int r = i / k, c = i % k;
            const double v = C[r * SB_KLD + c];
            if (r != c) off += v * v; else dg += v * v;
        }
        off = sb_warp_sum(off);
        dg = sb_warp_sum(dg);
        if (off <= 1e-34 * dg || off == 0.0) break;
        for (int p = 0; p < k - 1; ++p) {
            for (int q = p + 1; q < k; ++q) {
                const double apq = C[p * SB_KLD + q];
                const double app = C[p * SB_KLD + p], aqq = C[q * SB_KLD + q];
                // skip rotations that cannot change anything at double precision
                if (fabs(apq) <= 1e-300 || fabs(apq) <= 1e-19 * (fabs(app) + fabs(aqq))) {
                    __syncwarp();           // all lanes have read C[p][q] before lane 0 clears it
                    if (lane == 0 && apq != 0.0 && fabs(apq) <= 1e-19 * (fabs(app) + fabs(aqq))) {
                        C[p * SB_KLD + q] = 0.0; C[q * SB_KLD + p] = 0.0;
                    }
                    __syncwarp();
                    continue;
                }
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0);
                const double s = t * c;
                __syncwarp();
                if (lane < k) {               // columns p,q of row `lane`:  C <- C P,  Z <- Z P
                    const int r = lane;
                    const double ap = C[r * SB_KLD + p], aq = C[r * SB_KLD + q];
                    C[r * SB_KLD + p] = c * ap - s * aq;
                    C[r * SB_KLD + q] = s * ap + c * aq;
                    const double zp = Z[r * SB_KLD + p], zq = Z[r * SB_KLD + q];
                    Z[r * SB_KLD + p] = c * zp - s * zq;
                    Z[r * SB_KLD + q] = s * zp + c * zq;
                }
                __syncwarp();
                if (lane < k) {               // rows p,q of column `lane`:  C <- P^T C
                    const int j = lane;
                    const double a = C[p * SB_KLD + j], b = C[q * SB_KLD + j];
                    C[p * SB_KLD + j] = c * a - s * b;
                    C[q * SB_KLD + j] = s * a + c * b;
                }
                __syncwarp();
                if (lane == 0) { C[p * SB_KLD + q] = 0.0; C[q * SB_KLD + p] = 0.0; }
                __syncwarp();
            }
        }
    }
    // ascending order (rank sort; ties broken by index)
    if (lane < k) {
        const double di = C[lane * SB_KLD + lane];
        int rank = 0;
        for (int j = 0; j < k; ++j) {
            const double dj = C[j * SB_KLD + j];
            rank += (dj < di) || (dj == di && j < lane);
        }
        perm[rank] = lane;
        w[rank] = di;
    }
    __syncwarp();
    // permute columns of Z through C as scratch
    for (int i = lane; i < k * k; i += 32) {
        const int r = i / k, c = i % k;
        C[r * SB_KLD + c] = Z[r * SB_KLD + perm[c]];
    }
    __syncwarp();
    for (int i = lane; i < k * k; i += 32) {
        const int r = i / k, c = i % k;
        Z[r * SB_KLD + c] = C[r * SB_KLD + c];
    }
    __syncwarp();
}

// Generalised symmetric-definite eigenproblem  A z = lambda M z  (warp):
// uses the LOWER triangles of A and M, exactly as LAPACK dsygvd does for
// scipy.linalg.eigh(a, b).  On exit: w ascending, R (k x k) with columns z
// normalised so that R^T M R = I.  A, M are destroyed; T is k x k scratch.
__device__ inline bool sbs_gen_eigh_warp(double* A, double* M, int k, double* R, double* w, double* T,
                                         int* perm) {
    const int lane = threadIdx.x & 31;
    __shared__ int ok_flag;
    if (lane == 0) {
        bool ok = sbs_cholesky_serial(M, k);          // M <- L (lower)
        ok_flag = ok ? 1 : 0;
        if (ok) {
            // symmetrise A from its lower triangle
            for (int i = 0; i < k; ++i)
                for (int j = i + 1; j < k; ++j) A[i * SB_KLD + j] = A[j * SB_KLD + i];
            // T = L^-1 A   (forward substitution on each column)
            for (int c = 0; c < k; ++c)
                for (int i = 0; i < k; ++i) {
                    double v = A[i * SB_KLD + c];
                    for (int p = 0; p < i; ++p) v -= M[i * SB_KLD + p] * T[p * SB_KLD + c];
                    T[i * SB_KLD + c] = v / M[i * SB_KLD + i];
                }
            // A = T L^-T  ->  row r of A solves  L y = (row r of T)^T
            for (int r = 0; r < k; ++r)
                for (int i = 0; i < k; ++i) {
                    double v = T[r * SB_KLD + i];
                    for (int p = 0; p < i; ++p) v -= M[i * SB_KLD + p] * A[r * SB_KLD + p];
                    A[r * SB_KLD + i] = v / M[i * SB_KLD + i];
                }
            // enforce exact symmetry before Jacobi
            for (int i = 0; i < k; ++i)
                for (int j = i + 1; j < k; ++j) {
                    const double m = 0.5 * (A[i * SB_KLD + j] + A[j * SB_KLD + i]);
                    A[i * SB_KLD + j] = m; A[j * SB_KLD + i] = m;
                }
        }
    }
    __syncwarp();
    if (!ok_flag) return false;
    sbs_jacobi_warp(A, k, T, w, perm);               // T <- eigenvectors of the standard problem
    if (lane < k) {
        // R[:, lane] = L^-T T[:, lane]  (back substitution)
        const int c = lane;
        for (int i = k - 1; i >= 0; --i) {
            double v = T[i * SB_KLD + c];
            for (int p = i + 1; p < k; ++p) v -= M[p * SB_KLD + i] * R[p * SB_KLD + c];
            R[i * SB_KLD + c] = v / M[i * SB_KLD + i];
        }
    }
    __syncwarp();
    return true;
}
