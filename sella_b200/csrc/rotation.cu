// Rotation coordinate of a whole configuration: value (rotation vector, 3 components), Jacobian
// rows and the multiplier-contracted Hessian sum_k L_k d2v_k/dx2, batched over systems.
//
// Reference: sella/internal.py  _build_F_matrix_np :534-551, _stabilize_quaternion(_from_eigh)
// :554-585, _asinc_np / _expmap_np :588-604, _rotation_3axis_jacobian_np :607-648, _apply_dF
// :651-700, _rotation_hessian_single :703-800, class Rotation :1031-1078.  The unit quaternion q
// that best superimposes the centred positions on the centred reference is the top eigenvector
// of the 4x4 matrix F (linear in the positions); v = 2 asinc(q0) q[1:4].  Derivatives follow from
// first- and second-order eigenvector perturbation theory with the pseudo-inverse
// Minv = V diag(1/(w_k - w_top)) V^T restricted to the other eigenvectors:
//   dq_a   = -Minv (dF_a q)                                   a = (atom, Cartesian direction)
//   H_ab   = [d2f dq_a + 2 (w.q) dF_a q - dF_a w - (df.q) dq_a] . dq_b - (dF_b w).dq_a
//            + dE_a (dq_b.w) + dE_b (dq_a.w),   w = Minv df,  dE_a = (dF_a q).q,
// where df, d2f are the first / second derivatives of f(q) = sum_k L_k 2 asinc(q0) q_{k+1}.
// One CTA per system: the 4x4 algebra by one thread, the 3N per-coordinate 4-vectors by all
// threads, the 3N x 3N accumulation with warps over rows and lanes over columns.
#include "common.cuh"

namespace {

constexpr int ROT_THREADS = 256;

// cyclic Jacobi for a symmetric 4x4 (row-major A, destroyed); w ascending, V columns = eigenvectors
__device__ void jacobi4(double* A, double* w, double* V) {
    for (int i = 0; i < 16; ++i) V[i] = (i / 4 == i % 4) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) { if (i != j) off += A[4 * i + j] * A[4 * i + j]; else dg += A[4 * i + j] * A[4 * i + j]; }
        if (off <= 1e-36 * dg || off == 0.0) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[4 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[4 * q + q] - A[4 * p + p]) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < 4; ++r) {
                    const double ap = A[4 * r + p], aq = A[4 * r + q];
                    A[4 * r + p] = c * ap - s * aq; A[4 * r + q] = s * ap + c * aq;
                    const double vp = V[4 * r + p], vq = V[4 * r + q];
                    V[4 * r + p] = c * vp - s * vq; V[4 * r + q] = s * vp + c * vq;
                }
                for (int r = 0; r < 4; ++r) {
                    const double ap = A[4 * p + r], aq = A[4 * q + r];
                    A[4 * p + r] = c * ap - s * aq; A[4 * q + r] = s * ap + c * aq;
                }
                A[4 * p + q] = 0.0; A[4 * q + p] = 0.0;
            }
    }
    // ascending order by rank counting (ties by index)
    double Vs[16], dgn[4];
    for (int k = 0; k < 4; ++k) dgn[k] = A[5 * k];
    for (int k = 0; k < 4; ++k) {
        int rank = 0;
        for (int j = 0; j < 4; ++j) rank += (dgn[j] < dgn[k]) || (dgn[j] == dgn[k] && j < k);
        w[rank] = dgn[k];
        for (int r = 0; r < 4; ++r) Vs[4 * r + rank] = V[4 * r + k];
    }
    for (int i = 0; i < 16; ++i) V[i] = Vs[i];
}

__device__ double asinc_dev(double x) {                       // internal.py:588-597
    if (x < 0.97) return acos(x) / sqrt(1.0 - x * x);
    const double y = x - 1.0;
    return 1.0 - y / 3 + 2 * y * y / 15 - 2 * y * y * y / 35 + 8 * pow(y, 4) / 315 - 8 * pow(y, 5) / 693 +
           16 * pow(y, 6) / 3003 - 16 * pow(y, 7) / 6435 + 128 * pow(y, 8) / 109395 - 128 * pow(y, 9) / 230945;
}

// (dF/dx_{k,d}) v for the atom with centred reference position y (internal.py:651-700)
__device__ __forceinline__ void apply_dF(const double* y, int d, const double* v, double* out) {
    const int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    double top[3] = {0.0, 0.0, 0.0};
    top[d1] = -y[d2];
    top[d2] = y[d1];
    const double yv = y[0] * v[1] + y[1] * v[2] + y[2] * v[3];
    out[0] = y[d] * v[0] + top[0] * v[1] + top[1] * v[2] + top[2] * v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double val = -y[d] * v[1 + i] + y[i] * v[1 + d];
        if (i == d) val += yv;
        out[1 + i] = top[i] * v[0] + val;
    }
}

struct RotShared {
    double F[16], V[16], Minv[16], d2f[16];
    double ws[4], q[4], w[4], df[4];
    double a_j, da_j, wc, fdqc;
    double mean[3], R[9];
    double scratch[SB_SCRATCH_DOUBLES];
};

__global__ void __launch_bounds__(ROT_THREADS)
rotation_kernel(const double* __restrict__ x_, int natoms, const double* __restrict__ ref_, long long refstride,
                double* __restrict__ qprev_, double* __restrict__ vals_, long long valstride,
                double* __restrict__ J_, long long jstride, const double* __restrict__ L_, long long lstride,
                double* __restrict__ D_, double* __restrict__ work_, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    __shared__ RotShared S;
    const int n = 3 * natoms;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const double* x = x_ + (size_t)b * n;
    const double* y = ref_ + (size_t)b * refstride;
    for (int d = 0; d < 3; ++d) {
        double acc = 0.0;
        for (int i = tid; i < natoms; i += nt) acc += x[3 * i + d];
        const double s = sb_block_sum(acc, S.scratch);
        if (tid == 0) S.mean[d] = s / natoms;
    }
    __syncthreads();
    for (int e = 0; e < 9; ++e) {                      // R = dx^T y
        const int i = e / 3, j = e % 3;
        double acc = 0.0;
        for (int k = tid; k < natoms; k += nt) acc = fma(x[3 * k + i] - S.mean[i], y[3 * k + j], acc);
        const double s = sb_block_sum(acc, S.scratch);
        if (tid == 0) S.R[e] = s;
    }
    __syncthreads();
    if (tid == 0) {
        const double* R = S.R;
        const double tr = R[0] + R[4] + R[8];
        const double top[3] = {R[5] - R[7], R[6] - R[2], R[1] - R[3]};
        S.F[0] = tr;
        for (int i = 0; i < 3; ++i) { S.F[1 + i] = top[i]; S.F[4 * (1 + i)] = top[i]; }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) S.F[4 * (1 + i) + 1 + j] = (i == j ? -tr : 0.0) + R[3 * i + j] + R[3 * j + i];
        double A[16];
        for (int i = 0; i < 16; ++i) A[i] = S.F[i];
        jacobi4(A, S.ws, S.V);
        // branch-stable quaternion (:569-585)
        double qp[4] = {1.0, 0.0, 0.0, 0.0};
        if (qprev_) for (int i = 0; i < 4; ++i) qp[i] = qprev_[(size_t)b * 4 + i];
        double q[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < 4; ++k)
            if (S.ws[3] - S.ws[k] < 1e-10) {
                double c = 0.0;
                for (int r = 0; r < 4; ++r) c += S.V[4 * r + k] * qp[r];
                for (int r = 0; r < 4; ++r) q[r] += c * S.V[4 * r + k];
            }
        double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (nrm < 1e-14) { for (int r = 0; r < 4; ++r) q[r] = S.V[4 * r + 3]; nrm = 1.0; }
        const double sg = q[0] < 0.0 ? -1.0 : 1.0;
        for (int r = 0; r < 4; ++r) { q[r] = sg * q[r] / nrm; S.q[r] = q[r]; }
        if (qprev_) for (int r = 0; r < 4; ++r) qprev_[(size_t)b * 4 + r] = q[r];
        double inv[4];
        for (int k = 0; k < 4; ++k) { const double gap = S.ws[k] - S.ws[3]; inv[k] = fabs(gap) > 1e-14 ? 1.0 / gap : 0.0; }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double acc = 0.0;
                for (int k = 0; k < 4; ++k) acc += S.V[4 * i + k] * inv[k] * S.V[4 * j + k];
                S.Minv[4 * i + j] = acc;
            }
        const double q0 = q[0];
        const double a_j = asinc_dev(q0);
        double da_j = 0.0;                             // the Jacobian's own branches (:629-641)
        if (fabs(q0 - 1.0) < 1e-8) da_j = -1.0 / 3 + 4 * (q0 - 1.0) / 15;
        else if (fabs(q0) < 1.0 - 1e-12) { const double s2 = 1 - q0 * q0; da_j = -1.0 / s2 + q0 * acos(q0) / (sqrt(s2) * s2); }
        S.a_j = a_j; S.da_j = da_j;
        if (vals_) for (int k = 0; k < 3; ++k) vals_[(size_t)b * valstride + k] = 2.0 * q[k + 1] * a_j;
        if (L_) {
            double a, da, d2a;                         // the Hessian's branches (:739-764)
            if (fabs(q0 - 1.0) < 1e-8) { const double yy = q0 - 1.0; a = 1 - yy / 3 + 2 * yy * yy / 15; da = -1.0 / 3 + 4 * yy / 15; d2a = 4.0 / 15; }
            else if (fabs(q0) < 1.0 - 1e-12) {
                const double s2 = 1 - q0 * q0, s = sqrt(s2), ac = acos(q0);
                a = ac / s; da = -1.0 / s2 + q0 * ac / (s * s2);
                d2a = (3 * q0 / s2 - (1 + 2 * q0 * q0) * ac / (s * s2)) * (-1.0 / s2);
            } else { a = q0 > 0 ? 1.5707963267948966 : -1.5707963267948966; da = 0.0; d2a = 0.0; }
            for (int i = 0; i < 4; ++i) S.df[i] = 0.0;
            for (int i = 0; i < 16; ++i) S.d2f[i] = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double Lk = L_[(size_t)b * lstride + k];
                S.df[0] += Lk * 2 * q[k + 1] * da;
                S.df[k + 1] += Lk * 2 * a;
                S.d2f[0] += Lk * 2 * q[k + 1] * d2a;
                S.d2f[k + 1] += Lk * 2 * da;
                S.d2f[4 * (k + 1)] += Lk * 2 * da;
            }
            double wc = 0.0, fq = 0.0;
            for (int i = 0; i < 4; ++i) {
                double acc = 0.0;
                for (int j = 0; j < 4; ++j) acc += S.Minv[4 * i + j] * S.df[j];
                S.w[i] = acc;
            }
            for (int i = 0; i < 4; ++i) { wc += S.w[i] * q[i]; fq += S.df[i] * q[i]; }
            S.wc = wc; S.fdqc = fq;
        }
    }
    __syncthreads();
    const bool hess = L_ != nullptr && D_ != nullptr;
    // per-coordinate 4-vectors: work[b] = dc[n][4] | p[n][4] | dFw[n][4] | dE[n] | wdc[n]
    double* dc = work_ + (size_t)b * n * 14;
    double* pv = dc + (size_t)4 * n;
    double* dFw = pv + (size_t)4 * n;
    double* dE = dFw + (size_t)4 * n;
    double* wdc = dE + n;
    for (int a = tid; a < n; a += nt) {
        const int k = a / 3, d = a % 3;
        double dFq[4], dca[4];
        apply_dF(y + 3 * k, d, S.q, dFq);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = fma(S.Minv[4 * i + j], dFq[j], acc);
            dca[i] = -acc;
        }
        if (J_)
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
                J_[(size_t)b * jstride + (size_t)kk * n + a] = 2.0 * (dca[kk + 1] * S.a_j + S.q[kk + 1] * S.da_j * dca[0]);
        if (hess) {
            double fw[4];
            apply_dF(y + 3 * k, d, S.w, fw);
            double e = 0.0, wd = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { e = fma(dFq[i], S.q[i], e); wd = fma(dca[i], S.w[i], wd); }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc = fma(S.d2f[4 * i + j], dca[j], acc);
                pv[4 * a + i] = acc + 2.0 * S.wc * dFq[i] - fw[i] - S.fdqc * dca[i];
                dc[4 * a + i] = dca[i];
                dFw[4 * a + i] = fw[i];
            }
            dE[a] = e; wdc[a] = wd;
        }
    }
    if (!hess) return;
    __syncthreads();
    double* D = D_ + (size_t)b * n * n;
    for (int a = warp; a < n; a += nw) {
        double pa[4], dca[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { pa[i] = pv[4 * a + i]; dca[i] = dc[4 * a + i]; }
        const double ea = dE[a], wa = wdc[a];
        for (int c = lane; c < n; c += 32) {
            double h = ea * wdc[c] + dE[c] * wa;
#pragma unroll
            for (int i = 0; i < 4; ++i) h += pa[i] * dc[4 * c + i] - dFw[4 * c + i] * dca[i];
            D[(size_t)a * n + c] += h;
        }
    }
}

}  // namespace

extern "C" int sb_rotation_impl(const double* x, int natoms, const double* ref, long long refstride, double* qprev,
                                double* vals, long long valstride, double* J, long long jstride, const double* L,
                                long long lstride, double* D, double* work, const int* active, int batch,
                                cudaStream_t st) {
    SB_COUNT(1);
    rotation_kernel<<<batch, ROT_THREADS, 0, st>>>(x, natoms, ref, refstride, qprev, vals, valstride, J, jstride, L,
                                                   lstride, D, work, active);
    return SB_LAUNCH_CHECK();
}
