// Internal coordinates: values q(x), Wilson B-matrix dq/dx and the contraction of the
// per-coordinate second derivatives with a vector, batched over systems.
//
// Reference (file:line relative to the reference tree), where all derivatives come from
// JAX autodiff of the primal formulas on the CPU:
//   primal formulas       sella/internal.py:58-80  (_bond_value, _angle_value, _dihedral_value)
//                         sella/internal.py:466-470 (_translation: mean of pos[:, dim])
//   BaseInternals.calc    sella/internal.py:1735-1778   -> q
//   BaseInternals.jacobian sella/internal.py:1780-1902  -> B (nint x 3N), rows scattered per coordinate
//   BaseInternals.hessian + SparseInternalHessians.ldot   sella/internal.py:2189-2305,
//                         sella/linalg.py:601-618       -> D(v) = sum_i v_i d2q_i/dx2  (3N x 3N)
//   hessian_rdot / SparseInternalHessians.rdot            sella/internal.py:2307-2575,
//                         sella/linalg.py:620-646       -> R[i,:] = (d2q_i/dx2) w
// Here the same primal formulas are differentiated by forward-mode (hyper-)dual numbers
// inside the kernel, one thread per (system, coordinate); PBC shift vectors (tvec = ncvec @
// cell) enter exactly as in the reference.  Coordinate order: translations, bonds, angles,
// dihedrals (the reference's _names order, internal.py:1209ff).
//
// Index arrays (int32) and shift vectors are shared by the whole batch (same topology,
// different geometries), which is the batched-search use case.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ hyper-dual numbers
// value, gradient (NV entries) and, if H, the symmetric Hessian (NV*(NV+1)/2 entries).
template <int NV, bool H>
struct HD {
    static constexpr int NH = H ? NV * (NV + 1) / 2 : 1;
    double v;
    double g[NV];
    double h[NH];
    __device__ static int idx(int i, int j) { return i <= j ? i * NV - i * (i - 1) / 2 + (j - i) : idx(j, i); }
    __device__ HD() {}
    __device__ explicit HD(double c) : v(c) {
#pragma unroll
        for (int i = 0; i < NV; ++i) g[i] = 0.0;
        if (H) {
#pragma unroll
            for (int i = 0; i < NH; ++i) h[i] = 0.0;
        }
    }
    __device__ static HD var(double x, int k) { HD r(x); r.g[k] = 1.0; return r; }
};

template <int NV, bool H>
__device__ HD<NV, H> operator+(const HD<NV, H>& a, const HD<NV, H>& b) {
    HD<NV, H> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] + b.g[i];
    if (H) {
#pragma unroll
        for (int i = 0; i < HD<NV, H>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
    }
    return r;
}
template <int NV, bool H>
__device__ HD<NV, H> operator-(const HD<NV, H>& a, const HD<NV, H>& b) {
    HD<NV, H> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] - b.g[i];
    if (H) {
#pragma unroll
        for (int i = 0; i < HD<NV, H>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
    }
    return r;
}
template <int NV, bool H>
__device__ HD<NV, H> operator*(const HD<NV, H>& a, const HD<NV, H>& b) {
    HD<NV, H> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
    if (H) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = i; j < NV; ++j, ++k)
                r.h[k] = a.h[k] * b.v + a.v * b.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
    }
    return r;
}
// f(a) for a scalar function with derivatives f1 = f'(a.v), f2 = f''(a.v)
template <int NV, bool H>
__device__ HD<NV, H> chain(const HD<NV, H>& a, double f0, double f1, double f2) {
    HD<NV, H> r;
    r.v = f0;
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = f1 * a.g[i];
    if (H) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = i; j < NV; ++j, ++k) r.h[k] = f1 * a.h[k] + f2 * a.g[i] * a.g[j];
    }
    return r;
}
template <int NV, bool H>
__device__ HD<NV, H> hd_sqrt(const HD<NV, H>& a) {
    const double s = sqrt(a.v);
    return chain(a, s, 0.5 / s, -0.25 / (s * a.v));
}
template <int NV, bool H>
__device__ HD<NV, H> hd_inv(const HD<NV, H>& a) {
    const double r = 1.0 / a.v;
    return chain(a, r, -r * r, 2.0 * r * r * r);
}
template <int NV, bool H>
__device__ HD<NV, H> hd_acos(const HD<NV, H>& a) {
    // clip to [-1, 1] as the reference does (internal.py:69): at the clip the derivative is 0
    if (a.v >= 1.0) return HD<NV, H>(0.0);
    if (a.v <= -1.0) return HD<NV, H>(M_PI);
    const double om = 1.0 - a.v * a.v;
    const double s = sqrt(om);
    return chain(a, acos(a.v), -1.0 / s, -a.v / (om * s));
}
// atan2(y, x)
template <int NV, bool H>
__device__ HD<NV, H> hd_atan2(const HD<NV, H>& y, const HD<NV, H>& x) {
    const double r2 = x.v * x.v + y.v * y.v;
    const double fy = x.v / r2, fx = -y.v / r2;                  // d/dy, d/dx
    const double fyy = -2.0 * x.v * y.v / (r2 * r2), fxx = -fyy;
    const double fxy = (y.v * y.v - x.v * x.v) / (r2 * r2);
    HD<NV, H> r;
    r.v = atan2(y.v, x.v);
#pragma unroll
    for (int i = 0; i < NV; ++i) r.g[i] = fy * y.g[i] + fx * x.g[i];
    if (H) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = i; j < NV; ++j, ++k)
                r.h[k] = fy * y.h[k] + fx * x.h[k] + fyy * y.g[i] * y.g[j] + fxx * x.g[i] * x.g[j] +
                         fxy * (y.g[i] * x.g[j] + y.g[j] * x.g[i]);
    }
    return r;
}

template <class T>
struct V3 { T x, y, z; };
template <class T> __device__ V3<T> vsub(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> __device__ V3<T> vadd(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> __device__ T vdot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> __device__ V3<T> vcross(const V3<T>& a, const V3<T>& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// primal formulas, sella/internal.py:58-80, on hyper-dual coordinates
template <class T> __device__ T bond_value(const V3<T>* p, const V3<T>* t) {
    const V3<T> d = vadd(vsub(p[1], p[0]), t[0]);
    return hd_sqrt(vdot(d, d));
}
template <class T> __device__ T angle_value(const V3<T>* p, const V3<T>* t) {
    const V3<T> e = vadd(vsub(p[1], p[0]), t[0]);
    const V3<T> d1 = {T(0.0) - e.x, T(0.0) - e.y, T(0.0) - e.z};
    const V3<T> d2 = vadd(vsub(p[2], p[1]), t[1]);
    const T c = vdot(d1, d2) * hd_inv(hd_sqrt(vdot(d1, d1)) * hd_sqrt(vdot(d2, d2)));
    return hd_acos(c);
}
template <class T> __device__ T dihedral_value(const V3<T>* p, const V3<T>* t) {
    const V3<T> d1 = vadd(vsub(p[1], p[0]), t[0]);
    const V3<T> d2 = vadd(vsub(p[2], p[1]), t[1]);
    const V3<T> d3 = vadd(vsub(p[3], p[2]), t[2]);
    const V3<T> c12 = vcross(d1, d2), c23 = vcross(d2, d3);
    const T numer = vdot(d2, vcross(c12, c23));
    const T denom = hd_sqrt(vdot(d2, d2)) * vdot(c12, c23);
    return hd_atan2(numer, denom);
}

struct Topology {
    const int* trans;      // [nt, 2]  (atom, dim)
    const int* bonds;      // [nb, 2]
    const int* angles;     // [na, 3]
    const int* diheds;     // [nd, 4]
    const double* tb;      // [nb, 1, 3] shift vectors (may be null = zeros)
    const double* ta;      // [na, 2, 3]
    const double* td;      // [nd, 3, 3]
    int nt, nb, na, nd;
};

// evaluates coordinate `c` (global index) of system positions `x`; returns value, the
// atoms involved (m, atoms[4]) and derivatives wrt their 3m Cartesian components
template <bool H>
__device__ void eval_coord(const Topology& T, int c, const double* __restrict__ x, double* val, int* m_out,
                           int* atoms, double* grad, double* hess /* 12x12 full, row-major, if H */) {
    using D = HD<12, H>;
    if (c < T.nt) {
        const int a = T.trans[2 * c], dim = T.trans[2 * c + 1];
        *val = x[3 * a + dim];
        *m_out = 1; atoms[0] = a;
        for (int i = 0; i < 3; ++i) grad[i] = (i == dim) ? 1.0 : 0.0;
        if (H) for (int i = 0; i < 9; ++i) hess[(i / 3) * 12 + (i % 3)] = 0.0;
        return;
    }
    int m; const int* idx; const double* tv; int kind;
    if (c < T.nt + T.nb) { const int i = c - T.nt; m = 2; idx = T.bonds + 2 * i; tv = T.tb ? T.tb + 3 * i : nullptr; kind = 0; }
    else if (c < T.nt + T.nb + T.na) { const int i = c - T.nt - T.nb; m = 3; idx = T.angles + 3 * i; tv = T.ta ? T.ta + 6 * i : nullptr; kind = 1; }
    else { const int i = c - T.nt - T.nb - T.na; m = 4; idx = T.diheds + 4 * i; tv = T.td ? T.td + 9 * i : nullptr; kind = 2; }
    V3<D> p[4], t[3];
    for (int a = 0; a < 4; ++a) {
        const int at = a < m ? idx[a] : 0;
        if (a < m) atoms[a] = at;
        p[a].x = a < m ? D::var(x[3 * at], 3 * a) : D(0.0);
        p[a].y = a < m ? D::var(x[3 * at + 1], 3 * a + 1) : D(0.0);
        p[a].z = a < m ? D::var(x[3 * at + 2], 3 * a + 2) : D(0.0);
    }
    for (int a = 0; a < 3; ++a) {
        t[a].x = D(tv && a < m - 1 ? tv[3 * a] : 0.0);
        t[a].y = D(tv && a < m - 1 ? tv[3 * a + 1] : 0.0);
        t[a].z = D(tv && a < m - 1 ? tv[3 * a + 2] : 0.0);
    }
    D q = kind == 0 ? bond_value(p, t) : (kind == 1 ? angle_value(p, t) : dihedral_value(p, t));
    *val = q.v;
    *m_out = m;
    for (int i = 0; i < 3 * m; ++i) grad[i] = q.g[i];
    if (H)
        for (int i = 0; i < 3 * m; ++i)
            for (int j = 0; j < 3 * m; ++j) hess[i * 12 + j] = q.h[D::idx(i, j)];
}

// q[b, nint] and (optionally) the dense B-matrix Bmat[b, nint, n] (must be zero-filled).
__global__ void __launch_bounds__(128)
internals_qB_kernel(Topology T, const double* __restrict__ x_, int n, double* __restrict__ q_, double* __restrict__ B_,
                    const int* __restrict__ active, int batch) {
    const int nint = T.nt + T.nb + T.na + T.nd;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)batch * nint) return;
    const int b = (int)(gid / nint), c = (int)(gid % nint);
    if (active && !active[b]) return;
    double val, grad[12];
    int m, atoms[4];
    eval_coord<false>(T, c, x_ + (size_t)b * n, &val, &m, atoms, grad, nullptr);
    q_[(size_t)b * nint + c] = val;
    if (B_) {
        double* row = B_ + ((size_t)b * nint + c) * n;
        for (int a = 0; a < m; ++a)
            for (int d = 0; d < 3; ++d) row[3 * atoms[a] + d] = grad[3 * a + d];
    }
}

// D[b] += sum_c v[b,c] d2q_c/dx2  (dense n x n, must be zero-filled; ldot) and/or
// R[b,c,:] = (d2q_c/dx2) w[b,:]  (nint x n, must be zero-filled; rdot).
__global__ void __launch_bounds__(64)
internals_hess_kernel(Topology T, const double* __restrict__ x_, int n, const double* __restrict__ v_,
                      double* __restrict__ D_, const double* __restrict__ w_, double* __restrict__ R_,
                      const int* __restrict__ active, int batch) {
    const int nint = T.nt + T.nb + T.na + T.nd;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)batch * nint) return;
    const int b = (int)(gid / nint), c = (int)(gid % nint);
    if (active && !active[b]) return;
    if (c < T.nt) return;                       // translations have no curvature
    double val, grad[12], hess[144];
    int m, atoms[4];
    eval_coord<true>(T, c, x_ + (size_t)b * n, &val, &m, atoms, grad, hess);
    if (D_) {
        const double vc = v_[(size_t)b * nint + c];
        if (vc != 0.0) {
            double* Db = D_ + (size_t)b * n * n;
            for (int i = 0; i < 3 * m; ++i)
                for (int j = 0; j < 3 * m; ++j)
                    atomicAdd(&Db[(size_t)(3 * atoms[i / 3] + i % 3) * n + 3 * atoms[j / 3] + j % 3],
                              vc * hess[i * 12 + j]);
        }
    }
    if (R_) {
        const double* w = w_ + (size_t)b * n;
        double* row = R_ + ((size_t)b * nint + c) * n;
        for (int i = 0; i < 3 * m; ++i) {
            double acc = 0.0;
            for (int j = 0; j < 3 * m; ++j) acc = fma(hess[i * 12 + j], w[3 * atoms[j / 3] + j % 3], acc);
            row[3 * atoms[i / 3] + i % 3] = acc;
        }
    }
}

}  // namespace

extern "C" int sb_internals_qB_impl(const int* trans, int nt, const int* bonds, int nb, const int* angles, int na,
                                    const int* diheds, int nd, const double* tb, const double* ta, const double* td,
                                    const double* x, int n, double* q, double* Bmat, const int* active, int batch,
                                    cudaStream_t st) {
    Topology T{trans, bonds, angles, diheds, tb, ta, td, nt, nb, na, nd};
    const long long tot = (long long)batch * (nt + nb + na + nd);
    if (tot == 0) return 0;
    SB_COUNT(1);
    internals_qB_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(T, x, n, q, Bmat, active, batch);
    return SB_LAUNCH_CHECK();
}

extern "C" int sb_internals_hess_impl(const int* trans, int nt, const int* bonds, int nb, const int* angles, int na,
                                      const int* diheds, int nd, const double* tb, const double* ta, const double* td,
                                      const double* x, int n, const double* v, double* D, const double* w, double* R,
                                      const int* active, int batch, cudaStream_t st) {
    Topology T{trans, bonds, angles, diheds, tb, ta, td, nt, nb, na, nd};
    const long long tot = (long long)batch * (nt + nb + na + nd);
    if (tot == 0) return 0;
    SB_COUNT(1);
    internals_hess_kernel<<<(unsigned)((tot + 63) / 64), 64, 0, st>>>(T, x, n, v, D, w, R, active, batch);
    return SB_LAUNCH_CHECK();
}
