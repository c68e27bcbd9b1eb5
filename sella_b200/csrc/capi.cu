// extern "C" surface of libsella_b200.so (declarations: include/sella_b200.h).
#include "common.cuh"
#include "../../include/sella_b200.h"

extern "C" int sb_hv_impl(const double*, const double*, double*, const int*, int, int, int, int,
                          cudaStream_t);
extern "C" int sb_eigh_impl(const double*, double*, double*, double*, double*, double*, int*, const int*, int, int,
                            cudaStream_t);
extern "C" int sb_hv_ld_impl(const double*, const double*, double*, const int*, int, int, int, int, int,
                             cudaStream_t);
extern "C" int sb_mgs_impl(double*, int, const double*, double*, int, int, double, double, int, int*, int*,
                           const int*, int, cudaStream_t);
extern "C" int sb_davidson_init_impl(const double*, const double*, const double*, int, double*, int, int, int*,
                                     int*, int*, int*, int*, const int*, int, cudaStream_t);
extern "C" int sb_davidson_rr_impl(double*, double*, int, const int*, int, double, int, double*, double*, double*,
                                   int*, int*, int, cudaStream_t);
extern "C" int sb_davidson_jd_coeff_impl(const double*, const double*, const double*, double*, int, int,
                                         const int*, int, cudaStream_t);
extern "C" int sb_davidson_expand_impl(const double*, const double*, const double*, double*, double*, int,
                                       const int*, int, int, int, double*, int*, int*, int, cudaStream_t);
extern "C" int sb_davidson_mjd_coeff_impl(const double*, int, const int*, const double*, const double*, const double*,
                                          double*, int, const int*, int*, int, cudaStream_t);
extern "C" int sb_hvp_prepare_impl(const double*, long long, const double*, const double*, double, double*, double*,
                                   int, const int*, int, int, cudaStream_t);
extern "C" int sb_hvp_finish_impl(const double*, long long, const double*, const double*, const double*, double,
                                  double*, double*, double*, int, int*, int*, int, const int*, int, int,
                                  cudaStream_t);
extern "C" int sb_history_ritz_impl(double*, double*, int, const int*, int, int*, const int*, int*, const double*,
                                    int, cudaStream_t);
extern "C" int sb_cons_solve_impl(const double*, const double*, const double*, int, int, double*, double*, int*,
                                  const int*, int, cudaStream_t);
extern "C" int sb_update_prep_impl(const double*, const double*, double*, int, const int*, int, int, int, int,
                                   double*, int*, int*, const int*, int, cudaStream_t);
extern "C" int sb_fill_scaled_identity_impl(double*, double*, double*, const double*, int, int, const int*, int,
                                            cudaStream_t);
extern "C" int sb_abs_scale_impl(const double*, const double*, double*, int, int, const int*, int, cudaStream_t);
extern "C" int sb_update_mid_impl(const double*, const double*, const double*, const double*, double*, double*,
                                  double*, double*, int, const int*, int, int, const int*, int*, double*,
                                  const double*, int*, int, cudaStream_t);
extern "C" int sb_lowrank_factor_impl(const double*, const double*, const double*, int, const int*, int, double*,
                                      double*, int*, const int*, int, cudaStream_t);
extern "C" int sb_secular_update_impl(double*, double*, double*, int, const double*, const int*, int, double*,
                                      double*, int*, const int*, int, cudaStream_t);
extern "C" int sb_update_apply_impl(double*, const double*, const double*, const double*, int, const int*, int,
                                    const int*, int, cudaStream_t);
extern "C" int sb_qn_tr_impl(const double*, const double*, const double*, int, int, double*, double*, double*, int*,
                             const int*, const double*, int, cudaStream_t);
extern "C" int sb_qn_ras_impl(const double*, const double*, const double*, const double*, int, int, double*,
                              double*, double*, int*, const int*, const double*, int, cudaStream_t);
extern "C" int sb_rect_dots_impl(const double*, long long, int, const double*, long long, const double*, double*, int,
                                 const int*, int, cudaStream_t);
extern "C" int sb_rect_comb_impl(const double*, long long, int, const double*, double, const double*, long long,
                                 double, double*, long long, int, const int*, int, cudaStream_t);
extern "C" int sb_scons_measure_impl(const double*, const double*, int, int, double*, double*, int*, int*, int,
                                     cudaStream_t);
extern "C" int sb_combine_step_impl(const double*, const double*, const double*, const double*, const int*, double*,
                                    double*, int, const int*, int, cudaStream_t);
extern "C" int sb_converged_cons_impl(const double*, const double*, int, int, double, double, double*, double*, int*,
                                      int, cudaStream_t);
extern "C" int sb_axpy_impl(const double*, const double*, double*, int, const int*, int, cudaStream_t);
extern "C" int sb_rfo_tr_impl(const double*, const double*, const double*, int, int, int, double*, double*, double*,
                              int*, const int*, const double*, int, cudaStream_t);
extern "C" int sb_pack_coef_impl(const double*, const double*, double*, int, const int*, int, cudaStream_t);
extern "C" int sb_unpack2_impl(const double*, double*, double*, int, const int*, int, cudaStream_t);
extern "C" int sb_kick_finish_impl(double*, double*, double*, const double*, const double*, const double*,
                                   const double*, const double*, const double*, double*, double*, double*, int*,
                                   const double*, const int*, int, const int*, int, cudaStream_t);
extern "C" int sb_ev_decide_impl(const double*, int, int, int*, int*, const double*, const int*, const int*, int,
                                 cudaStream_t);
extern "C" int sb_converged_impl(const double*, int, double, double*, int*, int, cudaStream_t);

extern "C" int sb_hv_rect_impl(const double*, long long, int, const double*, double*, const int*, int, int, int, int,
                               int, cudaStream_t);
extern "C" int sb_secular_update_c_impl(double*, double*, double*, int, const double*, const int*, int, double*,
                                        double*, int*, const int*, const int*, int, long long, long long, int,
                                        cudaStream_t, int, int*);
extern "C" int sb_qn_ras_c_impl(const double*, const double*, const double*, const double*, int, int, double*,
                                double*, double*, int*, const int*, const double*, int, const int*, const double*,
                                const double*, long long, int, cudaStream_t);
extern "C" int sb_rfo_ras_c_impl(const double*, const double*, const double*, const double*, int, int, int, double*,
                                 double*, double*, int*, const int*, const double*, int, const int*, const double*,
                                 const double*, long long, int, cudaStream_t);
extern "C" int sb_davidson_init_c_impl(const double*, const double*, const double*, int, double*, int, int, int*,
                                       int*, int*, int*, int*, const int*, const int*, const double*, const double*,
                                       long long, long long, int, cudaStream_t);

std::atomic<long long> sb_launch_counter{0};

namespace {

__global__ void diff_kernel(const double* __restrict__ x, const double* __restrict__ x0,
                            double* __restrict__ d, const int* __restrict__ active, int n) {
    const int b = blockIdx.y;
    if (active && !active[b]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[(size_t)b * n + i] = x[(size_t)b * n + i] - x0[(size_t)b * n + i];
}

// f[b] = scale * x[b].y[b]   (deterministic block reduction)
__global__ void dot_kernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ f,
                           double scale, const int* __restrict__ active, int n) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    __shared__ double scratch[SB_SCRATCH_DOUBLES];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        acc = fma(x[(size_t)b * n + i], y[(size_t)b * n + i], acc);
    acc = sb_block_sum(acc, scratch);
    if (threadIdx.x == 0) f[b] = scale * acc;
}

// M[b] <- scale * M[b] + diag * I
__global__ void add_scaled_identity_kernel(double* __restrict__ M, double scale, double diag, int n) {
    const int b = blockIdx.y;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx / n), j = (int)(idx % n);
    double* p = M + (size_t)b * n * n + idx;
    *p = scale * *p + (i == j ? diag : 0.0);
}

}  // namespace

extern "C" {

int sb_version(void) { return 2; }

extern "C" int sb_secular_apply_bench_impl(const double*, const double*, double*, int*, int, int, long long, int, int,
                                           float*, cudaStream_t);
int sb_secular_apply_bench(const double* Vt, const double* qwork, double* work, int32_t* aux, int r, int n,
                           long long vstride, int batch, int reps, float* ms_host, void* stream) {
    if (r < 1 || r > n || (long long)r * n > vstride || reps < 1 || !ms_host) return -1;
    return sb_secular_apply_bench_impl(Vt, qwork, work, aux, r, n, vstride, batch, reps, ms_host,
                                       (cudaStream_t)stream);
}
extern "C" int sb_fp64_peak_impl(int, int, int, double*, double*);
int sb_fp64_peak(int kind, int iters, int ctas_per_sm, double* scratch, double* tflops_host) {
    if (kind < 0 || kind > 1 || iters < 1 || ctas_per_sm < 1 || !scratch || !tflops_host) return -1;
    return sb_fp64_peak_impl(kind, iters, ctas_per_sm, scratch, tflops_host);
}

int sb_add_scaled_identity(double* M, double scale, double diag, int n, int batch, void* stream) {
    if (n < 1 || batch < 1) return -1;
    dim3 grid((unsigned)(((size_t)n * n + 255) / 256), batch);
    SB_COUNT(1);
    add_scaled_identity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, scale, diag, n);
    return SB_LAUNCH_CHECK();
}

long long sb_launch_count(void) { return sb_launch_counter.load(std::memory_order_relaxed); }

int sb_device_sms(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sms;
}

int sb_hv(const double* A, const double* X, double* Y, const int32_t* active, int batch, int n, int nvec,
          int transposed, void* stream) {
    if (batch <= 0 || n <= 0 || nvec <= 0) return -1;
    return sb_hv_impl(A, X, Y, active, batch, n, nvec, transposed, (cudaStream_t)stream);
}

int sb_hv_ld(const double* A, const double* X, double* Y, const int32_t* active, int batch, int n, int nvec,
             int ldv, int transposed, void* stream) {
    if (batch <= 0 || n <= 0 || nvec <= 0 || ldv < nvec) return -1;
    return sb_hv_ld_impl(A, X, Y, active, batch, n, nvec, ldv, transposed, (cudaStream_t)stream);
}

int sb_hv_rect(const double* A, long long astride, int mrows, const double* X, double* Y, const int32_t* active,
               int batch, int n, int nvec, int ldv, int transposed, void* stream) {
    if (batch <= 0 || n <= 0 || nvec <= 0 || ldv < nvec || mrows < 0 || mrows > n) return -1;
    return sb_hv_rect_impl(A, astride, mrows, X, Y, active, batch, n, nvec, ldv, transposed, (cudaStream_t)stream);
}

int sb_quadratic_pes(const double* A, const double* xstar, const double* x, double* f, double* g,
                     double* dwork, const int32_t* active, int batch, int n, void* stream) {
    if (batch <= 0 || n <= 0) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((n + 255) / 256, batch);
    SB_COUNT(1);
    diff_kernel<<<grid, 256, 0, st>>>(x, xstar, dwork, active, n);
    int rc = sb_hv_impl(A, dwork, g, active, batch, n, 1, 0, st);
    if (rc) return rc;
    SB_COUNT(1);
    dot_kernel<<<batch, 256, 0, st>>>(dwork, g, f, 0.5, active, n);
    return SB_LAUNCH_CHECK();
}
extern "C" int sb_emt_pes_impl(const double*, int, const double*, long long, const int*, const double*, double*,
                               double*, const int*, int, cudaStream_t);
int sb_emt_pes(const double* x, int natoms, const double* cell, long long cellstride, const int32_t* nimg3,
               const double* par6, double* f, double* g, const int32_t* active, int batch, void* stream) {
    if (natoms < 1 || !par6) return -1;
    return sb_emt_pes_impl(x, natoms, cell, cellstride, nimg3, par6, f, g, active, batch, (cudaStream_t)stream);
}

int sb_eigh(const double* A, double* evals, double* Vt, double* work, double* small_work, int32_t* status,
            const int32_t* active, int batch, int n, void* stream) {
    if (batch <= 0 || n <= 0) return -1;
    return sb_eigh_impl(A, evals, Vt, work, small_work, nullptr, status, active, batch, n, (cudaStream_t)stream);
}
int sb_eigh_blocked(const double* A, double* evals, double* Vt, double* work, double* small_work, double* work2,
                    int32_t* status, const int32_t* active, int batch, int n, void* stream) {
    if (batch <= 0 || n <= 0 || !work2) return -1;
    return sb_eigh_impl(A, evals, Vt, work, small_work, work2, status, active, batch, n, (cudaStream_t)stream);
}

extern "C" int sb_internals_qB_impl(const int*, int, const int*, int, const int*, int, const int*, int, const double*,
                                    const double*, const double*, const double*, int, double*, double*, const int*,
                                    int, cudaStream_t);
extern "C" int sb_internals_hess_impl(const int*, int, const int*, int, const int*, int, const int*, int,
                                      const double*, const double*, const double*, const double*, int, const double*,
                                      double*, const double*, double*, const int*, int, cudaStream_t);
#define ST ((cudaStream_t)stream)
int sb_internals_qB(const int32_t* trans, int nt, const int32_t* bonds, int nb, const int32_t* angles, int na,
                    const int32_t* diheds, int nd, const double* tb, const double* ta, const double* td,
                    const double* x, int n, double* q, double* Bmat, const int32_t* active, int batch,
                    void* stream) {
    return sb_internals_qB_impl(trans, nt, bonds, nb, angles, na, diheds, nd, tb, ta, td, x, n, q, Bmat, active,
                                batch, ST);
}
int sb_internals_hess(const int32_t* trans, int nt, const int32_t* bonds, int nb, const int32_t* angles, int na,
                      const int32_t* diheds, int nd, const double* tb, const double* ta, const double* td,
                      const double* x, int n, const double* v, double* D, const double* w, double* R,
                      const int32_t* active, int batch, void* stream) {
    return sb_internals_hess_impl(trans, nt, bonds, nb, angles, na, diheds, nd, tb, ta, td, x, n, v, D, w, R, active,
                                  batch, ST);
}
int sb_mgs(double* X, int nx, const double* Y, double* Ywork, int ny, int n, double eps1, double eps2,
           int maxiter, int32_t* nkept, int32_t* status, const int32_t* active, int batch, void* stream) {
    if (batch <= 0 || n <= 0 || nx <= 0) return -1;
    return sb_mgs_impl(X, nx, Y, Ywork, ny, n, eps1, eps2, maxiter, nkept, status, active, batch, ST);
}
int sb_davidson_init(const double* v0, const double* pl, const double* Pvt, int mode, double* V, int kcap, int n,
                     int32_t* ksz, int32_t* ninit, int32_t* nhist, int32_t* dav_state, int32_t* status,
                     const int32_t* part, int batch, void* stream) {
    if (kcap > 32 || kcap < 2) return -1;
    return sb_davidson_init_impl(v0, pl, Pvt, mode, V, kcap, n, ksz, ninit, nhist, dav_state, status, part, batch,
                                 ST);
}
int sb_davidson_rr(double* V, double* AV, int kcap, const int32_t* ksz, int n, double gamma, int maxiter_eff,
                   double* lams, double* rv, double* theta, int32_t* dav_state, int32_t* status, int batch,
                   void* stream) {
    if (kcap > 32) return -1;
    return sb_davidson_rr_impl(V, AV, kcap, ksz, n, gamma, maxiter_eff, lams, rv, theta, dav_state, status, batch,
                               ST);
}
int sb_davidson_jd_coeff(const double* rvhat, const double* pl, const double* theta, double* that, int n,
                         int method, const int32_t* dav_state, int batch, void* stream) {
    return sb_davidson_jd_coeff_impl(rvhat, pl, theta, that, n, method, dav_state, batch, ST);
}
int sb_davidson_expand(const double* t, const double* rv, const double* theta, double* V, double* Ywork, int kcap,
                       const int32_t* ksz, int n, int p_identity, int lanczos, double* vnew, int32_t* dav_state,
                       int32_t* status, int batch, void* stream) {
    return sb_davidson_expand_impl(t, rv, theta, V, Ywork, kcap, ksz, n, p_identity, lanczos, vnew, dav_state,
                                   status, batch, ST);
}
int sb_hvp_prepare(const double* vfull, long long vstride, const double* x0, const double* g0, double eta,
                   double* xdisp, double* signnorm, int n, const int32_t* mask, int maskval, int batch,
                   void* stream) {
    return sb_hvp_prepare_impl(vfull, vstride, x0, g0, eta, xdisp, signnorm, n, mask, maskval, batch, ST);
}
int sb_hvp_finish(const double* vfull, long long vstride, const double* gplus, const double* g0,
                  const double* signnorm, double eta, double* AV, double* Vs, double* AVs, int kcap, int32_t* ksz,
                  int32_t* nhist, int n, const int32_t* mask, int maskval, int batch, void* stream) {
    return sb_hvp_finish_impl(vfull, vstride, gplus, g0, signnorm, eta, AV, Vs, AVs, kcap, ksz, nhist, n, mask,
                              maskval, batch, ST);
}
int sb_davidson_mjd_coeff(const double* Vhat, int kcap, const int32_t* ksz, const double* rvhat, const double* pl,
                          const double* theta, double* that, int n, const int32_t* dav_state, int32_t* status,
                          int batch, void* stream) {
    if (kcap > 32 || kcap < 1) return -1;
    return sb_davidson_mjd_coeff_impl(Vhat, kcap, ksz, rvhat, pl, theta, that, n, dav_state, status, batch, ST);
}
int sb_history_ritz(double* Vs, double* AVs, int kcap, const int32_t* nhist, int n, int32_t* nvec_out,
                    const int32_t* dav_state, int32_t* status, const double* HcVs, int batch, void* stream) {
    if (kcap > 32) return -1;
    return sb_history_ritz_impl(Vs, AVs, kcap, nhist, n, nvec_out, dav_state, status, HcVs, batch, ST);
}
int sb_cons_solve(const double* G, const double* Uc, const double* u, int nc, int n, double* Mr, double* L,
                  int32_t* status, const int32_t* active, int batch, void* stream) {
    if (nc < 1 || nc > 32 || n < 1) return -1;
    return sb_cons_solve_impl(G, Uc, u, nc, n, Mr, L, status, active, batch, ST);
}
int sb_update_prep(const double* S, const double* Y, double* Ytil, int kcap, const int32_t* kvec, int n, int ncart,
                   int first, int symm, double* lam0, int32_t* skip, int32_t* status, const int32_t* active,
                   int batch, void* stream) {
    if (kcap > 32 || kcap < 1) return -1;
    return sb_update_prep_impl(S, Y, Ytil, kcap, kvec, n, ncart, first, symm, lam0, skip, status, active, batch,
                               ST);
}
int sb_fill_scaled_identity(double* B, double* evals, double* Vt, const double* lam0, int n, int ncart,
                            const int32_t* skip, int batch, void* stream) {
    return sb_fill_scaled_identity_impl(B, evals, Vt, lam0, n, ncart, skip, batch, ST);
}
int sb_abs_scale(const double* VtS, const double* evals, double* out, int kcap, int n, const int32_t* skip,
                 int batch, void* stream) {
    return sb_abs_scale_impl(VtS, evals, out, kcap, n, skip, batch, ST);
}
int sb_update_mid(const double* S, const double* Ytil, const double* BS, const double* absBS, double* U, double* J,
                  double* W, double* Xwork, int kcap, const int32_t* kvec, int n, int method, const int32_t* skip,
                  int32_t* status, double* Cout, const double* evals, int32_t* kout, int batch, void* stream) {
    if (kcap > 32) return -1;
    return sb_update_mid_impl(S, Ytil, BS, absBS, U, J, W, Xwork, kcap, kvec, n, method, skip, status, Cout, evals,
                              kout, batch, ST);
}
int sb_lowrank_factor(const double* U, const double* J, const double* Cmat, int kcap, const int32_t* kvec, int n,
                      double* P, double* sig, int32_t* nterm, const int32_t* skip, int batch, void* stream) {
    if (kcap < 1 || kcap > 16) return -1;
    return sb_lowrank_factor_impl(U, J, Cmat, kcap, kvec, n, P, sig, nterm, skip, batch, ST);
}
extern "C" int sb_secular_profile_impl(unsigned long long*, int);
extern "C" int sb_gemm_impl(int, int, int, int, int, double, const double*, int, long long, const double*, int, long long,
                            double, double*, int, long long, const int*, int, cudaStream_t);
extern "C" int sb_qr_impl(double*, int, int, double*, double*, double*, const int*, int, cudaStream_t);
extern "C" int sb_trtri_impl(const double*, double*, double*, int, int*, const int*, int, cudaStream_t);
int sb_gemm(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda, long long strideA,
            const double* B, int ldb, long long strideB, double beta, double* C, int ldc, long long strideC,
            const int32_t* active, int batch, void* stream) {
    if (M < 1 || N < 1 || K < 1 || batch < 1) return -1;
    return sb_gemm_impl(transA, transB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, active,
                        batch, (cudaStream_t)stream);
}
int sb_qr(double* A, int m, int n, double* Q, double* R, double* work, const int32_t* active, int batch,
          void* stream) {
    if (m < n || n < 1 || batch < 1 || !work) return -1;
    return sb_qr_impl(A, m, n, Q, R, work, active, batch, (cudaStream_t)stream);
}
extern "C" int sb_potrf_impl(double*, int, int*, const int*, int, cudaStream_t);
int sb_potrf(double* A, int n, int32_t* status, const int32_t* active, int batch, void* stream) {
    if (n < 1 || batch < 1) return -1;
    return sb_potrf_impl(A, n, status, active, batch, (cudaStream_t)stream);
}
int sb_trtri(const double* R, double* Rinv, double* work, int n, int32_t* status, const int32_t* active, int batch,
             void* stream) {
    if (n < 1 || batch < 1 || !work) return -1;
    return sb_trtri_impl(R, Rinv, work, n, status, active, batch, (cudaStream_t)stream);
}
extern "C" int sb_rotation_impl(const double*, int, const double*, long long, double*, double*, long long, double*,
                                long long, const double*, long long, double*, double*, const int*, int, cudaStream_t);
int sb_rotation(const double* x, int natoms, const double* refpos, long long refstride, double* qprev, double* vals,
                long long valstride, double* J, long long jstride, const double* L, long long lstride, double* D,
                double* work, const int32_t* active, int batch, void* stream) {
    if (natoms < 2 || batch < 1 || !work) return -1;
    return sb_rotation_impl(x, natoms, refpos, refstride, qprev, vals, valstride, J, jstride, L, lstride, D, work,
                            active, batch, (cudaStream_t)stream);
}
extern "C" int sb_secular_timing_impl(float*, int);
int sb_secular_timing(float* out3, int enable) { return sb_secular_timing_impl(out3, enable); }
extern "C" int sb_rfo_profile_impl(unsigned long long*, int);
int sb_rfo_profile(unsigned long long* out8, int reset) { return sb_rfo_profile_impl(out8, reset); }
int sb_secular_profile(unsigned long long* out16, int reset) { return sb_secular_profile_impl(out16, reset); }
int sb_secular_update(double* evals, double* Vt, double* Z, int zcap, const double* sig, const int32_t* nterm,
                      int n, double* work, double* qwork, int32_t* status, const int32_t* skip, int batch,
                      void* stream) {
    return sb_secular_update_impl(evals, Vt, Z, zcap, sig, nterm, n, work, qwork, status, skip, batch, ST);
}
int sb_secular_update_c(double* evals, double* Vt, double* Z, int zcap, const double* sig, const int32_t* nterm,
                        int n, double* work, double* qwork, int32_t* status, const int32_t* skip,
                        const int32_t* mrows, int mcap, long long estride, long long vstride, int nterm_max,
                        int32_t* aux, int batch, void* stream) {
    if (!mrows || mcap < 0) return -1;
    return sb_secular_update_c_impl(evals, Vt, Z, zcap, sig, nterm, n, work, qwork, status, skip, mrows, mcap,
                                    estride, vstride, batch, ST, nterm_max, aux);
}
int sb_qn_ras_c(const double* Vg, const double* evals, const double* Vt, const double* delta, int order, int n,
                double* s, double* smag, double* alpha, int32_t* status, const int32_t* active, const double* sadd,
                int npole, const int32_t* rowmap, const double* gperp, const double* gam, long long vstride,
                int batch, void* stream) {
    if (n % 3 || npole < 1) return -1;
    return sb_qn_ras_c_impl(Vg, evals, Vt, delta, order, n, s, smag, alpha, status, active, sadd, npole, rowmap, gperp,
                            gam, vstride, batch, ST);
}
int sb_rfo_ras_c(const double* Vg, const double* evals, const double* Vt, const double* delta, int order, int n,
                 int mode, double* s, double* smag, double* alpha, int32_t* status, const int32_t* active,
                 const double* sadd, int npole, const int32_t* rowmap, const double* gperp, const double* gam,
                 long long vstride, int batch, void* stream) {
    if (n % 3 || mode < 0 || mode > 1 || npole < 1) return -1;
    return sb_rfo_ras_c_impl(Vg, evals, Vt, delta, order, n, mode, s, smag, alpha, status, active, sadd, npole, rowmap,
                             gperp, gam, vstride, batch, ST);
}
extern "C" int sb_qn_mis_impl(const double*, const double*, const double*, const double*, int, int, double*, double*,
                              double*, int*, const int*, const double*, int, long long, const double*, int,
                              cudaStream_t);
extern "C" int sb_rfo_mis_impl(const double*, const double*, const double*, const double*, int, int, int, double*,
                               double*, double*, int*, const int*, const double*, int, long long, const double*, int,
                               cudaStream_t);
int sb_qn_mis(const double* Vg, const double* evals, const double* Wt, const double* delta, int order, int n,
              double* s, double* smag, double* alpha, int32_t* status, const int32_t* active, const double* sadd,
              int npole, long long vstride, const double* w, int batch, void* stream) {
    if (npole < 1 || n < 1 || !w) return -1;
    return sb_qn_mis_impl(Vg, evals, Wt, delta, order, n, s, smag, alpha, status, active, sadd, npole, vstride, w, batch,
                          ST);
}
int sb_rfo_mis(const double* Vg, const double* evals, const double* Wt, const double* delta, int order, int n,
               int mode, double* s, double* smag, double* alpha, int32_t* status, const int32_t* active,
               const double* sadd, int npole, long long vstride, const double* w, int batch, void* stream) {
    if (npole < 1 || n < 1 || mode < 0 || mode > 1 || !w) return -1;
    return sb_rfo_mis_impl(Vg, evals, Wt, delta, order, n, mode, s, smag, alpha, status, active, sadd, npole, vstride, w,
                           batch, ST);
}
int sb_davidson_init_c(const double* v0, const double* pl, const double* Pvt, int mode, double* V, int kcap, int n,
                       int32_t* ksz, int32_t* ninit, int32_t* nhist, int32_t* dav_state, int32_t* status,
                       const int32_t* part, const int32_t* mrows, const double* lam0, const double* gperp,
                       long long estride, long long vstride, int batch, void* stream) {
    if (kcap > 32 || kcap < 2 || !mrows || !lam0 || !gperp) return -1;
    return sb_davidson_init_c_impl(v0, pl, Pvt, mode, V, kcap, n, ksz, ninit, nhist, dav_state, status, part, mrows,
                                   lam0, gperp, estride, vstride, batch, ST);
}
int sb_update_apply(double* B, const double* U, const double* J, const double* W, int kcap, const int32_t* kvec,
                    int n, const int32_t* skip, int batch, void* stream) {
    return sb_update_apply_impl(B, U, J, W, kcap, kvec, n, skip, batch, ST);
}
int sb_qn_tr(const double* Vg, const double* evals, const double* delta, int order, int n, double* coef,
             double* smag, double* alpha, int32_t* status, const int32_t* active, const double* extra2, int batch,
             void* stream) {
    return sb_qn_tr_impl(Vg, evals, delta, order, n, coef, smag, alpha, status, active, extra2, batch, ST);
}
int sb_rfo_tr(const double* Vg, const double* evals, const double* delta, int order, int n, int mode, double* coef,
              double* smag, double* alpha, int32_t* status, const int32_t* active, const double* extra2, int batch,
              void* stream) {
    if (mode < 0 || mode > 1 || order < 0) return -1;
    return sb_rfo_tr_impl(Vg, evals, delta, order, n, mode, coef, smag, alpha, status, active, extra2, batch, ST);
}
extern "C" int sb_rfo_ras_impl(const double*, const double*, const double*, const double*, int, int, int, double*,
                               double*, double*, int*, const int*, const double*, int, cudaStream_t);
int sb_rfo_ras(const double* Vg, const double* evals, const double* Vt, const double* delta, int order, int n,
               int mode, double* s, double* smag, double* alpha, int32_t* status, const int32_t* active,
               const double* sadd, int batch, void* stream) {
    if (n % 3 || mode < 0 || mode > 1) return -1;
    return sb_rfo_ras_impl(Vg, evals, Vt, delta, order, n, mode, s, smag, alpha, status, active, sadd, batch, ST);
}
int sb_qn_ras(const double* Vg, const double* evals, const double* Vt, const double* delta, int order, int n,
              double* s, double* smag, double* alpha, int32_t* status, const int32_t* active, const double* sadd,
              int batch, void* stream) {
    if (n % 3) return -1;
    return sb_qn_ras_impl(Vg, evals, Vt, delta, order, n, s, smag, alpha, status, active, sadd, batch, ST);
}
int sb_rect_dots(const double* R, long long rstride, int nr, const double* x, long long xstride, const double* c,
                 double* out, int n, const int32_t* active, int batch, void* stream) {
    if (nr <= 0) return 0;
    return sb_rect_dots_impl(R, rstride, nr, x, xstride, c, out, n, active, batch, ST);
}
int sb_rect_comb(const double* R, long long rstride, int nr, const double* coef, double scale, const double* base,
                 long long bstride, double beta, double* out, long long ostride, int n, const int32_t* active,
                 int batch, void* stream) {
    return sb_rect_comb_impl(R, rstride, nr, coef, scale, base, bstride, beta, out, ostride, n, active, batch, ST);
}
int sb_scons_measure(const double* scons, const double* delta, int kind, int n, double* scons2, double* consval,
                     int32_t* naive, int32_t* regular, int batch, void* stream) {
    return sb_scons_measure_impl(scons, delta, kind, n, scons2, consval, naive, regular, batch, ST);
}
int sb_combine_step(const double* slift, const double* scons, const double* consval, const double* delta,
                    const int32_t* naive, double* stot, double* smag, int n, const int32_t* active, int batch,
                    void* stream) {
    return sb_combine_step_impl(slift, scons, consval, delta, naive, stot, smag, n, active, batch, ST);
}
int sb_converged_cons(const double* pg, const double* res, int nr, int n, double fmax_tol, double cmax_tol,
                      double* fmax_out, double* cmax_out, int32_t* conv, int batch, void* stream) {
    return sb_converged_cons_impl(pg, res, nr, n, fmax_tol, cmax_tol, fmax_out, cmax_out, conv, batch, ST);
}
int sb_pack_coef(const double* coef, const double* evals, double* out2, int n, const int32_t* active, int batch,
                 void* stream) {
    return sb_pack_coef_impl(coef, evals, out2, n, active, batch, ST);
}
int sb_unpack2(const double* in2, double* s, double* absBs, int n, const int32_t* active, int batch,
               void* stream) {
    return sb_unpack2_impl(in2, s, absBs, n, active, batch, ST);
}
int sb_axpy(const double* x, const double* s, double* out, int n, const int32_t* active, int batch, void* stream) {
    return sb_axpy_impl(x, s, out, n, active, batch, ST);
}
int sb_kick_finish(double* x, double* f, double* g, const double* xnew, const double* fnew, const double* gnew,
                   const double* s, const double* Bs, const double* smag, double* dg, double* delta, double* rho,
                   int32_t* nsteps, const double* dpar, const int32_t* ipar, int n, const int32_t* active,
                   int batch, void* stream) {
    return sb_kick_finish_impl(x, f, g, xnew, fnew, gnew, s, Bs, smag, dg, delta, rho, nsteps, dpar, ipar, n,
                               active, batch, ST);
}
int sb_ev_decide(const double* evals, int n, int has_evals, int32_t* since_diag, int32_t* ev, const double* dpar,
                 const int32_t* ipar, const int32_t* active, int batch, void* stream) {
    return sb_ev_decide_impl(evals, n, has_evals, since_diag, ev, dpar, ipar, active, batch, ST);
}
int sb_converged(const double* g, int n, double fmax_tol, double* fmax_out, int32_t* conv, int batch,
                 void* stream) {
    return sb_converged_impl(g, n, fmax_tol, fmax_out, conv, batch, ST);
}
#undef ST

}  // extern "C"
