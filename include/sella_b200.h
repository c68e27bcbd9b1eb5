/* sella_b200 -- C ABI of the B200-native Sella saddle-search inner loop.
 *
 * Every entry point is a batched, device-resident replacement for one numerical
 * operator of the reference (zadorlab/sella); the reference interface each one
 * replaces is cited as file:line relative to the reference tree.  The binding a
 * maintainer of the reference would add (ctypes) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - all pointers are DEVICE pointers to contiguous fp64 (double) / int32 data;
 *   - matrices are row-major n x n ("numpy C order", as the reference stores B),
 *     batch-leading: M[b][i][j] at M + (b*n + i)*n + j;
 *   - vector blocks are vector-major: X[b][v][:] at X + (b*nvec + v)*n;
 *   - `active` (int32[batch], may be NULL) masks systems: 0 = leave untouched;
 *   - `status` (int32[batch]) receives OR-ed SB_ST_* bits (per-system error codes,
 *     the batched analogue of the reference's exceptions / negative returns);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - return value: 0 on success, a cudaError_t or a negative argument error.
 * No call synchronises the device or touches host memory.
 */
#ifndef SELLA_B200_H
#define SELLA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SB_ST_MGS_MAXITER 1     /* sella/utilities/math.pyx:132-133 (return -2) */
#define SB_ST_TR_NOCONV 2       /* sella/optimize/restricted_step.py:116-117    */
#define SB_ST_EIGH_NOCONV 4
#define SB_ST_DAVIDSON_CAP 8
#define SB_ST_SINGULAR 16
#define SB_ST_DAVIDSON_STALL 32 /* sella/eigensolvers.py:99-109                 */
#define SB_ST_CAPACITY 64       /* a vector block is too small for the operation */

/* library / device introspection */
int sb_version(void);
int sb_device_sms(void);
/* kernels launched by this library since load (bench.py: gpu_launches) */
long long sb_launch_count(void);

/* Y[b,v,:] = A[b] @ X[b,v,:]  (transposed=0)   or   A[b].T @ X[b,v,:]  (1).
 * Replaces A.dot(V) in rayleigh_ritz (sella/eigensolvers.py:52,112), the
 * gradient difference of NumericalHessian._matvec on a quadratic surface
 * (sella/linalg.py:82-87), B @ S (sella/hessian_update.py:119) and V.T @ g /
 * V @ c of QuasiNewton (sella/optimize/stepper.py:86,93-95).                     */
int sb_hv(const double* A, const double* X, double* Y, const int32_t* active,
          int batch, int n, int nvec, int transposed, void* stream);
/* same, for nvec vectors stored in blocks of ldv >= nvec slots per system
 * (X and Y both [b, ldv, n]; only the first nvec slots are touched).             */
int sb_hv_ld(const double* A, const double* X, double* Y, const int32_t* active,
             int batch, int n, int nvec, int ldv, int transposed, void* stream);

/* Quadratic surface evaluator (benchmark PES):
 * g[b] = A[b] (x[b]-xstar[b]),  f[b] = 1/2 (x-xstar).g ; dwork: batch*n doubles. */
int sb_quadratic_pes(const double* A, const double* xstar, const double* x,
                     double* f, double* g, double* dwork, const int32_t* active,
                     int batch, int n, void* stream);

/* EMT-form copper surface (energy f[b], gradient g[b,3N]) for a batch of configurations: the
 * on-device stand-in for the ASE calculator call of sella/peswrapper.py:413-418 in the EMT
 * configurations (README.md:14,29).  Functional form and constants: oracle/emt.py ("EMT-form";
 * ASE itself is not part of the reference tree).  par6 (HOST): E0, s0, V0, eta2, kappa, lambda
 * in eV / Angstrom; cell (device, may be NULL): lattice vectors as rows, shared by the batch
 * (cellstride 0) or per configuration (cellstride 9); nimg3 (HOST): periodic images to scan
 * along each lattice vector (0 = not periodic).                                            */
int sb_emt_pes(const double* x, int natoms, const double* cell, long long cellstride,
               const int32_t* nimg3, const double* par6, double* f, double* g,
               const int32_t* active, int batch, void* stream);

/* Symmetric eigendecomposition, replaces scipy.linalg.eigh at
 * sella/linalg.py:174-195, sella/optimize/stepper.py:79-83,
 * sella/eigensolvers.py:11, sella/_gpu.py:70-97 (gpu_eigh / gpu_eigh_t).
 * evals[b,:] ascending; Vt[b,i,:] = eigenvector i (row = eigenvector).
 * work: batch*n*n doubles, small: 3*batch*n doubles.
 * Vt == NULL: eigenvalues only (scipy eigvalsh; the re-diagonalisation test optimize.py:369-371
 * looks at eigenvalues alone) -- tridiagonalisation + QL on (d, e), no eigenvector work. */
int sb_eigh(const double* A, double* evals, double* Vt, double* work, double* small_work,
            int32_t* status, const int32_t* active, int batch, int n, void* stream);
/* same, with Q^T accumulated by blocked compact-WY GEMMs (faster for batches of large matrices);
 * work2: batch * (n*n + 32*n + 256*ceil(n/16)) doubles.                                        */
int sb_eigh_blocked(const double* A, double* evals, double* Vt, double* work, double* small_work,
                    double* work2, int32_t* status, const int32_t* active, int batch, int n, void* stream);

/* ---- orthogonalisation ------------------------------------------------------
 * modified_gram_schmidt(Xin, Yin, eps1, eps2, maxiter): sella/utilities/math.pyx:143-159
 * (cdef mgs :74-140).  X[b,nx,n] is orthonormalised in place against (an
 * orthonormalised copy of) Y[b,ny,n] and itself; dropped columns are compacted away,
 * leftovers zeroed.  nkept[b] = columns kept, or -2 (iteration limit, also sets
 * SB_ST_MGS_MAXITER).  Ywork: b*ny*n doubles (ny may be 0, Y/Ywork NULL).           */
int sb_mgs(double* X, int nx, const double* Y, double* Ywork, int ny, int n,
           double eps1, double eps2, int maxiter, int32_t* nkept, int32_t* status,
           const int32_t* active, int batch, void* stream);

/* ---- Davidson / Rayleigh-Ritz (sella/eigensolvers.py:31-153) --------------------
 * Device-resident state of one batched diagonalisation:
 *   V, AV   [b,kcap,n]  current basis and its image (free space)
 *   Vs, AVs [b,kcap,n]  operator history of NumericalHessian (linalg.py:89-90)
 *   ksz, ninit, nhist, dav_state  int32[b]
 * dav_state: 0 = expanding, 1 = finished, 2 = not participating.                   */
/* start vectors: mode 0 -> v0 (eigensolvers.py:43-44); mode 1 -> eigenvectors of the
 * preconditioner with negative eigenvalue, at least one (eigensolvers.py:46-50);
 * pl/Pvt: spectrum (ascending) and eigenvectors (rows) of the preconditioner.       */
int sb_davidson_init(const double* v0, const double* pl, const double* Pvt, int mode,
                     double* V, int kcap, int n, int32_t* ksz, int32_t* ninit, int32_t* nhist,
                     int32_t* dav_state, int32_t* status, const int32_t* part, int batch, void* stream);
/* one Rayleigh-Ritz step (eigensolvers.py:56-89): lams[b,kcap]; for systems that go
 * on, rv[b,0,:] = residual, rv[b,1,:] = Ritz vector of the chosen pair, theta[b].    */
int sb_davidson_rr(double* V, double* AV, int kcap, const int32_t* ksz, int n, double gamma,
                   int maxiter_eff, double* lams, double* rv, double* theta,
                   int32_t* dav_state, int32_t* status, int batch, void* stream);
/* correction equation in the eigenbasis of the preconditioner (eigensolvers.py:115-139):
 * rvhat = Pvt @ [r, v]; method 0 = jd0/jd0_alt, 1 = gd.  that[b,n]: t = Pvt.T @ that.  */
int sb_davidson_jd_coeff(const double* rvhat, const double* pl, const double* theta, double* that,
                         int n, int method, const int32_t* dav_state, int batch, void* stream);
/* 'mjd0' / 'mjd0_alt' (eigensolvers.py:140-151): correction orthogonal to all ksz[b] Ritz
 * vectors; Vhat[b,j,:] = Pvt @ V[b,j,:] (sb_hv_ld), rvhat[b,0,:] = Pvt @ r.               */
int sb_davidson_mjd_coeff(const double* Vhat, int kcap, const int32_t* ksz, const double* rvhat,
                          const double* pl, const double* theta, double* that, int n,
                          const int32_t* dav_state, int32_t* status, int batch, void* stream);
/* normalise / Lanczos fallback / MGS against V / append (eigensolvers.py:90-111);
 * vnew[b,n] = the appended direction.  p_identity: closed form for P = I.           */
int sb_davidson_expand(const double* t, const double* rv, const double* theta, double* V,
                       double* Ywork, int kcap, const int32_t* ksz, int n, int p_identity,
                       int lanczos, double* vnew, int32_t* dav_state, int32_t* status,
                       int batch, void* stream);
/* NumericalHessian._matvec, sella/linalg.py:39-95: displaced geometry with the
 * canonical sign (prepare), then Av and history bookkeeping (finish).  `mask`/`maskval`:
 * only systems with mask[b] == maskval take part (mask may be NULL).                 */
int sb_hvp_prepare(const double* vfull, long long vstride, const double* x0, const double* g0,
                   double eta, double* xdisp, double* signnorm, int n, const int32_t* mask,
                   int maskval, int batch, void* stream);
int sb_hvp_finish(const double* vfull, long long vstride, const double* gplus, const double* g0,
                  const double* signnorm, double eta, double* AV, double* Vs, double* AVs, int kcap,
                  int32_t* ksz, int32_t* nhist, int n, const int32_t* mask, int maskval, int batch,
                  void* stream);
/* vstride: distance in doubles between the direction vectors of consecutive systems. */
/* PES.diag tail, sella/peswrapper.py:541-551: Ritz rotation of the history (in place). */
/* HcVs (may be NULL): [b,kcap,n] = Hc @ Vs, the constraint-Hessian term of :546 (Atilde -= Vs^T Hc Vs). */
int sb_history_ritz(double* Vs, double* AVs, int kcap, const int32_t* nhist, int n,
                    int32_t* nvec_out, const int32_t* dav_state, int32_t* status, const double* HcVs,
                    int batch, void* stream);

/* ---- quasi-Newton update (sella/hessian_update.py:40-152, sella/linalg.py:274-304) ----
 * sb_update_prep: Ytil = symmetrize_Y(S, Y, symm) with symm in {0,1,2} (:27-37), the
 *   short-step no-op (:49-52 -> skip[b] = 1) and, if first != 0, lam0 of the scaled-identity
 *   start (:58-67).  kvec may be NULL (one secant pair).
 * sb_update_mid: method 0 TS-BFGS, 1 PSB, 2 Greenstadt, 3 DFP, 4 BFGS, 5 SR1, 6 BFGS_auto
 *   (:77-152).  absBS is needed by 0 and 6; evals (spectrum of B, ascending; may be NULL = "not
 *   positive definite") by 6; kout[b] (needed by 4 and 6, else may be NULL) = number of
 *   (U, J, W) triples written: k, or 2k for BFGS, which therefore needs 2k <= kcap.
 * sb_update_apply: B += sum_a (U_a J_a^T + J_a U_a^T) - 1/2 (W_a U_a^T + U_a W_a^T), i.e. the
 *   reference's Bplus = (Bplus + Bplus.T)/2 (:109); pass kout as kvec for methods 4/6.       */
int sb_update_prep(const double* S, const double* Y, double* Ytil, int kcap, const int32_t* kvec,
                   int n, int ncart, int first, int symm, double* lam0, int32_t* skip,
                   int32_t* status, const int32_t* active, int batch, void* stream);
int sb_fill_scaled_identity(double* B, double* evals, double* Vt, const double* lam0, int n,
                            int ncart, const int32_t* skip, int batch, void* stream);
int sb_abs_scale(const double* VtS, const double* evals, double* out, int kcap, int n,
                 const int32_t* skip, int batch, void* stream);
int sb_update_mid(const double* S, const double* Ytil, const double* BS, const double* absBS,
                  double* U, double* J, double* W, double* Xwork, int kcap, const int32_t* kvec,
                  int n, int method, const int32_t* skip, int32_t* status, double* Cout,
                  const double* evals, int32_t* kout, int batch, void* stream);
/* Cout (may be NULL): [b, 32*33] receives C = J^T S (k x k, leading dimension 33).     */
int sb_update_apply(double* B, const double* U, const double* J, const double* W, int kcap,
                    const int32_t* kvec, int n, const int32_t* skip, int batch, void* stream);

/* ---- eigendecomposition update after a low-rank Hessian update -------------------
 * Replaces the fresh scipy.linalg.eigh(B) the reference runs after every update
 * (sella/linalg.py:293 -> 174-195) by the diagonal-plus-rank-one update of the existing
 * eigenpairs (secular equation, LAPACK dlaed2/3/4 scheme).
 * sb_lowrank_factor: Delta = U J^T + J U^T - U sym(C) U^T = sum_t sig[t] p_t p_t^T,
 *   P [b, 2*kcap, n], sig [b, 2*kcap], nterm [b]; kcap <= 16.
 * caller: Z = Vt @ P  (sb_hv_ld with nvec = 2k, ldv = 2*kcap).
 * sb_secular_update: (evals, Vt) of B  ->  (evals, Vt) of B + Delta, in place
 *   (evals ascending, rows of Vt = eigenvectors); work, qwork: b*n*n doubles each.   */
int sb_lowrank_factor(const double* U, const double* J, const double* Cmat, int kcap,
                      const int32_t* kvec, int n, double* P, double* sig, int32_t* nterm,
                      const int32_t* skip, int batch, void* stream);
/* diagnostic: cycles spent per phase of sb_secular_update, summed over CTAs (host array
 * of 16 uint64; synchronises the device).                                              */
int sb_secular_profile(unsigned long long* out16, int reset);
/* diagnostic: enable > 0 switches on CUDA-event timing of the three kernels of the following
 * sb_secular_update calls; enable <= 0 reads the last call's {cluster_qr, cluster_reflect,
 * secular_update} milliseconds into out3 (synchronises; enable == 0 also switches it off).  */
int sb_secular_timing(float* out3, int enable);
int sb_secular_update(double* evals, double* Vt, double* Z, int zcap, const double* sig,
                      const int32_t* nterm, int n, double* work, double* qwork, int32_t* status,
                      const int32_t* skip, int batch, void* stream);

/* ---- restricted step (sella/optimize/restricted_step.py:72-121, stepper.py:75-96) ----
 * quasi-Newton model in the eigenbasis; Vg = Vt @ g.  tr: coef with s = Vt.T @ coef;
 * ras: the Cartesian step itself (max atomic displacement constraint).               */
int sb_qn_tr(const double* Vg, const double* evals, const double* delta, int order, int n,
             double* coef, double* smag, double* alpha, int32_t* status, const int32_t* active,
             const double* extra2, int batch, void* stream);
/* extra2 (may be NULL): |scons|^2 of the constraint-restoring part, added under the norm;
 * sadd (may be NULL): scons itself, added to the Cartesian step before the atomic norms. */
/* rational-function models (sella/optimize/stepper.py:114-185) with the spherical trust
 * region: mode 0 = rfo, 1 = prfo (Sella's default for saddles).  The bordered-matrix
 * eigenproblem per alpha is solved as an arrow-head secular equation in the eigenbasis. */
int sb_rfo_tr(const double* Vg, const double* evals, const double* delta, int order, int n, int mode,
              double* coef, double* smag, double* alpha, int32_t* status, const int32_t* active,
              const double* extra2, int batch, void* stream);
/* diagnostic: {alpha evaluations, secular iterations, secular roots, cycles, systems} summed
 * over the systems sb_rfo_tr has solved (host array of 8 uint64; synchronises the device).  */
int sb_rfo_profile(unsigned long long* out8, int reset);
int sb_rfo_ras(const double* Vg, const double* evals, const double* Vt, const double* delta,
               int order, int n, int mode, double* s, double* smag, double* alpha, int32_t* status,
               const int32_t* active, const double* sadd, int batch, void* stream);
int sb_qn_ras(const double* Vg, const double* evals, const double* Vt, const double* delta,
              int order, int n, double* s, double* smag, double* alpha, int32_t* status,
              const int32_t* active, const double* sadd, int batch, void* stream);

/* ---- linear constraints C x = c (sella/peswrapper.py:51-69, 395-407, 429-438, 558-568;
 * restricted_step.py:28-49).  Thin row-wise matrices R [nr, n]; rstride = nr*n for
 * per-system data, 0 when the batch shares one matrix.
 *   sb_rect_dots : out[b,j] = R[j,:].x[b,:] - (c ? c[b,j] : 0)
 *   sb_rect_comb : out[b,:] = beta*base[b,:] + scale * sum_j coef[b,j] R[j,:]
 *   sb_scons_measure / sb_combine_step : size of the constraint-restoring step, the
 *     "violation alone exceeds the radius" branch (NaiveStepper), s_tot = s_free + scons
 *   sb_converged_cons : fmax over atoms of the projected gradient, cmax = |res|.        */
/* Position-dependent constraints (peswrapper.py:395-407, 429-438, 467-481): Uc [b,nc,n] = the
 * orthonormalised rows of the constraint Jacobian drdx (sb_mgs), G [b,nc,nc] = drdx Uc^T (sb_gemm).
 *   Mr [b,nc,n] = G^-T Uc          (scons = -sum_j res_j Mr[j,:], as for linear constraints)
 *   L  [b,nc]   = G^-T u, u = Uc g (Lagrange multipliers, lstsq(drdx^T, g); u/L may be NULL)    */
int sb_cons_solve(const double* G, const double* Uc, const double* u, int nc, int n, double* Mr,
                  double* L, int32_t* status, const int32_t* active, int batch, void* stream);
int sb_rect_dots(const double* R, long long rstride, int nr, const double* x, long long xstride,
                 const double* c, double* out, int n, const int32_t* active, int batch, void* stream);
int sb_rect_comb(const double* R, long long rstride, int nr, const double* coef, double scale,
                 const double* base, long long bstride, double beta, double* out, long long ostride,
                 int n, const int32_t* active, int batch, void* stream);
int sb_scons_measure(const double* scons, const double* delta, int kind, int n, double* scons2,
                     double* consval, int32_t* naive, int32_t* regular, int batch, void* stream);
int sb_combine_step(const double* slift, const double* scons, const double* consval,
                    const double* delta, const int32_t* naive, double* stot, double* smag, int n,
                    const int32_t* active, int batch, void* stream);
int sb_converged_cons(const double* pg, const double* res, int nr, int n, double fmax_tol,
                      double cmax_tol, double* fmax_out, double* cmax_out, int32_t* conv, int batch,
                      void* stream);

/* s = V c and |B| s = V(|lam| * c) from ONE transposed pass over Vt: pack the two
 * coefficient vectors [b,2,n], run sb_hv_ld(transposed, nvec=2), unpack.               */
int sb_pack_coef(const double* coef, const double* evals, double* out2, int n,
                 const int32_t* active, int batch, void* stream);
int sb_unpack2(const double* in2, double* s, double* absBs, int n, const int32_t* active,
               int batch, void* stream);

/* ---- internal coordinates (sella/internal.py:58-80, 466-470, 1735-1902, 2189-2575;
 * sella/linalg.py:601-646).  Topology (int32 index arrays, optional PBC shift vectors) is
 * shared by the batch; coordinate order: translations (atom, dim), bonds, angles, dihedrals.
 * Derivatives by forward-mode (hyper-)dual numbers of the reference's primal formulas.
 *   sb_internals_qB  : q[b,nint] and, if Bmat != NULL (zero-filled), the Wilson B-matrix
 *                      Bmat[b,nint,n] = dq/dx
 *   sb_internals_hess: D[b,n,n] += sum_c v[b,c] d2q_c/dx2 (if D != NULL, zero-filled; "ldot"),
 *                      R[b,c,:] = (d2q_c/dx2) w[b,:]       (if R != NULL, zero-filled; "rdot")  */
int sb_internals_qB(const int32_t* trans, int nt, const int32_t* bonds, int nb, const int32_t* angles,
                    int na, const int32_t* diheds, int nd, const double* tb, const double* ta,
                    const double* td, const double* x, int n, double* q, double* Bmat,
                    const int32_t* active, int batch, void* stream);
int sb_internals_hess(const int32_t* trans, int nt, const int32_t* bonds, int nb, const int32_t* angles,
                      int na, const int32_t* diheds, int nd, const double* tb, const double* ta,
                      const double* td, const double* x, int n, const double* v, double* D,
                      const double* w, double* R, const int32_t* active, int batch, void* stream);

/* Rotation coordinate of a whole configuration (sella/internal.py:507-800, class Rotation
 * :1031-1078): refpos [natoms,3] CENTRED reference geometry (refstride 0 = shared, 3*natoms = per
 * system); qprev [b,4] (in/out, may be NULL = identity start) keeps the quaternion branch;
 * vals (may be NULL): 3 rotation-vector components at vals + b*valstride; J (may be NULL): three
 * Jacobian rows of length 3*natoms at J + b*jstride; if L (3 multipliers at L + b*lstride) and D
 * are given, D[b] (3N x 3N) += sum_k L_k d2v_k/dx2.  work: batch * 14 * 3*natoms doubles.         */
int sb_rotation(const double* x, int natoms, const double* refpos, long long refstride, double* qprev,
                double* vals, long long valstride, double* J, long long jstride, const double* L,
                long long lstride, double* D, double* work, const int32_t* active, int batch,
                void* stream);

/* ---- dense algebra of the internal-coordinate path (sella/peswrapper.py:674-736, 1011-1082,
 * 1124-1127, 1176-1183; sella/_gpu.py:100-132).  Row-major matrices with explicit leading
 * dimensions; batch strides in doubles (0 = one matrix shared by the batch).
 *   sb_gemm : C = alpha op(A) op(B) + beta C, op(A) M x K, op(B) K x N (transX != 0: X is stored
 *             transposed); fp64 tensor-core (DMMA) tiles.  Serves gpu_project (U^T H U), g_int =
 *             B+^T g_cart, Binv = R^-1 Q^T, Hc = B+^T (D_c - D_q) B+, get_df_pred.
 *   sb_qr   : economy Householder QR of A [b, m, n] (m >= n; A is overwritten with the reflectors),
 *             Q [b, m, n] with orthonormal columns, R [b, n, n] upper triangular (LAPACK signs):
 *             gpu_qr (_gpu.py:100-111), _get_jacobian_qr (peswrapper.py:674-709).  Blocked (panels of
 *             16 columns, compact WY, trailing updates as sb_gemm); work: batch * (m*n + 32*n +
 *             256*ceil(n/16)) doubles.  Q == NULL: R only (half the work; B+ w = R^-1 R^-T A^T w then
 *             needs no Q -- the stages of the geodesic integrator, peswrapper.py:1200-1221).
 *   sb_trtri: Rinv [b, n, n] = R^-1 (upper triangular; diagonal blocks of <= 48 by back substitution,
 *             assembled with sb_gemm); work: batch*n*n doubles; SB_ST_SINGULAR on a zero diagonal. */
int sb_gemm(int transA, int transB, int M, int N, int K, double alpha, const double* A, int lda,
            long long strideA, const double* B, int ldb, long long strideB, double beta, double* C,
            int ldc, long long strideC, const int32_t* active, int batch, void* stream);
int sb_qr(double* A, int m, int n, double* Q, double* R, double* work, const int32_t* active, int batch,
          void* stream);
int sb_trtri(const double* R, double* Rinv, double* work, int n, int32_t* status, const int32_t* active,
             int batch, void* stream);
/* sb_potrf: in-place upper Cholesky factor of symmetric positive definite A [b, n, n] (A = R^T R; panels of 32
 * in shared memory + one sb_gemm trailing update per panel; SB_ST_SINGULAR on a non-positive pivot).  With
 * A = Bw^T Bw it gives the R factor of the Wilson matrix (np.linalg.qr at sella/peswrapper.py:691, _gpu.py:
 * 100-111) where Q is not needed: the stages of the geodesic integrator (peswrapper.py:1200-1221) use
 * B+ w = R^-1 R^-T Bw^T w.                                                                               */
int sb_potrf(double* A, int n, int32_t* status, const int32_t* active, int batch, void* stream);

/* ---- compact representation of the approximate Hessian and of its spectrum --------------------
 * B = lam0 I + VR^T diag(theta - lam0) VR: mrows[b] explicit eigenpairs (theta ascending in
 * evals[b*estride + i], eigenvector i = row i of VR at VR + b*vstride + i*n) and the eigenvalue lam0[b]
 * on the whole orthogonal complement, whose eigenvectors are never stored (a quasi-Newton Hessian
 * is a scaled identity plus low rank, sella/hessian_update.py:58-111).  Replaces the dense
 * eigh(B) of sella/linalg.py:174-195 / 293 at O(n m) per pass; for mrows = n it is the dense
 * eigendecomposition.  numpy blueprint: tests/compact_proto.py.
 *   sb_hv_rect          : sb_hv_ld on the first mrows (HOST bound) rows of A, systems astride doubles apart
 *                         (transposed = 0: Y[b,v,:mrows] = A X; 1: Y[b,v,:] = sum_{i<mrows} X[b,v,i] A[i,:])
 *   sb_compact_append_a : candidates Qc (unit, mutually orthogonal) for the part of the update vectors
 *                         P [b,zcap,n] outside span(VR); W1 = VR^T (VR P) (two sb_hv_rect passes)
 *   sb_compact_append_b : candidates projected once more (W2 = VR^T (VR Qc)), re-orthonormalised,
 *                         appended as rows mrows.. with eigenvalue lam0; Z[b,t,mrows+j] = q_j.p_t;
 *                         mrows[b] += number appended
 *   sb_secular_update_c : sb_secular_update on the first mrows[b] rows (mcap: HOST bound on mrows).  With
 *                         aux (int32 [b, mcap + 4]) and nterm_max (HOST bound on nterm) the work is split:
 *                         per rank-one term one solve launch (deflation, secular roots, rotation matrix),
 *                         the rotation of the eigenvector rows as a batched fp64 tensor-core GEMM and a
 *                         copy back, then the final sort; aux == NULL: one launch does everything
 *   sb_compact_prepare  : merged ascending pole list (width entries, stride width) for sb_qn_tr /
 *                         sb_rfo_tr / the *_ras_c kernels: explicit eigenvalues with coefficients Vg = VR g,
 *                         the complement as ONE pole lam0 with coefficient |g_perp| (g_perp = g - Wg,
 *                         Wg = VR^T Vg), zero-weight copies of lam0 as padding; rowmap: explicit row, -1
 *                         complement, -2 padding
 *   sb_compact_finish   : pole coefficients -> C4[b,0..2,:] = c, |theta| c, theta c on the explicit rows,
 *                         kappa[b] = c_complement / |g_perp|
 *   sb_compact_finish2  : (T4 = VR^T C4)  s = T4[0] + kappa g_perp, |B| s, B s likewise, xnew = x + s
 *   sb_compact_scale / sb_compact_axpy : f(B) S = f(lam0) S + VR^T[(f(theta) - f(lam0)) (VR S)],
 *                         mode 0: f = |.| (the |B| S of TS-BFGS, hessian_update.py:118-125), 1: f = id
 *   sb_compact_jd_coeff / sb_compact_jd_finish : Jacobi-Davidson correction (eigensolvers.py:115-139)
 *                         in the eigenbasis of the preconditioner, method 0 jd0, 1 gd
 *   sb_compact_lowest   : the k lowest eigenvalues of B per system (out [b,k]; the test of optimize.py:369-371)
 *   sb_qn_ras_c / sb_rfo_ras_c / sb_davidson_init_c : the dense kernels of the same name on a pole
 *                         list / the compact rows                                                      */
int sb_hv_rect(const double* A, long long astride, int mrows, const double* X, double* Y, const int32_t* active,
               int batch, int n, int nvec, int ldv, int transposed, void* stream);
int sb_compact_append_a(const double* P, const double* W1, int zcap, const int32_t* nterm, const int32_t* mrows,
                        int n, double* Qc, int32_t* ncand, const int32_t* skip, int batch, void* stream);
int sb_compact_append_b(const double* P, const double* Qc, const double* W2, int zcap, const int32_t* nterm,
                        const int32_t* ncand, int n, double* evals, long long estride, double* VR, long long vstride,
                        int32_t* mrows, const double* lam0, double* Z, const int32_t* skip, int batch, void* stream);
int sb_secular_update_c(double* evals, double* Vt, double* Z, int zcap, const double* sig, const int32_t* nterm,
                        int n, double* work, double* qwork, int32_t* status, const int32_t* skip,
                        const int32_t* mrows, int mcap, long long estride, long long vstride, int nterm_max,
                        int32_t* aux, int batch, void* stream);
int sb_compact_prepare(const double* g, const double* Vg, const double* Wg, const double* evals, long long estride,
                       const int32_t* mrows, const double* lam0, int n, int width, double* gperp, double* gam,
                       double* cev, double* cvg, int32_t* rowmap, const int32_t* active, int batch, void* stream);
int sb_compact_finish(const double* ccoef, const int32_t* rowmap, int width, const double* evals, long long estride,
                      const double* gam, int n, int mbound, double* C4, double* kappa, const int32_t* active,
                      int batch, void* stream);
int sb_compact_finish2(const double* T4, const double* gperp, const double* kappa, const double* lam0,
                       const double* x, int n, double* s, double* absBs, double* Bs, double* xnew,
                       const int32_t* active, int batch, void* stream);
int sb_compact_scale(const double* VtS, const double* evals, long long estride, const int32_t* mrows,
                     const double* lam0, int kcap, int nvec, int n, int mbound, int mode, double* out,
                     const int32_t* skip, int batch, void* stream);
int sb_compact_axpy(const double* S, const double* T, const double* lam0, int kcap, int nvec, int n, int mode,
                    double* out, const int32_t* skip, int batch, void* stream);
int sb_compact_jd_coeff(const double* rvhat, const double* rv, const double* evals, long long estride,
                        const int32_t* mrows, const double* lam0, const double* theta, int n, int mbound, int method,
                        double* that, double* ed, const int32_t* dav_state, int batch, void* stream);
int sb_compact_jd_finish(double* t, const double* rv, const double* ed, int n, int method, const int32_t* dav_state,
                         int batch, void* stream);
int sb_compact_lowest(const double* evals, long long estride, const int32_t* mrows, const double* lam0, int n,
                      int k, double* out, int batch, void* stream);
/* out[b,v,:] = mask[:] * X[b,v,:], v < nvec: the projection onto the free coordinates when whole Cartesian
 * coordinates are fixed (Constraints.fix_translation(i): Ufree of sella/peswrapper.py:51-69 is a permutation) */
int sb_mask_vec(const double* X, const double* mask, double* out, int ldv, int nvec, int n, int batch, void* stream);
int sb_qn_ras_c(const double* Vg, const double* evals, const double* Vt, const double* delta, int order, int n,
                double* s, double* smag, double* alpha, int32_t* status, const int32_t* active, const double* sadd,
                int npole, const int32_t* rowmap, const double* gperp, const double* gam, long long vstride,
                int batch, void* stream);
int sb_rfo_ras_c(const double* Vg, const double* evals, const double* Vt, const double* delta, int order, int n,
                 int mode, double* s, double* smag, double* alpha, int32_t* status, const int32_t* active,
                 const double* sadd, int npole, const int32_t* rowmap, const double* gperp, const double* gam,
                 long long vstride, int batch, void* stream);
/* MaxInternalStep (sella/optimize/restricted_step.py:186-243, the default restricted step of
 * Sella(internal=True), optimize.py:166-168): cons(s) = max_j |s_j w_j| over the n internal coordinates.
 * The model is a list of npole poles (evals[b,npole], Vg[b,npole]); pole i's eigenvector, lifted to the
 * n internal coordinates, is row i of Wt[b] (vstride doubles per system, row length n); the step is
 * s = sum_i c_i(alpha) Wt[i,:] + sadd with alpha solving cons(s) = delta (interior step when cons < delta
 * at the model's alpha0).  w [n] is shared by the batch (wx/wb/wa/wd per coordinate kind, :222-243).   */
int sb_qn_mis(const double* Vg, const double* evals, const double* Wt, const double* delta, int order, int n,
              double* s, double* smag, double* alpha, int32_t* status, const int32_t* active, const double* sadd,
              int npole, long long vstride, const double* w, int batch, void* stream);
int sb_rfo_mis(const double* Vg, const double* evals, const double* Wt, const double* delta, int order, int n,
               int mode, double* s, double* smag, double* alpha, int32_t* status, const int32_t* active,
               const double* sadd, int npole, long long vstride, const double* w, int batch, void* stream);
int sb_davidson_init_c(const double* v0, const double* pl, const double* Pvt, int mode, double* V, int kcap, int n,
                       int32_t* ksz, int32_t* ninit, int32_t* nhist, int32_t* dav_state, int32_t* status,
                       const int32_t* part, const int32_t* mrows, const double* lam0, const double* gperp,
                       long long estride, long long vstride, int batch, void* stream);

/* diagnostic: average milliseconds (HOST float) of `reps` launches of the rotation GEMM of sb_secular_update_c's
 * split mode (work[j,:] = sum_i Qh[i,j] Vt[i,:], r rows per system, 2 r^2 n flops per system); aux: int32
 * [batch, r + 4] scratch.  Synchronises. */
int sb_secular_apply_bench(const double* Vt, const double* qwork, double* work, int32_t* aux, int r, int n,
                           long long vstride, int batch, int reps, float* ms_host, void* stream);

/* diagnostic: measured fp64 throughput of the device in TFLOP/s (HOST double), kind 0 = DFMA pipe, 1 = DMMA
 * (mma.sync m8n8k4 f64) tensor path; scratch: sms * ctas_per_sm * 256 doubles of device memory.  The
 * denominator of the compute rooflines in bench.py (MEASURED_PEAKS.json has no fp64 figure).  Synchronises. */
int sb_fp64_peak(int kind, int iters, int ctas_per_sm, double* scratch, double* tflops_host);

/* M[b] <- scale * M[b] + diag * I (n x n, in place): the identity step model on the free space of
 * position-dependent constraints, Bp = P_f + sigma P_c = I + (sigma - 1) Ucons Ucons^T (the projected
 * Hessian of an uninitialised ApproximateHessian is None, sella/peswrapper.py:363-386, sella/linalg.py:306-317) */
int sb_add_scaled_identity(double* M, double scale, double diag, int n, int batch, void* stream);

/* ---- per-step bookkeeping (sella/peswrapper.py:578-602, optimize/optimize.py:362-434) ----
 * dpar = {rho_inc, rho_dec, sigma_inc, sigma_dec, delta_min} (host array of 5 doubles),
 * ipar = {order, eig, nsteps_per_diag, diag_every_n(<0: never)} (host array of 4 ints). */
int sb_axpy(const double* x, const double* s, double* out, int n, const int32_t* active,
            int batch, void* stream);
int sb_kick_finish(double* x, double* f, double* g, const double* xnew, const double* fnew,
                   const double* gnew, const double* s, const double* Bs, const double* smag,
                   double* dg, double* delta, double* rho, int32_t* nsteps, const double* dpar,
                   const int32_t* ipar, int n, const int32_t* active, int batch, void* stream);
int sb_ev_decide(const double* evals, int n, int has_evals, int32_t* since_diag, int32_t* ev,
                 const double* dpar, const int32_t* ipar, const int32_t* active, int batch, void* stream);
int sb_converged(const double* g, int n, double fmax_tol, double* fmax_out, int32_t* conv,
                 int batch, void* stream);

#ifdef __cplusplus
}
#endif
#endif
