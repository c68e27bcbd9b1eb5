// Batched dense H.V kernels: the Davidson matvec / synthetic-PES gradient /
// B.s / V^T g contractions of the Sella inner loop.
//
//   hv   : Y[b,v,:] = A[b] . X[b,v,:]          (row dots;      "A @ X")
//   hv_t : Y[b,v,:] = A[b]^T . C[b,v,:]        (row combination "A.T @ C")
//
// Reference call sites these replace: A.dot(V) in rayleigh_ritz
// (sella/eigensolvers.py:52,112), the finite-difference gradient difference
// behind NumericalHessian._matvec on a quadratic surface (sella/linalg.py:82-87),
// B @ S (sella/hessian_update.py:119), V.T @ g / V @ c in QuasiNewton
// (sella/optimize/stepper.py:86,93-95).
//
// Layout: A[b] is n x n row-major fp64 (numpy C order, as the reference stores
// B), batch-leading; vectors are "vector-major" [b, nvec, n].
//
// Both kernels are HBM-bound (0.25 flop/byte at nvec=1).  Design: persistent
// CTAs (one per SM), a dedicated producer warp streams row tiles of A -- each a
// single contiguous run of R*n*8 bytes because rows are stored whole -- with the
// TMA engine (cp.async.bulk, completion on an mbarrier) into a ring of shared
// memory stages; 8 consumer warps do fp64 FMAs out of shared memory and hand the
// slot back through a second mbarrier.  Algorithmic bytes per system:
// 8*(n^2 + 2*n*nvec).
#include "common.cuh"

namespace {

constexpr int NCW = 8;                       // consumer warps
constexpr int HV_THREADS = (NCW + 1) * 32;   // + 1 producer warp

struct HvPlan {
    int rows;        // rows of A per stage
    int stages;      // ring depth
    int tile_doubles;
    int vec_doubles;
    size_t smem;
};

template <int NV>
__global__ void __launch_bounds__(HV_THREADS, 1)
hv_tma_kernel(const double* __restrict__ A, const double* __restrict__ X, double* __restrict__ Y,
              const int* __restrict__ active, int batch, int n, int ldv, int R, int S, int m, long long astride) {
    // A[b] is m x n (the first m rows of a matrix whose systems are astride doubles apart)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_d = R * n, vec_d = NV * n;
    const int stage_d = tile_d + vec_d;
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + (size_t)S * stage_d);
    uint64_t* empty = full + S;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            sb_mbar_init(&full[s], 1);
            sb_mbar_init(&empty[s], NCW);
        }
        sb_fence_mbar_init();
    }
    __syncthreads();

    const int tps = (m + R - 1) / R;             // row tiles per system
    const long long ntiles = (long long)batch * tps;

    if (warp == NCW) {
        // ------------------------------------------------ producer
        if (lane == 0) {
            int it = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int b = (int)(t / tps);
                if (active && !active[b]) continue;
                const int r0 = (int)(t % tps) * R;
                const int rows = min(R, m - r0);
                const int s = it % S;
                const uint32_t ph = (uint32_t)((it / S) & 1);
                sb_mbar_wait(&empty[s], ph ^ 1u);
                double* dst = stage_base + (size_t)s * stage_d;
                const uint32_t tb = (uint32_t)rows * n * 8u, vb = (uint32_t)vec_d * 8u;
                sb_mbar_expect_tx(&full[s], tb + vb);
                sb_tma_load_1d(dst, A + (size_t)b * astride + (size_t)r0 * n, tb, &full[s]);
                sb_tma_load_1d(dst + tile_d, X + (size_t)b * ldv * n, vb, &full[s]);
                ++it;
            }
        }
    } else {
        // ------------------------------------------------ consumers
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int b = (int)(t / tps);
            if (active && !active[b]) continue;
            const int r0 = (int)(t % tps) * R;
            const int rows = min(R, m - r0);
            const int s = it % S;
            const uint32_t ph = (uint32_t)((it / S) & 1);
            sb_mbar_wait(&full[s], ph);
            const double* tile = stage_base + (size_t)s * stage_d;
            const double* xs = tile + tile_d;
            // each warp takes rows (warp, warp+NCW) together: the vector is read once for
            // both rows and every row has two accumulators -> four independent FMA chains
            for (int r = warp; r < rows; r += 2 * NCW) {
                const int r2 = r + NCW;
                const bool two = r2 < rows;
                double a0[NV], a1[NV], b0[NV], b1[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) { a0[v] = 0.0; a1[v] = 0.0; b0[v] = 0.0; b1[v] = 0.0; }
                const double2* rowa = reinterpret_cast<const double2*>(tile + (size_t)r * n);
                const double2* rowb = reinterpret_cast<const double2*>(tile + (size_t)(two ? r2 : r) * n);
                for (int j = lane; j < (n >> 1); j += 32) {
                    const double2 a = rowa[j];
                    const double2 bq = rowb[j];
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const double2 x = reinterpret_cast<const double2*>(xs + (size_t)v * n)[j];
                        a0[v] = fma(a.x, x.x, a0[v]);
                        a1[v] = fma(a.y, x.y, a1[v]);
                        b0[v] = fma(bq.x, x.x, b0[v]);
                        b1[v] = fma(bq.y, x.y, b1[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double ta = sb_warp_sum(a0[v] + a1[v]);
                    const double tb = sb_warp_sum(b0[v] + b1[v]);
                    if (lane == 0) {
                        Y[((size_t)b * ldv + v) * n + r0 + r] = ta;
                        if (two) Y[((size_t)b * ldv + v) * n + r0 + r2] = tb;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) sb_mbar_arrive(&empty[s]);
            ++it;
        }
    }
}

// y = A^T c : whole systems per CTA so the column accumulators never leave the SM.
template <int NV>
__global__ void __launch_bounds__(HV_THREADS, 1)
hvt_tma_kernel(const double* __restrict__ A, const double* __restrict__ C, double* __restrict__ Y,
               const int* __restrict__ active, int batch, int n, int ldv, int R, int S, int m, long long astride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tile_d = R * n;
    double* stage_base = reinterpret_cast<double*>(smem_raw);
    double* coef = stage_base + (size_t)S * tile_d;            // NV * n coefficients of the system
    double* red = coef + (size_t)NV * n;                        // cross-group reduction, NV * n * (G-1)
    const int nslots = n >> 1;                                  // double2 column slots
    int G = (NCW * 32) / nslots;                                // row groups sharing the tile
    if (G < 1) G = 1;
    if (G > 8) G = 8;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + (size_t)NV * n * (G > 1 ? (G - 1) : 0));
    uint64_t* empty = full + S;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            sb_mbar_init(&full[s], 1);
            sb_mbar_init(&empty[s], NCW);
        }
        sb_fence_mbar_init();
    }
    __syncthreads();

    const int tps = (m + R - 1) / R;
    const int ctid = threadIdx.x;                               // consumer thread id (< NCW*32)
    const int gsize = (NCW * 32) / G;                           // threads per row group
    const int grp = ctid / gsize, gt = ctid % gsize;
    constexpr int MAXSLOT = 4;                                  // slots per thread (n <= 2*4*gsize)

    int it = 0;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        if (active && !active[b]) continue;
        if (warp == NCW) {
            if (lane == 0) {
                for (int tt = 0; tt < tps; ++tt, ++it) {
                    const int r0 = tt * R, rows = min(R, m - r0);
                    const int s = it % S;
                    const uint32_t ph = (uint32_t)((it / S) & 1);
                    sb_mbar_wait(&empty[s], ph ^ 1u);
                    const uint32_t tb = (uint32_t)rows * n * 8u;
                    sb_mbar_expect_tx(&full[s], tb);
                    sb_tma_load_1d(stage_base + (size_t)s * tile_d, A + (size_t)b * astride + (size_t)r0 * n, tb,
                                   &full[s]);
                }
            }
            else it += tps;
            __syncwarp();
            continue;
        }
        // consumers: stage the coefficients of this system
        for (int i = ctid; i < NV * n; i += NCW * 32) coef[i] = (i % n) < m ? C[(size_t)b * ldv * n + i] : 0.0;
        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32));

        double2 acc[NV][MAXSLOT];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int q = 0; q < MAXSLOT; ++q) acc[v][q] = make_double2(0.0, 0.0);

        for (int tt = 0; tt < tps; ++tt, ++it) {
            const int r0 = tt * R, rows = min(R, m - r0);
            const int s = it % S;
            const uint32_t ph = (uint32_t)((it / S) & 1);
            sb_mbar_wait(&full[s], ph);
            const double* tile = stage_base + (size_t)s * tile_d;
            for (int r = grp; r < rows && grp < G; r += G) {
                const double2* row2 = reinterpret_cast<const double2*>(tile + (size_t)r * n);
                double c[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) c[v] = coef[(size_t)v * n + r0 + r];
#pragma unroll
                for (int q = 0; q < MAXSLOT; ++q) {
                    const int slot = gt + q * gsize;
                    if (slot < nslots) {
                        const double2 a = row2[slot];
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            acc[v][q].x = fma(c[v], a.x, acc[v][q].x);
                            acc[v][q].y = fma(c[v], a.y, acc[v][q].y);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) sb_mbar_arrive(&empty[s]);
        }
        // combine the row groups (fixed order -> deterministic) and store
        if (grp > 0 && grp < G) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int q = 0; q < MAXSLOT; ++q) {
                    const int slot = gt + q * gsize;
                    if (slot < nslots)
                        reinterpret_cast<double2*>(red + ((size_t)(grp - 1) * NV + v) * n)[slot] = acc[v][q];
                }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32));
        if (grp == 0) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int q = 0; q < MAXSLOT; ++q) {
                    const int slot = gt + q * gsize;
                    if (slot < nslots) {
                        double2 tot = acc[v][q];
                        for (int g = 1; g < G; ++g) {
                            const double2 p =
                                reinterpret_cast<const double2*>(red + ((size_t)(g - 1) * NV + v) * n)[slot];
                            tot.x += p.x;
                            tot.y += p.y;
                        }
                        reinterpret_cast<double2*>(Y + ((size_t)b * ldv + v) * n)[slot] = tot;
                    }
                }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(NCW * 32));
    }
}

// Fallback for odd n (no 16-byte row alignment): warp per row / thread per column.
template <int NV>
__global__ void hv_ldg_kernel(const double* __restrict__ A, const double* __restrict__ X,
                              double* __restrict__ Y, const int* __restrict__ active, int batch, int n,
                              int ldv, int m, long long astride) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < m; r += nw) {
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.0;
        const double* row = A + (size_t)b * astride + (size_t)r * n;
        for (int j = lane; j < n; j += 32) {
            const double a = row[j];
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = fma(a, X[((size_t)b * ldv + v) * n + j], acc[v]);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double tot = sb_warp_sum(acc[v]);
            if (lane == 0) Y[((size_t)b * ldv + v) * n + r] = tot;
        }
    }
}

template <int NV>
__global__ void hvt_ldg_kernel(const double* __restrict__ A, const double* __restrict__ C,
                               double* __restrict__ Y, const int* __restrict__ active, int batch, int n,
                               int ldv, int m, long long astride) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.0;
        for (int r = 0; r < m; ++r) {
            const double a = A[(size_t)b * astride + (size_t)r * n + j];
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = fma(C[((size_t)b * ldv + v) * n + r], a, acc[v]);
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) Y[((size_t)b * ldv + v) * n + j] = acc[v];
    }
}

int g_sms = 0;
int g_smem_optin = 0;
int g_dev = -1;

// SM count and opt-in shared memory of the CURRENT device (one process may drive several devices)
void query_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev == g_dev) return;
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_dev = dev;
}

HvPlan plan_hv(int n, int nvec, bool transposed, int m) {
    HvPlan p;
    const int target = 48 * 1024;                 // bytes of A per stage
    int R = target / (n * 8);
    if (R >= 2 * NCW) R = (R / (2 * NCW)) * (2 * NCW);
    else if (R >= NCW) R = NCW;
    if (R < 1) R = 1;
    if (R > m) R = m < 1 ? 1 : m;
    p.rows = R;
    p.tile_doubles = R * n;
    p.vec_doubles = transposed ? 0 : nvec * n;
    const size_t stage_bytes = (size_t)(p.tile_doubles + p.vec_doubles) * 8;
    size_t fixed = 2 * 16 * 8 + 256;              // barriers + slack
    if (transposed) {
        int G = (NCW * 32) / (n / 2);
        if (G < 1) G = 1;
        if (G > 8) G = 8;
        fixed += (size_t)nvec * n * 8 * (1 + (G > 1 ? G - 1 : 0));
    }
    const size_t budget = (size_t)g_smem_optin - fixed - 1024;
    int S = (int)(budget / stage_bytes);
    if (S > 8) S = 8;
    if (S < 2) S = 2;
    p.stages = S;
    p.smem = (size_t)S * stage_bytes + fixed;
    return p;
}

template <int NV>
int launch_hv(const double* A, const double* X, double* Y, const int* active, int batch, int n,
              int ldv, int m, long long astride, cudaStream_t st) {
    query_device();
    if (m <= 0) return 0;
    // TMA bulk copies need 16-byte aligned sources: even n and an even system stride
    if ((n & 1) || n < 16 || (astride & 1)) {
        SB_COUNT(1);
        hv_ldg_kernel<NV><<<batch, 256, 0, st>>>(A, X, Y, active, batch, n, ldv, m, astride);
        return SB_LAUNCH_CHECK();
    }
    HvPlan p = plan_hv(n, NV, false, m);
    if (p.smem > (size_t)g_smem_optin) {
        SB_COUNT(1);
        hv_ldg_kernel<NV><<<batch, 256, 0, st>>>(A, X, Y, active, batch, n, ldv, m, astride);
        return SB_LAUNCH_CHECK();
    }
    cudaFuncSetAttribute(hv_tma_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    const long long ntiles = (long long)batch * ((m + p.rows - 1) / p.rows);
    const int grid = (int)(ntiles < g_sms ? ntiles : g_sms);
    SB_COUNT(1);
    hv_tma_kernel<NV><<<grid, HV_THREADS, p.smem, st>>>(A, X, Y, active, batch, n, ldv, p.rows, p.stages, m, astride);
    return SB_LAUNCH_CHECK();
}

template <int NV>
int launch_hvt(const double* A, const double* C, double* Y, const int* active, int batch, int n,
               int ldv, int m, long long astride, cudaStream_t st) {
    query_device();
    const bool tma_ok = !(n & 1) && n >= 16 && (n / 2) <= 4 * (NCW * 32) && !(astride & 1) && m > 0;
    HvPlan p = plan_hv(n, NV, true, m);
    if (!tma_ok || p.smem > (size_t)g_smem_optin) {
        SB_COUNT(1);
        hvt_ldg_kernel<NV><<<batch, 256, 0, st>>>(A, C, Y, active, batch, n, ldv, m, astride);
        return SB_LAUNCH_CHECK();
    }
    cudaFuncSetAttribute(hvt_tma_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    const int grid = batch < g_sms ? batch : g_sms;
    SB_COUNT(1);
    hvt_tma_kernel<NV><<<grid, HV_THREADS, p.smem, st>>>(A, C, Y, active, batch, n, ldv, p.rows, p.stages, m, astride);
    return SB_LAUNCH_CHECK();
}

}  // namespace

// X and Y are [b, ldv, n] with the first nvec slots in use; nvec is processed in
// chunks of 4/3/2/1 vectors per pass over A (a chunk is a contiguous sub-block of each
// system's vectors).  A[b]: the first m rows (length n) of a matrix with system stride astride.
extern "C" int sb_hv_rect_impl(const double* A, long long astride, int m, const double* X, double* Y,
                               const int* active, int batch, int n, int nvec, int ldv, int transposed,
                               cudaStream_t st) {
    int done = 0;
    while (done < nvec) {
        const int left = nvec - done;
        const int c = left >= 4 ? 4 : left;
        const double* x = X + (size_t)done * n;
        double* y = Y + (size_t)done * n;
        int rc;
        if (!transposed) {
            rc = c == 4   ? launch_hv<4>(A, x, y, active, batch, n, ldv, m, astride, st)
                 : c == 3 ? launch_hv<3>(A, x, y, active, batch, n, ldv, m, astride, st)
                 : c == 2 ? launch_hv<2>(A, x, y, active, batch, n, ldv, m, astride, st)
                          : launch_hv<1>(A, x, y, active, batch, n, ldv, m, astride, st);
        } else {
            rc = c == 4   ? launch_hvt<4>(A, x, y, active, batch, n, ldv, m, astride, st)
                 : c == 3 ? launch_hvt<3>(A, x, y, active, batch, n, ldv, m, astride, st)
                 : c == 2 ? launch_hvt<2>(A, x, y, active, batch, n, ldv, m, astride, st)
                          : launch_hvt<1>(A, x, y, active, batch, n, ldv, m, astride, st);
        }
        if (rc) return rc;
        done += c;
    }
    return 0;
}

extern "C" int sb_hv_ld_impl(const double* A, const double* X, double* Y, const int* active, int batch,
                             int n, int nvec, int ldv, int transposed, cudaStream_t st) {
    return sb_hv_rect_impl(A, (long long)n * n, n, X, Y, active, batch, n, nvec, ldv, transposed, st);
}

extern "C" int sb_hv_impl(const double* A, const double* X, double* Y, const int* active, int batch,
                          int n, int nvec, int transposed, cudaStream_t st) {
    return sb_hv_ld_impl(A, X, Y, active, batch, n, nvec, nvec, transposed, st);
}
