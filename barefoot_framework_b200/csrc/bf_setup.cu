// Per-set GP setup on the device: kernel matrix, batched Cholesky, L^-1 (+ tensor-core operand
// packing) and alpha.  Replaces gp_model.create_kernel/create_gp (gpModel.py:29-42), i.e.
// george GP.compute -> scipy cholesky, and the cho_solve for alpha behind gpModel.py:47,53.
// All matrices are np x np row-major, np = n rounded up to 32, padded with the identity so
// every 32 x 32 block is full.
#include <stdlib.h>

#include "bf_common.cuh"
#include "bf_block32.cuh"

// ---- K = amp k(X, X) + diag(yerr_eff^2) ------------------------------------------------------
// The covariance function and (for the common dimensions) nDim are compile-time constants: the distance
// loop unrolls and the four elements of a thread are straight-line code that the scheduler interleaves
// (D == 0: run-time nDim).
template <int KERN, int D>
__global__ void __launch_bounds__(256) bf_kernel_matrix_kernel(
    int n, int d_rt, int np, const double* __restrict__ X, const double* __restrict__ inv_l2,
    const double* __restrict__ amp, const int32_t* __restrict__ set_hp, const double* __restrict__ yerr,
    int yerr_n, long long yerr_stride, double* __restrict__ K, int lower_only) {
    if (lower_only && blockIdx.x > blockIdx.y) return;      // the factorisation never reads the blocks above the diagonal
    __shared__ double xi_s[32][BF_MAX_DIM + 1];
    __shared__ double xj_s[32][BF_MAX_DIM + 1];
    __shared__ double w_s[BF_MAX_DIM];
    const int d = D ? D : d_rt;
    const int set = blockIdx.z;
    const int hp = set_hp ? set_hp[set] : set;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < d) w_s[threadIdx.x] = bf_metric_scale(KERN, inv_l2[(size_t)hp * d + threadIdx.x]);
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * d; e += 256) {       // pre-scaled coordinates
        const int r = e / d, a = e - r * d;
        xi_s[r][a] = (i0 + r < n) ? X[(size_t)(i0 + r) * d + a] * w_s[a] : 0.0;
        xj_s[r][a] = (j0 + r < n) ? X[(size_t)(j0 + r) * d + a] * w_s[a] : 0.0;
    }
    __syncthreads();
    const double am = amp[hp];
    double* Kset = K + (size_t)set * np * np;
    const int j = j0 + tx;
    double q[4] = {0.0, 0.0, 0.0, 0.0};
    if (D) {
#pragma unroll
        for (int a = 0; a < (D ? D : 1); ++a) {
            const double xj = xj_s[tx][a];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const double df = xi_s[ty + 8 * rr][a] - xj;
                q[rr] = fma(df, df, q[rr]);
            }
        }
    } else {
        for (int a = 0; a < d; ++a) {
            const double xj = xj_s[tx][a];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const double df = xi_s[ty + 8 * rr][a] - xj;
                q[rr] = fma(df, df, q[rr]);
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int i = i0 + ty + 8 * rr;
        double v = am * bf_radial_q_k<KERN>(q[rr], bf_exp_tab_g);      // padded rows: q = 0, overwritten below
        if (i == j && i < n) {
            const double e = yerr_n == 1 ? yerr[0] : yerr[(size_t)set * yerr_stride + i];
            const double s = sqrt(e * e + BF_TINY);   // george folds the jitter into yerr
            v += s * s;
        }
        if (i >= n || j >= n) v = (i == j) ? 1.0 : 0.0;
        Kset[(size_t)i * np + j] = v;
    }
}

static int kernel_matrix(int kern, int n, int d, int S, const double* X, const double* inv_l2,
                         const double* amp, const int32_t* set_hp, const double* yerr,
                         int yerr_n, long long yerr_stride, double* K, int lower_only, void* stream);

extern "C" int bf_kernel_matrix(int kern, int n, int d, int S, const double* X, const double* inv_l2,
                                const double* amp, const int32_t* set_hp, const double* yerr,
                                int yerr_n, long long yerr_stride, double* K, void* stream) {
    return kernel_matrix(kern, n, d, S, X, inv_l2, amp, set_hp, yerr, yerr_n, yerr_stride, K, 0, stream);
}

extern "C" int bf_kernel_matrix_lower(int kern, int n, int d, int S, const double* X, const double* inv_l2,
                                      const double* amp, const int32_t* set_hp, const double* yerr,
                                      int yerr_n, long long yerr_stride, double* K, void* stream) {
    return kernel_matrix(kern, n, d, S, X, inv_l2, amp, set_hp, yerr, yerr_n, yerr_stride, K, 1, stream);
}

static int kernel_matrix(int kern, int n, int d, int S, const double* X, const double* inv_l2,
                         const double* amp, const int32_t* set_hp, const double* yerr,
                         int yerr_n, long long yerr_stride, double* K, int lower_only, void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_kernel_matrix: bad kern %d", kern);
    BF_CHECK_ARG(n > 0 && S >= 0 && X && inv_l2 && amp && yerr && K, BF_ERR_ARG, "bf_kernel_matrix: bad argument");
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_kernel_matrix: nDim %d outside 1..%d", d, BF_MAX_DIM);
    BF_CHECK_ARG(yerr_n == 1 || yerr_n == n, BF_ERR_ARG, "bf_kernel_matrix: yerr_n must be 1 or n");
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_kernel_matrix: at most 65535 sets per call");
    if (S == 0) return 0;
    const int np = bf_padded_n(n);
    dim3 grid(np / 32, np / 32, S);
    cudaStream_t st = (cudaStream_t)stream;
#define BF_KM_LAUNCH(KK, DD)                                                                                   \
    bf_kernel_matrix_kernel<KK, DD><<<grid, 256, 0, st>>>(n, d, np, X, inv_l2, amp, set_hp, yerr, yerr_n,       \
                                                          yerr_stride, K, lower_only)
#define BF_KM_DIM(KK)                                  \
    switch (d) {                                       \
        case 2: BF_KM_LAUNCH(KK, 2); break;            \
        case 3: BF_KM_LAUNCH(KK, 3); break;            \
        case 4: BF_KM_LAUNCH(KK, 4); break;            \
        case 6: BF_KM_LAUNCH(KK, 6); break;            \
        case 8: BF_KM_LAUNCH(KK, 8); break;            \
        case 10: BF_KM_LAUNCH(KK, 10); break;          \
        default: BF_KM_LAUNCH(KK, 0); break;           \
    }
    if (kern == BF_KERN_SE) { BF_KM_DIM(BF_KERN_SE) }
    else if (kern == BF_KERN_M32) { BF_KM_DIM(BF_KERN_M32) }
    else { BF_KM_DIM(BF_KERN_M52) }
#undef BF_KM_DIM
#undef BF_KM_LAUNCH
    BF_CHECK_LAUNCH("bf_kernel_matrix");
    return 0;
}

// ---- 32 x 32 tile product on the tensor cores -------------------------------------------------
// c[4][4][2] += Arows(32 x 4k) * Brows(32 x 4k)^T for k-steps [0, ksteps): both operands are
// row-major with leading dimension ld, fragment element (row = lane >> 2, col = lane & 3).
__device__ __forceinline__ void bf_tile_abt(double (&c)[4][4][2], const double* __restrict__ Arow,
                                            const double* __restrict__ Brow, int ld, int ksteps) {
    const int lane = threadIdx.x & 31;
    const double* ap = Arow + (size_t)(lane >> 2) * ld + (lane & 3);
    const double* bp = Brow + (size_t)(lane >> 2) * ld + (lane & 3);
    for (int ks = 0; ks < ksteps; ++ks) {
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a[t] = ap[(size_t)(8 * t) * ld + 4 * ks];
            b[t] = bp[(size_t)(8 * t) * ld + 4 * ks];
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], a[t], b[u]);
    }
}

// ---- panel-in-shared-memory variants (n <= BF_SETUP_RL_MAX_NP): 8 warps, one CTA per matrix ------
// The left-looking kernels below read both DMMA operands from L2 inside every k-step, which makes a
// single 500 x 500 factorisation latency-bound (1.7 ms + 0.5 ms).  Here the operand every tile of a
// step shares is staged in shared memory once per step:
//   Cholesky (right-looking): the freshly solved panel column P = A[k+1:, k] L_kk^-T; the trailing
//   update A[i, j] -= P_i P_j^T then runs with both operands in shared memory, up to 120 tiles
//   spread over the warps;
//   inverse (row by row): W = L_II^-1 L[I, 0:I], after which L^-1[I, J] = - sum_K W_IK L^-1[K, J].
#ifdef BF_PROFILE_SETUP
__device__ long long bf_prof_cycles[8];
#define BF_PROF_MARK(slot)                                                        \
    do {                                                                          \
        __syncthreads();                                                          \
        if (threadIdx.x == 0) {                                                   \
            const long long now_ = clock64();                                     \
            atomicAdd((unsigned long long*)&bf_prof_cycles[slot], (unsigned long long)(now_ - prof_t_)); \
            prof_t_ = now_;                                                       \
        }                                                                         \
    } while (0)
#else
#define BF_PROF_MARK(slot)
#endif
#define BF_RL_WARPS 8
#define BF_RL_THREADS (BF_RL_WARPS * 32)
#define BF_RL_PLD 36                    // panel leading dimension (doubles): conflict-free fragment reads
#define BF_SETUP_RL_MAX_NP 736

// c += A(32 x 4 ksteps, row-major, lda) * op(B); BKM = false: op(B) = B^T with B row-major 32 x K (ldb);
// BKM = true: op(B) = B with B row-major K x 32 (ldb).  One k-step of operands is fetched ahead.
template <bool BKM>
__device__ __forceinline__ void rl_tile(double (&c)[4][4][2], const double* __restrict__ A, int lda,
                                        const double* __restrict__ B, int ldb, int ksteps) {
    // ksteps is a multiple of 4; operands of the NEXT group of 4 k-steps are in flight while the
    // current group is multiplied (64 DMMAs cover an L2 round trip)
    const int lane = threadIdx.x & 31, r = lane >> 2, kk = lane & 3;
    const double* ap = A + (size_t)r * lda + kk;
    const double* bp = BKM ? B + (size_t)kk * ldb + r : B + (size_t)r * ldb + kk;
    double a[4][4], b[4][4];
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a[h][t] = ap[(size_t)(8 * t) * lda + 4 * h];
            b[h][t] = BKM ? bp[(size_t)(4 * h) * ldb + 8 * t] : bp[(size_t)(8 * t) * ldb + 4 * h];
        }
    for (int ks = 0; ks < ksteps; ks += 4) {
        double an[4][4], bn[4][4];
        if (ks + 4 < ksteps) {
#pragma unroll
            for (int h = 0; h < 4; ++h)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    an[h][t] = ap[(size_t)(8 * t) * lda + 4 * (ks + 4 + h)];
                    bn[h][t] = BKM ? bp[(size_t)(4 * (ks + 4 + h)) * ldb + 8 * t]
                                   : bp[(size_t)(8 * t) * ldb + 4 * (ks + 4 + h)];
                }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], a[h][t], b[h][u]);
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                a[h][t] = an[h][t];
                b[h][t] = bn[h][t];
            }
    }
}

__device__ __forceinline__ void rl_zero(double (&c)[4][4][2]) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
}

// c += A * B^T with both operands in shared memory (32 x 4 ksteps each, leading dimension ld)
__device__ __forceinline__ void rl_tile_smem(double (&c)[4][4][2], const double* __restrict__ A,
                                             const double* __restrict__ B, int ld, int ksteps) {
    const int lane = threadIdx.x & 31, r = lane >> 2, kk = lane & 3;
    const double* ap = A + (size_t)r * ld + kk;
    const double* bp = B + (size_t)r * ld + kk;
#pragma unroll 2
    for (int ks = 0; ks < ksteps; ++ks) {
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a[t] = ap[(size_t)(8 * t) * ld + 4 * ks];
            b[t] = bp[(size_t)(8 * t) * ld + 4 * ks];
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], a[t], b[u]);
    }
}

__global__ void __launch_bounds__(BF_RL_THREADS, 1) bf_cholesky_rl_kernel(int n, int np, double* __restrict__ Aall,
                                                                          int32_t* __restrict__ info,
                                                                          double* __restrict__ dinv_all) {
    extern __shared__ __align__(16) double rl_sm[];
    double* panel = rl_sm;                                          // (np - 32) x BF_RL_PLD
    double (*Ld)[33] = reinterpret_cast<double (*)[33]>(panel + (size_t)(np - 32) * BF_RL_PLD);
    double (*Li)[33] = Ld + 32;
    __shared__ int bad;
    __shared__ double rdiag[32];
    double* A = Aall + (size_t)blockIdx.x * np * np;
    const int nb = np / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
#ifdef BF_PROFILE_SETUP
    long long prof_t_ = clock64();
#endif
    // Diagonal block kb sits in Ld: factor it, write L_kk, invert it into Li (warp 0 only).
    auto factor_diag = [&](int kb) {
        const int r0 = 32 * kb;
        const int b = rl_factor_block_blocked(Ld, rdiag, lane);
        if (b && lane == 0 && bad == 0) bad = r0 + b;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) A[(size_t)(r0 + j) * np + r0 + lane] = Ld[j][lane];
        rl_invert_block_blocked(Ld, rdiag, Li, lane);
        if (dinv_all) {                 // inverses of the diagonal blocks, row-major 32 x 32 each (solve path)
            double* dv = dinv_all + ((size_t)blockIdx.x * (np / 32) + kb) * 1024;
#pragma unroll 8
            for (int j = 0; j < 32; ++j) dv[j * 32 + lane] = Li[j][lane];     // each lane wrote this column itself
        }
    };
    if (warp == 0) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) Ld[j][lane] = A[(size_t)j * np + lane];                   // coalesced rows
        __syncwarp();
        factor_diag(0);
    }
    __syncthreads();
    BF_PROF_MARK(0);
    for (int kb = 0; kb < nb - 1; ++kb) {
        const int r0 = 32 * kb;
        // 1. panel column: P_i = A[i, k] L_kk^-T for the row blocks below, kept in shared memory
        for (int ib = kb + 1 + warp; ib < nb; ib += BF_RL_WARPS) {
            double c[4][4][2];
            rl_zero(c);
            rl_tile<false>(c, A + (size_t)(32 * ib) * np + r0, np, &Li[0][0], 33, 8);
            double* prow = panel + (size_t)(32 * (ib - kb - 1)) * BF_RL_PLD;
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rr = 8 * t + (lane >> 2), cc = 8 * u + 2 * (lane & 3);
                    const double2 v = make_double2(c[t][u][0], c[t][u][1]);
                    *reinterpret_cast<double2*>(prow + (size_t)rr * BF_RL_PLD + cc) = v;
                    *reinterpret_cast<double2*>(A + (size_t)(32 * ib + rr) * np + r0 + cc) = v;
                }
        }
        __syncthreads();
        BF_PROF_MARK(1);
        // 2. trailing update: A[i, j] -= P_i P_j^T for kb < j <= i (both operands in shared memory; the
        //    old values are requested first so their L2 round trip overlaps the 64 DMMAs).  Look-ahead:
        //    warp 0 updates the next diagonal block, keeps it in shared memory and factors it while the
        //    other warps finish the rest of the update, so the serial 32 x 32 factorisation (a single-warp
        //    dependency chain) is off the critical path whenever the update takes longer.
        const int m = nb - kb - 1;
        const int ntiles = m * (m + 1) / 2;
        const int q_first = warp == 0 ? 0 : warp, q_step = warp == 0 ? ntiles : BF_RL_WARPS - 1;
        for (int q = q_first; q < ntiles; q += q_step) {
            int ii = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
            while (ii * (ii + 1) / 2 > q) --ii;
            while ((ii + 1) * (ii + 2) / 2 <= q) ++ii;
            const int jj = q - ii * (ii + 1) / 2;
            double* dst0 = A + (size_t)(32 * (kb + 1 + ii) + (lane >> 2)) * np + 32 * (kb + 1 + jj) + 2 * (lane & 3);
            double2 old[4][4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) old[t][u] = *reinterpret_cast<double2*>(dst0 + (size_t)(8 * t) * np + 8 * u);
            double c[4][4][2];
            rl_zero(c);
            rl_tile_smem(c, panel + (size_t)(32 * ii) * BF_RL_PLD, panel + (size_t)(32 * jj) * BF_RL_PLD, BF_RL_PLD, 8);
            if (q == 0) {                       // warp 0: the next diagonal block goes straight to the factorisation
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int rr = 8 * t + (lane >> 2), cc = 8 * u + 2 * (lane & 3);
                        Ld[rr][cc] = old[t][u].x - c[t][u][0];
                        Ld[rr][cc + 1] = old[t][u].y - c[t][u][1];
                    }
                __syncwarp();
                factor_diag(kb + 1);
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double2 v = old[t][u];
                        v.x -= c[t][u][0];
                        v.y -= c[t][u][1];
                        *reinterpret_cast<double2*>(dst0 + (size_t)(8 * t) * np + 8 * u) = v;
                    }
            }
        }
        __syncthreads();
        BF_PROF_MARK(2);
    }
    // the strictly upper 32 x 32 blocks still hold K: clear them so the output is L
    for (int e = threadIdx.x; e < np * (np / 2); e += BF_RL_THREADS) {
        const int i = e / (np / 2), j2 = (e - i * (np / 2)) * 2;
        if ((j2 >> 5) > (i >> 5)) *reinterpret_cast<double2*>(A + (size_t)i * np + j2) = make_double2(0.0, 0.0);
    }
    if (threadIdx.x == 0 && info) info[blockIdx.x] = bad;
}

__global__ void __launch_bounds__(BF_RL_THREADS, 1) bf_factor_finish_rl_kernel(
    int n, int np, const double* __restrict__ Lall, const double* __restrict__ y, long long y_stride,
    double* __restrict__ Linv_all, double* __restrict__ packed_all, size_t packed_stride,
    double* __restrict__ alpha_all) {
    extern __shared__ __align__(16) double rl_sm[];
    const int wld = np + 4;                                         // W row panel leading dimension
    double* tvec = rl_sm;                                           // np
    double* W = rl_sm + np;                                         // 32 x (np + 4)       (phase 2)
    double (*Ld)[33] = reinterpret_cast<double (*)[33]>(W);         // per-warp 2 x 32 x 33 (phase 1, aliased)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int set = blockIdx.x;
    const double* L = Lall + (size_t)set * np * np;
    double* Linv = Linv_all + (size_t)set * np * np;
    double* packed = packed_all + (size_t)set * packed_stride;
    const int nb = np / 32;
#ifdef BF_PROFILE_SETUP
    long long prof_t_ = clock64();
#endif

    // phase 1: inverses of the diagonal blocks; zero the blocks right of the diagonal
    for (int kb = warp; kb < nb; kb += BF_RL_WARPS) {
        const int r0 = 32 * kb;
        double (*Lw)[33] = Ld + (size_t)warp * 64;
        double (*Lo)[33] = Lw + 32;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) Lw[j][lane] = L[(size_t)(r0 + j) * np + r0 + lane];       // coalesced rows
        __syncwarp();
        rl_invert_block(Lw, nullptr, Lo, lane);
        __syncwarp();
        for (int i = 0; i < 32; ++i) {
            const double v = Lo[i][lane];
            Linv[(size_t)(r0 + i) * np + r0 + lane] = v;
            bf_pack_store(packed, r0 + i, r0 + lane, v, n, np);
        }
        for (int jb = kb + 1; jb < nb; ++jb)
            for (int i = 0; i < 32; ++i) Linv[(size_t)(r0 + i) * np + 32 * jb + lane] = 0.0;
    }
    __syncthreads();
    BF_PROF_MARK(3);

    // phase 2: block row I of the inverse
    for (int I = 1; I < nb; ++I) {
        // W[:, K] = Linv_II * L[I, K] for K < I (B operand k-major, straight from the row-major L)
        for (int K = warp; K < I; K += BF_RL_WARPS) {
            double c[4][4][2];
            rl_zero(c);
            rl_tile<true>(c, Linv + (size_t)(32 * I) * np + 32 * I, np, L + (size_t)(32 * I) * np + 32 * K, np, 8);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int rr = 8 * t + (lane >> 2), cc = 32 * K + 8 * u + 2 * (lane & 3);
                    *reinterpret_cast<double2*>(W + (size_t)rr * wld + cc) = make_double2(c[t][u][0], c[t][u][1]);
                }
        }
        __syncthreads();
        BF_PROF_MARK(4);
        // Linv[I, J] = - sum_{K = J}^{I - 1} W[:, K] Linv[K, J]
        for (int J = warp; J < I; J += BF_RL_WARPS) {
            double c[4][4][2];
            rl_zero(c);
            rl_tile<true>(c, W + 32 * J, wld, Linv + (size_t)(32 * J) * np + 32 * J, np, 8 * (I - J));
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = 32 * I + 8 * t + (lane >> 2), j = 32 * J + 8 * u + 2 * (lane & 3);
                    const double2 v = make_double2(-c[t][u][0], -c[t][u][1]);
                    *reinterpret_cast<double2*>(Linv + (size_t)i * np + j) = v;
                    bf_pack_store(packed, i, j, v.x, n, np);
                    bf_pack_store(packed, i, j + 1, v.y, n, np);
                }
        }
        __syncthreads();
        BF_PROF_MARK(5);
    }

    // phase 3: alpha = Linv^T (Linv y)
    if (alpha_all) {
        const double* yv = y + (size_t)set * y_stride;
        for (int i = warp; i < np; i += BF_RL_WARPS) {       // warp per row: coalesced, fixed-order lane tree
            double s = 0.0;
            if (i < n) {
                const double* row = Linv + (size_t)i * np;
                for (int j = lane; j <= i; j += 32) s = fma(row[j], yv[j], s);
            }
#pragma unroll
            for (int mm = 16; mm > 0; mm >>= 1) s += __shfl_xor_sync(0xffffffffu, s, mm);
            if (lane == 0) tvec[i] = s;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += BF_RL_THREADS) {
            double s = 0.0;
            for (int i = j; i < n; ++i) s = fma(Linv[(size_t)i * np + j], tvec[i], s);
            alpha_all[(size_t)set * n + j] = s;
        }
    }
    BF_PROF_MARK(6);
}

// ---- L^-1 by independent column groups (every column of the inverse is its own forward substitution) ----
// One CTA per (block column J, pair of 8-column groups): W warps, each owning 8 columns of L^-1 for all
// row blocks I >= J.  x_J = Linv_JJ[:, cols]; x_I = -Linv_II (sum_{K=J}^{I-1} L[I,K] x_K).  The A tiles
// (L[I,K], then Linv_II) are the same for every warp of the CTA and stream from L2 through a cp.async ring;
// the x_K history stays in shared memory in DMMA B-fragment order.  4 nb CTAs per set instead of one, so a
// single 500 x 500 inverse takes tens of microseconds instead of the 0.55 ms one SM needs for n^3/3 flop.
#define BF_COLS_RING 8
#define BF_COLS_TLD 36


// inverses of the diagonal blocks (warp 0) and zeros right of the diagonal (the other warps): grid (nb, S)
__global__ void __launch_bounds__(128) bf_diag_inverse_kernel(int n, int np, const double* __restrict__ Lall,
                                                              double* __restrict__ Linv_all,
                                                              double* __restrict__ packed_all, size_t packed_stride) {
    __shared__ double Lw[32][33];
    __shared__ double Lo[32][33];
    const int kb = blockIdx.x, set = blockIdx.y, nb = np / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* L = Lall + (size_t)set * np * np;
    double* Linv = Linv_all + (size_t)set * np * np;
    double* packed = packed_all + (size_t)set * packed_stride;
    const int r0 = 32 * kb;
    if (warp == 0) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) Lw[j][lane] = L[(size_t)(r0 + j) * np + r0 + lane];
        __syncwarp();
        rl_invert_block(Lw, nullptr, Lo, lane);
        __syncwarp();
        for (int i = 0; i < 32; ++i) {
            const double v = Lo[i][lane];
            Linv[(size_t)(r0 + i) * np + r0 + lane] = v;
            bf_pack_store(packed, r0 + i, r0 + lane, v, n, np);
        }
    } else {
        const int width = np - r0 - 32;                 // doubles right of the diagonal block, per row
        for (int e = threadIdx.x - 32; e < 32 * (width / 2); e += 96) {
            const int i = e / (width / 2), j2 = (e - i * (width / 2)) * 2;
            *reinterpret_cast<double2*>(Linv + (size_t)(r0 + i) * np + r0 + 32 + j2) = make_double2(0.0, 0.0);
        }
    }
}

// grid (4 nb, S): CTA = (block column J, 8-column group a); warp w owns rows 8w..8w+7 of every row block
__global__ void __launch_bounds__(128) bf_linv_cols_kernel(int n, int np, const double* __restrict__ Lall,
                                                          double* __restrict__ Linv_all,
                                                          double* __restrict__ packed_all, size_t packed_stride) {
    extern __shared__ __align__(16) double cs_sm[];
    constexpr int TLD = BF_COLS_TLD, RING = BF_COLS_RING;
    const int nb = np / 32;
    const int J = blockIdx.x >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, kk = lane & 3;
    const int c0 = 32 * J + 8 * (blockIdx.x & 3);              // this CTA's 8 columns
    const int set = blockIdx.y;
    const double* L = Lall + (size_t)set * np * np;
    double* Linv = Linv_all + (size_t)set * np * np;
    double* packed = packed_all + (size_t)set * packed_stride;
    double* ring = cs_sm;                                       // RING x 32 x TLD
    double* xh = ring + RING * 32 * TLD;                        // nb blocks of 32 x 8 (B-fragment order), + scratch
    double* sbuf = xh + (size_t)nb * 256;

    const int rows_below = nb - 1 - J;
    const int M = rows_below * (rows_below + 3) / 2;            // sum_{I = J+1}^{nb-1} (I - J + 1) tiles
    if (M == 0) return;
    // x_J: these columns of the diagonal-block inverse (written by bf_diag_inverse_kernel)
    for (int e = threadIdx.x; e < 256; e += 128) xh[e] = Linv[(size_t)(32 * J + (e >> 3)) * np + c0 + (e & 7)];

    int pI = J + 1, pK = J;                                     // next tile to request; K == I means Linv_II
    auto request = [&](int slot) {
        const double* src = (pK < pI ? L : Linv) + (size_t)(32 * pI) * np + 32 * pK;
        double* dst = ring + (size_t)slot * 32 * TLD;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = threadIdx.x + 128 * q, row = e >> 4, ch = e & 15;
            bf_cp_async16(dst + row * TLD + 2 * ch, src + (size_t)row * np + 2 * ch);
        }
        if (pK < pI) ++pK; else { ++pI; pK = J; }
    };
    for (int t = 0; t < RING - 1; ++t) {
        if (t < M) request(t);
        bf_cp_async_commit();
    }
    double acc[8][2];                                           // one chain per k-step: a tile adds one DMMA to each
#pragma unroll
    for (int h = 0; h < 8; ++h) acc[h][0] = acc[h][1] = 0.0;
    auto acc_sum = [&](int q) {
        return ((acc[0][q] + acc[1][q]) + (acc[2][q] + acc[3][q])) + ((acc[4][q] + acc[5][q]) + (acc[6][q] + acc[7][q]));
    };
    int I = J + 1, K = J;
    for (int t = 0; t < M; ++t) {
        bf_cp_async_wait<RING - 2>();
        __syncthreads();                                        // tile t and x_{I-1} visible; slot (t - 1) % RING free
        if (t + RING - 1 < M) request((t + RING - 1) % RING);
        bf_cp_async_commit();
        const double* A = ring + (size_t)(t % RING) * 32 * TLD + (8 * warp + r) * TLD + kk;
        if (K < I) {
            const double* B = xh + (size_t)(K - J) * 256 + kk * 8 + r;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) bf_dmma(acc[ks][0], acc[ks][1], A[4 * ks], B[ks * 32]);
            ++K;
        } else {
            // s = sum_K L[I,K] x_K is complete: x_I = -Linv_II s
            *reinterpret_cast<double2*>(sbuf + (8 * warp + r) * 8 + 2 * kk) = make_double2(acc_sum(0), acc_sum(1));
#pragma unroll
            for (int h = 0; h < 8; ++h) acc[h][0] = acc[h][1] = 0.0;
            __syncthreads();
            const double* B = sbuf + kk * 8 + r;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) bf_dmma(acc[ks][0], acc[ks][1], A[4 * ks], B[ks * 32]);
            const double2 v = make_double2(-acc_sum(0), -acc_sum(1));
#pragma unroll
            for (int h = 0; h < 8; ++h) acc[h][0] = acc[h][1] = 0.0;
            const int i = 32 * I + 8 * warp + r, j = c0 + 2 * kk;
            *reinterpret_cast<double2*>(xh + (size_t)(I - J) * 256 + (8 * warp + r) * 8 + 2 * kk) = v;
            *reinterpret_cast<double2*>(Linv + (size_t)i * np + j) = v;
            bf_pack_store(packed, i, j, v.x, n, np);
            bf_pack_store(packed, i, j + 1, v.y, n, np);
            ++I;
            K = J;
        }
    }
}

// alpha = Linv^T (Linv y): one CTA of 32 warps per set.  t = Linv y: warp per row.  alpha_j = sum_i Linv[i][j] t_i:
// warp w sums the rows i = w (mod 32) for all columns (lane = column within a 32-column block), then the 32
// partial rows are added in warp order.
#define BF_ALPHA_MAX_NB 23              // np <= BF_SETUP_RL_MAX_NP
#define BF_ALPHA_WARPS 16
__global__ void __launch_bounds__(32 * BF_ALPHA_WARPS) bf_alpha_kernel(int n, int np, const double* __restrict__ Linv_all,
                                                                       const double* __restrict__ y, long long y_stride,
                                                                       double* __restrict__ alpha_all) {
    extern __shared__ __align__(16) double al_sm[];
    double* tvec = al_sm;                                       // np
    double* part = al_sm + np;                                  // BF_ALPHA_WARPS x np
    const int set = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = np / 32;
    const double* Linv = Linv_all + (size_t)set * np * np;
    const double* yv = y + (size_t)set * y_stride;
    double* ys = part;                                          // y staged in shared memory (part is free until phase 2)
    for (int j = threadIdx.x; j < np; j += 32 * BF_ALPHA_WARPS) ys[j] = j < n ? yv[j] : 0.0;
    __syncthreads();
    for (int i = warp; i < np; i += BF_ALPHA_WARPS) {
        // all loads of the row are issued before the (order-preserving) FMA chain
        double v[BF_ALPHA_MAX_NB];
        const double* row = Linv + (size_t)i * np + lane;
#pragma unroll
        for (int cb = 0; cb < BF_ALPHA_MAX_NB; ++cb) v[cb] = (i < n && 32 * cb + lane <= i) ? row[32 * cb] : 0.0;
        double s = 0.0;
#pragma unroll
        for (int cb = 0; cb < BF_ALPHA_MAX_NB; ++cb)
            if (32 * cb + lane <= i) s = fma(v[cb], ys[32 * cb + lane], s);
#pragma unroll
        for (int mm = 16; mm > 0; mm >>= 1) s += __shfl_xor_sync(0xffffffffu, s, mm);
        if (lane == 0) tvec[i] = s;
    }
    __syncthreads();
    double acc[BF_ALPHA_MAX_NB];
#pragma unroll
    for (int cb = 0; cb < BF_ALPHA_MAX_NB; ++cb) acc[cb] = 0.0;
    for (int i = warp; i < np; i += BF_ALPHA_WARPS) {
        const int ib = i >> 5;
        const double ti = (i < n) ? tvec[i] : 0.0;
        const double* row = Linv + (size_t)i * np + lane;
#pragma unroll
        for (int cb = 0; cb < BF_ALPHA_MAX_NB; ++cb)
            if (cb <= ib) acc[cb] = fma(row[32 * cb], ti, acc[cb]);     // zeros above the diagonal inside block ib
    }
#pragma unroll
    for (int cb = 0; cb < BF_ALPHA_MAX_NB; ++cb)
        if (cb < nb) part[(size_t)warp * np + 32 * cb + lane] = acc[cb];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += 32 * BF_ALPHA_WARPS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < BF_ALPHA_WARPS; ++w) s += part[(size_t)w * np + j];
        alpha_all[(size_t)set * n + j] = s;
    }
}

static size_t linv_cols_smem_bytes(int np) {
    const int nb = np / 32;
    return sizeof(double) * ((size_t)BF_COLS_RING * 32 * BF_COLS_TLD + (size_t)(nb + 1) * 256);
}

// ---- batched Cholesky (left-looking, 32-wide panels, one CTA per matrix) ----------------------
__global__ void __launch_bounds__(BF_THREADS) bf_cholesky_kernel(int n, int np, double* __restrict__ Aall,
                                                                 int32_t* __restrict__ info) {
    __shared__ double Ld[32][33];
    __shared__ int bad;
    double* A = Aall + (size_t)blockIdx.x * np * np;
    const int nb = np / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int kb = 0; kb < nb; ++kb) {
        const int r0 = 32 * kb;
        // 1. panel update with the columns already factored: P -= A[:, :r0] A[r0:r0+32, :r0]^T
        if (kb > 0) {
            for (int rb = kb + warp; rb < nb; rb += BF_WARPS) {
                double c[4][4][2];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
                bf_tile_abt(c, A + (size_t)(32 * rb) * np, A + (size_t)r0 * np, np, r0 / 4);
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double2* dst = reinterpret_cast<double2*>(
                            A + (size_t)(32 * rb + 8 * t + (lane >> 2)) * np + r0 + 8 * u + 2 * (lane & 3));
                        double2 v = *dst;
                        v.x -= c[t][u][0];
                        v.y -= c[t][u][1];
                        *dst = v;
                    }
            }
        }
        __syncthreads();
        // 2. factor the diagonal block (warp 0, lane = row)
        if (warp == 0) {
            for (int j = 0; j < 32; ++j) Ld[lane][j] = A[(size_t)(r0 + lane) * np + r0 + j];
            __syncwarp();
            for (int j = 0; j < 32; ++j) {
                const double djj = Ld[j][j];
                if (!(djj > 0.0) && lane == 0 && bad == 0) bad = r0 + j + 1;
                const double dj = sqrt(djj);
                __syncwarp();
                if (lane == j) Ld[j][j] = dj;
                if (lane > j) Ld[lane][j] = Ld[lane][j] / dj;
                __syncwarp();
                if (lane > j) {
                    const double lij = Ld[lane][j];
                    for (int cc = j + 1; cc <= lane; ++cc) Ld[lane][cc] = fma(-lij, Ld[cc][j], Ld[lane][cc]);
                }
                __syncwarp();
            }
            for (int j = 0; j < 32; ++j) A[(size_t)(r0 + lane) * np + r0 + j] = (j <= lane) ? Ld[lane][j] : 0.0;
        }
        __syncthreads();
        // 3. rows below the block: X L_kk^T = P, one row per thread
        for (int i = r0 + 32 + threadIdx.x; i < np; i += BF_THREADS) {
            double x[32];
            double* row = A + (size_t)i * np + r0;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) x[cc] = row[cc];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) {
                double s = x[cc];
#pragma unroll
                for (int t = 0; t < cc; ++t) s = fma(-x[t], Ld[cc][t], s);
                x[cc] = s / Ld[cc][cc];
            }
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) row[cc] = x[cc];
        }
        __syncthreads();
    }
    // the strictly upper 32 x 32 blocks still hold K: clear them so the output is L
    for (int e = threadIdx.x; e < np * (np / 2); e += BF_THREADS) {
        const int i = e / (np / 2), j2 = (e - i * (np / 2)) * 2;
        if ((j2 >> 5) > (i >> 5)) *reinterpret_cast<double2*>(A + (size_t)i * np + j2) = make_double2(0.0, 0.0);
    }
    if (threadIdx.x == 0 && info) info[blockIdx.x] = bad;
}

extern "C" int bf_cholesky_dinv_supported(int n) { return bf_padded_n(n) <= BF_SETUP_RL_MAX_NP ? 1 : 0; }

static int cholesky_batched(int n, int S, double* K_inout_L, int32_t* info, double* dinv, void* stream);

extern "C" int bf_cholesky_batched(int n, int S, double* K_inout_L, int32_t* info, void* stream) {
    return cholesky_batched(n, S, K_inout_L, info, nullptr, stream);
}

extern "C" int bf_cholesky_batched_dinv(int n, int S, double* K_inout_L, int32_t* info, double* dinv, void* stream) {
    BF_CHECK_ARG(dinv, BF_ERR_ARG, "bf_cholesky_batched_dinv: null dinv");
    BF_CHECK_ARG(bf_cholesky_dinv_supported(n), BF_ERR_SIZE, "bf_cholesky_batched_dinv: n=%d too large (padded size above %d)",
                 n, BF_SETUP_RL_MAX_NP);
    return cholesky_batched(n, S, K_inout_L, info, dinv, stream);
}

static int cholesky_batched(int n, int S, double* K_inout_L, int32_t* info, double* dinv, void* stream) {
    BF_CHECK_ARG(n > 0 && S >= 0 && K_inout_L, BF_ERR_ARG, "bf_cholesky_batched: bad argument");
    if (S == 0) return 0;
    const int np = bf_padded_n(n);
    if (np <= BF_SETUP_RL_MAX_NP) {
        const size_t smem = sizeof(double) * ((size_t)(np - 32) * BF_RL_PLD + 2 * 32 * 33);
        cudaError_t e = cudaFuncSetAttribute(bf_cholesky_rl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            bf_set_error("bf_cholesky_batched: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        bf_cholesky_rl_kernel<<<S, BF_RL_THREADS, smem, (cudaStream_t)stream>>>(n, np, K_inout_L, info, dinv);
    } else {
        bf_cholesky_kernel<<<S, BF_THREADS, 0, (cudaStream_t)stream>>>(n, np, K_inout_L, info);
    }
    BF_CHECK_LAUNCH("bf_cholesky_batched");
    return 0;
}

// ---- L^-1, packing, alpha --------------------------------------------------------------------
__global__ void __launch_bounds__(BF_THREADS) bf_factor_finish_kernel(
    int n, int np, const double* __restrict__ Lall, const double* __restrict__ y, long long y_stride,
    double* __restrict__ Linv_all, double* __restrict__ packed_all, size_t packed_stride,
    double* __restrict__ alpha_all) {
    extern __shared__ __align__(16) double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* buf = sm + (size_t)warp * 32 * 33;           // per-warp 32 x 33 scratch
    double* tvec = sm + (size_t)BF_WARPS * 32 * 33;      // np doubles
    const int set = blockIdx.x;
    const double* L = Lall + (size_t)set * np * np;
    double* Linv = Linv_all + (size_t)set * np * np;
    double* packed = packed_all + (size_t)set * packed_stride;
    const int nb = np / 32;

    // phase 1: inverses of the diagonal blocks (lane = column of the inverse)
    for (int kb = warp; kb < nb; kb += BF_WARPS) {
        const int r0 = 32 * kb;
        for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = L[(size_t)(r0 + lane) * np + r0 + j];
        __syncwarp();
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int t = 0; t < i; ++t) s = fma(-buf[i * 33 + t], x[t], s);   // x[t] = 0 for t < lane
            x[i] = (i >= lane) ? s / buf[i * 33 + i] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            Linv[(size_t)(r0 + i) * np + r0 + lane] = x[i];
            bf_pack_store(packed, r0 + i, r0 + lane, x[i], n, np);
        }
        // zero the blocks to the right of the diagonal in the row-major copy
        for (int jb = kb + 1; jb < nb; ++jb)
            for (int i = 0; i < 32; ++i) Linv[(size_t)(r0 + i) * np + 32 * jb + lane] = 0.0;
        __syncwarp();
    }
    __syncthreads();

    // phase 2: block column J of the inverse, top to bottom; columns are independent, paired
    // (J, nb-1-J) per warp for balance.  Linv_IJ = -Linv_II * sum_{K=J}^{I-1} L_IK Linv_KJ.
    const int npairs = (nb + 1) >> 1;
    for (int pr = warp; pr < npairs; pr += BF_WARPS) {
        for (int side = 0; side < 2; ++side) {
            const int J = side == 0 ? pr : nb - 1 - pr;
            if (side == 1 && J == pr) break;
            for (int I = J + 1; I < nb; ++I) {
                double c[4][4][2];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
                // T = L[I, J..I-1] * Linv[J..I-1, J]  (B operand is NOT transposed here)
                {
                    const double* ap = L + (size_t)(32 * I + (lane >> 2)) * np + 32 * J + (lane & 3);
                    const double* bp = Linv + (size_t)(32 * J + (lane & 3)) * np + 32 * J + (lane >> 2);
                    const int ksteps = 8 * (I - J);
                    for (int ks = 0; ks < ksteps; ++ks) {
                        double a[4], b[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            a[t] = ap[(size_t)(8 * t) * np + 4 * ks];
                            b[t] = bp[(size_t)(4 * ks) * np + 8 * t];
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], a[t], b[u]);
                    }
                }
                // stage T in shared memory (row-major 32 x 33) to use it as the next B operand
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = 8 * t + (lane >> 2), cc = 8 * u + 2 * (lane & 3);
                        buf[r * 33 + cc] = c[t][u][0];
                        buf[r * 33 + cc + 1] = c[t][u][1];
                        c[t][u][0] = c[t][u][1] = 0.0;
                    }
                __syncwarp();
                {
                    const double* ap = Linv + (size_t)(32 * I + (lane >> 2)) * np + 32 * I + (lane & 3);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        double a[4], b[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            a[t] = ap[(size_t)(8 * t) * np + 4 * ks];
                            b[t] = buf[(4 * ks + (lane & 3)) * 33 + 8 * t + (lane >> 2)];
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t)
#pragma unroll
                            for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], a[t], b[u]);
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = 32 * I + 8 * t + (lane >> 2), j = 32 * J + 8 * u + 2 * (lane & 3);
                        double2 v;
                        v.x = -c[t][u][0];
                        v.y = -c[t][u][1];
                        *reinterpret_cast<double2*>(Linv + (size_t)i * np + j) = v;
                        bf_pack_store(packed, i, j, v.x, n, np);
                        bf_pack_store(packed, i, j + 1, v.y, n, np);
                    }
                __syncwarp();
                __threadfence_block();
            }
        }
    }
    __syncthreads();

    // phase 3: alpha = Linv^T (Linv y)
    if (alpha_all) {
        const double* yv = y + (size_t)set * y_stride;
        for (int i = threadIdx.x; i < np; i += BF_THREADS) {
            double s = 0.0;
            if (i < n) {
                const double* row = Linv + (size_t)i * np;
                for (int j = 0; j <= i; ++j) s = fma(row[j], yv[j], s);
            }
            tvec[i] = s;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += BF_THREADS) {
            double s = 0.0;
            for (int i = j; i < n; ++i) s = fma(Linv[(size_t)i * np + j], tvec[i], s);
            alpha_all[(size_t)set * n + j] = s;
        }
    }
}

// Sets up to which bf_factor_finish uses the column-group kernels (BF_SETUP_COLS_MAX_SETS overrides; measurement aid)
static int bf_setup_cols_max_sets() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("BF_SETUP_COLS_MAX_SETS");
        v = e ? atoi(e) : 96;
    }
    return v;
}

extern "C" int bf_factor_finish_kernels(int n, int S, int with_alpha) {
    const int np = bf_padded_n(n);
    const int clear = np != n ? 1 : 0;                  // bf_packed_clear_kernel
    if (np <= BF_SETUP_RL_MAX_NP && S <= bf_setup_cols_max_sets()) return clear + (with_alpha ? 3 : 2);
    return clear + 1;
}

// Zero the cells of the packed operand that the shifted writes of the inverse never reach: in every row block
// the 8 k-groups of its diagonal block (the n x n part is written over them afterwards) and the leading k-groups
// that hold padding columns.  Grid (nb, S).
__global__ void __launch_bounds__(256) bf_packed_clear_kernel(int lead_groups, double* __restrict__ packed_all,
                                                             size_t packed_stride) {
    const int b = blockIdx.x;
    double* blk = packed_all + (size_t)blockIdx.y * packed_stride + bf_linv_block_off(b);
    const int lead = lead_groups < 8 * b ? lead_groups : 8 * b;     // row block 0 is all diagonal block
    for (int e = threadIdx.x; e < lead * 128; e += 256) blk[e] = 0.0;
    double* diag = blk + (size_t)8 * b * 128;
    for (int e = threadIdx.x; e < 1024; e += 256) diag[e] = 0.0;
}

extern "C" int bf_factor_finish(int n, int S, const double* L, const double* y, long long y_stride,
                                double* Linv, double* linv_packed, double* alpha, void* stream) {
    BF_CHECK_ARG(n > 0 && S >= 0 && L && Linv && linv_packed, BF_ERR_ARG, "bf_factor_finish: bad argument");
    BF_CHECK_ARG(!alpha || y, BF_ERR_ARG, "bf_factor_finish: alpha requested without y");
    if (S == 0) return 0;
    const int np = bf_padded_n(n);
    if (np != n) {
        // the packed operand holds L^-1 shifted to the END of the padded index range (bf_pack_store): the
        // kernels below write the n x n part; what they never touch - the leading rows and columns, the
        // triangles above the diagonal inside the shifted diagonal blocks - is cleared first
        bf_packed_clear_kernel<<<dim3(np / 32, S), 256, 0, (cudaStream_t)stream>>>(
            (np - n + 3) / 4, linv_packed, bf_linv_packed_doubles(n));
        BF_CHECK_LAUNCH("bf_factor_finish");
    }
    if (np <= BF_SETUP_RL_MAX_NP && S <= bf_setup_cols_max_sets()) {
        // few sets: spread every inverse over 4 nb CTAs
        const size_t smem_c = linv_cols_smem_bytes(np);
        const size_t smem_a = sizeof(double) * (size_t)(BF_ALPHA_WARPS + 1) * np;
        cudaError_t e3 = cudaFuncSetAttribute(bf_linv_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
        if (e3 == cudaSuccess)
            e3 = cudaFuncSetAttribute(bf_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
        if (e3 != cudaSuccess) {
            bf_set_error("bf_factor_finish: cudaFuncSetAttribute: %s", cudaGetErrorString(e3));
            return (int)e3;
        }
        cudaStream_t st = (cudaStream_t)stream;
        const int nb = np / 32;
        const size_t pstride = bf_linv_packed_doubles(n);
        bf_diag_inverse_kernel<<<dim3(nb, S), 128, 0, st>>>(n, np, L, Linv, linv_packed, pstride);
        bf_linv_cols_kernel<<<dim3(4 * nb, S), 128, smem_c, st>>>(n, np, L, Linv, linv_packed, pstride);
        if (alpha) bf_alpha_kernel<<<S, 32 * BF_ALPHA_WARPS, smem_a, st>>>(n, np, Linv, y, y_stride, alpha);
        BF_CHECK_LAUNCH("bf_factor_finish");
        return 0;
    }
    if (np <= BF_SETUP_RL_MAX_NP) {
        const size_t w_sz = (size_t)32 * (np + 4), scratch = (size_t)BF_RL_WARPS * 64 * 33;
        const size_t smem_rl = sizeof(double) * (np + (w_sz > scratch ? w_sz : scratch));
        cudaError_t e2 = cudaFuncSetAttribute(bf_factor_finish_rl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem_rl);
        if (e2 != cudaSuccess) {
            bf_set_error("bf_factor_finish: cudaFuncSetAttribute: %s", cudaGetErrorString(e2));
            return (int)e2;
        }
        bf_factor_finish_rl_kernel<<<S, BF_RL_THREADS, smem_rl, (cudaStream_t)stream>>>(
            n, np, L, y, y_stride, Linv, linv_packed, bf_linv_packed_doubles(n), alpha);
        BF_CHECK_LAUNCH("bf_factor_finish");
        return 0;
    }
    const size_t smem = sizeof(double) * ((size_t)BF_WARPS * 32 * 33 + np);
    cudaError_t e = cudaFuncSetAttribute(bf_factor_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_factor_finish: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    bf_factor_finish_kernel<<<S, BF_THREADS, smem, (cudaStream_t)stream>>>(
        n, np, L, y, y_stride, Linv, linv_packed, bf_linv_packed_doubles(n), alpha);
    BF_CHECK_LAUNCH("bf_factor_finish");
    return 0;
}
