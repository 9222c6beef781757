// Full posterior covariance between two point sets under one factorised GP:
//   C[a, b] = amp k(xa, xb) - sum_j V[j, a] V[j, b],   V = L^-1 amp k(X, .)
// Replaces gp_model.predict_cov (gpModel.py:44-48, george predict(..., return_cov=True)), which the
// batch-only path calls for every hyper-parameter set (util.py:952), and provides the cross term
// the fantasy-point update of the ROM stage needs.  Three steps, all 32 x 32 DMMA tiles:
//   1. Ks  [Ap x np] = amp k(Xa, X) (zero padded)  and  mean = Ks alpha
//   2. Vt  [Ap x np] = Ks L^-T   (row a of Vt is column a of V; L^-1 is lower triangular)
//   3. C   [A x B]   = amp k(Xa, Xb) - Vt_a Vt_b^T
#include "bf_common.cuh"

// ---- step 1 -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bf_cross_matrix_kernel(int kern, long long A, int n, int d, int np,
                                                              const double* __restrict__ Xa,
                                                              const double* __restrict__ X,
                                                              const double* __restrict__ inv_l2, double amp,
                                                              double* __restrict__ Ks) {
    __shared__ double xa_s[32][BF_MAX_DIM + 1];
    __shared__ double xj_s[32][BF_MAX_DIM + 1];
    __shared__ double w_s[BF_MAX_DIM];
    const long long a0 = (long long)blockIdx.y * 32;
    const int j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < d) w_s[threadIdx.x] = bf_metric_scale(kern, inv_l2[threadIdx.x]);
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * d; e += 256) {       // pre-scaled coordinates
        const int r = e / d, c = e - r * d;
        xa_s[r][c] = (a0 + r < A) ? Xa[(size_t)(a0 + r) * d + c] * w_s[c] : 0.0;
        xj_s[r][c] = (j0 + r < n) ? X[(size_t)(j0 + r) * d + c] * w_s[c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = ty + 8 * rr;
        const long long a = a0 + r;
        const int j = j0 + tx;
        double v = 0.0;
        if (a < A && j < n) {
            double q = 0.0;
            for (int c = 0; c < d; ++c) {
                const double df = xa_s[r][c] - xj_s[tx][c];
                q = fma(df, df, q);
            }
            v = amp * bf_radial_q(kern, q, bf_exp_tab_g);
        }
        Ks[(size_t)a * np + j] = v;
    }
}

// mean[a] = Ks[a, :] . alpha  (one warp per row, fixed-order lane tree)
__global__ void __launch_bounds__(256) bf_rows_dot_kernel(long long A, int n, int np, const double* __restrict__ Ks,
                                                          const double* __restrict__ alpha,
                                                          double* __restrict__ mean) {
    const long long a = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (a >= A) return;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(Ks[(size_t)a * np + j], alpha[j], s);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) mean[a] = s;
}

// ---- 32 x 32 tile: c += Arows (32 x 4 ksteps) * Brows^T, operands row-major with ld --------------
__device__ __forceinline__ void cov_tile_abt(double (&c)[4][4][2], const double* __restrict__ Arow,
                                             const double* __restrict__ Brow, int ld, int ksteps) {
    const int lane = threadIdx.x & 31;
    const double* ap = Arow + (size_t)(lane >> 2) * ld + (lane & 3);
    const double* bp = Brow + (size_t)(lane >> 2) * ld + (lane & 3);
    int ks = 0;
    for (; ks + 2 <= ksteps; ks += 2) {
        double a[2][4], b[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                a[h][t] = ap[(size_t)(8 * t) * ld + 4 * (ks + h)];
                b[h][t] = bp[(size_t)(8 * t) * ld + 4 * (ks + h)];
            }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], a[h][t], b[h][u]);
    }
    for (; ks < ksteps; ++ks) {
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

// ---- step 2: Vt[a-block, i-block] = Ks[a-block, 0 .. 32(ib+1)) * Linv[i-block, same]^T --------------
__global__ void __launch_bounds__(256) bf_solve_rows_kernel(int np, const double* __restrict__ Ks,
                                                            const double* __restrict__ Linv,
                                                            double* __restrict__ Vt) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = np / 32;
    const int ib = blockIdx.x * 8 + warp;
    if (ib >= nb) return;
    const size_t a0 = (size_t)blockIdx.y * 32;
    double c[4][4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
    cov_tile_abt(c, Ks + a0 * np, Linv + (size_t)(32 * ib) * np, np, 8 * (ib + 1));
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double2 v = make_double2(c[t][u][0], c[t][u][1]);
            *reinterpret_cast<double2*>(Vt + (a0 + 8 * t + (lane >> 2)) * np + 32 * ib + 8 * u + 2 * (lane & 3)) = v;
        }
}

// ---- step 3 ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bf_cov_kernel(int kern, long long A, long long B, int d, int np,
                                                     const double* __restrict__ Xa, const double* __restrict__ Xb,
                                                     const double* __restrict__ inv_l2, double amp,
                                                     const double* __restrict__ Vta, const double* __restrict__ Vtb,
                                                     double* __restrict__ C, long long ldc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bb = (long long)blockIdx.x * 8 + warp;
    const long long b0 = bb * 32, a0 = (long long)blockIdx.y * 32;
    if (b0 >= B) return;
    double c[4][4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
    cov_tile_abt(c, Vta + (size_t)a0 * np, Vtb + (size_t)b0 * np, np, np / 4);
    double w[BF_MAX_DIM];
    for (int k = 0; k < d; ++k) w[k] = bf_metric_scale(kern, inv_l2[k]);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const long long a = a0 + 8 * t + (lane >> 2);
        if (a >= A) continue;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const long long b = b0 + 8 * u + 2 * (lane & 3) + q;
                if (b >= B) continue;
                double qq = 0.0;
                for (int k = 0; k < d; ++k) {
                    const double df = Xa[(size_t)a * d + k] * w[k] - Xb[(size_t)b * d + k] * w[k];
                    qq = fma(df, df, qq);
                }
                C[(size_t)a * ldc + b] = amp * bf_radial_q(kern, qq, bf_exp_tab_g) - c[t][u][q];
            }
    }
}

extern "C" int bf_gp_solve_rows(int kern, long long A, int n, int d, const double* Xa, const double* X,
                                const double* inv_l2, double amp, const double* Linv, const double* alpha,
                                double* Ks, double* Vt, double* mean, void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_gp_solve_rows: bad kern %d", kern);
    BF_CHECK_ARG(A >= 0 && n > 0 && Xa && X && inv_l2 && Linv && Ks && Vt, BF_ERR_ARG, "bf_gp_solve_rows: bad argument");
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_gp_solve_rows: nDim %d outside 1..%d", d, BF_MAX_DIM);
    BF_CHECK_ARG(!mean || alpha, BF_ERR_ARG, "bf_gp_solve_rows: mean requested without alpha");
    if (A == 0) return 0;
    const int np = bf_padded_n(n);
    const long long Ab = (A + 31) / 32;
    BF_CHECK_ARG(Ab <= 65535, BF_ERR_SIZE, "bf_gp_solve_rows: at most 2,097,120 rows per call");
    cudaStream_t st = (cudaStream_t)stream;
    bf_cross_matrix_kernel<<<dim3(np / 32, (unsigned)Ab), 256, 0, st>>>(kern, A, n, d, np, Xa, X, inv_l2, amp, Ks);
    BF_CHECK_LAUNCH("bf_gp_solve_rows(cross)");
    if (mean) {
        bf_rows_dot_kernel<<<(unsigned)((A + 7) / 8), 256, 0, st>>>(A, n, np, Ks, alpha, mean);
        BF_CHECK_LAUNCH("bf_gp_solve_rows(mean)");
    }
    bf_solve_rows_kernel<<<dim3((np / 32 + 7) / 8, (unsigned)Ab), 256, 0, st>>>(np, Ks, Linv, Vt);
    BF_CHECK_LAUNCH("bf_gp_solve_rows(solve)");
    return 0;
}

extern "C" int bf_gp_cov(int kern, long long A, long long B, int n, int d, const double* Xa, const double* Xb,
                         const double* inv_l2, double amp, const double* Vta, const double* Vtb, double* C,
                         long long ldc, void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_gp_cov: bad kern %d", kern);
    BF_CHECK_ARG(A >= 0 && B >= 0 && n > 0 && Xa && Xb && inv_l2 && Vta && Vtb && C && ldc >= B, BF_ERR_ARG,
                 "bf_gp_cov: bad argument");
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_gp_cov: nDim %d outside 1..%d", d, BF_MAX_DIM);
    if (A == 0 || B == 0) return 0;
    const int np = bf_padded_n(n);
    const long long Ab = (A + 31) / 32, Bb = (B + 31) / 32;
    BF_CHECK_ARG(Ab <= 65535, BF_ERR_SIZE, "bf_gp_cov: at most 2,097,120 rows per call");
    bf_cov_kernel<<<dim3((unsigned)((Bb + 7) / 8), (unsigned)Ab), 256, 0, (cudaStream_t)stream>>>(
        kern, A, B, d, np, Xa, Xb, inv_l2, amp, Vta, Vtb, C, ldc);
    BF_CHECK_LAUNCH("bf_gp_cov");
    return 0;
}
