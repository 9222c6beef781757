// Small on-device steps of model_reification.create_fused_GP between the GP posteriors and the
// fused-GP build (reificationFusion.py:145-159), so the fused targets never leave the GPU.
#include "bf_common.cuh"

__global__ void bf_affine_kernel(long long n, const double* __restrict__ a, double scale, double shift,
                                 double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * scale + shift;
}

extern "C" int bf_affine(long long n, const double* a, double scale, double shift, double* out, void* stream) {
    BF_CHECK_ARG(n >= 0 && a && out, BF_ERR_ARG, "bf_affine: bad argument");
    if (n == 0) return 0;
    bf_affine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, a, scale, shift, out);
    BF_CHECK_LAUNCH("bf_affine");
    return 0;
}

// (err_mean_pred)^2 + m_var with the reference's association:
//   m_var = var * (model_std ** 2);  err = err_mu * err_std + err_mean;  out = err ** 2 + m_var
__global__ void bf_total_variance_kernel(long long n, const double* __restrict__ err_mu, double err_std,
                                         double err_mean, const double* __restrict__ var, double model_std2,
                                         double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double e = err_mu[i] * err_std + err_mean;
        out[i] = e * e + var[i] * model_std2;
    }
}

extern "C" int bf_total_variance(long long n, const double* err_mu, double err_std, double err_mean,
                                 const double* var, double model_std, double* out, void* stream) {
    BF_CHECK_ARG(n >= 0 && err_mu && var && out, BF_ERR_ARG, "bf_total_variance: bad argument");
    if (n == 0) return 0;
    bf_total_variance_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, err_mu, err_std, err_mean, var, model_std * model_std, out);
    BF_CHECK_LAUNCH("bf_total_variance");
    return 0;
}

// block-wide sum in a fixed order (deterministic)
__device__ double bf_block_sum(double v, double* red) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(256) bf_zscore_kernel(int n, const double* __restrict__ f_mean,
                                                        const double* __restrict__ f_var, double* __restrict__ z,
                                                        double* __restrict__ zerr, double* __restrict__ stats) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += f_mean[i];
    const double mean = bf_block_sum(s, red) / n;
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double d = f_mean[i] - mean;
        q += d * d;
    }
    double sd = sqrt(bf_block_sum(q, red) / n);
    if (sd == 0.0) sd = 1.0;
    const double sd2 = sd * sd;
    for (int i = threadIdx.x; i < n; i += 256) {
        z[i] = (f_mean[i] - mean) / sd;
        zerr[i] = sqrt(fabs(f_var[i] / sd2));
    }
    if (threadIdx.x == 0) {
        stats[0] = mean;
        stats[1] = sd;
    }
}

extern "C" int bf_zscore_targets(int n, const double* f_mean, const double* f_var, double* z, double* zerr,
                                 double* stats, void* stream) {
    BF_CHECK_ARG(n > 0 && f_mean && f_var && z && zerr && stats, BF_ERR_ARG, "bf_zscore_targets: bad argument");
    bf_zscore_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(n, f_mean, f_var, z, zerr, stats);
    BF_CHECK_LAUNCH("bf_zscore_targets");
    return 0;
}

// ---- batched forms for the ROM stage (one launch for all (model, fantasy point) groups) ------------------
// G independent z-scorings of n values each: the arithmetic of bf_zscore_kernel per group, one CTA per group.
__global__ void __launch_bounds__(256) bf_zscore_batched_kernel(int n, const double* __restrict__ f_mean,
                                                                const double* __restrict__ f_var,
                                                                double* __restrict__ z, double* __restrict__ zerr,
                                                                double* __restrict__ stats) {
    const size_t off = (size_t)blockIdx.x * n;
    f_mean += off; f_var += off; z += off; zerr += off; stats += 2 * (size_t)blockIdx.x;
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += f_mean[i];
    const double mean = bf_block_sum(s, red) / n;
    double q = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double d = f_mean[i] - mean;
        q += d * d;
    }
    double sd = sqrt(bf_block_sum(q, red) / n);
    if (sd == 0.0) sd = 1.0;
    const double sd2 = sd * sd;
    for (int i = threadIdx.x; i < n; i += 256) {
        z[i] = (f_mean[i] - mean) / sd;
        zerr[i] = sqrt(fabs(f_var[i] / sd2));
    }
    if (threadIdx.x == 0) {
        stats[0] = mean;
        stats[1] = sd;
    }
}

extern "C" int bf_zscore_targets_batched(int G, int n, const double* f_mean, const double* f_var, double* z,
                                         double* zerr, double* stats, void* stream) {
    BF_CHECK_ARG(G >= 0 && n > 0 && f_mean && f_var && z && zerr && stats, BF_ERR_ARG, "bf_zscore_targets_batched: bad argument");
    if (G == 0) return 0;
    bf_zscore_batched_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(n, f_mean, f_var, z, zerr, stats);
    BF_CHECK_LAUNCH("bf_zscore_targets_batched");
    return 0;
}

// The ROM-stage fantasy update (util.py:243: reification object.update_GP(x_test[kk], new_mean, jj), i.e.
// gp_model.update, gpModel.py:56-64 - append ONE point and refactor) for K fantasy points at once, as the
// bordered (rank-one) form of the appended Cholesky factor.  With the un-updated GP's posterior at the fused
// points (mu_f, var_f), at the fantasy points (mu_t, var_t) and the posterior covariance C[f, k] between them,
// appending (x_k, y_k) with noise variance s2 = sigma_n^2 + TINY gives
//     s_k   = var_t[k] + s2                               (the new pivot squared: amp + s2 - |L^-1 k_new|^2)
//     mu'   = mu_f + C[f, k] (y_k - mu_t[k]) / s_k
//     var'  = var_f - C[f, k]^2 / s_k
// and the fused-GP inputs follow as in reificationFusion.py:147-152:
//     mean_row = (mu' + gp_mean) * model_std + model_mean
//     var_row  = err_term[f] + var' * model_std^2,   err_term = (err_mu * err_std + err_mean)^2 precomputed.
__global__ void bf_fantasy_rows_kernel(int F, int K, const double* __restrict__ mu_f, const double* __restrict__ var_f,
                                       const double* __restrict__ C, long long ldc, const double* __restrict__ mu_t,
                                       const double* __restrict__ var_t, const double* __restrict__ y_new, double s2,
                                       double gp_mean, double model_std, double model_mean,
                                       const double* __restrict__ err_term, double* __restrict__ mean_rows,
                                       double* __restrict__ var_rows) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (f >= F) return;
    const double s = var_t[k] + s2;
    const double c = C[(size_t)f * ldc + k];
    const double mu = mu_f[f] + c * (y_new[k] - mu_t[k]) / s;
    const double var = var_f[f] - c * c / s;
    mean_rows[(size_t)k * F + f] = (mu + gp_mean) * model_std + model_mean;
    var_rows[(size_t)k * F + f] = err_term[f] + var * (model_std * model_std);
}

extern "C" int bf_fantasy_rows(int F, int K, const double* mu_f, const double* var_f, const double* C, long long ldc,
                               const double* mu_t, const double* var_t, const double* y_new, double noise_var,
                               double gp_mean, double model_std, double model_mean, const double* err_term,
                               double* mean_rows, double* var_rows, void* stream) {
    BF_CHECK_ARG(F > 0 && K >= 0 && mu_f && var_f && C && ldc >= K && mu_t && var_t && y_new && err_term && mean_rows && var_rows,
                 BF_ERR_ARG, "bf_fantasy_rows: bad argument");
    if (K == 0) return 0;
    BF_CHECK_ARG(K <= 65535, BF_ERR_SIZE, "bf_fantasy_rows: at most 65535 fantasy points per call");
    dim3 grid((F + 127) / 128, K);
    bf_fantasy_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(F, K, mu_f, var_f, C, ldc, mu_t, var_t, y_new, noise_var,
                                                                  gp_mean, model_std, model_mean, err_term, mean_rows, var_rows);
    BF_CHECK_LAUNCH("bf_fantasy_rows");
    return 0;
}
