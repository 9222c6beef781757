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
