// Knowledge gradient exactly as the reference evaluates it: knowledge_gradient(M, 0.1, mu,
// np.diag(var)) (acquisitionFunc.py:12-96 called from util.py:259-262,618-621).  With a diagonal
// covariance the sort / unique / upper-envelope of Frazier's algorithm leaves two lines per
// candidate, so
//   beta_i = var_i / sqrt(sn^2 + var_i);  o_i = max_{j != i} mu_j;  c = (o_i - mu_i) / beta_i
//   nu_i   = log(|beta_i| * (pdf(-|c|) - |c| cdf(-|c|)))
// and the running maximum starts at (0, index 0) with strict '>' over finite values (quirk Q8).
// HBM-bound: 16 B read per candidate (+ 8 B if nu_out is requested).
#include "bf_common.cuh"

#define KG_THREADS 256

extern "C" int bf_kg_tiles(long long N) {
    long long t = (N + 4 * KG_THREADS - 1) / (4 * KG_THREADS);
    if (t < 1) t = 1;
    if (t > 592) t = 592;   // 4 CTAs per SM on 148 SMs
    return (int)t;
}

__global__ void __launch_bounds__(KG_THREADS) bf_kg_kernel(
    long long N, long long M, const double* __restrict__ mu, const double* __restrict__ var,
    const double* __restrict__ top, const long long* __restrict__ top_idx, double sn2,
    double* __restrict__ nu_out, double* __restrict__ part_val, long long* __restrict__ part_idx) {
    const int set = blockIdx.y, tiles = gridDim.x;
    const double t1 = top[(size_t)set * BF_NSLOT_VAL + BF_SLOT_GREEDY];
    const double t2 = top[(size_t)set * BF_NSLOT_VAL + BF_SLOT_TOP2];
    const long long i1 = top_idx[(size_t)set * BF_NSLOT_IDX + BF_SLOT_GREEDY];
    const double* m_ = mu + (size_t)set * N;
    const double* v_ = var + (size_t)set * N;
    BfBest best = bf_best_init();
    for (long long i = (long long)blockIdx.x * KG_THREADS + threadIdx.x; i < M;
         i += (long long)tiles * KG_THREADS) {
        const double m = m_[i], v = v_[i];
        const double beta = v / sqrt(sn2 + v);
        double nu = -INFINITY;
        if (beta != 0.0) {
            const double o = (i == i1) ? t2 : t1;
            const double c = beta > 0.0 ? (o - m) / (beta - 0.0) : (m - o) / (0.0 - beta);
            const double ac = -fabs(c);
            const double sig = fabs(beta) * (bf_norm_pdf(ac) + ac * bf_norm_cdf(ac));
            nu = log(sig);
        }
        if (nu_out) nu_out[(size_t)set * N + i] = nu;
        if (isfinite(nu) && nu > 0.0) bf_best_take(best, nu, i);
    }
    best = bf_best_warp(best);
    __shared__ BfBest red[KG_THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        BfBest a = red[0];
        for (int w = 1; w < KG_THREADS / 32; ++w) bf_best_take(a, red[w].v, red[w].i);
        part_val[(size_t)set * tiles + blockIdx.x] = a.v;
        part_idx[(size_t)set * tiles + blockIdx.x] = a.i;
    }
}

__global__ void __launch_bounds__(KG_THREADS) bf_kg_finalize_kernel(
    int tiles, const double* __restrict__ part_val, const long long* __restrict__ part_idx,
    double* __restrict__ kg_val, long long* __restrict__ kg_idx) {
    const int set = blockIdx.x;
    BfBest best = bf_best_init();
    for (int t = threadIdx.x; t < tiles; t += KG_THREADS)
        bf_best_take(best, part_val[(size_t)set * tiles + t], part_idx[(size_t)set * tiles + t]);
    best = bf_best_warp(best);
    __shared__ BfBest red[KG_THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        BfBest a = red[0];
        for (int w = 1; w < KG_THREADS / 32; ++w) bf_best_take(a, red[w].v, red[w].i);
        if (a.v > 0.0 && isfinite(a.v)) {
            kg_val[set] = a.v;
            kg_idx[set] = a.i;
        } else {   // nothing beat the floor: nu_star = 0, x_star = 0
            kg_val[set] = 0.0;
            kg_idx[set] = 0;
        }
    }
}

// top-2 of the mean "by position" for callers that hold mu in memory (the fused path gets it from
// the posterior epilogue): one CTA per set, fixed-order tree so the result is deterministic.
__global__ void __launch_bounds__(1024) bf_top2_kernel(long long N, const double* __restrict__ mu,
                                                       double* __restrict__ top, long long* __restrict__ top_idx) {
    const int set = blockIdx.x;
    const double* m_ = mu + (size_t)set * N;
    BfTop2 t = bf_top2_init();
    for (long long i = threadIdx.x; i < N; i += blockDim.x) bf_top2_merge(t, m_[i], i, -INFINITY);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double t1 = __shfl_xor_sync(0xffffffffu, t.t1, m);
        const long long i1 = __shfl_xor_sync(0xffffffffu, t.i1, m);
        const double t2 = __shfl_xor_sync(0xffffffffu, t.t2, m);
        bf_top2_merge(t, t1, i1, t2);
    }
    __shared__ BfTop2 red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        BfTop2 a = red[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) bf_top2_merge(a, red[w].t1, red[w].i1, red[w].t2);
        top[(size_t)set * BF_NSLOT_VAL + BF_SLOT_GREEDY] = a.t1;
        top[(size_t)set * BF_NSLOT_VAL + BF_SLOT_TOP2] = a.t2;
        top_idx[(size_t)set * BF_NSLOT_IDX + BF_SLOT_GREEDY] = a.i1;
    }
}

extern "C" int bf_top2(long long N, int S, const double* mu, double* top, long long* top_idx, void* stream) {
    BF_CHECK_ARG(N > 0 && S >= 0 && mu && top && top_idx, BF_ERR_ARG, "bf_top2: bad argument");
    if (S == 0) return 0;
    bf_top2_kernel<<<S, 1024, 0, (cudaStream_t)stream>>>(N, mu, top, top_idx);
    BF_CHECK_LAUNCH("bf_top2");
    return 0;
}

extern "C" int bf_kg_diag(long long N, long long M, int S, const double* mu, const double* var,
                          const double* top, const long long* top_idx, double sn, double* nu_out,
                          double* part_val, long long* part_idx, double* kg_val, long long* kg_idx,
                          void* stream) {
    BF_CHECK_ARG(N > 0 && M >= 0 && M <= N && S >= 0, BF_ERR_ARG, "bf_kg_diag: bad sizes N=%lld M=%lld S=%d", N, M, S);
    BF_CHECK_ARG(mu && var && top && top_idx && part_val && part_idx && kg_val && kg_idx, BF_ERR_ARG,
                 "bf_kg_diag: null argument");
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_kg_diag: at most 65535 sets per call");
    if (S == 0) return 0;
    const int tiles = bf_kg_tiles(N);
    cudaStream_t st = (cudaStream_t)stream;
    bf_kg_kernel<<<dim3(tiles, S), KG_THREADS, 0, st>>>(N, M, mu, var, top, top_idx, sn * sn, nu_out,
                                                        part_val, part_idx);
    BF_CHECK_LAUNCH("bf_kg_diag");
    bf_kg_finalize_kernel<<<S, KG_THREADS, 0, st>>>(tiles, part_val, part_idx, kg_val, kg_idx);
    BF_CHECK_LAUNCH("bf_kg_diag(finalize)");
    return 0;
}
