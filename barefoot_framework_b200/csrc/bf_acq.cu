// Stand-alone elementwise acquisitions on given mean / spread vectors: expected_improvement,
// probability_improvement, upper_conf_bound (acquisitionFunc.py:98-214) and the greedy first
// arg-max of the mean (util.py:655-665).  The hot path fuses these into the posterior epilogue
// (bf_posterior.cu); this entry point serves the reference-named host functions that receive
// arrays.  HBM-bound: 16 B read (+ 8 B written if the values are requested) per candidate.
#include "bf_common.cuh"

#define ACQ_THREADS 256

extern "C" int bf_acq_tiles(long long N) {
    long long t = (N + 4 * ACQ_THREADS - 1) / (4 * ACQ_THREADS);
    if (t < 1) t = 1;
    if (t > 592) t = 592;   // 4 CTAs per SM on 148 SMs
    return (int)t;
}

__global__ void __launch_bounds__(ACQ_THREADS) bf_acq_eval_kernel(
    int kind, long long N, const double* __restrict__ y, const double* __restrict__ spread, double curr_max,
    double xi, double kt, double* __restrict__ out, double* __restrict__ part_val,
    long long* __restrict__ part_idx) {
    const int set = blockIdx.y, tiles = gridDim.x;
    const double* y_ = y + (size_t)set * N;
    const double* s_ = spread ? spread + (size_t)set * N : nullptr;
    BfBest best = bf_best_init();
    for (long long i = (long long)blockIdx.x * ACQ_THREADS + threadIdx.x; i < N;
         i += (long long)tiles * ACQ_THREADS) {
        const double m = y_[i];
        const double s = s_ ? s_[i] : 0.0;
        double v;
        if (kind == BF_ACQ_EI) {
            v = (m - curr_max - xi) * bf_norm_pdf(m) + s * bf_norm_cdf(m);
        } else if (kind == BF_ACQ_PI) {
            v = bf_norm_cdf((m - curr_max - xi) / s);
        } else if (kind == BF_ACQ_UCB) {
            v = m + kt * s;
        } else {
            v = m;
        }
        if (out) out[(size_t)set * N + i] = v;
        bf_best_take(best, v, i);
    }
    best = bf_best_warp(best);
    __shared__ BfBest red[ACQ_THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        BfBest a = red[0];
        for (int w = 1; w < ACQ_THREADS / 32; ++w) bf_best_take(a, red[w].v, red[w].i);
        part_val[(size_t)set * tiles + blockIdx.x] = a.v;
        part_idx[(size_t)set * tiles + blockIdx.x] = a.i;
    }
}

__global__ void __launch_bounds__(ACQ_THREADS) bf_acq_eval_finalize_kernel(
    int tiles, const double* __restrict__ part_val, const long long* __restrict__ part_idx,
    double* __restrict__ best_val, long long* __restrict__ best_idx) {
    const int set = blockIdx.x;
    BfBest best = bf_best_init();
    for (int t = threadIdx.x; t < tiles; t += ACQ_THREADS)
        bf_best_take(best, part_val[(size_t)set * tiles + t], part_idx[(size_t)set * tiles + t]);
    best = bf_best_warp(best);
    __shared__ BfBest red[ACQ_THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        BfBest a = red[0];
        for (int w = 1; w < ACQ_THREADS / 32; ++w) bf_best_take(a, red[w].v, red[w].i);
        best_val[set] = a.v;
        best_idx[set] = a.i;     // 0x7fff... when every value was NaN (the reference raises IndexError)
    }
}

extern "C" int bf_acq_eval(int kind, long long N, int S, const double* y, const double* spread,
                           double curr_max, double xi, double kt, double* out, double* part_val,
                           long long* part_idx, double* best_val, long long* best_idx, void* stream) {
    BF_CHECK_ARG(kind == BF_ACQ_EI || kind == BF_ACQ_PI || kind == BF_ACQ_UCB || kind == BF_ACQ_GREEDY,
                 BF_ERR_ARG, "bf_acq_eval: kind must be exactly one BF_ACQ_* bit (got %d)", kind);
    BF_CHECK_ARG(N > 0 && S >= 0 && y && part_val && part_idx && best_val && best_idx, BF_ERR_ARG,
                 "bf_acq_eval: bad argument");
    BF_CHECK_ARG(spread || kind == BF_ACQ_GREEDY, BF_ERR_ARG, "bf_acq_eval: spread is required");
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_acq_eval: at most 65535 sets per call");
    if (S == 0) return 0;
    const int tiles = bf_acq_tiles(N);
    cudaStream_t st = (cudaStream_t)stream;
    bf_acq_eval_kernel<<<dim3(tiles, S), ACQ_THREADS, 0, st>>>(kind, N, y, spread, curr_max, xi, kt, out,
                                                               part_val, part_idx);
    BF_CHECK_LAUNCH("bf_acq_eval");
    bf_acq_eval_finalize_kernel<<<S, ACQ_THREADS, 0, st>>>(tiles, part_val, part_idx, best_val, best_idx);
    BF_CHECK_LAUNCH("bf_acq_eval(finalize)");
    return 0;
}

// ---- candidate-sharded stages: the per-rank (value, local index) records around the all-gather ------------------
// pack: out[0..S) = values, out[S..2S) = (index + offset) bit-cast to double - ONE buffer for ONE collective.
__global__ void bf_records_pack_kernel(int S, const double* __restrict__ val, const long long* __restrict__ idx,
                                       long long offset, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) {
        out[s] = val[s];
        out[S + s] = __longlong_as_double(idx[s] + offset);
    }
}

// reduce: gathered [world x 2 x S] -> per set the largest value and, among equal values, the lowest GLOBAL index
// (numpy's first arg-max over the concatenated candidate array, acquisitionFunc.py:135-138); NaN never wins.
__global__ void bf_records_reduce_kernel(int world, int S, const double* __restrict__ gathered,
                                         double* __restrict__ best_val, long long* __restrict__ best_idx) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    BfBest b = bf_best_init();
    for (int r = 0; r < world; ++r) {
        const double* rec = gathered + (size_t)r * 2 * S;
        bf_best_take(b, rec[s], __double_as_longlong(rec[S + s]));
    }
    best_val[s] = b.v;
    best_idx[s] = b.i;
}

extern "C" int bf_records_pack(int S, const double* val, const long long* idx, long long offset, double* out, void* stream) {
    BF_CHECK_ARG(S >= 0 && val && idx && out, BF_ERR_ARG, "bf_records_pack: bad argument");
    if (S == 0) return 0;
    bf_records_pack_kernel<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(S, val, idx, offset, out);
    BF_CHECK_LAUNCH("bf_records_pack");
    return 0;
}

extern "C" int bf_records_reduce(int world, int S, const double* gathered, double* best_val, long long* best_idx, void* stream) {
    BF_CHECK_ARG(world > 0 && S >= 0 && gathered && best_val && best_idx, BF_ERR_ARG, "bf_records_reduce: bad argument");
    if (S == 0) return 0;
    bf_records_reduce_kernel<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(world, S, gathered, best_val, best_idx);
    BF_CHECK_LAUNCH("bf_records_reduce");
    return 0;
}
