// Knowledge gradient on a DENSE covariance: Frazier, Powell & Dayanik's algorithm exactly as
// knowledge_gradient(M, sn, mu, sigma) runs it (acquisitionFunc.py:12-96).  The batch-only path
// calls it with the full posterior covariance of the truth GP (util.py:1078-1081).
//
// One CTA per candidate i (and per set):
//   1. b_j = sigma[j, i] / sqrt(sn^2 + sigma[i, i]),  a_j = mu_j                      (:62)
//   2. sort the lines by slope b; among equal slopes keep the largest intercept a      (:66-72)
//      -> bitonic sort on (b ascending, a descending) in shared memory, then a head-flag scan
//   3. upper envelope of the remaining lines: Algorithm 1, a sequential stack scan     (:37-53)
//   4. nu_i = log sum_k (b_{k+1} - b_k) (pdf(-|c_k|) - |c_k| cdf(-|c_k|))               (:82-84)
// and a second tiny kernel applies the reference's running maximum from the floor (0, index 0)
// with strict '>' over finite values (:56-57,87-91).  N <= BF_KG_FULL_MAX_N candidates (the
// reference itself holds the N x N matrix, so N is small on this path).
#include "bf_common.cuh"

#define KGF_THREADS 256

struct KgLine {
    double b, a;
};

// (b ascending, a descending); NaN slopes sort last
__device__ __forceinline__ bool kgf_before(const KgLine& x, const KgLine& y) {
    const bool xn = isnan(x.b), yn = isnan(y.b);
    if (xn || yn) return !xn && yn;
    if (x.b != y.b) return x.b < y.b;
    return x.a > y.a;
}

__global__ void __launch_bounds__(KGF_THREADS) bf_kg_full_kernel(int N, int NP2, const double* __restrict__ mu_all,
                                                                 const double* __restrict__ sigma_all, double sn2,
                                                                 double* __restrict__ nu_all, long long M) {
    extern __shared__ __align__(16) unsigned char kgf_smem[];
    KgLine* line = reinterpret_cast<KgLine*>(kgf_smem);                 // NP2
    double* cc = reinterpret_cast<double*>(line + NP2);                 // N   intersection abscissae
    int* stack = reinterpret_cast<int*>(cc + N);                        // N   envelope members
    __shared__ int s_count, s_env;
    __shared__ double s_red[KGF_THREADS / 32];

    const int i = blockIdx.x, set = blockIdx.y, tid = threadIdx.x;
    const double* mu = mu_all + (size_t)set * N;
    const double* sigma = sigma_all + (size_t)set * N * N;
    const double denom = sqrt(sn2 + sigma[(size_t)i * N + i]);
    for (int j = tid; j < NP2; j += KGF_THREADS) {
        KgLine l;
        if (j < N) {
            l.b = sigma[(size_t)j * N + i] / denom;      // column i, as the reference reads it
            l.a = mu[j];
        } else {
            l.b = INFINITY;                              // padding sorts behind every real line
            l.a = -INFINITY;
        }
        line[j] = l;
    }
    __syncthreads();
    // bitonic sort
    for (int k = 2; k <= NP2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < NP2; t += KGF_THREADS) {
                const int p = t ^ j;
                if (p > t) {
                    const KgLine x = line[t], y = line[p];
                    const bool up = (t & k) == 0;
                    if (up ? kgf_before(y, x) : kgf_before(x, y)) {
                        line[t] = y;
                        line[p] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
    // unique slopes, largest intercept first in each run; then Algorithm 1.  Both are order
    // dependent scans: one thread walks the sorted lines (N is small on this path).
    if (tid == 0) {
        int cnt = 0;
        for (int j = 0; j < N; ++j) {
            const KgLine l = line[j];
            if (j > 0 && (l.b == line[cnt - 1].b || (isnan(l.b) && isnan(line[cnt - 1].b)))) continue;
            line[cnt++] = l;
        }
        // a = line[.].a, b = line[.].b, MM = cnt
        int top = 0;
        stack[0] = 0;
        cc[0] = INFINITY;
        for (int q = 1; q < cnt; ++q) {
            cc[q] = INFINITY;
            while (true) {
                const int j = stack[top];
                cc[j] = (line[j].a - line[q].a) / (line[q].b - line[j].b);
                if (top != 0 && cc[j] <= cc[stack[top - 1]]) {
                    --top;
                } else {
                    break;
                }
            }
            stack[++top] = q;
        }
        s_count = cnt;
        s_env = top + 1;
    }
    __syncthreads();
    const int env = s_env;
    // sum over consecutive envelope members
    double part = 0.0;
    for (int k = tid; k + 1 < env; k += KGF_THREADS) {
        const int j0 = stack[k], j1 = stack[k + 1];
        const double ac = -fabs(cc[j0]);
        part += (line[j1].b - line[j0].b) * (bf_norm_pdf(ac) + ac * bf_norm_cdf(ac));
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double sig = 0.0;
        for (int w = 0; w < KGF_THREADS / 32; ++w) sig += s_red[w];
        nu_all[(size_t)set * M + i] = log(sig);
    }
}

// running maximum from (0, index 0): strict '>' over finite values
__global__ void __launch_bounds__(KGF_THREADS) bf_kg_full_select_kernel(long long M, const double* __restrict__ nu_all,
                                                                        double* __restrict__ kg_val,
                                                                        long long* __restrict__ kg_idx) {
    const int set = blockIdx.x;
    const double* nu = nu_all + (size_t)set * M;
    BfBest best = bf_best_init();
    for (long long i = threadIdx.x; i < M; i += KGF_THREADS) {
        const double v = nu[i];
        if (isfinite(v) && v > 0.0) bf_best_take(best, v, i);
    }
    best = bf_best_warp(best);
    __shared__ BfBest red[KGF_THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        BfBest a = red[0];
        for (int w = 1; w < KGF_THREADS / 32; ++w) bf_best_take(a, red[w].v, red[w].i);
        if (a.v > 0.0 && isfinite(a.v)) {
            kg_val[set] = a.v;
            kg_idx[set] = a.i;
        } else {
            kg_val[set] = 0.0;
            kg_idx[set] = 0;
        }
    }
}

extern "C" int bf_kg_full_max_n(void) { return BF_KG_FULL_MAX_N; }

extern "C" int bf_kg_full(int N, long long M, int S, const double* mu, const double* sigma, double sn,
                          double* nu_out, double* kg_val, long long* kg_idx, void* stream) {
    BF_CHECK_ARG(N > 0 && M >= 0 && M <= N && S >= 0 && mu && sigma && nu_out && kg_val && kg_idx, BF_ERR_ARG,
                 "bf_kg_full: bad argument");
    BF_CHECK_ARG(N <= BF_KG_FULL_MAX_N, BF_ERR_SIZE, "bf_kg_full: N=%d above the supported %d", N, BF_KG_FULL_MAX_N);
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_kg_full: at most 65535 sets per call");
    if (S == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (M > 0) {
        int np2 = 1;
        while (np2 < N) np2 <<= 1;
        const size_t smem = (size_t)np2 * sizeof(KgLine) + (size_t)N * (sizeof(double) + sizeof(int)) + 16;
        cudaError_t e = cudaFuncSetAttribute(bf_kg_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            bf_set_error("bf_kg_full: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
        bf_kg_full_kernel<<<dim3((unsigned)M, S), KGF_THREADS, smem, st>>>(N, np2, mu, sigma, sn * sn, nu_out, M);
        BF_CHECK_LAUNCH("bf_kg_full");
    }
    bf_kg_full_select_kernel<<<S, KGF_THREADS, 0, st>>>(M, nu_out, kg_val, kg_idx);
    BF_CHECK_LAUNCH("bf_kg_full(select)");
    return 0;
}
