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

// ---- large N (4096 < N <= BF_KG_FULL_BIG_MAX_N): the same algorithm with the lines in a per-CTA global scratch ----
// 16 N bytes of lines no longer fit in shared memory (256 kB at N = 16384), so every CTA owns a scratch slice
// (lines, abscissae, stack: 28 N bytes, L2-resident for the ~450 CTAs in flight) and the bitonic network runs
// hybrid: strides >= KGB_CHUNK as passes over the scratch, all smaller strides of a stage inside shared memory
// one KGB_CHUNK-line block at a time.  The comparator, the unique-with-max scan, Algorithm 1 and the summation
// tree are the ones of bf_kg_full_kernel, so both kernels return the same bits.  Persistent grid over the
// (candidate, set) work items: ONE launch for all sets.
#define KGB_CHUNK 4096

__device__ __forceinline__ void kgb_cmpswap(KgLine* __restrict__ line, int t, int p, bool up) {
    const KgLine x = line[t], y = line[p];
    if (up ? kgf_before(y, x) : kgf_before(x, y)) {
        line[t] = y;
        line[p] = x;
    }
}

__global__ void __launch_bounds__(KGF_THREADS) bf_kg_full_big_kernel(int N, int NP2, int S, const double* __restrict__ mu_all,
                                                                     const double* __restrict__ sigma_all, double sn2,
                                                                     double* __restrict__ nu_all, long long M,
                                                                     unsigned char* __restrict__ workspace,
                                                                     size_t ws_stride) {
    extern __shared__ __align__(16) unsigned char kgf_smem[];
    KgLine* sl = reinterpret_cast<KgLine*>(kgf_smem);                   // KGB_CHUNK lines
    __shared__ int s_env;
    __shared__ double s_red[KGF_THREADS / 32];
    unsigned char* ws = workspace + (size_t)blockIdx.x * ws_stride;
    KgLine* line = reinterpret_cast<KgLine*>(ws);                        // NP2
    double* cc = reinterpret_cast<double*>(line + NP2);                  // N
    int* stack = reinterpret_cast<int*>(cc + N);                         // N
    const int tid = threadIdx.x;
    const long long items = M * (long long)S;
    for (long long w = blockIdx.x; w < items; w += gridDim.x) {
        const int i = (int)(w % M), set = (int)(w / M);
        const double* mu = mu_all + (size_t)set * N;
        const double* sigma = sigma_all + (size_t)set * N * N;
        const double denom = sqrt(sn2 + sigma[(size_t)i * N + i]);
        // 1 + 2a. lines -> sorted blocks of KGB_CHUNK (each block built and sorted in shared memory)
        for (int c0 = 0; c0 < NP2; c0 += KGB_CHUNK) {
            for (int tl = tid; tl < KGB_CHUNK; tl += KGF_THREADS) {
                const int j = c0 + tl;
                KgLine l;
                if (j < N) {
                    l.b = sigma[(size_t)j * N + i] / denom;              // column i, as the reference reads it
                    l.a = mu[j];
                } else {
                    l.b = INFINITY;
                    l.a = -INFINITY;
                }
                sl[tl] = l;
            }
            __syncthreads();
            for (int k = 2; k <= KGB_CHUNK; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int tl = tid; tl < KGB_CHUNK; tl += KGF_THREADS) {
                        const int pl = tl ^ j;
                        if (pl > tl) kgb_cmpswap(sl, tl, pl, ((c0 + tl) & k) == 0);
                    }
                    __syncthreads();
                }
            }
            for (int tl = tid; tl < KGB_CHUNK; tl += KGF_THREADS) line[c0 + tl] = sl[tl];
            __syncthreads();
        }
        // 2b. the stages above the block size: long strides on the scratch, the rest per block in shared memory
        for (int k = 2 * KGB_CHUNK; k <= NP2; k <<= 1) {
            for (int j = k >> 1; j >= KGB_CHUNK; j >>= 1) {
                for (int t = tid; t < NP2; t += KGF_THREADS) {
                    const int pp = t ^ j;
                    if (pp > t) kgb_cmpswap(line, t, pp, (t & k) == 0);
                }
                __syncthreads();
            }
            for (int c0 = 0; c0 < NP2; c0 += KGB_CHUNK) {
                for (int tl = tid; tl < KGB_CHUNK; tl += KGF_THREADS) sl[tl] = line[c0 + tl];
                __syncthreads();
                for (int j = KGB_CHUNK >> 1; j > 0; j >>= 1) {
                    for (int tl = tid; tl < KGB_CHUNK; tl += KGF_THREADS) {
                        const int pl = tl ^ j;
                        if (pl > tl) kgb_cmpswap(sl, tl, pl, ((c0 + tl) & k) == 0);
                    }
                    __syncthreads();
                }
                for (int tl = tid; tl < KGB_CHUNK; tl += KGF_THREADS) line[c0 + tl] = sl[tl];
                __syncthreads();
            }
        }
        // 3. unique slopes (largest intercept first in each run), then Algorithm 1: order-dependent scans
        if (tid == 0) {
            int cnt = 0;
            for (int j = 0; j < N; ++j) {
                const KgLine l = line[j];
                if (j > 0 && (l.b == line[cnt - 1].b || (isnan(l.b) && isnan(line[cnt - 1].b)))) continue;
                line[cnt++] = l;
            }
            int top = 0;
            stack[0] = 0;
            cc[0] = INFINITY;
            for (int q = 1; q < cnt; ++q) {
                cc[q] = INFINITY;
                const KgLine lq = line[q];
                while (true) {
                    const int j = stack[top];
                    const KgLine lj = line[j];
                    cc[j] = (lj.a - lq.a) / (lq.b - lj.b);
                    if (top != 0 && cc[j] <= cc[stack[top - 1]]) {
                        --top;
                    } else {
                        break;
                    }
                }
                stack[++top] = q;
            }
            s_env = top + 1;
        }
        __syncthreads();
        const int env = s_env;
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
            for (int ww = 0; ww < KGF_THREADS / 32; ++ww) sig += s_red[ww];
            nu_all[(size_t)set * M + i] = log(sig);
        }
        __syncthreads();                        // scratch and s_red are reused by the next work item
    }
}

extern "C" int bf_kg_full_max_n(void) { return BF_KG_FULL_BIG_MAX_N; }

static int kgb_np2(int N) {
    int np2 = KGB_CHUNK;
    while (np2 < N) np2 <<= 1;
    return np2;
}
static size_t kgb_ws_stride(int N) {
    const size_t b = (size_t)kgb_np2(N) * sizeof(KgLine) + (size_t)N * (sizeof(double) + sizeof(int));
    return (b + 255) & ~(size_t)255;
}
static int kgb_ctas(long long items) {
    const long long want = 148LL * 3;           // 64 kB of shared memory per CTA: three per SM
    return (int)(items < want ? items : want);
}

/* bytes of device scratch bf_kg_full needs for N candidates (0 for N <= BF_KG_FULL_MAX_N) */
extern "C" size_t bf_kg_full_workspace(int N, long long M, int S) {
    if (N <= BF_KG_FULL_MAX_N || M <= 0 || S <= 0) return 0;
    return kgb_ws_stride(N) * (size_t)kgb_ctas(M * (long long)S);
}

extern "C" int bf_kg_full(int N, long long M, int S, const double* mu, const double* sigma, double sn,
                          double* nu_out, double* kg_val, long long* kg_idx, void* workspace, void* stream) {
    BF_CHECK_ARG(N > 0 && M >= 0 && M <= N && S >= 0 && mu && sigma && nu_out && kg_val && kg_idx, BF_ERR_ARG,
                 "bf_kg_full: bad argument");
    BF_CHECK_ARG(N <= BF_KG_FULL_BIG_MAX_N, BF_ERR_SIZE, "bf_kg_full: N=%d above the supported %d", N, BF_KG_FULL_BIG_MAX_N);
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_kg_full: at most 65535 sets per call");
    if (S == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (M > 0 && N > BF_KG_FULL_MAX_N) {
        BF_CHECK_ARG(workspace != nullptr, BF_ERR_ARG, "bf_kg_full: N=%d needs bf_kg_full_workspace(N, M, S) bytes of scratch", N);
        const size_t smem = (size_t)KGB_CHUNK * sizeof(KgLine);
        cudaError_t e = cudaFuncSetAttribute(bf_kg_full_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            bf_set_error("bf_kg_full: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
        bf_kg_full_big_kernel<<<kgb_ctas(M * (long long)S), KGF_THREADS, smem, st>>>(
            N, kgb_np2(N), S, mu, sigma, sn * sn, nu_out, M, (unsigned char*)workspace, kgb_ws_stride(N));
        BF_CHECK_LAUNCH("bf_kg_full(big)");
    } else if (M > 0) {
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
