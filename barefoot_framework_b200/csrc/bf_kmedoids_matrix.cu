// k-medoids on an EXPLICIT distance matrix: the reference's vendored kMedoids(D, k, tmax) (kmedoids.py:168-227;
// every call site in barefoot.py is commented out, the function is kept for completeness of the hot-path table,
// SURVEY.md 8a row a-K).  Same primitives as the feature-row kernels of bf_kmedoids.cu, with D read instead of
// recomputed:
//   assign   J = argmin(D[:, M], axis=1)                                  (kmedoids.py:208)
//   cost     mean(D[ix_(C, C)], axis=1) per cluster                       (:213)   (after the stable label sort)
//   update   Mnew[kappa] = C[kappa][argmin]  - always moves, first minimum (:214-215)
#include "bf_common.cuh"

__global__ void __launch_bounds__(256) bf_kmd_assign_kernel(long long n, int k, const double* __restrict__ D, long long ld,
                                                            const long long* __restrict__ M, int32_t* __restrict__ labels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double best = INFINITY;
    int lab = 0;
    bool any = false;
    for (int c = 0; c < k; ++c) {
        const double v = D[(size_t)i * ld + M[c]];
        // np.argmin: the first minimum; a NaN is returned as the minimum when it comes first
        if (!any || v < best || (isnan(v) && !isnan(best))) {
            if (!any || v < best || isnan(v)) {
                best = v;
                lab = c;
            }
            any = true;
        }
        if (isnan(best)) break;
    }
    labels[i] = lab;
}

// one warp per position p of the label-sorted order: cost[p] = (sum_q D[perm[p], perm[q]]) / n_c over its cluster
__global__ void __launch_bounds__(256) bf_kmd_cost_kernel(long long n, int k, const double* __restrict__ D, long long ld,
                                                          const long long* __restrict__ perm,
                                                          const long long* __restrict__ seg, double* __restrict__ cost) {
    const long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n) return;
    int c = 0;
    while (c + 1 < k && seg[c + 1] <= p) ++c;
    const long long q0 = seg[c], q1 = seg[c + 1];
    const double* row = D + (size_t)perm[p] * ld;
    double s = 0.0;
    for (long long q = q0 + lane; q < q1; q += 32) s += row[perm[q]];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) cost[p] = s / (double)(q1 - q0);
}

// one CTA per cluster: the first minimum of the cluster's mean distances becomes the medoid; empty[0] counts
// empty clusters (the reference's np.argmin of an empty array raises)
__global__ void __launch_bounds__(256) bf_kmd_update_kernel(const long long* __restrict__ perm, const long long* __restrict__ seg,
                                                            const double* __restrict__ cost, long long* __restrict__ Mnew,
                                                            int32_t* __restrict__ empty) {
    const int c = blockIdx.x;
    const long long q0 = seg[c], q1 = seg[c + 1];
    if (q1 <= q0) {
        if (threadIdx.x == 0) atomicAdd(empty, 1);
        return;
    }
    double bv = INFINITY;
    long long bp = 0x7fffffffffffffffLL;
    for (long long q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const double v = cost[q];
        if (v < bv || (v == bv && q < bp) || (bp == 0x7fffffffffffffffLL)) {
            if (v < bv || bp == 0x7fffffffffffffffLL || (v == bv && q < bp)) {
                bv = v;
                bp = q;
            }
        }
    }
    __shared__ double sv[256];
    __shared__ long long sp[256];
    sv[threadIdx.x] = bv;
    sp[threadIdx.x] = bp;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sv[threadIdx.x + s];
            const long long q = sp[threadIdx.x + s];
            const bool have = sp[threadIdx.x] != 0x7fffffffffffffffLL, other = q != 0x7fffffffffffffffLL;
            if (other && (!have || v < sv[threadIdx.x] || (v == sv[threadIdx.x] && q < sp[threadIdx.x]))) {
                sv[threadIdx.x] = v;
                sp[threadIdx.x] = q;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) Mnew[c] = perm[sp[0]];
}

extern "C" int bf_kmedoids_matrix_assign(long long n, int k, const double* D, long long ld, const long long* M,
                                         int32_t* labels, void* stream) {
    BF_CHECK_ARG(n > 0 && k > 0 && D && ld >= n && M && labels, BF_ERR_ARG, "bf_kmedoids_matrix_assign: bad argument");
    bf_kmd_assign_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, k, D, ld, M, labels);
    BF_CHECK_LAUNCH("bf_kmedoids_matrix_assign");
    return 0;
}

extern "C" int bf_kmedoids_matrix_update(long long n, int k, const double* D, long long ld, const long long* perm,
                                         const long long* seg, double* cost, long long* Mnew, int32_t* empty,
                                         void* stream) {
    BF_CHECK_ARG(n > 0 && k > 0 && D && ld >= n && perm && seg && cost && Mnew && empty, BF_ERR_ARG,
                 "bf_kmedoids_matrix_update: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    bf_kmd_cost_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(n, k, D, ld, perm, seg, cost);
    bf_kmd_update_kernel<<<k, 256, 0, st>>>(perm, seg, cost, Mnew, empty);
    BF_CHECK_LAUNCH("bf_kmedoids_matrix_update");
    return 0;
}
