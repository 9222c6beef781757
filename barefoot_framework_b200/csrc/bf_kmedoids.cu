// k-medoids primitives behind kmedoids_max (util.py:38-49 -> sklearn_extra KMedoids with the
// defaults: euclidean, "alternate", "heuristic" init).  The N' x N' distance matrix the
// reference materialises (80 GB at N' = 1e5) is never formed: distances are recomputed from
// the N' x f feature rows (f = 2 or 3, barefoot.py:1832,2020,2152) on the FP64 pipe.
//
// Distance form.  sklearn_extra calls sklearn.metrics.pairwise_distances(X, metric="euclidean"),
// which evaluates the EXPANDED form (sklearn/metrics/pairwise.py, _euclidean_distances, float64 branch):
//     D = -2 * (X @ X.T);  D += XX[:, None];  D += XX[None, :];  maximum(D, 0);  diag(D) = 0;  sqrt(D)
// with XX = einsum("ij,ij->i", X, X).  The reference's rows are (value, candidate index up to 1e6, model
// index), so XX ~ 1e12 and the cancellation in the expanded form moves a small distance by up to 5e-3
// relative against sqrt(sum (x - y)^2): medoids and labels are DEFINED by the expanded form, and the
// kernels below evaluate exactly that, operation for operation:
//   * dot   = x0*y0, then fma(x_c, y_c, dot) for c = 1..f-1 (OpenBLAS syrk/gemm micro-kernel order on
//             x86-64 with FMA3; exact for the reference's integer-valued trailing features, where every
//             product but the first is exact and every summation order gives the same bits),
//   * XX    = NumPy einsum's two-lane order: lane 0 sums the even-indexed squares (and the odd tail),
//             lane 1 the odd-indexed ones, XX = lane0 + lane1   (verified here against NumPy 2.3 for f <= 7),
//   * -2 * dot is formed exactly by scaling the position's coordinates by -2 beforehand.
// Sums of distances (row sums, in-cluster costs) use a fixed tree that depends only on the member run.
#include "bf_common.cuh"

#define KM_THREADS 128
#define KM_MAX_F 8

// einsum("ij,ij->i") order (see above)
template <int F>
__device__ __forceinline__ double km_norm(const double* __restrict__ x) {
    double l0 = x[0] * x[0];
    if (F == 1) return l0;
    double l1 = x[1] * x[1];
#pragma unroll
    for (int c = 2; c + 1 < F; c += 2) {
        l0 = l0 + x[c] * x[c];
        l1 = l1 + x[c + 1] * x[c + 1];
    }
    if (F & 1) l0 = l0 + x[F - 1] * x[F - 1];
    return l0 + l1;
}
__device__ __forceinline__ double km_norm_rt(const double* __restrict__ x, int f) {
    switch (f) {
#define BF_CASE(FF) case FF: return km_norm<FF>(x);
        BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8)
#undef BF_CASE
    }
    return 0.0;
}

// one off-diagonal element D[r, c] of the expanded form: xm2 = -2 * x (either point), y = the other point,
// n_row = XX of the ROW point r, n_col = XX of the COLUMN point c:  D[r, c] = sqrt(max((-2 x.y + XX_r) + XX_c, 0)).
// The two additions are not commutative in floating point, so D is not bit-symmetric and every caller
// states which of its two points is the row of the reference's matrix:
//   row sums      D.sum(axis=1)                        row = the position, column = the member
//   labels        argmin(D[medoids, :], axis=0)        row = the medoid,   column = the point
//   cluster costs sum(D[idx, idx[:, None]], axis=1)    row = the member,   column = the position
template <int F>
__device__ __forceinline__ double km_dist_exp(const double (&xm2)[F], const double* __restrict__ y, double n_row,
                                              double n_col) {
    double dot = xm2[0] * y[0];
#pragma unroll
    for (int c = 1; c < F; ++c) dot = fma(xm2[c], y[c], dot);
    const double d2 = (dot + n_row) + n_col;
    return sqrt(fmax(d2, 0.0));
}

// ---- sums of distances from a block of 128 "positions" to a run of "members" ----------------------
// One CTA = 128 positions x KM_GROUPS member groups (1024 threads).  The members are cut into tiles of
// KM_KT rows (coordinates + squared norm); group g takes tiles g, g + 8, ...; every thread keeps four
// interleaved partial sums over its tiles, and the eight group totals of a position are added in a fixed
// tree.  The summation order depends only on the member run, never on how rows / chunks are split over
// launches or ranks (PERM = in-cluster costs: the members are the rows of the reference's matrix; otherwise
// row sums: the position is the row), and the eight groups give a latency-bound single-position loop eight times the warps
// (a rank of an 8-GPU job has only ~100 position blocks per launch).  `self` is the member number (in
// [q0, q1)) that IS this position - its distance is the zero of the filled diagonal - or -1.
#define KM_GROUPS 8
#define KM_KT 256
#define KM_BLOCK (KM_THREADS * KM_GROUPS)

template <int F, bool PERM>
__device__ __forceinline__ double km_block_sum(const double (&xm2)[F], double ni, long long self, long long q0,
                                               long long q1, const double* __restrict__ X,
                                               const long long* __restrict__ perm, double* __restrict__ sm) {
    constexpr int FS = F + 1;                                      // coordinates + squared norm
    const int i = threadIdx.x & (KM_THREADS - 1), g = threadIdx.x / KM_THREADS;
    double* tile = sm + (size_t)g * KM_KT * FS;
    double* part = sm + (size_t)KM_GROUPS * KM_KT * FS;            // KM_GROUPS x KM_THREADS
    const long long ntile = (q1 - q0 + KM_KT - 1) / KM_KT;
    const long long rounds = (ntile + KM_GROUPS - 1) / KM_GROUPS;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long rd = 0; rd < rounds; ++rd) {
        const long long j0 = q0 + (rd * KM_GROUPS + g) * KM_KT;
        const int cnt = (int)(j0 >= q1 ? 0 : (q1 - j0 < KM_KT ? q1 - j0 : KM_KT));
        __syncthreads();
        for (int e = i; e < cnt; e += KM_THREADS) {
            const long long row = PERM ? perm[j0 + e] : j0 + e;
            double xr[F];
#pragma unroll
            for (int c = 0; c < F; ++c) xr[c] = X[(size_t)row * F + c];
#pragma unroll
            for (int c = 0; c < F; ++c) tile[e * FS + c] = xr[c];
            tile[e * FS + F] = km_norm<F>(xr);
        }
        __syncthreads();
        const long long sj = self - j0;                             // tile-local number of the diagonal element
        int j = 0;
        for (; j + 4 <= cnt; j += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double nj = tile[(j + u) * FS + F];
                const double dd = km_dist_exp<F>(xm2, tile + (j + u) * FS, PERM ? nj : ni, PERM ? ni : nj);
                acc[u] += (sj == j + u) ? 0.0 : dd;
            }
        }
        for (; j < cnt; ++j) {
            const double nj = tile[j * FS + F];
            const double dd = km_dist_exp<F>(xm2, tile + j * FS, PERM ? nj : ni, PERM ? ni : nj);
            acc[0] += (sj == j) ? 0.0 : dd;
        }
    }
    part[g * KM_THREADS + i] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    return ((part[i] + part[KM_THREADS + i]) + (part[2 * KM_THREADS + i] + part[3 * KM_THREADS + i])) +
           ((part[4 * KM_THREADS + i] + part[5 * KM_THREADS + i]) + (part[6 * KM_THREADS + i] + part[7 * KM_THREADS + i]));
}

static size_t km_block_smem(int f) { return sizeof(double) * ((size_t)KM_GROUPS * KM_KT * (f + 1) + (size_t)KM_GROUPS * KM_THREADS); }

// rowsum[i - row0] = sum_j |x_i - x_j| for rows [row0, row0 + rows) against all N rows
template <int F>
__global__ void __launch_bounds__(KM_BLOCK) bf_km_rowsum_kernel(long long N, const double* __restrict__ X,
                                                                long long row0, long long rows,
                                                                double* __restrict__ rowsum) {
    extern __shared__ __align__(16) double km_sm[];
    const long long r = (long long)blockIdx.x * KM_THREADS + (threadIdx.x & (KM_THREADS - 1));
    const bool live = r < rows;
    double xi[F], xm2[F];
#pragma unroll
    for (int c = 0; c < F; ++c) xi[c] = live ? X[(size_t)(row0 + r) * F + c] : 0.0;
#pragma unroll
    for (int c = 0; c < F; ++c) xm2[c] = -2.0 * xi[c];
    const double total = km_block_sum<F, false>(xm2, km_norm<F>(xi), live ? row0 + r : -1, 0, N, X, nullptr, km_sm);
    if (live && threadIdx.x < KM_THREADS) rowsum[r] = total;
}

template <int F>
static int km_launch_rowsum(long long N, const double* X, long long row0, long long rows, double* rowsum, cudaStream_t st) {
    const size_t smem = km_block_smem(F);
    cudaError_t e = cudaFuncSetAttribute(bf_km_rowsum_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_kmedoids_rowsum: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const unsigned grid = (unsigned)((rows + KM_THREADS - 1) / KM_THREADS);
    bf_km_rowsum_kernel<F><<<grid, KM_BLOCK, smem, st>>>(N, X, row0, rows, rowsum);
    return 0;
}

extern "C" int bf_kmedoids_rowsum(long long N, int f, const double* X, long long row0, long long rows,
                                  double* rowsum, void* stream) {
    BF_CHECK_ARG(N > 0 && X && rowsum && row0 >= 0 && rows >= 0 && row0 + rows <= N, BF_ERR_ARG,
                 "bf_kmedoids_rowsum: bad argument");
    BF_CHECK_ARG(f >= 1 && f <= KM_MAX_F, BF_ERR_SIZE, "bf_kmedoids_rowsum: f=%d outside 1..%d", f, KM_MAX_F);
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    switch (f) {
#define BF_CASE(FF) case FF: rc = km_launch_rowsum<FF>(N, X, row0, rows, rowsum, st); break;
        BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8)
#undef BF_CASE
    }
    if (rc) return rc;
    BF_CHECK_LAUNCH("bf_kmedoids_rowsum");
    return 0;
}

// labels[i] = first argmin over the medoids (np.argmin(D[medoids, :], axis=0))
__global__ void __launch_bounds__(KM_THREADS) bf_km_assign_kernel(long long N, int f, int k,
                                                                  const double* __restrict__ X,
                                                                  const long long* __restrict__ medoid_idx,
                                                                  int32_t* __restrict__ labels) {
    extern __shared__ double med[];   // k * (f + 1): coordinates + squared norm
    const int fs = f + 1;
    for (int c = threadIdx.x; c < k; c += KM_THREADS) {
        const double* xr = X + (size_t)medoid_idx[c] * f;
        for (int a = 0; a < f; ++a) med[c * fs + a] = xr[a];
        med[c * fs + f] = km_norm_rt(xr, f);
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * KM_THREADS + threadIdx.x;
    if (i >= N) return;
    double xm2[KM_MAX_F];
    for (int a = 0; a < f; ++a) xm2[a] = -2.0 * X[(size_t)i * f + a];
    const double ni = km_norm_rt(X + (size_t)i * f, f);
    double best = INFINITY;
    int lab = 0;
    for (int c = 0; c < k; ++c) {
        double dot = xm2[0] * med[c * fs];
        for (int a = 1; a < f; ++a) dot = fma(xm2[a], med[c * fs + a], dot);
        const double d = (medoid_idx[c] == i) ? 0.0 : sqrt(fmax((dot + med[c * fs + f]) + ni, 0.0));   // row = medoid
        if (d < best) {
            best = d;
            lab = c;
        }
    }
    labels[i] = lab;
}

extern "C" int bf_kmedoids_assign(long long N, int f, int k, const double* X, const long long* medoid_idx,
                                  int32_t* labels, void* stream) {
    BF_CHECK_ARG(N > 0 && k > 0 && X && medoid_idx && labels, BF_ERR_ARG, "bf_kmedoids_assign: bad argument");
    BF_CHECK_ARG(f >= 1 && f <= KM_MAX_F, BF_ERR_SIZE, "bf_kmedoids_assign: f=%d outside 1..%d", f, KM_MAX_F);
    BF_CHECK_ARG((size_t)k * (f + 1) * sizeof(double) <= 48 * 1024, BF_ERR_SIZE, "bf_kmedoids_assign: k*f too large");
    const unsigned grid = (unsigned)((N + KM_THREADS - 1) / KM_THREADS);
    bf_km_assign_kernel<<<grid, KM_THREADS, (size_t)k * (f + 1) * sizeof(double), (cudaStream_t)stream>>>(
        N, f, k, X, medoid_idx, labels);
    BF_CHECK_LAUNCH("bf_kmedoids_assign");
    return 0;
}

// ---- stable counting sort of the labels: perm = rows ordered by (label, row), seg = cluster offsets ----
// np.where(labels == c)[0] for every c in one pass (the member order defines cost ties and the
// "first argmin").  Three small launches: per-block histograms, one scan over (cluster, block),
// and a scatter in which every row counts the equal labels before it inside its block.
#define KM_SORT_BLOCK 1024

__global__ void __launch_bounds__(256) bf_km_hist_kernel(long long N, int k, const int32_t* __restrict__ labels,
                                                         int nblk, long long* __restrict__ counts) {
    extern __shared__ int hist[];          // k
    for (int c = threadIdx.x; c < k; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * KM_SORT_BLOCK;
    for (int t = threadIdx.x; t < KM_SORT_BLOCK; t += blockDim.x)
        if (i0 + t < N) atomicAdd(&hist[labels[i0 + t]], 1);
    __syncthreads();
    for (int c = threadIdx.x; c < k; c += blockDim.x) counts[(size_t)c * nblk + blockIdx.x] = hist[c];
}

// exclusive scan over the cluster-major (cluster, block) counts; one CTA, chunked Hillis-Steele
__global__ void __launch_bounds__(1024) bf_km_scan_kernel(int k, int nblk, long long* __restrict__ counts,
                                                          long long* __restrict__ seg) {
    __shared__ long long buf[1024];
    __shared__ long long carry;
    const long long total = (long long)k * nblk;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long e0 = 0; e0 < total; e0 += 1024) {
        const long long e = e0 + threadIdx.x;
        const long long v = e < total ? counts[e] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int s = 1; s < 1024; s <<= 1) {
            const long long add = threadIdx.x >= s ? buf[threadIdx.x - s] : 0;
            __syncthreads();
            buf[threadIdx.x] += add;
            __syncthreads();
        }
        const long long excl = carry + buf[threadIdx.x] - v;
        if (e < total) {
            counts[e] = excl;
            if (e % nblk == 0) seg[e / nblk] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) seg[k] = carry;
}

__global__ void __launch_bounds__(256) bf_km_scatter_kernel(long long N, const int32_t* __restrict__ labels, int nblk,
                                                            const long long* __restrict__ offsets,
                                                            long long* __restrict__ perm) {
    __shared__ int lab[KM_SORT_BLOCK];
    const long long i0 = (long long)blockIdx.x * KM_SORT_BLOCK;
    for (int t = threadIdx.x; t < KM_SORT_BLOCK; t += blockDim.x) lab[t] = (i0 + t < N) ? labels[i0 + t] : -1;
    __syncthreads();
    for (int t = threadIdx.x; t < KM_SORT_BLOCK; t += blockDim.x) {
        if (i0 + t >= N) continue;
        const int me = lab[t];
        int before = 0;
        for (int u = 0; u < t; ++u) before += (lab[u] == me);
        perm[offsets[(size_t)me * nblk + blockIdx.x] + before] = i0 + t;
    }
}

extern "C" size_t bf_kmedoids_sort_workspace(long long N, int k) {
    const long long nblk = (N + KM_SORT_BLOCK - 1) / KM_SORT_BLOCK;
    return sizeof(long long) * (size_t)(nblk > 0 ? nblk : 1) * (size_t)(k > 0 ? k : 1);
}

extern "C" int bf_kmedoids_sort_labels(long long N, int k, const int32_t* labels, long long* perm, long long* seg,
                                       void* workspace, void* stream) {
    BF_CHECK_ARG(N > 0 && k > 0 && labels && perm && seg && workspace, BF_ERR_ARG, "bf_kmedoids_sort_labels: bad argument");
    BF_CHECK_ARG((size_t)k * sizeof(int) <= 48 * 1024, BF_ERR_SIZE, "bf_kmedoids_sort_labels: k too large");
    const long long nblk = (N + KM_SORT_BLOCK - 1) / KM_SORT_BLOCK;
    BF_CHECK_ARG(nblk <= 0x7fffffffLL, BF_ERR_SIZE, "bf_kmedoids_sort_labels: N too large");
    cudaStream_t st = (cudaStream_t)stream;
    long long* counts = (long long*)workspace;
    bf_km_hist_kernel<<<(unsigned)nblk, 256, (size_t)k * sizeof(int), st>>>(N, k, labels, (int)nblk, counts);
    bf_km_scan_kernel<<<1, 1024, 0, st>>>(k, (int)nblk, counts, seg);
    bf_km_scatter_kernel<<<(unsigned)nblk, 256, 0, st>>>(N, labels, (int)nblk, counts, perm);
    BF_CHECK_LAUNCH("bf_kmedoids_sort_labels");
    return 0;
}

// ---- in-cluster cost: cost[p] = sum over the segment of p of |x_perm[p] - x_perm[q]|, q in segment order ----
// Work is cut into chunks of KM_THREADS consecutive positions of ONE cluster, numbered cluster-major;
// chunk b is computed by the call with b % nparts == part (ranks of a multi-GPU job take one part
// each; untouched positions keep the caller's zeros so the parts combine by addition).  Every chunk is
// one km_block_sum over the members of its cluster.
template <int F>
__global__ void __launch_bounds__(KM_BLOCK) bf_km_cost_kernel(int k, const double* __restrict__ X,
                                                              const long long* __restrict__ perm,
                                                              const long long* __restrict__ seg, int part, int nparts,
                                                              double* __restrict__ cost) {
    extern __shared__ __align__(16) double km_sm[];
    __shared__ long long s_q0, s_q1, s_p0;
    if (threadIdx.x == 0) {
        // chunk index -> (cluster, first position): walk the clusters (k is small)
        const long long b = (long long)blockIdx.x * nparts + part;
        long long acc = 0;
        s_q0 = s_q1 = s_p0 = -1;
        for (int c = 0; c < k; ++c) {
            const long long n_c = seg[c + 1] - seg[c];
            const long long chunks = (n_c + KM_THREADS - 1) / KM_THREADS;
            if (b < acc + chunks) {
                s_q0 = seg[c];
                s_q1 = seg[c + 1];
                s_p0 = seg[c] + (b - acc) * KM_THREADS;
                break;
            }
            acc += chunks;
        }
    }
    __syncthreads();
    const long long q0 = s_q0, q1 = s_q1;
    if (q0 < 0) return;                         // past the last chunk
    const long long p = s_p0 + (threadIdx.x & (KM_THREADS - 1));
    const bool live = p < q1;
    double xi[F], xm2[F];
    {
        const long long me = live ? perm[p] : 0;
#pragma unroll
        for (int c = 0; c < F; ++c) xi[c] = live ? X[(size_t)me * F + c] : 0.0;
#pragma unroll
        for (int c = 0; c < F; ++c) xm2[c] = -2.0 * xi[c];
    }
    const double total = km_block_sum<F, true>(xm2, km_norm<F>(xi), live ? p : -1, q0, q1, X, perm, km_sm);
    if (live && threadIdx.x < KM_THREADS) cost[p] = total;
}

template <int F>
static int km_launch_cost(unsigned grid, int k, const double* X, const long long* perm, const long long* seg, int part,
                          int nparts, double* cost, cudaStream_t st) {
    const size_t smem = km_block_smem(F);
    cudaError_t e = cudaFuncSetAttribute(bf_km_cost_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_kmedoids_cluster_cost: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    bf_km_cost_kernel<F><<<grid, KM_BLOCK, smem, st>>>(k, X, perm, seg, part, nparts, cost);
    return 0;
}

extern "C" int bf_kmedoids_cluster_cost_part(long long N, int f, int k, const double* X, const long long* perm,
                                             const long long* seg, int part, int nparts, double* cost, void* stream) {
    BF_CHECK_ARG(N > 0 && k > 0 && X && perm && seg && cost, BF_ERR_ARG, "bf_kmedoids_cluster_cost: bad argument");
    BF_CHECK_ARG(f >= 1 && f <= KM_MAX_F, BF_ERR_SIZE, "bf_kmedoids_cluster_cost: f=%d outside 1..%d", f, KM_MAX_F);
    BF_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, BF_ERR_ARG, "bf_kmedoids_cluster_cost: bad part %d of %d", part, nparts);
    // at most ceil(N / KM_THREADS) + k chunks in total; blocks past the last chunk exit
    const long long max_chunks = (N + KM_THREADS - 1) / KM_THREADS + k;
    const unsigned grid = (unsigned)((max_chunks + nparts - 1) / nparts);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = 0;
    switch (f) {
#define BF_CASE(FF) case FF: rc = km_launch_cost<FF>(grid, k, X, perm, seg, part, nparts, cost, st); break;
        BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8)
#undef BF_CASE
    }
    if (rc) return rc;
    BF_CHECK_LAUNCH("bf_kmedoids_cluster_cost");
    return 0;
}

extern "C" int bf_kmedoids_cluster_cost(long long N, int f, int k, const double* X, const long long* perm,
                                        const long long* seg, double* cost, void* stream) {
    return bf_kmedoids_cluster_cost_part(N, f, k, X, perm, seg, 0, 1, cost, stream);
}

// one CTA per cluster: first argmin of the in-cluster costs; move the medoid only if strictly
// better than the current one (sklearn_extra _update_medoid_idxs_in_place)
__global__ void __launch_bounds__(256) bf_km_update_kernel(const long long* __restrict__ perm,
                                                           const long long* __restrict__ seg,
                                                           const double* __restrict__ cost,
                                                           long long* __restrict__ medoid_idx,
                                                           int32_t* __restrict__ changed) {
    const int c = blockIdx.x;
    const long long q0 = seg[c], q1 = seg[c + 1];
    if (q1 <= q0) return;         // empty cluster: medoid kept
    const long long cur = medoid_idx[c];
    double bv = INFINITY;
    long long bp = 0x7fffffffffffffffLL;
    long long curp = 0x7fffffffffffffffLL;
    for (long long q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const double v = cost[q];
        if (v < bv || (v == bv && q < bp)) {
            bv = v;
            bp = q;
        }
        if (perm[q] == cur && q < curp) curp = q;
    }
    __shared__ double sv[256];
    __shared__ long long sp[256], sc[256];
    sv[threadIdx.x] = bv;
    sp[threadIdx.x] = bp;
    sc[threadIdx.x] = curp;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sv[threadIdx.x + s];
            const long long q = sp[threadIdx.x + s];
            if (v < sv[threadIdx.x] || (v == sv[threadIdx.x] && q < sp[threadIdx.x])) {
                sv[threadIdx.x] = v;
                sp[threadIdx.x] = q;
            }
            if (sc[threadIdx.x + s] < sc[threadIdx.x]) sc[threadIdx.x] = sc[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // np.argmax(cluster_idxs == medoid) is 0 when the medoid is not in its own cluster
        const long long cp = (sc[0] == 0x7fffffffffffffffLL) ? q0 : sc[0];
        if (sv[0] < cost[cp]) {
            medoid_idx[c] = perm[sp[0]];
            atomicAdd(changed, 1);
        }
    }
}

extern "C" int bf_kmedoids_update(long long N, int k, const long long* perm, const long long* seg,
                                  const double* cost, long long* medoid_idx, int32_t* changed, void* stream) {
    BF_CHECK_ARG(N > 0 && k > 0 && perm && seg && cost && medoid_idx && changed, BF_ERR_ARG,
                 "bf_kmedoids_update: bad argument");
    bf_km_update_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(perm, seg, cost, medoid_idx, changed);
    BF_CHECK_LAUNCH("bf_kmedoids_update");
    return 0;
}

// ---- medoid update when the in-cluster cost chunks are split over `nparts` ranks -----------
// Every rank reduces ITS chunks to one record per cluster, the ranks all-gather the k records (3 doubles
// each: 24 k bytes per rank instead of an all-reduce over the N' costs), and a final pass picks the global
// first argmin.  record[c] = (smallest cost among my positions of cluster c, that position, cost at the
// current medoid's position if that position is mine) with +inf for "none"; positions are < 2^53 and travel
// as doubles.  The winner is the smallest cost and, among equal costs, the smallest position - numpy's first
// argmin over the whole cluster - whatever the number of ranks.
__global__ void __launch_bounds__(256) bf_km_update_partial_kernel(int k, const long long* __restrict__ perm,
                                                                   const long long* __restrict__ seg,
                                                                   const double* __restrict__ cost,
                                                                   const long long* __restrict__ medoid_idx,
                                                                   int part, int nparts, double* __restrict__ rec) {
    const int c = blockIdx.x;
    const long long q0 = seg[c], q1 = seg[c + 1];
    __shared__ long long chunk_base;
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int cc = 0; cc < c; ++cc) acc += (seg[cc + 1] - seg[cc] + KM_THREADS - 1) / KM_THREADS;
        chunk_base = acc;
    }
    __syncthreads();
    const long long cur = medoid_idx[c];
    double bv = INFINITY;
    long long bp = 0x7fffffffffffffffLL, curp = 0x7fffffffffffffffLL;
    for (long long q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const bool mine = ((chunk_base + (q - q0) / KM_THREADS) % nparts) == part;
        if (mine) {
            const double v = cost[q];
            if (v < bv || (v == bv && q < bp)) {
                bv = v;
                bp = q;
            }
        }
        if (perm[q] == cur && q < curp) curp = q;
    }
    __shared__ double sv[256];
    __shared__ long long sp[256], sc[256];
    sv[threadIdx.x] = bv;
    sp[threadIdx.x] = bp;
    sc[threadIdx.x] = curp;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sv[threadIdx.x + s];
            const long long q = sp[threadIdx.x + s];
            if (v < sv[threadIdx.x] || (v == sv[threadIdx.x] && q < sp[threadIdx.x])) {
                sv[threadIdx.x] = v;
                sp[threadIdx.x] = q;
            }
            if (sc[threadIdx.x + s] < sc[threadIdx.x]) sc[threadIdx.x] = sc[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double cur_cost = INFINITY;
        if (q1 > q0) {
            // np.argmax(cluster_idxs == medoid) is 0 when the medoid is not in its own cluster
            const long long cp = (sc[0] == 0x7fffffffffffffffLL) ? q0 : sc[0];
            if (((chunk_base + (cp - q0) / KM_THREADS) % nparts) == part) cur_cost = cost[cp];
        }
        rec[c * 3 + 0] = sv[0];
        rec[c * 3 + 1] = sp[0] == 0x7fffffffffffffffLL ? INFINITY : (double)sp[0];
        rec[c * 3 + 2] = cur_cost;
    }
}

__global__ void bf_km_update_final_kernel(int k, int nparts, const double* __restrict__ recs,
                                          const long long* __restrict__ perm, long long* __restrict__ medoid_idx,
                                          int32_t* __restrict__ changed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double bv = INFINITY, bp = INFINITY, cur = INFINITY;
    for (int r = 0; r < nparts; ++r) {
        const double* rec = recs + ((size_t)r * k + c) * 3;
        if (rec[0] < bv || (rec[0] == bv && rec[1] < bp)) {
            bv = rec[0];
            bp = rec[1];
        }
        cur = fmin(cur, rec[2]);          // exactly one rank owns the current medoid's position
    }
    if (bv < cur) {
        medoid_idx[c] = perm[(long long)bp];
        atomicAdd(changed, 1);
    }
}

extern "C" int bf_kmedoids_update_partial(long long N, int k, const long long* perm, const long long* seg,
                                          const double* cost, const long long* medoid_idx, int part, int nparts,
                                          double* records, void* stream) {
    BF_CHECK_ARG(N > 0 && k > 0 && perm && seg && cost && medoid_idx && records, BF_ERR_ARG,
                 "bf_kmedoids_update_partial: bad argument");
    BF_CHECK_ARG(nparts >= 1 && part >= 0 && part < nparts, BF_ERR_ARG, "bf_kmedoids_update_partial: bad part %d of %d", part, nparts);
    bf_km_update_partial_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(k, perm, seg, cost, medoid_idx, part, nparts, records);
    BF_CHECK_LAUNCH("bf_kmedoids_update_partial");
    return 0;
}

extern "C" int bf_kmedoids_update_final(int k, int nparts, const double* records, const long long* perm,
                                        long long* medoid_idx, int32_t* changed, void* stream) {
    BF_CHECK_ARG(k > 0 && nparts >= 1 && records && perm && medoid_idx && changed, BF_ERR_ARG,
                 "bf_kmedoids_update_final: bad argument");
    bf_km_update_final_kernel<<<(k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(k, nparts, records, perm, medoid_idx, changed);
    BF_CHECK_LAUNCH("bf_kmedoids_update_final");
    return 0;
}

// One iteration of the alternate loop behind ONE host call (6 launches): assign -> stable label sort ->
// in-cluster costs of this part -> medoid update (nparts == 1) or this part's records (nparts > 1; the
// caller all-gathers them and calls bf_kmedoids_update_final).
extern "C" int bf_kmedoids_iteration(long long N, int f, int k, const double* X, long long* medoid_idx,
                                     int32_t* labels, long long* perm, long long* seg, void* workspace, int part,
                                     int nparts, double* cost, double* records, int32_t* changed, void* stream) {
    int rc = bf_kmedoids_assign(N, f, k, X, medoid_idx, labels, stream);
    if (rc) return rc;
    rc = bf_kmedoids_sort_labels(N, k, labels, perm, seg, workspace, stream);
    if (rc) return rc;
    rc = bf_kmedoids_cluster_cost_part(N, f, k, X, perm, seg, part, nparts, cost, stream);
    if (rc) return rc;
    if (nparts == 1) return bf_kmedoids_update(N, k, perm, seg, cost, medoid_idx, changed, stream);
    BF_CHECK_ARG(records != nullptr, BF_ERR_ARG, "bf_kmedoids_iteration: records buffer needed when nparts > 1");
    return bf_kmedoids_update_partial(N, k, perm, seg, cost, medoid_idx, part, nparts, records, stream);
}
