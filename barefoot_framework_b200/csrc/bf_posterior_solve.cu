// GP posterior + acquisitions for FEW candidates per set - the literal ROM stage of the reference
// (barefoot.py:1001-1073: every work item factorises its own fused GP and evaluates only
// sampleCount ~ 50 candidates, util.py:243-262).  With N << n the explicit L^-1 that the sweep kernel
// (bf_posterior.cu) amortises over 1e5..1e6 candidates costs more than everything else, so this path
// never forms it:
//     V = L^-1 [ k*(X, candidates) | y ]          blocked forward substitution on the FP64 tensor cores
//     mean_c = v_c . t,   var_c = amp - v_c . v_c     with t = L^-1 y the last column of V
// (k*^T K^-1 y = (L^-1 k*) . (L^-1 y); gpModel.py:50-54 via george's cho_solve).  Inputs are the Cholesky
// factor and the inverses of its 32 x 32 diagonal blocks as bf_cholesky_batched_dinv leaves them.
//
// One CTA = 8 NA - 1 candidates (+ the y column) of one set, 8 warps: warp w owns rows 8 (w & 3) .. + 7 of
// every row block and the 8-column groups (w >> 2), (w >> 2) + 2, ...  Row block I:
//     s = sum_{K < I} L_IK v_K;  v_I = L_II^-1 (rhs_I - s)
// The A operand streams from L2 through a cp.async ring in UNITS of up to two horizontally adjacent 32 x 32
// tiles of a row block (L[I, K..K+1], contiguous in the row-major factor), then L_II^-1: one wait + barrier per
// unit, i.e. 80 instead of 136 synchronisations per CTA at n = 500.  (Measured in round 2: the kernel time did
// not move - 1.7 us per set at n = 500, 50 candidates - so the barriers were not what holds the DMMA pipe at
// 30 %: a CTA's 126 us split into ~35 us of DMMA issue at 8 warps, ~23 us of right-hand-side generation at the
// FP64 latency of 8 warps per SM, 16 serial diagonal steps and the L2 stream; 222 kB of shared memory - 131 kB
// of it the solution block V - keep it at one CTA per SM.)  The right-hand sides are generated into shared memory in DMMA B-fragment order and overwritten in
// place by the solution.
#include "bf_common.cuh"

#define SV_THREADS 256
#define SV_RING 4                       // ring slots, one unit each
#define SV_TLD 68                       // unit leading dimension (doubles): 64 columns + pad, conflict-free fragment reads

struct SolveParams {
    int kern, n, d, np, nb, S;
    long long N;
    const double *Xs, *X, *inv_l2, *amp, *L, *dinv, *y, *scale, *shift;
    long long y_stride;
    const int32_t* set_hp;
    double *mu_out, *var_out;
    int acq_mask, spread_is_sqrt;
    double curr_max, xi, kt;
    double* part_val;
    long long* part_idx;
    int tiles;             // CTAs per set
};

template <int NA>
__global__ void __launch_bounds__(SV_THREADS, 1) bf_posterior_solve_kernel(SolveParams p) {
    constexpr int NCOL = 8 * NA, CPC = NCOL - 1, NAW = NA / 2, PARTS = SV_THREADS / NCOL;
    extern __shared__ __align__(16) double sv_sm[];
    double* ring = sv_sm;                                        // SV_RING x 32 x SV_TLD
    double* V = ring + SV_RING * 32 * SV_TLD;                    // np x NCOL: [(K NA + na) 256 + r 8 + c]
    double* sbuf = V + (size_t)p.np * NCOL;                      // NA x 256
    double* xs_s = sbuf + NA * 256;                              // BF_MAX_DIM x NCOL, pre-scaled candidates
    double* xt_s = xs_s + BF_MAX_DIM * NCOL;                     // 32 x BF_MAX_DIM, pre-scaled training rows
    double* red = xt_s + 32 * BF_MAX_DIM;                        // 2 x PARTS x NCOL
    double* w_s = red + 2 * PARTS * NCOL;                        // BF_MAX_DIM
    double* tab_s = w_s + BF_MAX_DIM;                            // BF_EXP_TAB_N

    const int set = blockIdx.y, tile = blockIdx.x;
    const int hp = p.set_hp ? p.set_hp[set] : set;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, kk = lane & 3;
    const int tt = warp & 3, nh = warp >> 2;
    const int np = p.np, nb = p.nb, dd = p.d;
    const long long c_base = (long long)tile * CPC;
    const double amp = p.amp[hp];
    const double* L = p.L + (size_t)set * np * np;
    const double* dinv = p.dinv + (size_t)set * nb * 1024;
    const double* yv = p.y + (size_t)set * p.y_stride;

    // ---- unit stream: row block I = units (I, K, w) with w = min(2, I - K) for K = 0, 2, .. < I, then the
    //      inverse of the diagonal block ----
    int U = 0;
    for (int i = 0; i < nb; ++i) U += ((i + 1) >> 1) + 1;
    int pI = 0, pK = 0;
    auto request = [&](int slot) {
        const bool diag = pK >= pI;
        const int w = diag ? 1 : (pI - pK >= 2 ? 2 : 1);
        const double* src = diag ? dinv + (size_t)pI * 1024 : L + (size_t)(32 * pI) * np + 32 * pK;
        const int ld = diag ? 32 : np;
        double* dst = ring + (size_t)slot * 32 * SV_TLD;
        const int cpr = 16 * w;                              // 16-byte chunks per row
        for (int e = threadIdx.x; e < 32 * cpr; e += SV_THREADS) {
            const int row = e / cpr, ch = e - row * cpr;
            bf_cp_async16(dst + row * SV_TLD + 2 * ch, src + (size_t)row * ld + 2 * ch);
        }
        if (diag) { ++pI; pK = 0; } else { pK += w; }
    };
    for (int t = 0; t < SV_RING - 1; ++t) {          // in flight while the right-hand sides are generated
        if (t < U) request(t);
        bf_cp_async_commit();
    }

    // ---- right-hand sides: k*(X, candidates) and y ----
    if (threadIdx.x < dd) w_s[threadIdx.x] = bf_metric_scale(p.kern, p.inv_l2[(size_t)hp * dd + threadIdx.x]);
    for (int e = threadIdx.x; e < BF_EXP_TAB_N; e += SV_THREADS) tab_s[e] = bf_exp_tab_g[e];
    __syncthreads();
    for (int e = threadIdx.x; e < dd * NCOL; e += SV_THREADS) {
        const int a = e / NCOL, col = e - a * NCOL;
        const long long gi = c_base + col;
        xs_s[a * NCOL + col] = (col < CPC && gi < p.N) ? p.Xs[gi * dd + a] * w_s[a] : 0.0;
    }
    for (int K = 0; K < nb; ++K) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * dd; e += SV_THREADS) {
            const int rr = e / dd, a = e - rr * dd;
            const int j = 32 * K + rr;
            xt_s[rr * BF_MAX_DIM + a] = j < p.n ? p.X[(size_t)j * dd + a] * w_s[a] : 0.0;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * NCOL; e += SV_THREADS) {
            const int rr = e / NCOL, col = e - rr * NCOL;
            const int j = 32 * K + rr;
            const long long gi = c_base + col;
            double val = 0.0;
            if (j < p.n) {
                if (col == CPC) {
                    val = yv[j];
                } else if (gi < p.N) {
                    double q = 0.0;
                    for (int a = 0; a < dd; ++a) {
                        const double df = xs_s[a * NCOL + col] - xt_s[rr * BF_MAX_DIM + a];
                        q = fma(df, df, q);
                    }
                    val = amp * bf_radial_q(p.kern, q, tab_s);
                }
            }
            V[(size_t)(K * NA + (col >> 3)) * 256 + rr * 8 + (col & 7)] = val;
        }
    }

    // ---- forward substitution ----
    double acc[NAW][4][2];
#pragma unroll
    for (int i = 0; i < NAW; ++i)
#pragma unroll
        for (int h = 0; h < 4; ++h) acc[i][h][0] = acc[i][h][1] = 0.0;
    int I = 0, K = 0;
    for (int t = 0; t < U; ++t) {
        bf_cp_async_wait<SV_RING - 2>();
        __syncthreads();                               // unit t, the right-hand sides and v_{I-1} visible
        if (t + SV_RING - 1 < U) request((t + SV_RING - 1) % SV_RING);
        bf_cp_async_commit();
        const double* A = ring + (size_t)(t % SV_RING) * 32 * SV_TLD + (8 * tt + r) * SV_TLD + kk;
        if (K < I) {
            const int w = I - K >= 2 ? 2 : 1;
            for (int kt = 0; kt < w; ++kt) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const double a = A[32 * kt + 4 * ks];
#pragma unroll
                    for (int i = 0; i < NAW; ++i) {
                        const double b = V[(size_t)((K + kt) * NA + nh + 2 * i) * 256 + (4 * ks + kk) * 8 + r];
                        bf_dmma(acc[i][ks & 3][0], acc[i][ks & 3][1], a, b);
                    }
                }
            }
            K += w;
        } else {
#pragma unroll
            for (int i = 0; i < NAW; ++i) {
                const int na = nh + 2 * i;
                const double2 rhs = *reinterpret_cast<const double2*>(V + (size_t)(I * NA + na) * 256 + (8 * tt + r) * 8 + 2 * kk);
                const double s0 = (acc[i][0][0] + acc[i][1][0]) + (acc[i][2][0] + acc[i][3][0]);
                const double s1 = (acc[i][0][1] + acc[i][1][1]) + (acc[i][2][1] + acc[i][3][1]);
                *reinterpret_cast<double2*>(sbuf + na * 256 + (8 * tt + r) * 8 + 2 * kk) = make_double2(rhs.x - s0, rhs.y - s1);
#pragma unroll
                for (int h = 0; h < 4; ++h) acc[i][h][0] = acc[i][h][1] = 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double a = A[4 * ks];
#pragma unroll
                for (int i = 0; i < NAW; ++i) {
                    const double b = sbuf[(nh + 2 * i) * 256 + (4 * ks + kk) * 8 + r];
                    bf_dmma(acc[i][ks & 3][0], acc[i][ks & 3][1], a, b);
                }
            }
#pragma unroll
            for (int i = 0; i < NAW; ++i) {
                const int na = nh + 2 * i;
                const double v0 = (acc[i][0][0] + acc[i][1][0]) + (acc[i][2][0] + acc[i][3][0]);
                const double v1 = (acc[i][0][1] + acc[i][1][1]) + (acc[i][2][1] + acc[i][3][1]);
                *reinterpret_cast<double2*>(V + (size_t)(I * NA + na) * 256 + (8 * tt + r) * 8 + 2 * kk) = make_double2(v0, v1);
#pragma unroll
                for (int h = 0; h < 4; ++h) acc[i][h][0] = acc[i][h][1] = 0.0;
            }
            ++I;
            K = 0;
        }
    }
    __syncthreads();

    // ---- mean = v . t, |v|^2; rows split over PARTS thread groups, added in group order ----
    {
        const int col = threadIdx.x % NCOL, part = threadIdx.x / NCOL;
        double ss = 0.0, dot = 0.0;
        for (int j = part; j < p.n; j += PARTS) {
            const int Kb = j >> 5, rr = j & 31;
            const double v = V[(size_t)(Kb * NA + (col >> 3)) * 256 + rr * 8 + (col & 7)];
            const double tj = V[(size_t)(Kb * NA + NA - 1) * 256 + rr * 8 + 7];
            ss = fma(v, v, ss);
            dot = fma(v, tj, dot);
        }
        red[part * NCOL + col] = ss;
        red[(PARTS + part) * NCOL + col] = dot;
    }
    __syncthreads();
    if (warp != 0) return;
    BfBest best[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) best[k] = bf_best_init();
    BfTop2 top = bf_top2_init();
    const double sc = p.scale ? p.scale[set] : 1.0;
    const double sh = p.shift ? p.shift[set] : 0.0;
    {
        const int col = lane;
        const long long gi = c_base + col;
        if (col < CPC && gi < p.N) {
            double s = 0.0, mean = 0.0;
#pragma unroll
            for (int q = 0; q < PARTS; ++q) {
                s += red[q * NCOL + col];
                mean += red[(PARTS + q) * NCOL + col];
            }
            const double mu = mean * sc + sh;
            const double var = (amp - s) * (sc * sc);
            if (p.mu_out) p.mu_out[(size_t)set * p.N + gi] = mu;
            if (p.var_out) p.var_out[(size_t)set * p.N + gi] = var;
            if (p.acq_mask) {
                const double spread = p.spread_is_sqrt ? sqrt(var) : var;
                if (p.acq_mask & BF_ACQ_EI)
                    bf_best_take(best[BF_SLOT_EI], (mu - p.curr_max - p.xi) * bf_norm_pdf(mu) + var * bf_norm_cdf(mu), gi);
                if (p.acq_mask & BF_ACQ_PI)
                    bf_best_take(best[BF_SLOT_PI], bf_norm_cdf((mu - p.curr_max - p.xi) / spread), gi);
                if (p.acq_mask & BF_ACQ_UCB) bf_best_take(best[BF_SLOT_UCB], mu + p.kt * spread, gi);
                if (p.acq_mask & BF_ACQ_GREEDY) bf_top2_merge(top, mu, gi, -INFINITY);
            }
        }
    }
    if (!p.acq_mask) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) best[k] = bf_best_warp(best[k]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double t1 = __shfl_xor_sync(0xffffffffu, top.t1, m);
        const long long i1 = __shfl_xor_sync(0xffffffffu, top.i1, m);
        const double t2 = __shfl_xor_sync(0xffffffffu, top.t2, m);
        bf_top2_merge(top, t1, i1, t2);
    }
    if (lane == 0) {
        double* pv = p.part_val + ((size_t)set * p.tiles + tile) * BF_NSLOT_VAL;
        long long* pi = p.part_idx + ((size_t)set * p.tiles + tile) * BF_NSLOT_IDX;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pv[k] = best[k].v;
            pi[k] = best[k].i;
        }
        pv[BF_SLOT_GREEDY] = top.t1;
        pi[BF_SLOT_GREEDY] = top.i1;
        pv[BF_SLOT_TOP2] = top.t2;
    }
}

// ---- host side ------------------------------------------------------------------------------
static size_t solve_smem_bytes(int np, int NA) {
    const int ncol = 8 * NA, parts = SV_THREADS / ncol;
    return sizeof(double) * ((size_t)SV_RING * 32 * SV_TLD + (size_t)np * ncol + (size_t)NA * 256 + (size_t)BF_MAX_DIM * ncol +
                             32 * BF_MAX_DIM + 2 * (size_t)parts * ncol + BF_MAX_DIM + BF_EXP_TAB_N);
}

static int solve_na(int n) {
    if (!bf_cholesky_dinv_supported(n)) return 0;
    const int np = bf_padded_n(n);
    const size_t cap = (size_t)227 * 1024 - 256;
    if (solve_smem_bytes(np, 4) <= cap) return 4;
    if (solve_smem_bytes(np, 2) <= cap) return 2;
    return 0;
}

extern "C" int bf_posterior_solve_cols(int n) {
    const int na = solve_na(n);
    return na ? 8 * na - 1 : 0;
}

extern "C" int bf_posterior_solve_parts(int n, long long N) {
    const int cols = bf_posterior_solve_cols(n);
    if (cols == 0 || N <= 0) return 0;
    return (int)((N + cols - 1) / cols);
}

template <int NA>
static int launch_solve(const SolveParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(bf_posterior_solve_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_gp_posterior_solve: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    bf_posterior_solve_kernel<NA><<<grid, SV_THREADS, smem, st>>>(p);
    BF_CHECK_LAUNCH("bf_gp_posterior_solve");
    return 0;
}

extern "C" int bf_gp_posterior_solve(int kern, long long N, int n, int d, int S, const double* Xs, const double* X,
                                     const double* inv_l2, const double* amp, const int32_t* set_hp, const double* L,
                                     const double* dinv, const double* y, long long y_stride, const double* scale,
                                     const double* shift, double* mu_out, double* var_out, int acq_mask,
                                     int spread_is_sqrt, double curr_max, double xi, double kt, double* part_val,
                                     long long* part_idx, void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_gp_posterior_solve: bad kern %d", kern);
    BF_CHECK_ARG(N >= 0 && n > 0 && S >= 0, BF_ERR_ARG, "bf_gp_posterior_solve: bad sizes N=%lld n=%d S=%d", N, n, S);
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_gp_posterior_solve: nDim %d outside 1..%d", d, BF_MAX_DIM);
    BF_CHECK_ARG(Xs && X && inv_l2 && amp && L && dinv && y, BF_ERR_ARG, "bf_gp_posterior_solve: null input");
    BF_CHECK_ARG(y_stride == 0 || y_stride >= n, BF_ERR_ARG, "bf_gp_posterior_solve: bad y_stride");
    BF_CHECK_ARG(!acq_mask || (part_val && part_idx), BF_ERR_ARG, "bf_gp_posterior_solve: acq_mask without partial buffers");
    BF_CHECK_ARG(((uintptr_t)L & 15) == 0 && ((uintptr_t)dinv & 15) == 0, BF_ERR_ALIGN,
                 "bf_gp_posterior_solve: L and dinv must be 16-byte aligned");
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_gp_posterior_solve: at most 65535 sets per call (got %d)", S);
    if (N == 0 || S == 0) return 0;
    const int na = solve_na(n);
    BF_CHECK_ARG(na > 0, BF_ERR_SIZE, "bf_gp_posterior_solve: n=%d too large for the solve path", n);
    SolveParams p;
    p.kern = kern; p.n = n; p.d = d; p.np = bf_padded_n(n); p.nb = p.np / 32; p.S = S; p.N = N;
    p.Xs = Xs; p.X = X; p.inv_l2 = inv_l2; p.amp = amp; p.L = L; p.dinv = dinv; p.y = y; p.y_stride = y_stride;
    p.scale = scale; p.shift = shift; p.set_hp = set_hp; p.mu_out = mu_out; p.var_out = var_out;
    p.acq_mask = acq_mask; p.spread_is_sqrt = spread_is_sqrt; p.curr_max = curr_max; p.xi = xi; p.kt = kt;
    p.part_val = part_val; p.part_idx = part_idx;
    p.tiles = bf_posterior_solve_parts(n, N);
    dim3 grid(p.tiles, S);
    const size_t smem = solve_smem_bytes(p.np, na);
    if (na == 4) return launch_solve<4>(p, grid, smem, (cudaStream_t)stream);
    return launch_solve<2>(p, grid, smem, (cudaStream_t)stream);
}
