// Fused kernel-matrix generation + batched Cholesky for MANY sets (the literal ROM stage factorises
// 150 000 fused GPs; gpModel.py:38-42 / reificationFusion.py:160-166 per work item).
//
// Why another factorisation kernel.  The right-looking kernel in bf_setup.cu re-reads and re-writes every
// trailing 32 x 32 tile once per panel: 680 tiles x 16 kB + the panel columns + the kernel matrix written by
// bf_kernel_matrix_lower and read back = ~14 MB of HBM traffic per 500 x 500 set, and with one CTA per SM the
// 148 matrices in flight (2 MB each) do not fit the 126 MB L2 - measured 3.8 us per set at S = 1000, i.e.
// ~3.7 TB/s of traffic for 8 TFLOP/s of DMMA work (profiles/r1c_setup_throughput.json).  This kernel is
// LEFT-looking by block column and never stores anything but the finished factor:
//   * K tiles are generated straight into the accumulator registers (no K in HBM at all);
//   * tile (i, j) accumulates sum_{k<j} L[i,k] L[j,k]^T in registers over the whole k range, streaming the
//     finished blocks of L (read-only, 680 tile reads of 8 kB per set, half the bytes and no read-modify-write)
//     through a register ring; the row block L[j, :] every tile of a column shares hits L1;
//   * the diagonal tile is factorised in registers by warp 0 (bf_block32.cuh) and L_jj^-1 applied to the
//     column's tiles with the accumulators re-used as the DMMA A operand through shuffles (no staging);
//   * warp 0 runs one block column AHEAD: after tile (j+1, j) it factorises the next diagonal tile while the
//     other warps finish column j, and with ~35 kB of shared memory and 4 warps per CTA THREE sets share an
//     SM, so a set's serial 32 x 32 factorisations (single-warp chains of ~10 us, 16 per set) overlap the
//     other sets' tensor work.
// Algorithmic work per set: n^3 / 3 flop (DMMA) + n^2 / 2 kernel evaluations (FP64 pipe, ~35 instructions each);
// HBM: ~5.4 MB read + 1.1 MB written at n = 500.
#include "bf_block32.cuh"
#include "bf_common.cuh"

#define BF_LL_WARPS 4
#define BF_LL_THREADS (BF_LL_WARPS * 32)
#define BF_LL_RING 2
#define BF_LL_CTAS_PER_SM 2
#define BF_LL_MAX_NP 1024

// Phase profile (tools/prof_chol_ll.cu builds with -DBF_LL_PROFILE): cycles of warp 0 (slots 0-6) and warp 1 (7-9)
#ifdef BF_LL_PROFILE
__device__ long long bf_ll_prof[16];
#define LL_T0() long long ll_t_ = clock64()
#define LL_MARK(slot)                                                                              \
    do {                                                                                           \
        const long long now_ = clock64();                                                          \
        if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)&bf_ll_prof[slot], (unsigned long long)(now_ - ll_t_)); \
        ll_t_ = now_;                                                                              \
    } while (0)
#else
#define LL_T0()
#define LL_MARK(slot)
#endif

struct LLParams {
    int kern, n, d, np, nb;
    const double *X, *inv_l2, *amp, *yerr;
    const int32_t* set_hp;
    int yerr_n;
    long long yerr_stride;
    double* L;
    int32_t* info;
    double* dinv;
    int zero_upper;
};

// acc += L[i-block rows, cols 0 .. 8 ngroups) * L[j-block rows, same cols]^T, both row-major in global memory
// (ngroups is a multiple of 4: whole 32-column blocks).  Lane (r = lane >> 2, q = lane & 3) loads the column
// PAIR (8 g + 2 q, 8 g + 2 q + 1) of its rows as one 16-byte word; the x halves of a group form one DMMA k-step
// (columns 8g + {0,2,4,6}), the y halves the next ({1,3,5,7}) - any permutation of the contraction index is
// allowed as long as A and B use the same one.  Two-slot register ring: group g + 1 is in flight while group g
// is multiplied (32 DMMAs); the other CTAs of the SM cover the rest of the L2 latency.
__device__ __forceinline__ void ll_accumulate(double (&c)[4][4][2], const double* __restrict__ Ai,
                                              const double* __restrict__ Bj, int np, int ngroups) {
    if (ngroups <= 0) return;
    double2 ra[BF_LL_RING][4], rb[BF_LL_RING][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        ra[0][t] = *reinterpret_cast<const double2*>(Ai + (size_t)(8 * t) * np);
        rb[0][t] = *reinterpret_cast<const double2*>(Bj + (size_t)(8 * t) * np);
    }
    for (int g0 = 0; g0 < ngroups; g0 += BF_LL_RING) {
#pragma unroll
        for (int s = 0; s < BF_LL_RING; ++s) {
            const int gn = g0 + s + 1;
            if (gn < ngroups) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    ra[(s + 1) % BF_LL_RING][t] = *reinterpret_cast<const double2*>(Ai + (size_t)(8 * t) * np + 8 * gn);
                    rb[(s + 1) % BF_LL_RING][t] = *reinterpret_cast<const double2*>(Bj + (size_t)(8 * t) * np + 8 * gn);
                }
            }
            // 16 independent accumulators per k-step: never two DMMAs on the same accumulator back to back
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], ra[s][t].x, rb[s][u].x);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 4; ++u) bf_dmma(c[t][u][0], c[t][u][1], ra[s][t].y, rb[s][u].y);
        }
    }
}

// Pull rows [32 ib, 32 ib + 32) x columns [0, ncols) of the factor into L2 ahead of the register ring: with three
// hundred sets in flight the finished blocks of a set have left the L2 by the time they are needed again, and a
// one-group register ring cannot cover HBM latency.  One bulk prefetch per lane (= per row), no registers held.
__device__ __forceinline__ void ll_prefetch_rows(const double* __restrict__ A, int np, int ib, int ncols) {
    if (ncols <= 0) return;
    const int lane = threadIdx.x & 31;
    const double* row = A + (size_t)(32 * ib + lane) * np;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(ncols * 8) : "memory");
}

__device__ __forceinline__ void ll_zero(double (&c)[4][4][2]) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) c[t][u][0] = c[t][u][1] = 0.0;
}

// c <- K[i-block, j-block] - c, K generated from the pre-scaled coordinates in shared memory
// (xs[a * np + row]); identity padding beyond n; the noise of george's compute() on the diagonal.
// D > 0: nDim known at compile time - the lane's 4 row points stay in registers and the 8 elements of a column
// atom (4 rows x 2 columns) are independent evaluation chains (the FP64 pipe, not latency, is then the limit);
// D == 0: run-time nDim (9..16), coordinates re-read from shared memory.
template <int D>
__device__ __forceinline__ void ll_kernel_tile(double (&c)[4][4][2], const LLParams& p, const double* __restrict__ xs,
                                               const double* __restrict__ tab, double amp,
                                               const double* __restrict__ yerr, int ib, int jb) {
    constexpr int DR = D ? D : 1;
    const int dd = D ? D : p.d;
    const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
    const int np = p.np, n = p.n, kern = p.kern;
    double xr[4][DR];
    if (D) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int a = 0; a < DR; ++a) xr[t][a] = xs[a * np + 32 * ib + 8 * t + r];
    }
    // all 32 squared distances of the lane first, then 32 independent radial evaluations: a lone warp per
    // scheduler has nothing else to hide the FP64 latency with
    double qq[4][4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int col0 = 32 * jb + 8 * u + 2 * q;
        if (D) {
            double xc[2][DR];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int a = 0; a < DR; ++a) xc[e][a] = xs[a * np + col0 + e];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < DR; ++a) {
                        const double df = xr[t][a] - xc[e][a];
                        s = fma(df, df, s);
                    }
                    qq[t][u][e] = s;
                }
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int e = 0; e < 2; ++e) qq[t][u][e] = 0.0;
            for (int a = 0; a < dd; ++a) {
                const double x0 = xs[a * np + col0], x1 = xs[a * np + col0 + 1];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double xrow = xs[a * np + 32 * ib + 8 * t + r];
                    const double d0 = xrow - x0, d1 = xrow - x1;
                    qq[t][u][0] = fma(d0, d0, qq[t][u][0]);
                    qq[t][u][1] = fma(d1, d1, qq[t][u][1]);
                }
            }
        }
    }
    const bool interior = 32 * ib + 31 < n && 32 * jb + 31 < n && ib != jb;
    if (interior) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < 2; ++e) c[t][u][e] = amp * bf_radial_q(kern, qq[t][u][e], tab) - c[t][u][e];
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = 32 * ib + 8 * t + r, col = 32 * jb + 8 * u + 2 * q + e;
                    double v = amp * bf_radial_q(kern, qq[t][u][e], tab);
                    if (row >= n || col >= n) v = (row == col) ? 1.0 : 0.0;
                    else if (row == col) {
                        const double er = p.yerr_n == 1 ? yerr[0] : yerr[row];
                        const double sd = sqrt(er * er + BF_TINY);          // george folds the jitter into yerr
                        v += sd * sd;
                    }
                    c[t][u][e] = v - c[t][u][e];
                }
    }
}

// c <- K[i-block, j-block] - c with K read from the factor buffer itself (in place: every tile of K is read exactly
// once, right before the same warp overwrites it with the finished block of L)
__device__ __forceinline__ void ll_load_tile(double (&c)[4][4][2], const double* __restrict__ A, int np, int ib, int jb) {
    const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
    const double* src = A + (size_t)(32 * ib + r) * np + 32 * jb + 2 * q;
    double2 k[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) k[t][u] = *reinterpret_cast<const double2*>(src + (size_t)(8 * t) * np + 8 * u);
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            c[t][u][0] = k[t][u].x - c[t][u][0];
            c[t][u][1] = k[t][u].y - c[t][u][1];
        }
}

// P = C L_jj^-T for a tile below the diagonal, stored as L[ib, jb]: the accumulators become the DMMA A operand
// through shuffles (C layout: row r, columns 2q, 2q+1 of atom u -> A layout: row r, column q of k-step ks).
__device__ __forceinline__ void ll_solve_store(double (&c)[4][4][2], const double (*Li)[33], double* __restrict__ A,
                                               int np, int ib, int jb) {
    const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
    double pacc[4][4][2];
    ll_zero(pacc);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int src = (lane & ~3) | (2 * (ks & 1) + (q >> 1));
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const double v0 = __shfl_sync(0xffffffffu, c[t][ks >> 1][0], src);
            const double v1 = __shfl_sync(0xffffffffu, c[t][ks >> 1][1], src);
            a[t] = (q & 1) ? v1 : v0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = Li[8 * u + r][4 * ks + q];      // (L_jj^-T)[k][col] = Li[col][k]
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) bf_dmma(pacc[t][u][0], pacc[t][u][1], a[t], b[u]);
    }
    double* dst = A + (size_t)(32 * ib + r) * np + 32 * jb + 2 * q;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<double2*>(dst + (size_t)(8 * t) * np + 8 * u) = make_double2(pacc[t][u][0], pacc[t][u][1]);
}

template <int D, bool GEN>
__global__ void __launch_bounds__(BF_LL_THREADS, BF_LL_CTAS_PER_SM) bf_cholesky_ll_kernel(LLParams p) {
    extern __shared__ __align__(16) double ll_sm[];
    double* xs = ll_sm;                                             // d x np, pre-scaled by the metric
    double* tab = xs + (size_t)p.d * p.np;                          // BF_EXP_TAB_N
    double (*Ld)[33] = reinterpret_cast<double (*)[33]>(tab + BF_EXP_TAB_N);
    double (*Li0)[33] = Ld + 32;                                    // L_jj^-1, double buffered by column parity
    __shared__ double rdiag[32];
    __shared__ double w_s[BF_MAX_DIM];
    __shared__ int bad;
    const int set = blockIdx.x;
    const int hp = (GEN && p.set_hp) ? p.set_hp[set] : set;
    const int np = p.np, nb = p.nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
    double* A = p.L + (size_t)set * np * np;
    const double* yerr = GEN ? p.yerr + (p.yerr_n == 1 ? 0 : (size_t)set * p.yerr_stride) : nullptr;
    const double amp = GEN ? p.amp[hp] : 0.0;

    if (threadIdx.x == 0) bad = 0;
    if (GEN && threadIdx.x < p.d) w_s[threadIdx.x] = bf_metric_scale(p.kern, p.inv_l2[(size_t)hp * p.d + threadIdx.x]);
    for (int e = threadIdx.x; e < BF_EXP_TAB_N; e += BF_LL_THREADS) tab[e] = bf_exp_tab_g[e];
    __syncthreads();
    if (GEN) {
        for (int e = threadIdx.x; e < np * p.d; e += BF_LL_THREADS) {
            const int i = e / p.d, a = e - i * p.d;
            xs[a * np + i] = i < p.n ? p.X[(size_t)i * p.d + a] * w_s[a] : 0.0;
        }
    }
    __syncthreads();

    // Diagonal tile jd (warp 0): accumulate over the finished columns, generate K, factorise in registers,
    // store L_jd,jd and leave its inverse in the Li buffer of that column's parity.
    auto diagonal = [&](int jd) {
        double c[4][4][2];
        ll_zero(c);
        const double* Bd = A + (size_t)(32 * jd + r) * np + 2 * q;
        LL_T0();
        ll_accumulate(c, Bd, Bd, np, 4 * jd);
        LL_MARK(3);
        if (GEN) ll_kernel_tile<D>(c, p, xs, tab, amp, yerr, jd, jd); else ll_load_tile(c, A, np, jd, jd);
        LL_MARK(1);
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                Ld[8 * t + r][8 * u + 2 * q] = c[t][u][0];
                Ld[8 * t + r][8 * u + 2 * q + 1] = c[t][u][1];
            }
        __syncwarp();
        const int b = rl_factor_block_blocked(Ld, rdiag, lane);
        LL_MARK(4);
        if (b && lane == 0 && bad == 0) bad = 32 * jd + b;
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) A[(size_t)(32 * jd + jj) * np + 32 * jd + lane] = Ld[jj][lane];
        double (*Li)[33] = Li0 + 32 * (jd & 1);
        rl_invert_block_blocked(Ld, rdiag, Li, lane);
        LL_MARK(5);
        if (p.dinv) {
            double* dv = p.dinv + ((size_t)set * nb + jd) * 1024;
#pragma unroll 8
            for (int jj = 0; jj < 32; ++jj) dv[jj * 32 + lane] = Li[jj][lane];      // each lane wrote this column itself
        }
        __syncwarp();
    };

    if (warp == 0) diagonal(0);
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        // Column j below the diagonal: tile (j + 1, j) belongs to warp 0, which then runs AHEAD - it factorises
        // the next diagonal tile (a single-warp dependency chain of ~10 us) while warps 1..3 finish the column.
        const double (*Li)[33] = Li0 + 32 * (j & 1);
        const double* Bj = A + (size_t)(32 * j + r) * np + 2 * q;
        if (warp == 0) {
            if (j + 1 < nb) {
                double c[4][4][2];
                ll_zero(c);
                if (j + 2 < nb) ll_prefetch_rows(A, np, j + 2, 32 * j);     // next column's look-ahead tile (its last block comes later)
                LL_T0();
                ll_accumulate(c, A + (size_t)(32 * (j + 1) + r) * np + 2 * q, Bj, np, 4 * j);
                LL_MARK(0);
                if (GEN) ll_kernel_tile<D>(c, p, xs, tab, amp, yerr, j + 1, j); else ll_load_tile(c, A, np, j + 1, j);
                LL_MARK(1);
                ll_solve_store(c, Li, A, np, j + 1, j);
                LL_MARK(2);
                __syncwarp();                   // the tile just stored is read back by the other lanes below
                diagonal(j + 1);
            }
            {
                LL_T0();
                __syncthreads();
                LL_MARK(6);
            }
            continue;
        } else {
            LL_T0();
            for (int i = j + 1 + warp; i < nb; i += BF_LL_WARPS - 1) {
                double c[4][4][2];
                ll_zero(c);
                // the rows of this warp's next tile: of this column, or its first one of the next column
                const int inext = i + BF_LL_WARPS - 1 < nb ? i + BF_LL_WARPS - 1 : j + 2 + warp;
                if (inext < nb) ll_prefetch_rows(A, np, inext, 32 * j);
                ll_accumulate(c, A + (size_t)(32 * i + r) * np + 2 * q, Bj, np, 4 * j);
                if (GEN) ll_kernel_tile<D>(c, p, xs, tab, amp, yerr, i, j); else ll_load_tile(c, A, np, i, j);
                ll_solve_store(c, Li, A, np, i, j);
            }
            if (warp == 1) LL_MARK(7);
        }
        {
            LL_T0();
            __syncthreads();                    // column j of L complete and visible; L_{j+1,j+1}^-1 ready
            if (warp == 1) LL_MARK(8);
        }
    }
    if (p.zero_upper) {                         // dense consumers (keep_dense): zeros in the blocks above the diagonal
        for (int e = threadIdx.x; e < np * (np / 2); e += BF_LL_THREADS) {
            const int i = e / (np / 2), j2 = (e - i * (np / 2)) * 2;
            if ((j2 >> 5) > (i >> 5)) *reinterpret_cast<double2*>(A + (size_t)i * np + j2) = make_double2(0.0, 0.0);
        }
    }
    if (threadIdx.x == 0 && p.info) p.info[set] = bad;
}

static size_t ll_smem_bytes(int np, int d) { return sizeof(double) * ((size_t)d * np + BF_EXP_TAB_N + 3 * 32 * 33); }

template <int D, bool GEN>
static int ll_launch(const LLParams& p, int S, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(bf_cholesky_ll_kernel<D, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_cholesky_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    bf_cholesky_ll_kernel<D, GEN><<<S, BF_LL_THREADS, smem, st>>>(p);
    return 0;
}

// In-place variant: L holds K (lower blocks, bf_kernel_matrix_lower) on entry.
extern "C" int bf_cholesky_ll_inplace(int n, int S, double* K_inout_L, int32_t* info, double* dinv, int zero_upper,
                                      void* stream) {
    BF_CHECK_ARG(n > 0 && S >= 0 && K_inout_L, BF_ERR_ARG, "bf_cholesky_ll_inplace: bad argument");
    BF_CHECK_ARG(bf_padded_n(n) <= BF_LL_MAX_NP, BF_ERR_SIZE, "bf_cholesky_ll_inplace: n=%d too large", n);
    BF_CHECK_ARG(((uintptr_t)K_inout_L & 15) == 0, BF_ERR_ALIGN, "bf_cholesky_ll_inplace: buffer must be 16-byte aligned");
    if (S == 0) return 0;
    LLParams p;
    p.kern = 0; p.n = n; p.d = 1; p.np = bf_padded_n(n); p.nb = p.np / 32;
    p.X = nullptr; p.inv_l2 = nullptr; p.amp = nullptr; p.yerr = nullptr; p.set_hp = nullptr;
    p.yerr_n = 1; p.yerr_stride = 0;
    p.L = K_inout_L; p.info = info; p.dinv = dinv; p.zero_upper = zero_upper;
    const size_t smem = ll_smem_bytes(p.np, 1);      // the (unused) coordinate area keeps the carve-up of the generating variant
    int rc = ll_launch<1, false>(p, S, smem, (cudaStream_t)stream);
    if (rc) return rc;
    BF_CHECK_LAUNCH("bf_cholesky_ll_inplace");
    return 0;
}

extern "C" int bf_cholesky_fused_supported(int n, int d) {
    const int np = bf_padded_n(n);
    return (np <= BF_LL_MAX_NP && ll_smem_bytes(np, d) <= (size_t)200 * 1024) ? 1 : 0;
}

extern "C" int bf_cholesky_fused(int kern, int n, int d, int S, const double* X, const double* inv_l2,
                                 const double* amp, const int32_t* set_hp, const double* yerr, int yerr_n,
                                 long long yerr_stride, double* L, int32_t* info, double* dinv, int zero_upper,
                                 void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_cholesky_fused: bad kern %d", kern);
    BF_CHECK_ARG(n > 0 && S >= 0 && X && inv_l2 && amp && yerr && L, BF_ERR_ARG, "bf_cholesky_fused: bad argument");
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_cholesky_fused: nDim %d outside 1..%d", d, BF_MAX_DIM);
    BF_CHECK_ARG(yerr_n == 1 || yerr_n == n, BF_ERR_ARG, "bf_cholesky_fused: yerr_n must be 1 or n");
    BF_CHECK_ARG(bf_cholesky_fused_supported(n, d), BF_ERR_SIZE, "bf_cholesky_fused: n=%d too large", n);
    BF_CHECK_ARG(((uintptr_t)L & 15) == 0, BF_ERR_ALIGN, "bf_cholesky_fused: L must be 16-byte aligned");
    if (S == 0) return 0;
    LLParams p;
    p.kern = kern; p.n = n; p.d = d; p.np = bf_padded_n(n); p.nb = p.np / 32;
    p.X = X; p.inv_l2 = inv_l2; p.amp = amp; p.yerr = yerr; p.set_hp = set_hp;
    p.yerr_n = yerr_n; p.yerr_stride = yerr_stride;
    p.L = L; p.info = info; p.dinv = dinv; p.zero_upper = zero_upper;
    const size_t smem = ll_smem_bytes(p.np, d);
    int rc = 0;
    switch (d) {
        // this kernel is an option, not the default path: one compile-time instantiation (nDim = 6, the shape all
        // measurements quote) and the run-time-dimension form keep the build short
        case 6: rc = ll_launch<6, true>(p, S, smem, (cudaStream_t)stream); break;
        default: rc = ll_launch<0, true>(p, S, smem, (cudaStream_t)stream); break;
    }
    if (rc) return rc;
    BF_CHECK_LAUNCH("bf_cholesky_fused");
    return 0;
}
