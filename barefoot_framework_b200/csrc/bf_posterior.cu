// GP posterior mean/variance + elementwise acquisitions, fused (subsystems 1, 2 and 4 of the
// north star).  Replaces gp_model.predict_var (gpModel.py:50-54), predict_low_order /
// predict_fused_GP (reificationFusion.py:189-206) and expected_improvement /
// probability_improvement / upper_conf_bound / greedy (acquisitionFunc.py:98-214,
// util.py:655-665) for all hyper-parameter sets x all candidates in one launch.
//
// One tile = NT = 8 NTA candidates of one set (NT = 24 at n = 500).  Phases per tile:
//   phase A  cross-covariance tile k*[np x NT] generated straight into shared memory in the
//            DMMA B-fragment layout (never touches HBM); mean = k* . alpha folded in.
//   phase B  V = L^-1 k* on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4).  L^-1 is
//            block lower triangular in a pre-packed fragment layout streamed from L2 through a
//            register ring; warps own balanced (b, nb-1-b) row-block pairs.
//            Only the column sums of V^2 are kept.
//   phase C  var = amp - |V|^2, de-normalise, acquisition values, running (max, first index)
//            and top-2 of the mean.
// Two kernels: bf_posterior_dual_kernel (two tiles per CTA in lock step, used whenever two B tiles
// fit in shared memory) and bf_posterior_kernel (one tile per CTA, larger n).
#include <stdlib.h>

#include "bf_common.cuh"

// Phase timing (tools/dbg_skip.sh) needs a build with -DBF_PHASE_PROFILE; the shipped library has no such switch.
#ifdef BF_PHASE_PROFILE
#define BF_SKIP(p, bit) ((p).debug_skip & (bit))
// cycle accounting per warp (BF_DEBUG_SKIP=64): slot 0 = wait at the loop-top barrier, 1 = phase C, 2 = phase A,
// 3 = wait at the barrier before phase B, 4 = phase B; one CTA prints its 16 warps when it retires
#define BF_PROF_DECL long long prof_t[5] = {0, 0, 0, 0, 0}; long long prof_c = clock64(); const long long prof_start = prof_c
#define BF_PROF_MARK(k) do { const long long c_ = clock64(); prof_t[k] += c_ - prof_c; prof_c = c_; } while (0)
#define BF_PROF_REPORT                                                                                              \
    do {                                                                                                            \
        if ((p.debug_skip & 64) && blockIdx.x == 700 && blockIdx.y == 0 && lane == 0)                               \
            printf("prof warp %2d pairs %lld total %lld wait_top %lld C %lld A %lld wait_mid %lld B %lld\n", warp,   \
                   (long long)cap, clock64() - prof_start, prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4]); \
    } while (0)
#else
#define BF_SKIP(p, bit) 0
#define BF_PROF_DECL
#define BF_PROF_MARK(k)
#define BF_PROF_REPORT
#endif

struct BfTrue { static constexpr bool value = true; };
struct BfFalse { static constexpr bool value = false; };

struct PostParams {
    int kern, n, d, np, nb, NT, NTA, S;
    int sh, g_skip;       // np - n leading padding rows of the packed operand (bf_pack_store); whole k-groups of them
    long long N;
    const double *Xs, *Xw, *inv_l2, *amp, *linv, *alpha, *scale, *shift;
    const int32_t* set_hp;
    unsigned* pair_ctr;   // per-set work counters of the dual kernel (NULL: static split)
    size_t linv_stride;
    double *mu_out, *var_out;
    int acq_mask, spread_is_sqrt;
    double curr_max, xi, kt;
    double* part_val;
    long long* part_idx;
    int tiles;            // partial slots per set (= CTAs per set)
    int tpc;              // tiles per CTA
    long long ntiles;     // candidate tiles per set
#ifdef BF_PHASE_PROFILE
    int debug_skip;       // phase timing builds only: 1 = no phase B, 2 = no phase A, 4 = no phase C (results are wrong)
#endif
};

// ---- phase A ------------------------------------------------------------------------------
// B-tile layout: cell ((g * NTA + a) * 32 + lane) holds k*[4g + (lane & 3)][8a + (lane >> 2)], the
// DMMA B fragment of atom a at k-group g.  The calling warp generates k-groups g_first, g_first +
// g_stride, ... < g_end and leaves its share of the mean in row `slot` of mpart.
// xw: the training inputs of this set's hyper-parameter row, pre-scaled by the metric (bf_prescale_train);
// the candidates in xs_s are pre-scaled too, so a squared distance is one subtraction and one FMA per
// dimension and nothing per training row.  The loop body is straight-line code (padding rows are computed
// from row 0 and masked, not branched around), so the NTA evaluation chains of a lane interleave.
// The padding is in front (bf_pack_store): k-groups below p.g_skip are all padding and are never generated
// (the callers start at p.g_skip; those rows of the B tile are zeroed once per CTA).
template <int D, int NTA, int KERN>
__device__ __forceinline__ void gen_tile_k(const PostParams& p, const double* __restrict__ xw,
                                           const double* __restrict__ xs_s, double* __restrict__ Bt,
                                           double* __restrict__ mpart, const double* __restrict__ alpha,
                                           double amp, const double* __restrict__ tab, int g_first, int g_stride,
                                           int g_end, int slot) {
    constexpr int NT = 8 * NTA;
    constexpr int DR = D ? D : BF_MAX_DIM;   // D == 0: nDim is a run-time value
    const int dd = D ? D : p.d;
    const int lane = threadIdx.x & 31;
    const int kk = lane & 3, nn = lane >> 2;
    const double* xc = xs_s + nn;            // this lane's candidates (pre-scaled): xc[a * NT + 8 * at]
    double macc[NTA];
#pragma unroll
    for (int at = 0; at < NTA; ++at) macc[at] = 0.0;

    // training row and weight of a k-group (padding rows: row 0, masked below)
    auto load_row = [&](int g, double (&xj)[DR], double& aj0) {
        const int j = 4 * g + kk - p.sh;
        const int jc = j >= 0 ? j : 0;
#pragma unroll
        for (int a = 0; a < DR; ++a) xj[a] = a < dd ? __ldg(xw + (size_t)jc * dd + a) : 0.0;
        aj0 = __ldg(alpha + jc);
    };
    // NTA independent branch-free evaluation chains per lane.  With the padding in front only ONE k-group can
    // hold padding rows - group g_skip, when np - n is not a multiple of 4 - so only that step carries the row
    // masks (MASKED); every other step is the bare evaluation (the loop is as much issue-bound as FP64-bound).
    // (Requesting the next row one iteration ahead was measured at +0.3-0.5 % but costs 14 registers, which
    // this form spills - not kept.)
    int g = g_first;
    auto step = [&](auto masked) {
        constexpr bool MASKED = decltype(masked)::value;
        const int j = 4 * g + kk - p.sh;                                 // padded row 4g + kk is training row j
        const long long keep = (!MASKED || j >= 0) ? -1LL : 0LL;         // padding rows: value and weight 0
        double xj[DR], a0;
        load_row(g, xj, a0);
        const double aj = MASKED ? __longlong_as_double(__double_as_longlong(a0) & keep) : a0;
        double* out = Bt + (size_t)g * NTA * 32 + lane;
        double q[NTA];
#pragma unroll
        for (int at = 0; at < NTA; ++at) q[at] = 0.0;
#pragma unroll
        for (int a = 0; a < DR; ++a) {
            if (a < dd) {
#pragma unroll
                for (int at = 0; at < NTA; ++at) {
                    const double df = xc[a * NT + 8 * at] - xj[a];                  // both pre-scaled
                    q[at] = fma(df, df, q[at]);
                }
            }
        }
#pragma unroll
        for (int at = 0; at < NTA; ++at) {
            const double v = amp * bf_radial_q_k<KERN>(q[at], tab);
            const double val = MASKED ? __longlong_as_double(__double_as_longlong(v) & keep) : v;
            macc[at] = fma(aj, val, macc[at]);
            out[at * 32] = val;
        }
    };
    if (g < g_end && g == p.g_skip && (p.sh & 3)) {
        step(BfTrue{});
        g += g_stride;
    }
    for (; g < g_end; g += g_stride) step(BfFalse{});
    // mean partials: sum over the 4 kk lanes, then one slot per warp (fixed-order sum later)
#pragma unroll
    for (int at = 0; at < NTA; ++at) {
        double v = macc[at];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (kk == 0) mpart[slot * NT + 8 * at + nn] = v;
    }
}

// the covariance function is a compile-time constant of the loop (one copy of the loop per function)
template <int D, int NTA>
__device__ __forceinline__ void gen_tile(const PostParams& p, const double* __restrict__ xw,
                                         const double* __restrict__ xs_s, double* __restrict__ Bt,
                                         double* __restrict__ mpart, const double* __restrict__ alpha,
                                         double amp, const double* __restrict__ tab, int g_first, int g_stride,
                                         int g_end, int slot) {
    if (p.kern == BF_KERN_SE)
        gen_tile_k<D, NTA, BF_KERN_SE>(p, xw, xs_s, Bt, mpart, alpha, amp, tab, g_first, g_stride, g_end, slot);
    else if (p.kern == BF_KERN_M32)
        gen_tile_k<D, NTA, BF_KERN_M32>(p, xw, xs_s, Bt, mpart, alpha, amp, tab, g_first, g_stride, g_end, slot);
    else
        gen_tile_k<D, NTA, BF_KERN_M52>(p, xw, xs_s, Bt, mpart, alpha, amp, tab, g_first, g_stride, g_end, slot);
}

// ---- phase B ------------------------------------------------------------------------------
// One warp, one row block b, NA column atoms starting at atom a0: accumulates the column sums
// of (L^-1[b,:] k*)^2 into cs.  The last 8 k-groups are the diagonal 32x32 block: atoms whose
// rows lie entirely above the columns of the group are skipped.
template <int NA, int R, int SREM>
__device__ __forceinline__ void mma_block(const double* __restrict__ Aset, const double* __restrict__ Bt,
                                          int NTA, int a0, int b, int g_skip, double (&cs)[NA][2]) {
    const int lane = threadIdx.x & 31;
    double c[4][NA][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int a = 0; a < NA; ++a) c[t][a][0] = c[t][a][1] = 0.0;

    const int gstride = NTA * 32;           // doubles per k-group in the B tile
    int ng_full = 8 * b;                    // groups left of the diagonal block
    // the first g_skip k-groups of every row are padding (zero columns of the operand): whole rounds of R
    // groups are stepped over, the rest (s_rem groups) is left out of the first round below
    const int skip_rounds = min(g_skip / R, ng_full / R);
    const int s_rem = g_skip - skip_rounds * R;
    ng_full -= skip_rounds * R;
    const double2* Ap = reinterpret_cast<const double2*>(Aset + bf_linv_block_off(b)) + lane + (size_t)skip_rounds * R * 64;
    const double* Bp = Bt + (size_t)a0 * 32 + lane + (size_t)skip_rounds * R * gstride;

    // A-operand ring: R (4 or 8) k-groups in registers; group g + R - 1 is requested while group g
    // is multiplied, so R - 1 groups (4 NA DMMAs each) cover the L2 latency.  All trip counts and
    // ring slots are compile-time so no DMMA of the main loop sits behind a run-time predicate.
    static_assert(R == 4 || R == 8, "ring depth must divide the 8 k-groups of a row block");
    double2 ring[R][2];
    auto load_group = [&](int slot, int g) {
        ring[slot][0] = __ldg(Ap + (size_t)g * 64);
        ring[slot][1] = __ldg(Ap + (size_t)g * 64 + 32);
    };
    auto mul_group = [&](int slot, int g) {
        double bb[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) bb[a] = Bp[(size_t)g * gstride + a * 32];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            bf_dmma(c[0][a][0], c[0][a][1], ring[slot][0].x, bb[a]);
            bf_dmma(c[1][a][0], c[1][a][1], ring[slot][0].y, bb[a]);
            bf_dmma(c[2][a][0], c[2][a][1], ring[slot][1].x, bb[a]);
            bf_dmma(c[3][a][0], c[3][a][1], ring[slot][1].y, bb[a]);
        }
    };
    if (SREM > 0 && ng_full > 0) {
        // first round with a compile-time number of padding groups (s_rem == SREM for every row block that
        // has full groups): the ring starts at the first real group, nothing is loaded for the padding
#pragma unroll
        for (int r = 0; r < R - 1; ++r) load_group((SREM + r) % R, SREM + r);
#pragma unroll
        for (int r = SREM; r < R; ++r) {
            load_group((r + R - 1) % R, r + R - 1);
            mul_group(r, r);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R - 1; ++r) load_group(r, r);      // at least the 8 diagonal groups exist
        if (ng_full > 0) {                  // first round; SREM < 0: groups below s_rem are padding (warp-uniform test)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                load_group((r + R - 1) % R, r + R - 1);
                if (SREM == 0 || r >= s_rem) mul_group(r, r);
            }
        }
    }
    for (int g0 = R; g0 < ng_full; g0 += R) {           // ng_full is a multiple of R
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int g = g0 + r;
            ring[(r + R - 1) % R][0] = __ldg(Ap + (size_t)(g + R - 1) * 64);   // in range: the diagonal block follows
            ring[(r + R - 1) % R][1] = __ldg(Ap + (size_t)(g + R - 1) * 64 + 32);
            double bb[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) bb[a] = Bp[(size_t)g * gstride + a * 32];
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                bf_dmma(c[0][a][0], c[0][a][1], ring[r][0].x, bb[a]);
                bf_dmma(c[1][a][0], c[1][a][1], ring[r][0].y, bb[a]);
                bf_dmma(c[2][a][0], c[2][a][1], ring[r][1].x, bb[a]);
                bf_dmma(c[3][a][0], c[3][a][1], ring[r][1].y, bb[a]);
            }
        }
    }
    // diagonal block: group gg touches columns 4gg..4gg+3 of the block; atom t (rows 8t..8t+7) is
    // non-zero only for gg < 2t + 2.  In row block 0 the leading padding groups sit inside the diagonal block:
    // stepping over them there too keeps the warp that owns the pair (nb - 1, 0) level with the others, which
    // skip them twice (SREM >= 0 only: the specialised copies of the 24-candidate tiles).
    const int dskip = (SREM >= 0 && b == 0) ? g_skip : 0;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) {
        const int g = ng_full + gg;
        if (gg + R - 1 < 8) {
            ring[(gg + R - 1) % R][0] = __ldg(Ap + (size_t)(g + R - 1) * 64);
            ring[(gg + R - 1) % R][1] = __ldg(Ap + (size_t)(g + R - 1) * 64 + 32);
        }
        if (SREM < 0 || gg >= dskip) {
            double bb[NA];
#pragma unroll
            for (int a = 0; a < NA; ++a) bb[a] = Bp[(size_t)g * gstride + a * 32];
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                if (gg < 2) bf_dmma(c[0][a][0], c[0][a][1], ring[gg % R][0].x, bb[a]);
                if (gg < 4) bf_dmma(c[1][a][0], c[1][a][1], ring[gg % R][0].y, bb[a]);
                if (gg < 6) bf_dmma(c[2][a][0], c[2][a][1], ring[gg % R][1].x, bb[a]);
                bf_dmma(c[3][a][0], c[3][a][1], ring[gg % R][1].y, bb[a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double s = cs[a][q];
#pragma unroll
            for (int t = 0; t < 4; ++t) s = fma(c[t][a][q], c[t][a][q], s);
            cs[a][q] = s;
        }
}

// Work units = (row-block pair) x (column slice of NA atoms); warps take units round-robin.  Unit
// u: slice = u % nslices, pair = u / nslices.  Each warp owns one slice for the whole phase so its
// column sums stay in registers; warps beyond the unit count idle.
template <int NA, int R, int SREM>
__device__ __forceinline__ void mma_phase_body(int nb, int NTA, int g_skip, const double* __restrict__ Aset,
                                       const double* __restrict__ Bt, double* __restrict__ spart, int warp) {
    const int lane = threadIdx.x & 31;
    const int NT = 8 * NTA;
    const int nslices = NTA / NA;
    const int nq = BF_WARPS / nslices;            // pair queues per slice
    const int slice = warp % nslices, wq = warp / nslices;
    if (wq >= nq) return;                         // spart was zeroed: idle warps contribute nothing
    const int npairs = (nb + 1) >> 1;
    const int a0 = slice * NA;
    double cs[NA][2];
#pragma unroll
    for (int a = 0; a < NA; ++a) cs[a][0] = cs[a][1] = 0.0;
    for (int pr = wq; pr < npairs; pr += nq) {
        const int b_hi = nb - 1 - pr;
        mma_block<NA, R, SREM>(Aset, Bt, NTA, a0, b_hi, g_skip, cs);
        if (pr != b_hi) mma_block<NA, R, SREM>(Aset, Bt, NTA, a0, pr, g_skip, cs);
    }
    // column c = 8 (a0 + a) + 2 (lane & 3) + q; reduce over the 8 row lanes (lane >> 2)
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double s = cs[a][q];
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            if ((lane >> 2) == 0) spart[wq * NT + 8 * (a0 + a) + 2 * (lane & 3) + q] = s;
        }
}

// single-buffer kernel: out-of-line copy per register budget
template <int NA, int MINB>
__device__ __noinline__ void mma_phase(int nb, int NTA, int g_skip, const double* __restrict__ Aset,
                                       const double* __restrict__ Bt, double* __restrict__ spart, int warp) {
    mma_phase_body<NA, 4, -1>(nb, NTA, g_skip, Aset, Bt, spart, warp);
}

// the 24-candidate tiles of the dual kernel (n <= 544): one copy per number of leading padding groups mod 4, so
// the first round of a row block is compile-time code like the rest (SREM < 0 above tests it at run time)
template <int SREM>
__device__ __noinline__ void mma_phase_3(int nb, int NTA, int g_skip, const double* __restrict__ Aset,
                                         const double* __restrict__ Bt, double* __restrict__ spart, int warp) {
    mma_phase_body<3, 4, SREM>(nb, NTA, g_skip, Aset, Bt, spart, warp);
}

template <int D, int NTA, int MINB>
__global__ void __launch_bounds__(BF_THREADS, MINB) bf_posterior_kernel(PostParams p) {
    constexpr int NT = 8 * NTA;
    extern __shared__ __align__(16) double smem[];
    const int dd = D ? D : p.d;
    double* Bt = smem;                                   // np * NT
    double* xs_s = Bt + (size_t)p.np * NT;               // d * NT
    double* mpart = xs_s + dd * NT;                      // BF_WARPS * NT
    double* spart = mpart + BF_WARPS * NT;               // BF_WARPS * NT
    double* w_s = spart + BF_WARPS * NT;                 // BF_MAX_DIM
    double* tab_s = w_s + BF_MAX_DIM;                    // BF_EXP_TAB_N
    __shared__ BfBest red_best[4][2];
    __shared__ BfTop2 red_top[2];

    const int tile = blockIdx.x, set = blockIdx.y;
    const int hp = p.set_hp ? p.set_hp[set] : set;
    const long long c_base = (long long)tile * NT;
    const double amp = p.amp[hp];

    for (int e = threadIdx.x; e < BF_WARPS * NT; e += BF_THREADS) spart[e] = 0.0;
    for (int e = threadIdx.x; e < p.g_skip * NTA * 32; e += BF_THREADS) Bt[e] = 0.0;   // all-padding k-groups
    if (threadIdx.x < dd) w_s[threadIdx.x] = bf_metric_scale(p.kern, p.inv_l2[(size_t)hp * dd + threadIdx.x]);
    for (int e = threadIdx.x; e < BF_EXP_TAB_N; e += BF_THREADS) tab_s[e] = bf_exp_tab_g[e];
    __syncthreads();
    for (int e = threadIdx.x; e < dd * NT; e += BF_THREADS) {        // candidates, pre-scaled by the metric
        const int a = e / NT, c = e - a * NT;
        const long long gi = c_base + c;
        xs_s[e] = gi < p.N ? p.Xs[gi * dd + a] * w_s[a] : 0.0;
    }
    __syncthreads();

    gen_tile<D, NTA>(p, p.Xw + (size_t)hp * p.n * dd, xs_s, Bt, mpart, p.alpha + (size_t)set * p.n, amp, tab_s, p.g_skip + (threadIdx.x >> 5), BF_WARPS,
                     p.np >> 2, threadIdx.x >> 5);
    __syncthreads();

    const double* Aset = p.linv + (size_t)set * p.linv_stride;
    {
        // widest column slice (a divisor of NTA) that still gives every pair queue a row-block pair
        const int npairs = (p.nb + 1) >> 1;
        int NA = 1;
        for (int cand = (MINB == 2 && NTA > 4) ? NTA / 2 : NTA; cand >= 1; --cand)   // 128-register budget
            if (NTA % cand == 0 && (cand == 1 || BF_WARPS / (NTA / cand) <= npairs)) { NA = cand; break; }
        switch (NA) {
            case 6: if (NTA == 6 && MINB == 1) mma_phase<6, MINB>(p.nb, NTA, p.g_skip, Aset, Bt, spart, threadIdx.x >> 5); break;
            case 3: if (NTA % 3 == 0) mma_phase<3, MINB>(p.nb, NTA, p.g_skip, Aset, Bt, spart, threadIdx.x >> 5); break;
            case 2: if (NTA % 2 == 0) mma_phase<2, MINB>(p.nb, NTA, p.g_skip, Aset, Bt, spart, threadIdx.x >> 5); break;
            default: mma_phase<1, MINB>(p.nb, NTA, p.g_skip, Aset, Bt, spart, threadIdx.x >> 5); break;
        }
    }
    __syncthreads();

    // ---- phase C ----
    const int c = threadIdx.x;
    BfBest best[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) best[k] = bf_best_init();
    BfTop2 top = bf_top2_init();
    if (c < 64) {
        const long long gi = c_base + c;
        if (c < NT && gi < p.N) {
            double mean = 0.0, s = 0.0;
#pragma unroll
            for (int w = 0; w < BF_WARPS; ++w) mean += mpart[w * NT + c];
            for (int w = 0; w < BF_WARPS; ++w) s += spart[w * NT + c];
            const double sc = p.scale ? p.scale[set] : 1.0;
            const double sh = p.shift ? p.shift[set] : 0.0;
            const double mu = mean * sc + sh;
            const double var = (amp - s) * (sc * sc);
            if (p.mu_out) p.mu_out[(size_t)set * p.N + gi] = mu;
            if (p.var_out) p.var_out[(size_t)set * p.N + gi] = var;
            if (p.acq_mask) {
                const double spread = p.spread_is_sqrt ? sqrt(var) : var;
                if (p.acq_mask & BF_ACQ_EI)
                    bf_best_take(best[BF_SLOT_EI],
                                 (mu - p.curr_max - p.xi) * bf_norm_pdf(mu) + var * bf_norm_cdf(mu), gi);
                if (p.acq_mask & BF_ACQ_PI)
                    bf_best_take(best[BF_SLOT_PI], bf_norm_cdf((mu - p.curr_max - p.xi) / spread), gi);
                if (p.acq_mask & BF_ACQ_UCB) bf_best_take(best[BF_SLOT_UCB], mu + p.kt * spread, gi);
                if (p.acq_mask & BF_ACQ_GREEDY) bf_top2_merge(top, mu, gi, -INFINITY);
            }
        }
        if (p.acq_mask) {
#pragma unroll
            for (int k = 0; k < 3; ++k) best[k] = bf_best_warp(best[k]);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                const double t1 = __shfl_xor_sync(0xffffffffu, top.t1, m);
                const long long i1 = __shfl_xor_sync(0xffffffffu, top.i1, m);
                const double t2 = __shfl_xor_sync(0xffffffffu, top.t2, m);
                bf_top2_merge(top, t1, i1, t2);
            }
            if ((c & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) red_best[k][c >> 5] = best[k];
                red_top[c >> 5] = top;
            }
        }
    }
    if (p.acq_mask) {
        __syncthreads();
        if (c == 0) {
            double* pv = p.part_val + ((size_t)set * p.tiles + tile) * BF_NSLOT_VAL;
            long long* pi = p.part_idx + ((size_t)set * p.tiles + tile) * BF_NSLOT_IDX;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                BfBest a = red_best[k][0];
                bf_best_take(a, red_best[k][1].v, red_best[k][1].i);
                pv[k] = a.v;
                pi[k] = a.i;
            }
            BfTop2 t = red_top[0];
            bf_top2_merge(t, red_top[1].t1, red_top[1].i1, red_top[1].t2);
            pv[BF_SLOT_GREEDY] = t.t1;
            pi[BF_SLOT_GREEDY] = t.i1;
            pv[BF_SLOT_TOP2] = t.t2;
        }
    }
}

// ---- dual-tile variant (the one used whenever two B tiles fit in shared memory) ---------------
// Measured on B200 (tools/dmma_probe.cu, profiles/): FP64 FMA and DMMA share one SM pipe, and an FP64
// instruction issued between DMMAs costs about three times its nominal pipe time, so overlapping the
// generator of one tile with the DMMA phase of another (two CTAs per SM, or warp-specialised
// producer/consumer warps) loses more than it hides.  This kernel therefore keeps the phases
// separate SM-wide: one CTA per SM, 16 warps in two groups of 8; each group owns one candidate tile
// and one shared-memory buffer, and all 16 warps move through phase A, phase B and phase C in
// lock step.  Phase A then runs at the full FP64 rate with 16 warps of branch-free evaluation
// chains, and phase B has 4 DMMA warps per scheduler to cover the L2 latency of the L^-1 stream.
// A CTA takes tile pairs of its set one at a time from the set's work counter (or walks a static chunk of
// consecutive pairs when the caller passes no counters) and keeps its running maxima in registers.
#define BF_DUAL_THREADS (2 * BF_THREADS)

template <int NTA>
__device__ __forceinline__ void dual_mma(int nb, int g_skip, const double* __restrict__ Aset,
                                         const double* __restrict__ Bt, double* __restrict__ spart, int warp) {
    // widest column slice (a divisor of NTA, at most 3 atoms: 128-register budget) that still gives
    // every pair queue a row-block pair
    const int npairs = (nb + 1) >> 1;
    int NA = 1;
    for (int cand = NTA > 4 ? NTA / 2 : NTA; cand >= 1; --cand)
        if (NTA % cand == 0 && (cand == 1 || BF_WARPS / (NTA / cand) <= npairs)) { NA = cand; break; }
    switch (NA) {
        case 3:
            if (NTA % 3 == 0) {
                switch (g_skip & 3) {
                    case 0: mma_phase_3<0>(nb, NTA, g_skip, Aset, Bt, spart, warp); break;
                    case 1: mma_phase_3<1>(nb, NTA, g_skip, Aset, Bt, spart, warp); break;
                    case 2: mma_phase_3<2>(nb, NTA, g_skip, Aset, Bt, spart, warp); break;
                    default: mma_phase_3<3>(nb, NTA, g_skip, Aset, Bt, spart, warp); break;
                }
            }
            break;
        case 2: if (NTA % 2 == 0) mma_phase<2, 2>(nb, NTA, g_skip, Aset, Bt, spart, warp); break;
        default: mma_phase<1, 2>(nb, NTA, g_skip, Aset, Bt, spart, warp); break;
    }
}

template <int D, int NTA>
__global__ void __launch_bounds__(BF_DUAL_THREADS, 1) bf_posterior_dual_kernel(PostParams p) {
    constexpr int NT = 8 * NTA;
    extern __shared__ __align__(16) double smem[];
    const int dd = D ? D : p.d;
    const size_t bt_sz = (size_t)p.np * NT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp >> 3, gw = warp & 7, gtid = threadIdx.x & (BF_THREADS - 1);
    double* Bt = smem + grp * bt_sz;                                    // 2 x np * NT
    double* xs_s = smem + 2 * bt_sz + grp * dd * NT;                    // 2 x d * NT
    double* mpart = smem + 2 * bt_sz + 2 * dd * NT + grp * BF_WARPS * NT;          // 2 buffers x 2 groups x BF_WARPS * NT
    double* spart = smem + 2 * bt_sz + 2 * dd * NT + (4 + grp) * BF_WARPS * NT;    // 2 groups x BF_WARPS * NT
    double* w_s = smem + 2 * bt_sz + 2 * dd * NT + 6 * BF_WARPS * NT;   // BF_MAX_DIM
    double* tab_s = w_s + BF_MAX_DIM;                                   // BF_EXP_TAB_N

    const int set = blockIdx.y;
    const int hp = p.set_hp ? p.set_hp[set] : set;
    // The tile PAIRS of a set are handed out one at a time from the set's counter (p.pair_ctr, zeroed by the
    // launcher): a CTA takes at most `cap` pairs, so an SM is released regularly (the set-up of the next work
    // item waits for one on a high-priority stream), and every SM stays busy until the set's pairs are exhausted -
    // the launch ends within one pair of a perfect balance whatever the block scheduler, the 17/18 split of a
    // static partition or a concurrent kernel do.  Without a counter (p.pair_ctr == NULL) the pairs are split
    // statically: chunk sizes differ by at most one pair.
    const long long npairs_all = (p.ntiles + 1) >> 1;
    const long long pr0 = npairs_all * blockIdx.x / p.tiles, pr1 = npairs_all * (blockIdx.x + 1) / p.tiles;
    const int cap = (int)((npairs_all + p.tiles - 1) / p.tiles);
    unsigned* ctr = p.pair_ctr ? p.pair_ctr + set : nullptr;
    __shared__ long long s_next[2];
    const double amp = p.amp[hp];
    const double* alpha = p.alpha + (size_t)set * p.n;
    const double* Aset = p.linv + (size_t)set * p.linv_stride;
    const double* xw = p.Xw + (size_t)hp * p.n * dd;
    const double sc = p.scale ? p.scale[set] : 1.0;
    const double sh = p.shift ? p.shift[set] : 0.0;

    for (int e = gtid; e < BF_WARPS * NT; e += BF_THREADS) spart[e] = 0.0;   // idle-queue slots stay 0
    for (int e = gtid; e < p.g_skip * NTA * 32; e += BF_THREADS) Bt[e] = 0.0;   // all-padding k-groups: never generated
    if (threadIdx.x < dd) w_s[threadIdx.x] = bf_metric_scale(p.kern, p.inv_l2[(size_t)hp * dd + threadIdx.x]);
    for (int e = threadIdx.x; e < BF_EXP_TAB_N; e += BF_DUAL_THREADS) tab_s[e] = bf_exp_tab_g[e];
    __syncthreads();

    BfBest best[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) best[k] = bf_best_init();
    BfTop2 top = bf_top2_init();

    // k-groups of phase A per warp: contiguous ranges; warp 0 of each group also finishes the previous
    // tile (phase C) while the others generate, so it takes a smaller share
    const int ng = (p.np >> 2) - p.g_skip;                           // k-groups that hold training rows
    int g_lo, g_hi;
    {
        const int per = ng / BF_WARPS;
        int w0 = per - (per >= 6 ? (per + 5) / 8 : 0);
        if (w0 < 1) w0 = per;
        const int rest = ng - w0;                                  // shared by warps 1..7
        if (gw == 0) {
            g_lo = 0;
            g_hi = w0;
        } else {
            g_lo = w0 + (int)((long long)rest * (gw - 1) / (BF_WARPS - 1));
            g_hi = w0 + (int)((long long)rest * gw / (BF_WARPS - 1));
        }
        g_lo += p.g_skip;
        g_hi += p.g_skip;
    }

    // phase C of one tile: one lane per candidate (two rounds when the tile has 48)
    auto finish_tile = [&](long long tile, const double* mp) {
        const long long cb = tile * NT;
#pragma unroll
        for (int c = lane; c < NT; c += 32) {
            const long long gi = cb + c;
            if (gi < p.N) {
                double mean = 0.0, s = 0.0;
#pragma unroll
                for (int w = 0; w < BF_WARPS; ++w) mean += mp[w * NT + c];
#pragma unroll
                for (int w = 0; w < BF_WARPS; ++w) s += spart[w * NT + c];
                const double mu = mean * sc + sh;
                const double var = (amp - s) * (sc * sc);
                if (p.mu_out) p.mu_out[(size_t)set * p.N + gi] = mu;
                if (p.var_out) p.var_out[(size_t)set * p.N + gi] = var;
                if (p.acq_mask) {
                    const double spread = p.spread_is_sqrt ? sqrt(var) : var;
                    if (p.acq_mask & BF_ACQ_EI)
                        bf_best_take(best[BF_SLOT_EI],
                                     (mu - p.curr_max - p.xi) * bf_norm_pdf(mu) + var * bf_norm_cdf(mu), gi);
                    if (p.acq_mask & BF_ACQ_PI)
                        bf_best_take(best[BF_SLOT_PI], bf_norm_cdf((mu - p.curr_max - p.xi) / spread), gi);
                    if (p.acq_mask & BF_ACQ_UCB) bf_best_take(best[BF_SLOT_UCB], mu + p.kt * spread, gi);
                    if (p.acq_mask & BF_ACQ_GREEDY) bf_top2_merge(top, mu, gi, -INFINITY);
                }
            }
        }
    };

    // candidate coordinates of a tile: global -> registers (issued before phase B) -> shared memory
    constexpr int XR = (BF_MAX_DIM * NT + BF_THREADS - 1) / BF_THREADS;
    double xr[XR];
    auto fetch_xs = [&](long long tile) {
        const long long cb = tile * NT;
#pragma unroll
        for (int k = 0; k < XR; ++k) {
            const int e = gtid + k * BF_THREADS;
            const int a = e / NT, c = e - a * NT;
            const long long gi = cb + c;
            xr[k] = (e < dd * NT && tile < p.ntiles && gi < p.N) ? p.Xs[gi * dd + a] * w_s[a] : 0.0;   // pre-scaled
        }
    };
    auto store_xs = [&]() {
#pragma unroll
        for (int k = 0; k < XR; ++k) {
            const int e = gtid + k * BF_THREADS;
            if (e < dd * NT) xs_s[e] = xr[k];
        }
    };

    if (threadIdx.x == 0) s_next[0] = ctr ? (long long)atomicAdd(ctr, 1u) : pr0;
    __syncthreads();
    long long cur = s_next[0];
    if (!ctr && cur >= pr1) cur = npairs_all;
    fetch_xs(2 * cur + grp);
    store_xs();
    long long prev_tile = -1;
    int buf = 0;
    BF_PROF_DECL;
    for (int it = 0; cur < npairs_all; ++it, buf ^= 1) {
        const long long tile = 2 * cur + grp;
        const bool active = tile < p.ntiles;
        double* mp_cur = mpart + buf * 2 * BF_WARPS * NT;
        __syncthreads();     // candidates staged; phase B of the previous pair complete (column sums ready)
        BF_PROF_MARK(0);
        // thread 0 asks for the pair after this one now and publishes it before the next barrier, so the
        // latency of the atomic hides behind phase A
        long long grabbed = npairs_all;
        if (threadIdx.x == 0) {
            if (ctr) {
                if (it + 1 < cap) grabbed = (long long)atomicAdd(ctr, 1u);
            } else if (cur + 1 < pr1) {
                grabbed = cur + 1;
            }
        }
        if (gw == 0 && prev_tile >= 0 && !BF_SKIP(p, 4)) finish_tile(prev_tile, mpart + (buf ^ 1) * 2 * BF_WARPS * NT);
        BF_PROF_MARK(1);
        if (active && !BF_SKIP(p, 2)) gen_tile<D, NTA>(p, xw, xs_s, Bt, mp_cur, alpha, amp, tab_s, g_lo, 1, g_hi, gw);
        BF_PROF_MARK(2);
        if (threadIdx.x == 0) s_next[(it + 1) & 1] = grabbed < npairs_all ? grabbed : npairs_all;
        __syncthreads();
        BF_PROF_MARK(3);
        const long long nxt = s_next[(it + 1) & 1];
        fetch_xs(2 * nxt + grp);                             // in flight during phase B
        if (active && !BF_SKIP(p, 1)) dual_mma<NTA>(p.nb, p.g_skip, Aset, Bt, spart, gw);
        store_xs();
        BF_PROF_MARK(4);
        prev_tile = active ? tile : -1;
        cur = nxt;
    }
    __syncthreads();
    BF_PROF_REPORT;
    if (gw == 0 && prev_tile >= 0 && !BF_SKIP(p, 4)) finish_tile(prev_tile, mpart + (buf ^ 1) * 2 * BF_WARPS * NT);
    if (!p.acq_mask) return;
    // the two finishing warps (warp 0 of each group) merge through shared memory
    __shared__ BfBest red_best[3][2];
    __shared__ BfTop2 red_top[2];
    if (gw == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) best[k] = bf_best_warp(best[k]);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const double t1v = __shfl_xor_sync(0xffffffffu, top.t1, m);
            const long long i1 = __shfl_xor_sync(0xffffffffu, top.i1, m);
            const double t2 = __shfl_xor_sync(0xffffffffu, top.t2, m);
            bf_top2_merge(top, t1v, i1, t2);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) red_best[k][grp] = best[k];
            red_top[grp] = top;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* pv = p.part_val + ((size_t)set * p.tiles + blockIdx.x) * BF_NSLOT_VAL;
        long long* pi = p.part_idx + ((size_t)set * p.tiles + blockIdx.x) * BF_NSLOT_IDX;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            BfBest a = red_best[k][0];
            bf_best_take(a, red_best[k][1].v, red_best[k][1].i);
            pv[k] = a.v;
            pi[k] = a.i;
        }
        BfTop2 t = red_top[0];
        bf_top2_merge(t, red_top[1].t1, red_top[1].i1, red_top[1].t2);
        pv[BF_SLOT_GREEDY] = t.t1;
        pi[BF_SLOT_GREEDY] = t.i1;
        pv[BF_SLOT_TOP2] = t.t2;
    }
}

// ---- per-set reduction of the tile partials --------------------------------------------------
__global__ void bf_acq_finalize_kernel(int tiles, const double* __restrict__ part_val,
                                       const long long* __restrict__ part_idx,
                                       double* __restrict__ best_val, long long* __restrict__ best_idx) {
    const int set = blockIdx.x;
    __shared__ BfBest red_best[3][BF_WARPS];
    __shared__ BfTop2 red_top[BF_WARPS];
    BfBest best[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) best[k] = bf_best_init();
    BfTop2 top = bf_top2_init();
    for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
        const double* pv = part_val + ((size_t)set * tiles + t) * BF_NSLOT_VAL;
        const long long* pi = part_idx + ((size_t)set * tiles + t) * BF_NSLOT_IDX;
#pragma unroll
        for (int k = 0; k < 3; ++k) bf_best_take(best[k], pv[k], pi[k]);
        bf_top2_merge(top, pv[BF_SLOT_GREEDY], pi[BF_SLOT_GREEDY], pv[BF_SLOT_TOP2]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) best[k] = bf_best_warp(best[k]);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const double t1 = __shfl_xor_sync(0xffffffffu, top.t1, m);
        const long long i1 = __shfl_xor_sync(0xffffffffu, top.i1, m);
        const double t2 = __shfl_xor_sync(0xffffffffu, top.t2, m);
        bf_top2_merge(top, t1, i1, t2);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) red_best[k][warp] = best[k];
        red_top[warp] = top;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 0; k < 3; ++k) {
            BfBest a = red_best[k][0];
            for (int w = 1; w < nw; ++w) bf_best_take(a, red_best[k][w].v, red_best[k][w].i);
            best_val[(size_t)set * BF_NSLOT_VAL + k] = a.v;
            best_idx[(size_t)set * BF_NSLOT_IDX + k] = a.i;
        }
        BfTop2 t = red_top[0];
        for (int w = 1; w < nw; ++w) bf_top2_merge(t, red_top[w].t1, red_top[w].i1, red_top[w].t2);
        best_val[(size_t)set * BF_NSLOT_VAL + BF_SLOT_GREEDY] = t.t1;
        best_idx[(size_t)set * BF_NSLOT_IDX + BF_SLOT_GREEDY] = t.i1;
        best_val[(size_t)set * BF_NSLOT_VAL + BF_SLOT_TOP2] = t.t2;
    }
}

// ---- host side ------------------------------------------------------------------------------
static size_t posterior_smem_bytes(int np, int NT, int d) {
    return sizeof(double) * ((size_t)np * NT + (size_t)d * NT + 2 * (size_t)BF_WARPS * NT + BF_MAX_DIM + BF_EXP_TAB_N);
}

extern "C" int bf_padded_n(int n) { return (n + 31) & ~31; }

extern "C" size_t bf_linv_packed_doubles(int n) {
    int nb = bf_padded_n(n) / 32;
    return bf_linv_block_off(nb);
}

// Tile / launch choice.  Dual-tile kernel (two B tiles in shared memory, one 16-warp CTA per SM)
// whenever it fits; otherwise the single-buffer kernel with one CTA per SM.
struct TileChoice { int NT; int dual; };
static size_t posterior_dual_smem_bytes(int np, int NT, int d) {
    return sizeof(double) * (2 * ((size_t)np * NT + (size_t)d * NT) + 6 * (size_t)BF_WARPS * NT + BF_MAX_DIM + BF_EXP_TAB_N);
}
static TileChoice choose_tile(int n) {
    const int np = bf_padded_n(n);
    const size_t cap = (size_t)227 * 1024 - 256;
    const int dual_nt[3] = {48, 24, 16};
    for (int i = 0; i < 3; ++i)
        if (posterior_dual_smem_bytes(np, dual_nt[i], BF_MAX_DIM) <= cap) return TileChoice{dual_nt[i], 1};
    const int one_nt[2] = {24, 16};
    for (int i = 0; i < 2; ++i)
        if (posterior_smem_bytes(np, one_nt[i], BF_MAX_DIM) <= cap) return TileChoice{one_nt[i], 0};
    return TileChoice{0, 0};
}

extern "C" int bf_posterior_tile(int n) { return choose_tile(n).NT; }

// Partial-result slots (= CTAs) per set.  A CTA of the dual-tile kernel processes at most ceil(pairs / CTAs)
// tile pairs of its set: about 8 CTAs per SM over the whole launch, at least 8 tiles per CTA.
static void posterior_chunks(int n, long long N, int S, long long* ntiles, int* tpc, int* parts) {
    const TileChoice tc = choose_tile(n);
    *ntiles = 0; *tpc = 1; *parts = 0;
    if (tc.NT == 0 || N <= 0) return;
    const long long nt = (N + tc.NT - 1) / tc.NT;
    *ntiles = nt;
    if (!tc.dual) { *parts = (int)nt; return; }
    const long long sets = S > 0 ? S : 1;
    const long long npairs = (nt + 1) / 2;
    // CTAs per set.  Few sets: 8 CTAs per SM in total (the kernel splits the tile pairs evenly, so every
    // SM gets the same work and there is no partial last wave).  Many sets: at least 8 CTAs per set, so CTAs that run at the
    // same time share few sets and the L^-1 operands in flight (1 MB per set at n = 500) stay L2-resident
    // instead of streaming from HBM.  Never fewer than 4 tile pairs per CTA (prologue amortisation).
    long long want = (148LL * 8) / sets;               // rounded down: the launch never spills into a ninth wave
    if (want < 8) want = 8;
    // CTAs take equal time, so a launch of sets * want CTAs costs ceil(sets * want / 148) waves: among up to four
    // times as many CTAs per set pick the count that wastes the least of its last wave (125 sets per GPU - config #5
    // on 8 GPUs - would otherwise run 7.6 waves in the time of 8)
    {
        long long best = want;
        double best_eff = 0.0;
        for (long long w = want; w <= 4 * want; ++w) {
            const long long ctas = sets * w, waves = (ctas + 147) / 148;
            const double eff = (double)ctas / (double)(waves * 148);
            if (eff > best_eff + 1e-3) {
                best_eff = eff;
                best = w;
            }
        }
        want = best;
    }
    if (want > (npairs + 3) / 4) want = (npairs + 3) / 4;
    if (want < 1) want = 1;
    *tpc = (int)(2 * ((npairs + want - 1) / want));    // upper bound of tiles per CTA (informational)
    *parts = (int)want;
}

extern "C" int bf_posterior_parts(int n, long long N, int S) {
    long long nt; int tpc, parts;
    posterior_chunks(n, N, S, &nt, &tpc, &parts);
    return parts;
}

template <int D, int NTA, int MINB>
static int launch_posterior(const PostParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(bf_posterior_kernel<D, NTA, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_gp_posterior: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    bf_posterior_kernel<D, NTA, MINB><<<grid, BF_THREADS, smem, st>>>(p);
    BF_CHECK_LAUNCH("bf_gp_posterior");
    return 0;
}

template <int D, int NTA>
static int launch_posterior_dual(const PostParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(bf_posterior_dual_kernel<D, NTA>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_gp_posterior: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    bf_posterior_dual_kernel<D, NTA><<<grid, BF_DUAL_THREADS, smem, st>>>(p);
    BF_CHECK_LAUNCH("bf_gp_posterior");
    return 0;
}

template <int D>
static int launch_posterior_tile(const PostParams& p, int dual, dim3 grid, size_t smem, cudaStream_t st) {
    if (dual) {
        if (p.NT == 48) return launch_posterior_dual<D, 6>(p, grid, smem, st);
        if (p.NT == 24) return launch_posterior_dual<D, 3>(p, grid, smem, st);
        return launch_posterior_dual<D, 2>(p, grid, smem, st);
    }
    if (p.NT == 24) return launch_posterior<D, 3, 1>(p, grid, smem, st);
    return launch_posterior<D, 2, 1>(p, grid, smem, st);
}

// X' = X * bf_metric_scale per hyper-parameter row: the training-side operand of the cross-covariance generator
__global__ void bf_prescale_train_kernel(int kern, long long per, int d, long long total,
                                         const double* __restrict__ X, const double* __restrict__ inv_l2,
                                         double* __restrict__ Xw) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long h = e / per, r = e - h * per;
    Xw[e] = X[r] * bf_metric_scale(kern, inv_l2[h * d + (r % d)]);
}

extern "C" int bf_prescale_train(int kern, int n, int d, int H, const double* X, const double* inv_l2,
                                 double* Xw, void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_prescale_train: bad kern %d", kern);
    BF_CHECK_ARG(n > 0 && H >= 0 && d >= 1 && d <= BF_MAX_DIM, BF_ERR_ARG, "bf_prescale_train: bad sizes n=%d d=%d H=%d", n, d, H);
    BF_CHECK_ARG(X && inv_l2 && Xw, BF_ERR_ARG, "bf_prescale_train: null argument");
    if (H == 0) return 0;
    const long long per = (long long)n * d, total = per * H;
    bf_prescale_train_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kern, per, d, total, X,
                                                                                                 inv_l2, Xw);
    BF_CHECK_LAUNCH("bf_prescale_train");
    return 0;
}

extern "C" int bf_gp_posterior(int kern, long long N, int n, int d, int S, const double* Xs,
                               const double* Xw, const double* inv_l2, const double* amp,
                               const int32_t* set_hp, const double* linv_packed, const double* alpha,
                               const double* scale, const double* shift, double* mu_out,
                               double* var_out, int acq_mask, int spread_is_sqrt, double curr_max,
                               double xi, double kt, double* part_val, long long* part_idx,
                               uint32_t* pair_ctr, void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_gp_posterior: bad kern %d", kern);
    BF_CHECK_ARG(N >= 0 && n > 0 && S >= 0, BF_ERR_ARG, "bf_gp_posterior: bad sizes N=%lld n=%d S=%d", N, n, S);
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_gp_posterior: nDim %d outside 1..%d", d, BF_MAX_DIM);
    BF_CHECK_ARG(Xs && Xw && inv_l2 && amp && linv_packed && alpha, BF_ERR_ARG, "bf_gp_posterior: null input");
    BF_CHECK_ARG(!acq_mask || (part_val && part_idx), BF_ERR_ARG, "bf_gp_posterior: acq_mask without partial buffers");
    BF_CHECK_ARG(((uintptr_t)linv_packed & 15) == 0, BF_ERR_ALIGN, "bf_gp_posterior: linv_packed must be 16-byte aligned");
    if (N == 0 || S == 0) return 0;
    PostParams p;
    p.kern = kern; p.n = n; p.d = d; p.np = bf_padded_n(n); p.nb = p.np / 32;
    p.sh = p.np - n; p.g_skip = p.sh >> 2;
    const TileChoice tc = choose_tile(n);
    p.NT = tc.NT;
    BF_CHECK_ARG(p.NT > 0, BF_ERR_SIZE, "bf_gp_posterior: n=%d too large for the shared-memory tile", n);
    p.NTA = p.NT / 8; p.S = S; p.N = N;
    p.Xs = Xs; p.Xw = Xw; p.inv_l2 = inv_l2; p.amp = amp; p.linv = linv_packed; p.alpha = alpha;
    p.scale = scale; p.shift = shift; p.set_hp = set_hp;
    p.linv_stride = bf_linv_packed_doubles(n);
    p.mu_out = mu_out; p.var_out = var_out;
    p.acq_mask = acq_mask; p.spread_is_sqrt = spread_is_sqrt;
    p.curr_max = curr_max; p.xi = xi; p.kt = kt;
    p.part_val = part_val; p.part_idx = part_idx;
    posterior_chunks(n, N, S, &p.ntiles, &p.tpc, &p.tiles);
#ifdef BF_PHASE_PROFILE
    {
        const char* dbg = getenv("BF_DEBUG_SKIP");
        p.debug_skip = dbg ? atoi(dbg) : 0;
    }
#endif
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_gp_posterior: at most 65535 sets per call (got %d)", S);
    dim3 grid(p.tiles, S);
    cudaStream_t st = (cudaStream_t)stream;
    p.pair_ctr = tc.dual ? pair_ctr : nullptr;
    if (p.pair_ctr) {
        cudaError_t e = cudaMemsetAsync(p.pair_ctr, 0, sizeof(uint32_t) * (size_t)S, st);
        if (e != cudaSuccess) {
            bf_set_error("bf_gp_posterior: cudaMemsetAsync(pair_ctr): %s", cudaGetErrorString(e));
            return (int)e;
        }
    }
    const size_t smem = tc.dual ? posterior_dual_smem_bytes(p.np, p.NT, d) : posterior_smem_bytes(p.np, p.NT, d);
    switch (d) {
#define BF_CASE(DD) case DD: return launch_posterior_tile<DD>(p, tc.dual, grid, smem, st);
        BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8)
        BF_CASE(9) BF_CASE(10) BF_CASE(11) BF_CASE(12) BF_CASE(13) BF_CASE(14) BF_CASE(15) BF_CASE(16)
#undef BF_CASE
        default: return launch_posterior_tile<0>(p, tc.dual, grid, smem, st);   // unreachable (nDim <= BF_MAX_DIM); kept as the run-time-dimension form
    }
    return BF_ERR_SIZE;
}

extern "C" int bf_acq_finalize(int S, int tiles, int acq_mask, const double* part_val,
                               const long long* part_idx, double* best_val, long long* best_idx,
                               void* stream) {
    (void)acq_mask;
    BF_CHECK_ARG(S >= 0 && tiles > 0 && part_val && part_idx && best_val && best_idx, BF_ERR_ARG,
                 "bf_acq_finalize: bad argument");
    if (S == 0) return 0;
    bf_acq_finalize_kernel<<<S, BF_THREADS, 0, (cudaStream_t)stream>>>(tiles, part_val, part_idx, best_val, best_idx);
    BF_CHECK_LAUNCH("bf_acq_finalize");
    return 0;
}
