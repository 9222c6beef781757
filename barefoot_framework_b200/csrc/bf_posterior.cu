// GP posterior mean/variance + elementwise acquisitions, fused (subsystems 1, 2 and 4 of the
// north star).  Replaces gp_model.predict_var (gpModel.py:50-54), predict_low_order /
// predict_fused_GP (reificationFusion.py:189-206) and expected_improvement /
// probability_improvement / upper_conf_bound / greedy (acquisitionFunc.py:98-214,
// util.py:655-665) for all hyper-parameter sets x all candidates in one launch.
//
// One CTA = one tile of NT candidates of one set:
//   phase A  cross-covariance tile k*[np x NT] generated straight into shared memory in the
//            DMMA B-fragment layout (never touches HBM); mean = k* . alpha folded in.
//   phase B  V = L^-1 k* on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4).  L^-1 is
//            block lower triangular in a pre-packed fragment layout streamed from L2 with
//            register double-buffering; warps own balanced (b, nb-1-b) row-block pairs.
//            Only the column sums of V^2 are kept.
//   phase C  var = amp - |V|^2, de-normalise, acquisition values, per-tile (max, first index)
//            and top-2 of the mean.
#include "bf_common.cuh"

struct PostParams {
    int kern, n, d, np, nb, NT, NTA, S;
    long long N;
    const double *Xs, *X, *inv_l2, *amp, *linv, *alpha, *scale, *shift;
    const int32_t* set_hp;
    size_t linv_stride;
    double *mu_out, *var_out;
    int acq_mask, spread_is_sqrt;
    double curr_max, xi, kt;
    double* part_val;
    long long* part_idx;
    int tiles;
};

// ---- phase A ------------------------------------------------------------------------------
// Thread (warp w, lane): k-group g = w, w+8, ...; j = 4g + (lane & 3); candidate pair
// c0 = 16 ap + (lane >> 2), c1 = c0 + 8.  Output cell ((g * NTA/2 + ap) * 32 + lane) * 2.
template <int D>
__device__ __forceinline__ void gen_tile(const PostParams& p, const double* __restrict__ w_s,
                                         const double* __restrict__ xs_s, double* __restrict__ Bt,
                                         double* __restrict__ mpart, const double* __restrict__ alpha,
                                         double amp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = lane & 3, nn = lane >> 2;
    const int NT = p.NT, NAP = p.NTA >> 1;
    double w[D];
#pragma unroll
    for (int a = 0; a < D; ++a) w[a] = w_s[a];
    double macc[4][2];
#pragma unroll
    for (int ap = 0; ap < 4; ++ap) macc[ap][0] = macc[ap][1] = 0.0;

    for (int g = warp; g < (p.np >> 2); g += BF_WARPS) {
        const int j = 4 * g + kk;
        const bool valid = j < p.n;
        double xj[D];
#pragma unroll
        for (int a = 0; a < D; ++a) xj[a] = valid ? __ldg(p.X + (size_t)j * D + a) : 0.0;
        const double aj = valid ? __ldg(alpha + j) : 0.0;
        double2* out = reinterpret_cast<double2*>(Bt) + (size_t)g * NAP * 32 + lane;
#pragma unroll
        for (int ap = 0; ap < 4; ++ap) {
            if (ap < NAP) {
                const int c0 = 16 * ap + nn;
                double r0 = 0.0, r1 = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const double d0 = xs_s[a * NT + c0] - xj[a];
                    const double d1 = xs_s[a * NT + c0 + 8] - xj[a];
                    r0 = fma(d0 * d0, w[a], r0);
                    r1 = fma(d1 * d1, w[a], r1);
                }
                double2 v;
                v.x = valid ? amp * bf_radial(p.kern, r0) : 0.0;
                v.y = valid ? amp * bf_radial(p.kern, r1) : 0.0;
                macc[ap][0] = fma(aj, v.x, macc[ap][0]);
                macc[ap][1] = fma(aj, v.y, macc[ap][1]);
                out[ap * 32] = v;
            }
        }
    }
    // mean partials: sum over the 4 kk lanes, then one slot per warp (fixed-order sum later)
#pragma unroll
    for (int ap = 0; ap < 4; ++ap) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double v = macc[ap][q];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (kk == 0 && ap < NAP) mpart[warp * NT + 16 * ap + 8 * q + nn] = v;
        }
    }
}

// ---- phase B ------------------------------------------------------------------------------
// One warp, one row block b, NA column atoms starting at atom a0: accumulates the column sums
// of (L^-1[b,:] k*)^2 into cs.  The last 8 k-groups are the diagonal 32x32 block: atoms whose
// rows lie entirely above the columns of the group are skipped.
template <int NA>
__device__ __forceinline__ void mma_block(const double* __restrict__ Aset, const double* __restrict__ Bt,
                                          int NAP_total, int a0, int b, double (&cs)[NA][2]) {
    const int lane = threadIdx.x & 31;
    double c[4][NA][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int a = 0; a < NA; ++a) c[t][a][0] = c[t][a][1] = 0.0;

    const double2* Ap = reinterpret_cast<const double2*>(Aset + bf_linv_block_off(b)) + lane;
    const double2* Bp = reinterpret_cast<const double2*>(Bt) + (size_t)(a0 >> 1) * 32 + lane;
    const int gstride = NAP_total * 32;     // double2 per k-group in the B tile
    const int ng_full = 8 * b;              // groups left of the diagonal block

    double2 A0[2], A1[2];                   // ring of two k-groups: [tpair]
    A0[0] = __ldg(Ap);
    A0[1] = __ldg(Ap + 32);
    A1[0] = __ldg(Ap + 64);
    A1[1] = __ldg(Ap + 96);

#define BF_MMA_STEP(AR, G)                                                          \
    {                                                                               \
        double2 bb[NA / 2];                                                         \
        _Pragma("unroll") for (int ap = 0; ap < NA / 2; ++ap) bb[ap] =              \
            Bp[(size_t)(G) * gstride + ap * 32];                                    \
        _Pragma("unroll") for (int ap = 0; ap < NA / 2; ++ap) {                     \
            bf_dmma(c[0][2 * ap][0], c[0][2 * ap][1], AR[0].x, bb[ap].x);           \
            bf_dmma(c[1][2 * ap][0], c[1][2 * ap][1], AR[0].y, bb[ap].x);           \
            bf_dmma(c[2][2 * ap][0], c[2][2 * ap][1], AR[1].x, bb[ap].x);           \
            bf_dmma(c[3][2 * ap][0], c[3][2 * ap][1], AR[1].y, bb[ap].x);           \
            bf_dmma(c[0][2 * ap + 1][0], c[0][2 * ap + 1][1], AR[0].x, bb[ap].y);   \
            bf_dmma(c[1][2 * ap + 1][0], c[1][2 * ap + 1][1], AR[0].y, bb[ap].y);   \
            bf_dmma(c[2][2 * ap + 1][0], c[2][2 * ap + 1][1], AR[1].x, bb[ap].y);   \
            bf_dmma(c[3][2 * ap + 1][0], c[3][2 * ap + 1][1], AR[1].y, bb[ap].y);   \
        }                                                                           \
    }

    int g = 0;
    for (; g < ng_full; g += 2) {
        double2 N0[2], N1[2];
        const double2* An = Ap + (size_t)(g + 2) * 64;   // always in range: diagonal block follows
        N0[0] = __ldg(An);
        N0[1] = __ldg(An + 32);
        N1[0] = __ldg(An + 64);
        N1[1] = __ldg(An + 96);
        BF_MMA_STEP(A0, g);
        BF_MMA_STEP(A1, g + 1);
        A0[0] = N0[0]; A0[1] = N0[1];
        A1[0] = N1[0]; A1[1] = N1[1];
    }
#undef BF_MMA_STEP
    // diagonal block: groups g .. g+7, group gg touches columns 4gg..4gg+3 of the block;
    // atom t (rows 8t..8t+7) is non-zero only for gg < 2t + 2.
#pragma unroll
    for (int gg = 0; gg < 8; gg += 2) {
        double2 N0[2], N1[2];
        if (gg + 2 < 8) {
            const double2* An = Ap + (size_t)(g + gg + 2) * 64;
            N0[0] = __ldg(An);
            N0[1] = __ldg(An + 32);
            N1[0] = __ldg(An + 64);
            N1[1] = __ldg(An + 96);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int G = gg + h;
            double2 bb[NA / 2];
#pragma unroll
            for (int ap = 0; ap < NA / 2; ++ap) bb[ap] = Bp[(size_t)(g + G) * gstride + ap * 32];
#pragma unroll
            for (int ap = 0; ap < NA / 2; ++ap) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double bv = q ? bb[ap].y : bb[ap].x;
                    const double2 a01 = h ? A1[0] : A0[0];
                    const double2 a23 = h ? A1[1] : A0[1];
                    if (G < 2) bf_dmma(c[0][2 * ap + q][0], c[0][2 * ap + q][1], a01.x, bv);
                    if (G < 4) bf_dmma(c[1][2 * ap + q][0], c[1][2 * ap + q][1], a01.y, bv);
                    if (G < 6) bf_dmma(c[2][2 * ap + q][0], c[2][2 * ap + q][1], a23.x, bv);
                    bf_dmma(c[3][2 * ap + q][0], c[3][2 * ap + q][1], a23.y, bv);
                }
            }
        }
        if (gg + 2 < 8) {
            A0[0] = N0[0]; A0[1] = N0[1];
            A1[0] = N1[0]; A1[1] = N1[1];
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

template <int NA>
__device__ __noinline__ void mma_phase(int nb, int NT, int NTA, const double* __restrict__ Aset,
                                       const double* __restrict__ Bt, double* __restrict__ spart) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = NTA / NA;                   // warps sharing a row-block pair
    const int slice = warp % split;               // which NA atoms of the tile
    const int wq = warp / split;                  // pair queue of this warp
    const int nq = BF_WARPS / split;
    const int npairs = (nb + 1) >> 1;
    const int a0 = slice * NA;
    double cs[NA][2];
#pragma unroll
    for (int a = 0; a < NA; ++a) cs[a][0] = cs[a][1] = 0.0;
    for (int pr = wq; pr < npairs; pr += nq) {
        const int b_hi = nb - 1 - pr;
        mma_block<NA>(Aset, Bt, NTA >> 1, a0, b_hi, cs);
        if (pr != b_hi) mma_block<NA>(Aset, Bt, NTA >> 1, a0, pr, cs);
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
    // queue slots >= nq are never written: the kernel zeroes spart before phase A
}

template <int D>
__global__ void __launch_bounds__(BF_THREADS, 1) bf_posterior_kernel(PostParams p) {
    extern __shared__ __align__(16) double smem[];
    double* Bt = smem;                                   // np * NT
    double* xs_s = Bt + (size_t)p.np * p.NT;             // D * NT
    double* mpart = xs_s + D * p.NT;                     // BF_WARPS * NT
    double* spart = mpart + BF_WARPS * p.NT;             // BF_WARPS * NT
    double* w_s = spart + BF_WARPS * p.NT;               // BF_MAX_DIM
    __shared__ BfBest red_best[4][2];
    __shared__ BfTop2 red_top[2];

    const int tile = blockIdx.x, set = blockIdx.y;
    const int hp = p.set_hp ? p.set_hp[set] : set;
    const int NT = p.NT;
    const long long c_base = (long long)tile * NT;
    const double amp = p.amp[hp];

    for (int e = threadIdx.x; e < D * NT; e += BF_THREADS) {
        const int a = e / NT, c = e - a * NT;
        const long long gi = c_base + c;
        xs_s[e] = gi < p.N ? p.Xs[gi * D + a] : 0.0;
    }
    for (int e = threadIdx.x; e < BF_WARPS * NT; e += BF_THREADS) spart[e] = 0.0;
    if (threadIdx.x < D) w_s[threadIdx.x] = p.inv_l2[(size_t)hp * D + threadIdx.x];
    __syncthreads();

    gen_tile<D>(p, w_s, xs_s, Bt, mpart, p.alpha + (size_t)set * p.n, amp);
    __syncthreads();

    const double* Aset = p.linv + (size_t)set * p.linv_stride;
    {
        // widest column slice such that every warp has a row-block pair to work on
        const int npairs = (p.nb + 1) >> 1;
        int NA = p.NTA;
        while (NA > 2 && (NA % 4 == 0) && (BF_WARPS / (p.NTA / NA)) > npairs) NA >>= 1;
        switch (NA) {
            case 8: mma_phase<8>(p.nb, NT, p.NTA, Aset, Bt, spart); break;
            case 6: mma_phase<6>(p.nb, NT, p.NTA, Aset, Bt, spart); break;
            case 4: mma_phase<4>(p.nb, NT, p.NTA, Aset, Bt, spart); break;
            default: mma_phase<2>(p.nb, NT, p.NTA, Aset, Bt, spart); break;
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
    return sizeof(double) * ((size_t)np * NT + (size_t)d * NT + 2 * (size_t)BF_WARPS * NT + BF_MAX_DIM);
}

extern "C" int bf_padded_n(int n) { return (n + 31) & ~31; }

extern "C" size_t bf_linv_packed_doubles(int n) {
    int nb = bf_padded_n(n) / 32;
    return bf_linv_block_off(nb);
}

extern "C" int bf_posterior_tile(int n) {
    const int np = bf_padded_n(n);
    const int cand[4] = {64, 48, 32, 16};
    for (int i = 0; i < 4; ++i)
        if (posterior_smem_bytes(np, cand[i], BF_MAX_DIM) + 256 <= 227 * 1024) return cand[i];
    return 0;
}

extern "C" int bf_posterior_tiles(int n, long long N) {
    int NT = bf_posterior_tile(n);
    if (NT == 0 || N <= 0) return 0;
    return (int)((N + NT - 1) / NT);
}

template <int D>
static int launch_posterior(const PostParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(bf_posterior_kernel<D>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        bf_set_error("bf_gp_posterior: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    bf_posterior_kernel<D><<<grid, BF_THREADS, smem, st>>>(p);
    BF_CHECK_LAUNCH("bf_gp_posterior");
    return 0;
}

extern "C" int bf_gp_posterior(int kern, long long N, int n, int d, int S, const double* Xs,
                               const double* X, const double* inv_l2, const double* amp,
                               const int32_t* set_hp, const double* linv_packed, const double* alpha,
                               const double* scale, const double* shift, double* mu_out,
                               double* var_out, int acq_mask, int spread_is_sqrt, double curr_max,
                               double xi, double kt, double* part_val, long long* part_idx,
                               void* stream) {
    BF_CHECK_ARG(kern >= BF_KERN_SE && kern <= BF_KERN_M52, BF_ERR_ARG, "bf_gp_posterior: bad kern %d", kern);
    BF_CHECK_ARG(N >= 0 && n > 0 && S >= 0, BF_ERR_ARG, "bf_gp_posterior: bad sizes N=%lld n=%d S=%d", N, n, S);
    BF_CHECK_ARG(d >= 1 && d <= BF_MAX_DIM, BF_ERR_SIZE, "bf_gp_posterior: nDim %d outside 1..%d", d, BF_MAX_DIM);
    BF_CHECK_ARG(Xs && X && inv_l2 && amp && linv_packed && alpha, BF_ERR_ARG, "bf_gp_posterior: null input");
    BF_CHECK_ARG(!acq_mask || (part_val && part_idx), BF_ERR_ARG, "bf_gp_posterior: acq_mask without partial buffers");
    BF_CHECK_ARG(((uintptr_t)linv_packed & 15) == 0, BF_ERR_ALIGN, "bf_gp_posterior: linv_packed must be 16-byte aligned");
    if (N == 0 || S == 0) return 0;
    PostParams p;
    p.kern = kern; p.n = n; p.d = d; p.np = bf_padded_n(n); p.nb = p.np / 32;
    p.NT = bf_posterior_tile(n);
    BF_CHECK_ARG(p.NT > 0, BF_ERR_SIZE, "bf_gp_posterior: n=%d too large for the shared-memory tile", n);
    p.NTA = p.NT / 8; p.S = S; p.N = N;
    p.Xs = Xs; p.X = X; p.inv_l2 = inv_l2; p.amp = amp; p.linv = linv_packed; p.alpha = alpha;
    p.scale = scale; p.shift = shift; p.set_hp = set_hp;
    p.linv_stride = bf_linv_packed_doubles(n);
    p.mu_out = mu_out; p.var_out = var_out;
    p.acq_mask = acq_mask; p.spread_is_sqrt = spread_is_sqrt;
    p.curr_max = curr_max; p.xi = xi; p.kt = kt;
    p.part_val = part_val; p.part_idx = part_idx;
    p.tiles = bf_posterior_tiles(n, N);
    BF_CHECK_ARG(S <= 65535, BF_ERR_SIZE, "bf_gp_posterior: at most 65535 sets per call (got %d)", S);
    dim3 grid(p.tiles, S);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = posterior_smem_bytes(p.np, p.NT, d);
    switch (d) {
#define BF_CASE(DD) case DD: return launch_posterior<DD>(p, grid, smem, st);
        BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8)
        BF_CASE(9) BF_CASE(10) BF_CASE(11) BF_CASE(12) BF_CASE(13) BF_CASE(14) BF_CASE(15) BF_CASE(16)
#undef BF_CASE
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
