// 32 x 32 block primitives shared by the set-up kernels (one warp each): in-register Cholesky of a diagonal
// block and the inverse of a lower-triangular block.
#pragma once
#include "bf_common.cuh"

// lower Cholesky of the 32 x 32 block in Ld (one warp).  Lane = row, the row lives in registers and
// column j of the factor is broadcast with shuffles: ~1500 instructions instead of ~1000 dependent
// shared-memory round trips.  Template recursion keeps every register index a compile-time constant.
template <int J>
__device__ __forceinline__ void rl_factor_step(double (&row)[32], int lane, int& bad, double& my_rdiag) {
    if constexpr (J < 32) {
        const double djj = __shfl_sync(0xffffffffu, row[J], J);
        if (!(djj > 0.0) && bad == 0) bad = J + 1;
        // one reciprocal square root per pivot (Newton-refined, < 1 ulp) instead of a square root and 31
        // divisions on the critical path of a single warp
        double rinv = rsqrt(djj);
        rinv = fma(rinv * 0.5, fma(-djj * rinv, rinv, 1.0), rinv);
        if (lane == J) {
            row[J] = djj * rinv;
            my_rdiag = rinv;
        }
        if (lane > J) row[J] = row[J] * rinv;
        const double lij = row[J];
#pragma unroll
        for (int cc = J + 1; cc < 32; ++cc) {
            const double lcj = __shfl_sync(0xffffffffu, lij, cc);           // L[cc][J]
            if (lane >= cc) row[cc] = fma(-lij, lcj, row[cc]);
        }
        rl_factor_step<J + 1>(row, lane, bad, my_rdiag);
    }
}

// Returns the first bad pivot + 1 or 0; Ld receives L (zeros above the diagonal), rdiag[j] = 1 / L[j][j].
__device__ __forceinline__ int rl_factor_block(double (*Ld)[33], double* rdiag, int lane) {
    double row[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) row[j] = Ld[lane][j];
    int bad = 0;
    double my_rdiag = 0.0;
    rl_factor_step<0>(row, lane, bad, my_rdiag);
#pragma unroll
    for (int j = 0; j < 32; ++j) Ld[lane][j] = (j <= lane) ? row[j] : 0.0;
    rdiag[lane] = my_rdiag;
    __syncwarp();
    return bad;
}

// inverse of the lower-triangular block Ld into Li (one warp, lane = column of the inverse); four
// partial sums per row shorten the dependent FMA chain.  rdiag = reciprocal diagonal, or NULL.
__device__ __forceinline__ void rl_invert_block(const double (*Ld)[33], const double* rdiag, double (*Li)[33],
                                                int lane) {
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int t = 0; t + 3 < i; t += 4) {
            s0 = fma(-Ld[i][t], x[t], s0);                                  // x[t] = 0 for t < lane
            s1 = fma(-Ld[i][t + 1], x[t + 1], s1);
            s2 = fma(-Ld[i][t + 2], x[t + 2], s2);
            s3 = fma(-Ld[i][t + 3], x[t + 3], s3);
        }
#pragma unroll
        for (int t = (i / 4) * 4; t < i; ++t) s0 = fma(-Ld[i][t], x[t], s0);
        const double s = (s0 + s1) + (s2 + s3);
        const double v = rdiag ? s * rdiag[i] : s / Ld[i][i];
        x[i] = (i >= lane) ? v : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) Li[i][lane] = x[i];
}


// ---- blocked variants (round 2): the same two operations by 8-column panels --------------------------------
// The in-register column-by-column forms above put ~31 kcycles (factor + inverse) on a single warp's dependency
// chain per diagonal block, and that chain is what the right-looking kernel waits for in its last ~8 panel steps
// and at its start (tools/prof_setup).  Here a pivot only updates the remaining columns of its own 8-wide panel
// (28 shuffle + FMA pairs per panel instead of ~200), and everything to the right of a panel is a rank-8 update
// on the tensor cores (10 DMMA atoms in total); the inverse is four independent 8 x 8 triangular inverses (eight
// lanes each) plus six 8 x 8 off-diagonal blocks, -L_tt^-1 sum_k L_tk L^-1_ku, as DMMA products.
__device__ __forceinline__ int rl_factor_block_blocked(double (*Ld)[33], double* rdiag, int lane) {
    const int r = lane >> 2, q = lane & 3;
    int bad = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = Ld[lane][8 * p + j];
        double my_rdiag = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * p + j;
            const double d = __shfl_sync(0xffffffffu, a[j], c);
            if (!(d > 0.0) && bad == 0) bad = c + 1;
            double rinv = rsqrt(d);
            rinv = fma(rinv * 0.5, fma(-d * rinv, rinv, 1.0), rinv);
            if (lane == c) {
                a[j] = d * rinv;
                my_rdiag = rinv;
            }
            if (lane > c) a[j] = a[j] * rinv;
            const double lij = a[j];
#pragma unroll
            for (int cc = j + 1; cc < 8; ++cc) {
                const double lcj = __shfl_sync(0xffffffffu, lij, 8 * p + cc);       // L[8p + cc][c]
                if (lane >= 8 * p + cc) a[cc] = fma(-lij, lcj, a[cc]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) Ld[lane][8 * p + j] = (lane >= 8 * p + j) ? a[j] : 0.0;
        if ((lane >> 3) == p) rdiag[lane] = my_rdiag;
        __syncwarp();
        // rank-8 update of the blocks to the right of the panel (atoms t >= u > p): C -= L[t, p] L[u, p]^T
#pragma unroll
        for (int t = p + 1; t < 4; ++t)
#pragma unroll
            for (int u = p + 1; u <= t; ++u) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    bf_dmma(c0, c1, Ld[8 * t + r][8 * p + 4 * ks + q], Ld[8 * u + r][8 * p + 4 * ks + q]);
                Ld[8 * t + r][8 * u + 2 * q] -= c0;
                Ld[8 * t + r][8 * u + 2 * q + 1] -= c1;
            }
        __syncwarp();
    }
    return bad;
}

// Li = Ld^-1 for the lower-triangular Ld (rdiag[i] = 1 / Ld[i][i]); Li is written completely (zeros above).
__device__ __forceinline__ void rl_invert_block_blocked(const double (*Ld)[33], const double* rdiag, double (*Li)[33],
                                                        int lane) {
    const int r = lane >> 2, q = lane & 3;
    {   // the four diagonal 8 x 8 blocks: lane = (block, column) solves one column by forward substitution
        const int p = lane >> 3, c = lane & 7;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int t = 0; t < i; ++t) s = fma(-Ld[8 * p + i][8 * p + t], x[t], s);   // x[t] = 0 for t < c
            x[i] = (i >= c) ? s * rdiag[8 * p + i] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Li[8 * p + i][8 * p + c] = x[i];
    }
    // zeros in the blocks above the diagonal
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int u = t + 1; u < 4; ++u) {
            Li[8 * t + r][8 * u + 2 * q] = 0.0;
            Li[8 * t + r][8 * u + 2 * q + 1] = 0.0;
        }
    __syncwarp();
    // block rows below: Li[t, u] = -Li[t, t] sum_{k = u}^{t - 1} L[t, k] Li[k, u]
#pragma unroll
    for (int t = 1; t < 4; ++t) {
        double res[3][2];
#pragma unroll
        for (int u = 0; u < t; ++u) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = u; k < t; ++k)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)       // A = L[t, k] (row r, col q), B = Li[k, u] (row q of the k-step, col r)
                    bf_dmma(s0, s1, Ld[8 * t + r][8 * k + 4 * ks + q], Li[8 * k + 4 * ks + q][8 * u + r]);
            // T = Li[t, t] S: S moves from the accumulator layout (row r, cols 2q, 2q + 1) to the B operand
            // (row 4 ks + q, col r) through shuffles
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int src = (4 * ks + q) * 4 + (r >> 1);
                const double v0 = __shfl_sync(0xffffffffu, s0, src);
                const double v1 = __shfl_sync(0xffffffffu, s1, src);
                bf_dmma(t0, t1, Li[8 * t + r][8 * t + 4 * ks + q], (r & 1) ? v1 : v0);
            }
            res[u][0] = -t0;
            res[u][1] = -t1;
        }
#pragma unroll
        for (int u = 0; u < t; ++u) {
            Li[8 * t + r][8 * u + 2 * q] = res[u][0];
            Li[8 * t + r][8 * u + 2 * q + 1] = res[u][1];
        }
        __syncwarp();
    }
}
