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

