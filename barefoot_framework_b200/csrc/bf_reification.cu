// Reification/fusion of m information sources per point (reificationFusion.py:26-57).
// One thread per point: the m x m correlation-weighted covariance is built in registers,
// pseudo-inverted through a cyclic Jacobi eigen-decomposition (numpy.linalg.pinv semantics:
// eigenvalues with |lambda| <= 1e-15 |lambda|max are dropped, reificationFusion.py:47) and
// reduced to the fused mean and variance.  Stand-alone it is HBM-bound: 16 m B read + 16 B
// written per point.
#include "bf_common.cuh"

template <int M>
__global__ void __launch_bounds__(128) bf_reification_kernel(long long P, const double* __restrict__ y,
                                                             const double* __restrict__ var,
                                                             double* __restrict__ mean_f,
                                                             double* __restrict__ var_f) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double yv[M], sv[M], rs[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        yv[i] = y[(size_t)i * P + p];
        sv[i] = var[(size_t)i * P + p];
        rs[i] = sqrt(sv[i]);
    }
    double A[M][M], V[M][M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
            V[i][j] = (i == j) ? 1.0 : 0.0;
            if (i == j) {
                A[i][j] = sv[i];
            } else if (j > i) {
                // element [i][j] of the reference tensor: "M" quantities carry j, "MT" carry i
                const double dy = yv[j] - yv[i];
                const double rho1 = rs[j] / sqrt(dy * dy + sv[j]);
                const double dyt = yv[i] - yv[j];
                const double rho2 = rs[i] / sqrt(dyt * dyt + sv[i]);
                const double rho = (sv[i] / (sv[i] + sv[j])) * rho1 + (sv[j] / (sv[i] + sv[j])) * rho2;
                A[i][j] = rho * sqrt(sv[j] * sv[i]);
            }
        }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) A[i][j] = A[j][i];

    // cyclic Jacobi
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = i + 1; j < M; ++j) off += fabs(A[i][j]);
        if (off == 0.0) break;
#pragma unroll
        for (int pi = 0; pi < M; ++pi)
#pragma unroll
            for (int q = pi + 1; q < M; ++q) {
                const double apq = A[pi][q];
                if (apq != 0.0) {
                    const double g = 100.0 * fabs(apq);
                    // negligible against both diagonal entries: annihilate
                    if (sweep > 3 && fabs(A[pi][pi]) + g == fabs(A[pi][pi]) &&
                        fabs(A[q][q]) + g == fabs(A[q][q])) {
                        A[pi][q] = A[q][pi] = 0.0;
                    } else {
                        const double h = A[q][q] - A[pi][pi];
                        double t;
                        if (fabs(h) + g == fabs(h)) {
                            t = apq / h;
                        } else {
                            const double theta = 0.5 * h / apq;
                            t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                            if (theta < 0.0) t = -t;
                        }
                        const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                        const double tau = s / (1.0 + c);
                        A[pi][pi] -= t * apq;
                        A[q][q] += t * apq;
                        A[pi][q] = A[q][pi] = 0.0;
#pragma unroll
                        for (int r = 0; r < M; ++r) {
                            if (r != pi && r != q) {
                                const double arp = A[r][pi], arq = A[r][q];
                                const double nrp = arp - s * (arq + arp * tau);
                                const double nrq = arq + s * (arp - arq * tau);
                                A[r][pi] = A[pi][r] = nrp;
                                A[r][q] = A[q][r] = nrq;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < M; ++r) {
                            const double vrp = V[r][pi], vrq = V[r][q];
                            V[r][pi] = vrp - s * (vrq + vrp * tau);
                            V[r][q] = vrq + s * (vrp - vrq * tau);
                        }
                    }
                }
            }
    }
    double lmax = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) lmax = fmax(lmax, fabs(A[k][k]));
    const double cutoff = 1e-15 * lmax;
    // pinv = V diag(1/lambda) V^T ; row sums and grand total of it
    double cw[M];
    double total = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
        double cs = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) cs += V[j][k];
        const double lam = A[k][k];
        const double inv = (fabs(lam) > cutoff) ? 1.0 / lam : 0.0;
        cw[k] = cs * inv;
        total += cs * cw[k];
    }
    double mean = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) r += V[i][k] * cw[k];
        mean += (r / total) * yv[i];
    }
    mean_f[p] = mean;
    var_f[p] = 1.0 / total;
}

extern "C" int bf_reification(int m, long long P, const double* y, const double* var, double* mean_f,
                              double* var_f, void* stream) {
    BF_CHECK_ARG(P >= 0 && y && var && mean_f && var_f, BF_ERR_ARG, "bf_reification: bad argument");
    BF_CHECK_ARG(m >= 1 && m <= BF_MAX_MODELS, BF_ERR_SIZE, "bf_reification: m=%d outside 1..%d", m, BF_MAX_MODELS);
    if (P == 0) return 0;
    const unsigned grid = (unsigned)((P + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    switch (m) {
#define BF_CASE(MM) case MM: bf_reification_kernel<MM><<<grid, 128, 0, st>>>(P, y, var, mean_f, var_f); break;
        BF_CASE(1) BF_CASE(2) BF_CASE(3) BF_CASE(4) BF_CASE(5) BF_CASE(6) BF_CASE(7) BF_CASE(8)
#undef BF_CASE
    }
    BF_CHECK_LAUNCH("bf_reification");
    return 0;
}
