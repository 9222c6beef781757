// Phase profile of the setup kernels (build with -DBF_PROFILE_SETUP): cycles per phase for one 500 x 500 matrix.
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../barefoot_framework_b200/csrc/bf_setup.cu"
#include "../barefoot_framework_b200/csrc/bf_api.cu"
extern "C" int bf_padded_n(int n) { return (n + 31) & ~31; }
extern "C" size_t bf_linv_packed_doubles(int n) { int nb = bf_padded_n(n) / 32; return bf_linv_block_off(nb); }
int main() {
    const int n = 500, d = 6, np = 512;
    std::vector<double> X(n * d), il(d, 4.0), amp(1, 0.1667), yerr(1, 0.05);
    for (int i = 0; i < n * d; ++i) X[i] = (double)((i * 2654435761u) % 100000) / 100000.0;
    double *dX, *dil, *damp, *dyerr, *dK; int* dinfo;
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dil, d * 8); cudaMalloc(&damp, 8); cudaMalloc(&dyerr, 8);
    cudaMalloc(&dK, (size_t)np * np * 8); cudaMalloc(&dinfo, 4);
    double *dy, *dLinv, *dpacked, *dalpha;
    cudaMalloc(&dy, np * 8); cudaMemset(dy, 0, np * 8); cudaMalloc(&dLinv, (size_t)np * np * 8);
    cudaMalloc(&dpacked, bf_linv_packed_doubles(n) * 8); cudaMalloc(&dalpha, np * 8);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dil, il.data(), d * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(damp, amp.data(), 8, cudaMemcpyHostToDevice); cudaMemcpy(dyerr, yerr.data(), 8, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) {
        bf_kernel_matrix(0, n, d, 1, dX, dil, damp, nullptr, dyerr, 1, 0, dK, nullptr);
        long long zero[8] = {0};
        cudaMemcpyToSymbol(bf_prof_cycles, zero, sizeof(zero));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int rc = bf_cholesky_batched(n, 1, dK, dinfo, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long prof[8]; cudaMemcpyFromSymbol(prof, bf_prof_cycles, sizeof(prof));
        int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
        printf("rc %d info %d  cholesky %.1f us : first diagonal block %lld, panels %lld, trailing updates incl. look-ahead diagonal blocks %lld cycles (%s)\n", rc, info, ms * 1e3, prof[0], prof[1],
               prof[2], cudaGetErrorString(cudaGetLastError()));
        cudaMemcpyToSymbol(bf_prof_cycles, zero, sizeof(zero));
        cudaEventRecord(e0);
        rc = bf_factor_finish(n, 1, dK, dy, 0, dLinv, dpacked, dalpha, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpyFromSymbol(prof, bf_prof_cycles, sizeof(prof));
        printf("rc %d finish %.1f us (S = 1 runs the column-group kernels, which carry no phase marks; one-CTA-per-set kernel: diag-inverses %lld W %lld rows %lld alpha %lld cycles) (%s)\n", rc, ms * 1e3, prof[3], prof[4],
               prof[5], prof[6], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
