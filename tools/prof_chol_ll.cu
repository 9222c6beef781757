// Phase profile of the fused left-looking Cholesky (build with -DBF_LL_PROFILE): cycles per phase of warp 0 and warp 1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -DBF_LL_PROFILE -o tools/prof_chol_ll tools/prof_chol_ll.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../barefoot_framework_b200/csrc/bf_cholesky_ll.cu"
#include "../barefoot_framework_b200/csrc/bf_api.cu"
#include "../barefoot_framework_b200/csrc/bf_setup.cu"
extern "C" size_t bf_linv_packed_doubles(int n) { int nb = bf_padded_n(n) / 32; return bf_linv_block_off(nb); }
extern "C" int bf_padded_n(int n) { return (n + 31) & ~31; }
int main(int argc, char** argv) {
    const int n = 500, d = 6, np = 512;
    const int S = argc > 1 ? atoi(argv[1]) : 1;
    std::vector<double> X(n * d), il(d * S, 4.0), amp(S, 0.1667), yerr(1, 0.05);
    for (int i = 0; i < n * d; ++i) X[i] = (double)((i * 2654435761u) % 100000) / 100000.0;
    double *dX, *dil, *damp, *dyerr, *dL; int* dinfo;
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dil, il.size() * 8); cudaMalloc(&damp, S * 8); cudaMalloc(&dyerr, 8);
    cudaMalloc(&dL, (size_t)S * np * np * 8); cudaMalloc(&dinfo, 4 * S);
    cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dil, il.data(), il.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(damp, amp.data(), S * 8, cudaMemcpyHostToDevice); cudaMemcpy(dyerr, yerr.data(), 8, cudaMemcpyHostToDevice);
    const char* names[9] = {"w0 tile accumulate", "w0 kernel tiles", "w0 solve+store", "w0 diag accumulate", "w0 factor", "w0 invert",
                            "w0 barrier wait", "w1 column work", "w1 barrier wait"};
    for (int rep = 0; rep < 3; ++rep) {
        long long zero[16] = {0};
        cudaMemcpyToSymbol(bf_ll_prof, zero, sizeof(zero));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int rc;
        if (argc > 2) {     // in-place variant: K from the stand-alone kernel first (not timed)
            bf_kernel_matrix_lower(2, n, d, S, dX, dil, damp, nullptr, dyerr, 1, 0, dL, nullptr);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            rc = bf_cholesky_ll_inplace(n, S, dL, dinfo, nullptr, 0, nullptr);
        } else
        rc = bf_cholesky_fused(2, n, d, S, dX, dil, damp, nullptr, dyerr, 1, 0, dL, dinfo, nullptr, 0, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long prof[16]; cudaMemcpyFromSymbol(prof, bf_ll_prof, sizeof(prof));
        printf("rc %d S %d: %.1f us (%s)\n", rc, S, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        for (int k = 0; k < 9; ++k) printf("   %-22s %10.1f kcycles per set\n", names[k], prof[k] / 1e3 / S);
    }
    return 0;
}
