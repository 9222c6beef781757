// Micro-benchmark: FP64 DMMA (mma.sync m8n8k4) throughput per SM versus warps per SM and independent
// accumulator chains per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void probe(double* out, int iters) {
    double c[CH][2];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(int warps, double* out) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<CH><<<148, warps * 32>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<CH><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 148.0 * warps * iters * CH * 512.0;   // 8*8*4*2 flop per DMMA
    printf("chains %2d warps/SM %2d : %7.2f TFLOP/s  (%.1f clk per DMMA per SM at 1.965 GHz)\n", CH, warps,
           flops / ms / 1e9, ms * 1e-3 * 1.965e9 / (double(warps) * iters * CH));
}

int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
    for (int w : {4, 8, 12, 16, 24, 32}) { run<4>(w, out); run<12>(w, out); run<24>(w, out); }
    return 0;
}
