// Version / error plumbing of the C ABI (include/barefoot_sm100.h).
#include <stdarg.h>
#include <string.h>

#include "bf_common.cuh"

static thread_local char g_err[512] = "";

void bf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int bf_version(void) { return BF_VERSION; }
extern "C" const char* bf_last_error(void) { return g_err; }

// ---- FP64 tensor issue-rate probe: the second roofline denominator bench.py reports ----------------------
// MEASURED_PEAKS.json holds no FP64 figure and cuBLAS DGEMM on this image dispatches an sm_80 kernel
// (cutlass_80_tensorop_d884gemm), so next to the DGEMM rate the bench line quotes what the DMMA pipe itself
// sustains: one CTA per SM, 8 warps, 12 independent accumulator chains of mma.sync.m8n8k4.f64 per warp.
__global__ void __launch_bounds__(256) bf_probe_dmma_kernel(double* __restrict__ out, int iters) {
    double c[12][2];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double bf_probe_dmma_flop(int ctas, int iters) { return (double)ctas * 8.0 * 12.0 * 512.0 * (double)iters; }

extern "C" int bf_probe_dmma(int ctas, int iters, double* out, void* stream) {
    BF_CHECK_ARG(ctas > 0 && iters > 0 && out, BF_ERR_ARG, "bf_probe_dmma: bad argument");
    bf_probe_dmma_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(out, iters);
    BF_CHECK_LAUNCH("bf_probe_dmma");
    return 0;
}
