// Bit-identity check of the branch-free device exp / sqrt (csrc/bf_common.cuh) against their first forms
// (fmax clamp + FP64 NaN test; CUDA library rsqrt): every q in [0, 707) must give the same bits (beyond, both
// forms are below 9e-308).  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/check_radial tools/check_radial.cu && /tmp/check_radial
// Prints "OK <count>" or the first mismatches.  tests/test_gpu_kernels.py::test_radial_forms_bit_identical runs it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../barefoot_framework_b200/csrc/bf_common.cuh"

void bf_set_error(const char*, ...) {}

__device__ double ref_exp_nonpos(double x, const double* tab) {
    const double xc = fmax(x, -800.0);
    const double kd0 = fma(xc, 92.33248261689366, 6755399441055744.0);
    const int ki = __double2loint(kd0);
    const double kd = kd0 - 6755399441055744.0;
    double r = fma(kd, -0.01083042469326756, xc);
    r = fma(kd, -2.9815858269852933e-12, r);
    double q = fma(r, 0.008333333333333333, 0.041666666666666664);
    q = fma(q, r, 0.16666666666666666);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;
    const double t = tab[ki & (BF_EXP_TAB_N - 1)];
    const double e = fma(t, q, t);
    const int m = ki >> 6;
    const double res = __hiloint2double(__double2hiint(e) + (m << 20), __double2loint(e));
    return (x != x) ? x : (m < -1021 ? 0.0 : res);
}
__device__ double ref_sqrt_nonneg(double q) {
    const double y = rsqrt(q);
    double r = q * y;
    r = fma(fma(-r, r, q), 0.5 * y, r);
    return q > 0.0 ? r : (q == 0.0 ? 0.0 : q * y);
}

__global__ void check(const double* q, long long n, unsigned long long* bad, double* first) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = q[i];
    const double e0 = ref_exp_nonpos(-v, bf_exp_tab_g), e1 = bf_exp_neg(v, bf_exp_tab_g);
    const double s0 = ref_sqrt_nonneg(v), s1 = bf_sqrt_nonneg(v);
    const bool nan_in = v != v;
    bool ok_e = nan_in ? (e1 != e1) : (__double_as_longlong(e0) == __double_as_longlong(e1));
    // q >= 707: the new form is clamped there (exp(-707) = 8.99e-308, a normal number: no flush test), the
    // first form returns the true value (<= 8.99e-308) or 0 - both below 9e-308
    if (!nan_in && v >= 707.0) ok_e = e1 >= 0.0 && e1 <= 9.0e-308 && e0 >= 0.0 && e0 <= 9.0e-308;
    bool ok_s = (s0 != s0) ? (s1 != s1) : (__double_as_longlong(s0) == __double_as_longlong(s1));
    // subnormal q: the new form returns exactly 0 (documented; sqrt < 1.5e-154)
    if (v > 0.0 && v < 2.2250738585072014e-308) ok_s = s1 == 0.0;
    if (!ok_e || !ok_s) {
        if (atomicAdd(bad, 1ULL) == 0) { first[0] = v; first[1] = e0; first[2] = e1; first[3] = s0; first[4] = s1; }
    }
}

int main() {
    const long long n = 1LL << 26;
    std::vector<double> h(n);
    unsigned long long s = 88172645463325252ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (long long i = 0; i < n; ++i) {
        const unsigned long long r = rnd();
        const int mode = (int)(i & 7);
        double v;
        if (mode < 3) v = (double)(r >> 11) * (1.0 / 9007199254740992.0) * 40.0;                   // [0, 40)
        else if (mode < 5) v = (double)(r >> 11) * (1.0 / 9007199254740992.0) * 900.0;             // [0, 900): clamp and flush
        else if (mode == 5) v = ldexp((double)(r >> 11) * (1.0 / 9007199254740992.0), (int)(rnd() % 120) - 100);  // tiny .. 2^20
        else { unsigned long long b = r & 0x7fffffffffffffffULL; memcpy(&v, &b, 8); }                // any non-negative pattern (incl. subnormal, inf, NaN)
        h[i] = v;
    }
    const double edge[] = {0.0, 800.0, 799.9999999999999, 800.0000000000001, 745.2, 708.4, 1e-300, 1e-320, 1e300};
    for (size_t k = 0; k < sizeof(edge) / sizeof(edge[0]); ++k) h[k] = edge[k];
    { unsigned long long b = 0x7ff0000000000000ULL; memcpy(&h[16], &b, 8); b = 0x7ff8000000000000ULL; memcpy(&h[17], &b, 8); b = 0x7ff0000000000001ULL; memcpy(&h[18], &b, 8); }
    double *d, *first; unsigned long long* bad;
    cudaMalloc(&d, n * 8); cudaMalloc(&first, 5 * 8); cudaMalloc(&bad, 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemset(bad, 0, 8);
    check<<<(unsigned)((n + 255) / 256), 256>>>(d, n, bad, first);
    unsigned long long nb = 0; double f[5];
    if (cudaMemcpy(&nb, bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 2; }
    cudaMemcpy(f, first, 40, cudaMemcpyDeviceToHost);
    if (nb) { printf("MISMATCH %llu first q=%.17g exp %.17g vs %.17g sqrt %.17g vs %.17g\n", nb, f[0], f[1], f[2], f[3], f[4]); return 1; }
    printf("OK %lld\n", n);
    return 0;
}
