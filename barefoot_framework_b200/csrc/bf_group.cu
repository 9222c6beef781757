// Segmented "first strict maximum" over work-item records: the device form of the dictionary loop of
// barefoot.__kg_calc_clustering (barefoot.py:2102-2140), which keeps for every (selected candidate x*,
// model) pair the largest acquisition value and the row it came from.
//
// Reference scan per group, records in row order: the first record sets (value, row); a later record
// replaces them only if its value is STRICTLY greater (barefoot.py:2116-2118).  Hence: a NaN never
// replaces anything and a NaN first record is never replaced; among equal maxima the first row wins;
// -0.0 and +0.0 compare equal.  Three passes with 64-bit atomics on a G-entry table reproduce exactly that.
#include "bf_common.cuh"

#define GB_THREADS 256

// order-preserving map double -> uint64 (NaN excluded by the callers); -0.0 is folded into +0.0
__device__ __forceinline__ unsigned long long gb_encode(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}

__global__ void __launch_bounds__(GB_THREADS) bf_gb_init_kernel(long long G, long long R, long long* __restrict__ first_row,
                                                                unsigned long long* __restrict__ enc,
                                                                long long* __restrict__ best_row) {
    const long long g = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
    if (g >= G) return;
    first_row[g] = R;
    enc[g] = 0ULL;                 // below every encoded value
    best_row[g] = R;
}

__global__ void __launch_bounds__(GB_THREADS) bf_gb_max_kernel(long long R, long long G, const long long* __restrict__ key,
                                                               const double* __restrict__ value,
                                                               long long* __restrict__ first_row,
                                                               unsigned long long* __restrict__ enc,
                                                               int32_t* __restrict__ bad_key) {
    const long long r = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
    if (r >= R) return;
    const long long g = key[r];
    if (g < 0 || g >= G) {
        atomicAdd(bad_key, 1);
        return;
    }
    atomicMin((unsigned long long*)&first_row[g], (unsigned long long)r);
    const double v = value[r];
    if (v == v) atomicMax(&enc[g], gb_encode(v));
}

__global__ void __launch_bounds__(GB_THREADS) bf_gb_row_kernel(long long R, long long G, const long long* __restrict__ key,
                                                               const double* __restrict__ value,
                                                               const unsigned long long* __restrict__ enc,
                                                               long long* __restrict__ best_row) {
    const long long r = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
    if (r >= R) return;
    const long long g = key[r];
    if (g < 0 || g >= G) return;
    const double v = value[r];
    if (v == v && gb_encode(v) == enc[g]) atomicMin((unsigned long long*)&best_row[g], (unsigned long long)r);
}

// enc and best_val alias (same 8-byte slots): decode in place
__global__ void __launch_bounds__(GB_THREADS) bf_gb_final_kernel(long long G, long long R, const double* __restrict__ value,
                                                                 const long long* __restrict__ first_row,
                                                                 double* __restrict__ best_val,
                                                                 long long* __restrict__ best_row) {
    const long long g = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
    if (g >= G) return;
    const long long f = first_row[g];
    if (f >= R) {                  // empty group
        best_val[g] = 0.0;
        best_row[g] = -1;
        return;
    }
    const double v0 = value[f];
    if (v0 != v0) {                // a NaN first record is never replaced
        best_val[g] = v0;
        best_row[g] = f;
    } else {
        best_val[g] = value[best_row[g]];
    }
}

extern "C" int bf_group_best(long long R, long long G, const long long* key, const double* value,
                             long long* first_row, double* best_val, long long* best_row, int32_t* bad_key,
                             void* stream) {
    BF_CHECK_ARG(R >= 0 && G > 0 && first_row && best_val && best_row && bad_key, BF_ERR_ARG, "bf_group_best: bad argument");
    BF_CHECK_ARG(R == 0 || (key && value), BF_ERR_ARG, "bf_group_best: null records");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned gg = (unsigned)((G + GB_THREADS - 1) / GB_THREADS);
    const unsigned gr = (unsigned)((R + GB_THREADS - 1) / GB_THREADS);
    unsigned long long* enc = reinterpret_cast<unsigned long long*>(best_val);
    bf_gb_init_kernel<<<gg, GB_THREADS, 0, st>>>(G, R, first_row, enc, best_row);
    if (R > 0) {
        bf_gb_max_kernel<<<gr, GB_THREADS, 0, st>>>(R, G, key, value, first_row, enc, bad_key);
        bf_gb_row_kernel<<<gr, GB_THREADS, 0, st>>>(R, G, key, value, enc, best_row);
    }
    bf_gb_final_kernel<<<gg, GB_THREADS, 0, st>>>(G, R, value, first_row, best_val, best_row);
    BF_CHECK_LAUNCH("bf_group_best");
    return 0;
}
