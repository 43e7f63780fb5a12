// xp_debug.cu -- device self-checks of the arithmetic building blocks of xp_detect_x (xp_narrowphase.cuh).
//
// xp_detect_x is bit-identical to the reference only if (1) x_div / x_sqrt -- the fast halves of the compiler's own
// IEEE sequences, without the per-call range check -- are correctly rounded on the whole guarded operand range, and
// (2) the decisions it takes on the OPERANDS of a division equal the decisions the reference takes on the QUOTIENT.
// Both are checked here on the device against __fdiv_rn / __fsqrt_rn: every square-root input of the guarded range
// (1.7e9 bit patterns), and counter-hashed plus structured operand pairs for the rest.  tests/test_gpu_arith.py
// asserts that every counter comes back zero.
#include "xp_internal.h"
#include "xp_narrowphase.cuh"

namespace xp {

__device__ __forceinline__ uint32_t mix32(uint64_t x) {       // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((x ^ (x >> 31)) >> 16);
}
// a float of random sign and mantissa whose exponent is uniform in [-span, span]
__device__ __forceinline__ float rnd_float(uint64_t key, int span) {
    const uint32_t r = mix32(key), e = mix32(key ^ 0xABCDEF1234567ull);
    const int ex = (int)(e % (uint32_t)(2 * span + 1)) - span + 127;
    return __uint_as_float((r & 0x807FFFFFu) | ((uint32_t)ex << 23));
}

enum { CHK_SQRT = 0, CHK_DIV_RANDOM = 1, CHK_DIV_MANTISSA = 2, CHK_EDGE_PARAM = 3, CHK_FACE_INTERVAL = 4, CHK_VERTEX_MIN = 5, CHK_COUNT = 8 };

// (1a) every bit pattern in [2^-100, 2^100]
__global__ void __launch_bounds__(256) k_chk_sqrt(unsigned long long *bad) {
    const uint32_t lo = __float_as_uint(XP_SQ_LO), hi = __float_as_uint(XP_SQ_HI);
    unsigned n = 0;
    for (uint64_t b = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= hi; b += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)b);
        n += __float_as_uint(x_sqrt(x)) != __float_as_uint(__fsqrt_rn(x));
    }
    if (n) atomicAdd(&bad[CHK_SQRT], (unsigned long long)n);
}

// (1b) quotients: operands of either sign with magnitudes in [2^-50, 2^50]
__global__ void __launch_bounds__(256) k_chk_div(uint64_t n_pairs, uint64_t seed, unsigned long long *bad) {
    unsigned n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        const float a = rnd_float(seed + 2 * i, 50), b = rnd_float(seed + 2 * i + 1, 50);
        n += __float_as_uint(x_div(a, b, x_rcp(b))) != __float_as_uint(__fdiv_rn(a, b));
    }
    if (n) atomicAdd(&bad[CHK_DIV_RANDOM], (unsigned long long)n);
}
// every mantissa of the denominator (the reciprocal's error depends on it alone), several numerators each,
// among them the hard cases a = b * (1 + k ulp) and a = nextafter(b)
__global__ void __launch_bounds__(256) k_chk_div_mantissa(int reps, uint64_t seed, unsigned long long *bad) {
    unsigned n = 0;
    for (uint64_t m = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; m < (1u << 23); m += (uint64_t)gridDim.x * blockDim.x) {
        for (int ex = -1; ex <= 1; ++ex) {
            const float b = __uint_as_float((uint32_t)m | ((uint32_t)(127 + 17 * ex) << 23));
            const float r = x_rcp(b);
            for (int k = 0; k < reps; ++k) {
                float a = rnd_float(seed + m * 64 + k, 30);
                if (k == 0) a = b;
                if (k == 1) a = __uint_as_float(__float_as_uint(b) + 1);
                if (k == 2) a = __uint_as_float(__float_as_uint(b) - 1);
                if (k == 3) a = __fmul_rn(b, rnd_float(seed + m, 20));            // near-exact quotients
                n += __float_as_uint(x_div(a, b, r)) != __float_as_uint(__fdiv_rn(a, b));
                n += __float_as_uint(x_div(-a, b, r)) != __float_as_uint(__fdiv_rn(-a, b));
            }
        }
    }
    if (n) atomicAdd(&bad[CHK_DIV_MANTISSA], (unsigned long long)n);
}

// (2a) edge parameter: 0 <= RN(f / els) <= 1  <=>  0 <= f <= els for els > 0 (collision_detect.rs:106-107)
__global__ void __launch_bounds__(256) k_chk_edge_param(uint64_t n_pairs, uint64_t seed, unsigned long long *bad) {
    unsigned n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        const float els = fabsf(rnd_float(seed + 2 * i, 40));
        float f = rnd_float(seed + 2 * i + 1, 50);
        const uint32_t sel = mix32(seed ^ (i * 7919)) & 15u;
        if (sel < 8) f = __uint_as_float(__float_as_uint(els) + (int)sel - 4);    // els -4 .. +3 ulp
        if (sel == 8) f = 0.0f;
        if (sel == 9) f = -0.0f;
        if (sel == 10) f = -f * 1e-30f;
        const float q = __fdiv_rn(f, els);
        const bool ref = q >= 0.0f && q <= 1.0f;
        const bool mine = f >= 0.0f && f <= els;
        // |f| below 2^-50 (and not zero) is outside the guarded range: xp_detect_x flags it and never decides on it
        const bool guarded = f == 0.0f || fabsf(f) >= XP_DIV_LO;
        n += guarded && ref != mine;
    }
    if (n) atomicAdd(&bad[CHK_EDGE_PARAM], (unsigned long long)n);
}

// (2b) face interval (collision_detect.rs:28-33): the sorted (t0, t1) of u0 / p, u1 / p is rejected iff ... see xp_detect_x
__global__ void __launch_bounds__(256) k_chk_face_interval(uint64_t n_pairs, uint64_t seed, unsigned long long *bad) {
    unsigned n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        float sd = rnd_float(seed + 2 * i, 12);
        float p = rnd_float(seed + 2 * i + 1, 10);
        const uint32_t sel = mix32(seed ^ (i * 104729)) & 15u;
        if (sel < 6) sd = __uint_as_float(__float_as_uint(1.0f) + sel);                    // |sd| = 1 + k ulp
        if (sel == 6) sd = -1.0f;
        if (sel == 7) sd = -__uint_as_float(__float_as_uint(1.0f) + 3);
        if (!(fabsf(sd) >= 1.0f)) continue;                                                // the block runs only for |sd| >= 1
        const float u0 = __fsub_rn(1.0f, sd), u1 = __fsub_rn(-1.0f, sd);
        if (sel >= 8 && sel < 12) p = __uint_as_float(__float_as_uint(sel & 1 ? u0 : u1) + (int)(sel & 2) - 1);   // quotient next to 1
        if (p == 0.0f) continue;
        float t0 = __fdiv_rn(u0, p), t1 = __fdiv_rn(u1, p);
        if (!(t0 < t1)) { const float tmp = t0; t0 = t1; t1 = tmp; }
        const bool ref = t0 > 1.0f || t1 < 0.0f;
        const bool mine = (p > 0.0f) ? (u1 > p || u0 < 0.0f) : (u0 < p || u1 > 0.0f);
        n += ref != mine;
    }
    if (n) atomicAdd(&bad[CHK_FACE_INTERVAL], (unsigned long long)n);
}

// (2c) the three vertex roots share 2a > 0: min_i RN(n_i / den) == RN(min_i n_i / den)
__global__ void __launch_bounds__(256) k_chk_vertex_min(uint64_t n_pairs, uint64_t seed, unsigned long long *bad) {
    unsigned n = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        const float den = fabsf(rnd_float(seed + 4 * i, 40));
        const float a = rnd_float(seed + 4 * i + 1, 30);
        float b = rnd_float(seed + 4 * i + 2, 30), c = rnd_float(seed + 4 * i + 3, 30);
        if (i & 1) b = __uint_as_float(__float_as_uint(a) + 1);                            // neighbours: quotients may tie
        if (i & 2) c = __uint_as_float(__float_as_uint(a) - 1);
        const float ref = fminf(fminf(__fdiv_rn(a, den), __fdiv_rn(b, den)), __fdiv_rn(c, den));
        const float mine = x_div(x_min3(a, b, c), den, x_rcp(den));
        n += ref != mine;                                                                  // values (a zero's sign is invisible downstream)
    }
    if (n) atomicAdd(&bad[CHK_VERTEX_MIN], (unsigned long long)n);
}

#define XP_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t check_arith(uint64_t n_random, uint64_t seed, unsigned long long *h_bad) {
    int dev = 0, n_sm = 1;
    XP_TRY(cudaGetDevice(&dev));
    XP_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long *d = nullptr;
    XP_TRY(cudaMalloc(&d, CHK_COUNT * sizeof(unsigned long long)));
    cudaError_t e = cudaMemset(d, 0, CHK_COUNT * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)n_sm * 8;
        k_chk_sqrt<<<grid, 256>>>(d);
        k_chk_div<<<grid, 256>>>(n_random, seed, d);
        k_chk_div_mantissa<<<grid, 256>>>(24, seed ^ 0x5555, d);
        k_chk_edge_param<<<grid, 256>>>(n_random, seed ^ 0x1111, d);
        k_chk_face_interval<<<grid, 256>>>(n_random, seed ^ 0x2222, d);
        k_chk_vertex_min<<<grid, 256>>>(n_random / 4, seed ^ 0x3333, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(h_bad, d, CHK_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e;
}

}  // namespace xp
