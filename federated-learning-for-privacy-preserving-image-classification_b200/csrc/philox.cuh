// Philox4x32-10 counter-based generator + Box-Muller (device side).
// Stream layout (restated by oracle/philox.py for the tests):
//   counter = (block_lo, block_hi, stream_lo, stream_hi), key = (seed_lo, seed_hi)
//   block   = element_index / 4; the four 32-bit outputs feed elements 4*block .. 4*block+3
//   u       = fma(x, 2^-32, 2^-33) clamped below 1      (so u is in (0, 1))
//   z0, z1  = sqrt(-2 ln u0) * (cos 2 pi u1, sin 2 pi u1);  z2, z3 likewise from (u2, u3)
// The result for an element depends only on (seed, stream, element index), never on the grid
// shape or on how clients are spread over GPUs.
#pragma once
#include <stdint.h>

struct flb_u4 { uint32_t x, y, z, w; };

__device__ __forceinline__ flb_u4 flb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    flb_u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

__device__ __forceinline__ flb_u4 flb_philox_block(unsigned long long seed, unsigned long long stream,
                                                  unsigned long long block) {
    return flb_philox4x32_10((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)stream,
                             (uint32_t)(stream >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ float flb_u01(uint32_t x) {
    float u = __fmaf_rn((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    return fminf(u, 0.99999994f);
}

// four standard normals for one Philox block.  The clip + noise kernel is ALU-bound on this function (Philox is ~100
// integer instructions per block), so the transcendental part uses the hardware approximations: lg2.approx (absolute
// error < 2^-22 in log2 u), x * rsqrt(x) for the square root, and sin/cos.approx on an argument reduced to [-pi, pi)
// (absolute error < 2^-20.9); the results agree with the double-precision transform of oracle/philox.py to ~2e-6.
__device__ __forceinline__ float4 flb_normal4(unsigned long long seed, unsigned long long stream,
                                              unsigned long long block) {
    const flb_u4 r = flb_philox_block(seed, stream, block);
    // -2 ln u = -2 ln2 * log2 u; clamped away from 0 (u < 1, but the approximation may return +0 next to 1)
    const float x0 = fmaxf(-1.3862943611198906f * __log2f(flb_u01(r.x)), 1e-30f);
    const float x1 = fmaxf(-1.3862943611198906f * __log2f(flb_u01(r.z)), 1e-30f);
    const float r0 = -(x0 * rsqrtf(x0)), r1 = -(x1 * rsqrtf(x1));       // minus: the angle below is shifted by pi
    const float t0 = fmaf(flb_u01(r.y), 6.283185307179586f, -3.141592653589793f);
    const float t1 = fmaf(flb_u01(r.w), 6.283185307179586f, -3.141592653589793f);
    return make_float4(r0 * __cosf(t0), r0 * __sinf(t0), r1 * __cosf(t1), r1 * __sinf(t1));
}
