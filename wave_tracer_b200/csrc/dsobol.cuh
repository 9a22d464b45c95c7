// dsobol.cuh -- the reference's "sobolld" sampler (base-3 Sobol' sequence with nested uniform / Owen scrambling) computed in registers.
//
// Reference: include/wt/sampler/sobolld/sobolld_sampler.hpp:59-207 generates a whole batch of 3^11 points x 47 dimensions
// incrementally (point i from point i-1, :156-176) into a thread-local vector.  A GPU thread wants the value of ONE (point, dimension)
// from its index, so the point is computed directly: digit j of dimension `dim` of point i is
//     x_j = sum_k C_dim[M-1-j][k] * i_k  (mod 3)                                   (what :166-171 accumulates over the changed digits)
// with the generator matrices C (gen_mat, :140-154) kept in constant memory as two 11-bit masks per row (entries == 1, entries == 2),
// the index digits i_k as the same two masks, and the dot product over GF(3) as four popcounts.  The nested scrambling (:181-207) walks
// the permutation tree from the most significant digit with the reference's 64-bit multiplicative hash (:124-131).
// Integer arithmetic only: bit-exact with the oracle's literal restatement (tests/test_sobol.py, tests/test_gpu_parity.py).
#pragma once
#include <stdint.h>
#include "../../include/wtgpu.h"

namespace wt {

struct SobolTables {
    uint16_t ones[WTGPU_SOBOL_DIMS][WTGPU_SOBOL_DIGITS];    // bit k of ones[dim][j]: C_dim[M-1-j][k] == 1
    uint16_t twos[WTGPU_SOBOL_DIMS][WTGPU_SOBOL_DIGITS];    //                               ... == 2
};
__constant__ SobolTables c_sobol;

constexpr uint32_t kSobolPoints = 177147u;                  // 3^11
constexpr uint32_t kSobolStreamFlag = 0x80000000u;          // Sampler::stream = flag | spp selects the sobolld scene sampler

// Philox4x32-10, lane 0 (same rounds as philox_lane() in dscene.cuh; kept local so this header stands alone)
__device__ __forceinline__ uint32_t sobol_philox0(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// rng_t::hash (sobolld_sampler.hpp:124-131)
__device__ __forceinline__ uint64_t sobol_hash(uint64_t x) {
    x ^= x >> 16; x *= 0x21f0aaadull; x ^= x >> 15; x *= 0xd35a2d97ull; x ^= x >> 15;
    return x;
}

// Numerator (value * 3^11) of dimension `dim` of point `g` of the global sequence; batch g / 3^11 has its own scrambling seeds.
__device__ __noinline__ uint32_t sobol_numerator(uint32_t k0, uint32_t k1, uint64_t g, uint32_t dim) {
    const uint64_t batch = g / kSobolPoints;
    uint32_t i = (uint32_t)(g - batch * kSobolPoints);
    // index digits as masks (integer3.hpp:29-32)
    uint32_t i1 = 0u, i2 = 0u;
#pragma unroll
    for (uint32_t k = 0; k < WTGPU_SOBOL_DIGITS; ++k) {
        const uint32_t q = i / 3u, dg = i - 3u * q; i = q;
        i1 |= (dg == 1u ? 1u : 0u) << k; i2 |= (dg == 2u ? 1u : 0u) << k;
    }
    // per-(batch, dimension) scrambling seed: a 32-bit draw, as the reference's uniform_int_distribution<unsigned> (src/sampler/sobolld.cpp:56-58)
    const uint64_t seed = sobol_philox0(k0, k1, dim, (uint32_t)batch, (uint32_t)(batch >> 32), 0x50B01Du);
    const uint64_t key = (seed << 1) | 1ull;                                    // rng_t ctor (:104-105)
    const uint64_t divisor = ((0ull - 6ull) / 6ull) + 1ull;                     // sample_range(6) (:115-122)
    // permutations of {0,1,2} in the reference's order (:183-190), 2 bits per entry, entry = flip*3 + digit
    const uint64_t perms = 0ull | (1ull << 2) | (2ull << 4)       // {0,1,2}
                         | (0ull << 6) | (2ull << 8) | (1ull << 10)   // {0,2,1}
                         | (1ull << 12) | (0ull << 14) | (2ull << 16) // {1,0,2}
                         | (1ull << 18) | (2ull << 20) | (0ull << 22) // {1,2,0}
                         | (2ull << 24) | (0ull << 26) | (1ull << 28) // {2,0,1}
                         | (2ull << 30) | (1ull << 32) | (0ull << 34);// {2,1,0}
    uint64_t node = 0ull;
    uint32_t value = 0u;
    uint32_t p3 = 59049u;                                                       // 3^(M-1)
#pragma unroll 1
    for (int j = (int)WTGPU_SOBOL_DIGITS - 1; j >= 0; --j) {
        const uint32_t r1 = c_sobol.ones[dim][j], r2 = c_sobol.twos[dim][j];
        const uint32_t s = __popc(r1 & i1) + __popc(r2 & i2) + 2u * (__popc(r1 & i2) + __popc(r2 & i1));
        const uint32_t digit = s % 3u;
        uint64_t n = node, x;
        do { x = sobol_hash(++n * key) / divisor; } while (x >= 6ull);          // rng.index(node).sample_range(6)
        const uint32_t y = (uint32_t)(perms >> (2u * ((uint32_t)x * 3u + digit))) & 3u;
        value += y * p3; p3 /= 3u;
        node = 3ull * node + 1ull + digit;
    }
    return value;
}
// integer3_t::value_fp (integer3.hpp:45-48)
__device__ __forceinline__ float sobol_value(uint32_t numerator) { return (float)numerator / 177147.f; }

// scene-sampler draw d of the sample whose first point is g (flat layout of the reference's batch: point-major, 47 values per point)
__device__ __forceinline__ float sobol_draw(uint32_t k0, uint32_t k1, uint64_t g, uint32_t d) {
    const uint32_t q = d / WTGPU_SOBOL_DIMS;
    return sobol_value(sobol_numerator(k0, k1, g + q, d - q * WTGPU_SOBOL_DIMS));
}

} // namespace wt
