// zb_gf2.h — GF(2) polynomial arithmetic modulo the CRC-32 polynomial, shared
// by host code (crc32_combine family, table generation) and device code
// (per-thread partial weighting in the checksum kernels).
//
// Replaces, for the B200 path, the reference's multmodp / x2nmodp /
// x2n_table (crc32.c:155-187, crc32.h:9439-9446).  Representation follows the
// reference: reflected polynomials, bit 31 holds the x^0 coefficient.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZB_HD __host__ __device__ __forceinline__
#else
#define ZB_HD inline
#endif

namespace zb {

constexpr uint32_t kCrcPoly = 0xedb88320u;

// a(x) * b(x) mod p(x)
ZB_HD uint32_t gf2_mul(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    // walk the 32 coefficients of `a` from x^0 (bit 31) upward; b is multiplied
    // by x each step.  Branch-free so that a warp stays converged.
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        p ^= b & (0u - ((a >> (31 - i)) & 1u));
        b = (b >> 1) ^ (kCrcPoly & (0u - (b & 1u)));
    }
    return p;
}

// x^(2^k) mod p for k = 0..31, filled by gf2_fill_x2n().
struct X2nTable { uint32_t v[32]; };

inline void gf2_fill_x2n(X2nTable &t) {
    uint32_t p = 0x40000000u;   // x^1
    t.v[0] = p;
    for (int k = 1; k < 32; ++k) t.v[k] = p = gf2_mul(p, p);
}

// x^(n * 2^k) mod p(x)
ZB_HD uint32_t gf2_xpow(const uint32_t *x2n, uint64_t n, unsigned k) {
    uint32_t p = 0x80000000u;   // 1
    while (n) {
        if (n & 1) p = gf2_mul(x2n[k & 31], p);
        n >>= 1;
        ++k;
    }
    return p;
}

// One byte through the plain (init 0, no final xor) CRC register.
ZB_HD uint32_t crc_byte_bitwise(uint32_t c, uint32_t byte) {
    c ^= byte;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (kCrcPoly & (0u - (c & 1u)));
    return c;
}

constexpr uint32_t kAdlerBase = 65521u;   // adler32.c:10

}  // namespace zb
