// zb_checksum.cuh — per-thread core of the CRC-32 / Adler-32 folding kernels.
//
// B200 replacement for crc32.c:694-1010 (crc32_z, braided N=5/W=8) and
// adler32.c:61-125 (adler32_z).  NVIDIA SMs have no carry-less multiply, so
// CRC folding is done as a strided Horner evaluation over GF(2):
//
//   a part of the buffer (16-byte aligned body) is read as uint4 blocks; thread
//   t owns blocks t, t+T, t+2T, ... (T = threads per CTA, so a warp reads 512
//   contiguous bytes per step and the CTA reads T*16 contiguous bytes).  For
//   each of the four 32-bit word slots of its blocks the thread keeps
//       V_k  <-  V_k * x^(8*16*T)  xor  w_k            (one step per block)
//   where the multiplication by the fixed power of x is four 256-entry table
//   lookups (tables in shared memory; in the big-buffer kernel every lane owns
//   a private bank-conflict-free copy).  At the end V is weighted by
//   x^(8 * bytes-after-it) with the reference's own x2nmodp/multmodp algebra
//   (crc32.c:155-187) and all partials are XOR-ed together, exactly the
//   crc32_combine identity (crc32.c:1021) applied per thread.
//
//   Adler-32: per block s = sum(b_i), t = sum(i*b_i) via dp4a; the position
//   weights (len - offset) are folded in once per thread from running sums, and
//   partial (s1, s2) pairs add up modulo 65521 (the adler32_combine identity,
//   adler32.c:133-155).
//
// The per-thread routine is __host__ __device__ so that tests can replay the
// exact thread decomposition on the CPU before a GPU is involved
// (tests/emul/).  The product only ever runs it on the device.
#pragma once
#include "zb_gf2.h"

namespace zb {

struct U4 { uint32_t x, y, z, w; };

ZB_HD uint32_t dot4(uint32_t word, uint32_t weights, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(word, weights, acc);
#else
    for (int i = 0; i < 4; ++i) acc += ((word >> (8 * i)) & 0xff) * ((weights >> (8 * i)) & 0xff);
    return acc;
#endif
}

ZB_HD U4 load_block(const U4 *p) {
#if defined(__CUDA_ARCH__)
    U4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return *p;
#endif
}

struct CkPartial { uint32_t crc, a, b; };

// Horner table for a CTA of T threads: row j, entry b = (b << 8j) * x^(8*16*T) mod p.
inline void ck_fill_horner(uint32_t out[4][256], const uint32_t *x2n, uint32_t T) {
    const uint32_t xs = gf2_xpow(x2n, 16ull * T, 3);
    for (int j = 0; j < 4; ++j)
        for (uint32_t b = 0; b < 256; ++b) out[j][b] = gf2_mul(b << (8 * j), xs);
}

// Geometry of one part [s, e) of a segment that starts at address seg and is
// seg_len bytes long: unaligned head, 16-byte aligned body, unaligned tail.
struct CkPart {
    const uint8_t *seg;
    uint64_t seg_len;
    uint64_t s, e;            // byte offsets within the segment
    uint64_t body_s, body_e;  // 16-byte aligned (by address) sub-range
};

ZB_HD CkPart ck_make_part(const uint8_t *seg, uint64_t seg_len, uint32_t part, uint32_t parts) {
    CkPart p;
    p.seg = seg; p.seg_len = seg_len;
    const uint64_t addr = (uint64_t)(uintptr_t)seg;
    auto cut = [&](uint32_t i) -> uint64_t {
        if (i == 0) return 0;
        if (i >= parts) return seg_len;
        uint64_t raw = (uint64_t)(((unsigned __int128)seg_len * i) / parts);
        uint64_t a = (addr + raw) & ~(uint64_t)15;          // align the cut by address
        return a <= addr ? 0 : (a - addr > seg_len ? seg_len : a - addr);
    };
    p.s = cut(part); p.e = cut(part + 1);
    uint64_t bs = ((addr + p.s + 15) & ~(uint64_t)15) - addr;
    if (bs > p.e) bs = p.e;
    uint64_t be = ((addr + p.e) & ~(uint64_t)15);
    be = be <= addr + bs ? bs : be - addr;
    p.body_s = bs; p.body_e = be;
    return p;
}

// Contribution of thread t (of T) to the checksums of the segment, over the
// body blocks of part `p`.  TabFn(j, v) returns the Horner table entry for byte j of v.
template <bool DO_CRC, bool DO_ADLER, int UNROLL = 4, class TabFn>
ZB_HD CkPartial ck_thread_body(const CkPart &p, uint32_t t, uint32_t T, TabFn tab,
                               const uint32_t *x2n, uint32_t x32) {
    CkPartial out{0, 0, 0};
    const uint64_t nblk = (p.body_e - p.body_s) >> 4;
    if (t >= nblk) return out;
    const uint64_t M = (nblk - t + T - 1) / T;            // blocks owned by this thread
    const U4 *blk = reinterpret_cast<const U4 *>(p.seg + p.body_s) + t;

    uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;              // CRC word-slot states
    uint64_t asum = 0;                                    // sum of block byte sums
    uint64_t csum = 0, tsum = 0;                          // sum_m sum_{j<m} s_j ; sum of in-block weighted sums

    auto step = [&](const U4 &w) {
        if (DO_CRC) {
            v0 = tab(0, v0) ^ tab(1, v0) ^ tab(2, v0) ^ tab(3, v0) ^ w.x;
            v1 = tab(0, v1) ^ tab(1, v1) ^ tab(2, v1) ^ tab(3, v1) ^ w.y;
            v2 = tab(0, v2) ^ tab(1, v2) ^ tab(2, v2) ^ tab(3, v2) ^ w.z;
            v3 = tab(0, v3) ^ tab(1, v3) ^ tab(2, v3) ^ tab(3, v3) ^ w.w;
        }
        if (DO_ADLER) {
            uint32_t s = dot4(w.x, 0x01010101u, 0);
            s = dot4(w.y, 0x01010101u, s);
            s = dot4(w.z, 0x01010101u, s);
            s = dot4(w.w, 0x01010101u, s);
            uint32_t ws = dot4(w.x, 0x03020100u, 0);
            ws = dot4(w.y, 0x07060504u, ws);
            ws = dot4(w.z, 0x0b0a0908u, ws);
            ws = dot4(w.w, 0x0f0e0d0cu, ws);
            csum += asum;
            asum += s;
            tsum += ws;
        }
    };

    uint64_t m = 0;
    for (; m + UNROLL <= M; m += UNROLL) {                // UNROLL x 16 B in flight per thread
        U4 w[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) w[j] = load_block(blk + (m + j) * T);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) step(w[j]);
    }
    for (; m < M; ++m) step(load_block(blk + m * T));

    const uint64_t o_first = p.body_s + 16ull * t;        // segment offset of the first owned block
    const uint64_t o_end = o_first + 16ull * T * (M - 1) + 16;
    if (DO_CRC) {
        // slots sit 4 bytes apart: fold them into one value relative to the block end
        uint32_t u = gf2_mul(v0, x32) ^ v1;
        u = gf2_mul(u, x32) ^ v2;
        u = gf2_mul(u, x32) ^ v3;
        u = gf2_mul(u, x32);                              // data word -> CRC register scale (x^32)
        out.crc = gf2_mul(u, gf2_xpow(x2n, p.seg_len - o_end, 3));
    }
    if (DO_ADLER) {
        const uint64_t P = kAdlerBase;
        const uint64_t a = asum % P;
        // sum_m m*s_m = (M-1)*asum - csum
        const uint64_t ms = (((M - 1) % P) * a + P - (csum % P)) % P;
        const uint64_t stride = (16ull * T) % P;
        uint64_t b = (((p.seg_len - o_first) % P) * a) % P;
        b = (b + P - (stride * ms) % P) % P;
        b = (b + P - (tsum % P)) % P;
        out.a = (uint32_t)a; out.b = (uint32_t)b;
    }
    return out;
}

// The (at most 15 + 15) unaligned head / tail bytes of a part, done by one thread.
template <bool DO_CRC, bool DO_ADLER>
ZB_HD CkPartial ck_edge_bytes(const CkPart &p, const uint32_t *x2n) {
    CkPartial out{0, 0, 0};
    const uint64_t P = kAdlerBase;
    uint64_t a = 0, b = 0;
    for (int side = 0; side < 2; ++side) {
        const uint64_t lo = side ? p.body_e : p.s, hi = side ? p.e : p.body_s;
        if (hi <= lo) continue;
        uint32_t c = 0;
        for (uint64_t j = lo; j < hi; ++j) {
            const uint32_t byte = p.seg[j];
            if (DO_CRC) c = crc_byte_bitwise(c, byte);
            if (DO_ADLER) { a += byte; b += ((p.seg_len - j) % P) * byte; }
        }
        if (DO_CRC) out.crc ^= gf2_mul(c, gf2_xpow(x2n, p.seg_len - hi, 3));
    }
    out.a = (uint32_t)(a % P); out.b = (uint32_t)(b % P);
    return out;
}

// Final fix-up for one segment: fold in the caller's running values.
// crc32(init, D) = ~( R(D) xor (~init) * x^(8*len) ),   R = pure CRC register
// adler32(init, D): s1 = a0 + sum, s2 = b0 + len*a0 + weighted sum  (mod 65521)
ZB_HD void ck_finish(uint32_t acc_crc, uint64_t acc_a, uint64_t acc_b, uint64_t len,
                     uint32_t init_crc, uint32_t init_adler, const uint32_t *x2n,
                     uint32_t *crc, uint32_t *adler) {
    if (crc) *crc = ~(acc_crc ^ gf2_mul(~init_crc, gf2_xpow(x2n, len, 3)));
    if (adler) {
        const uint64_t P = kAdlerBase;
        const uint64_t a0 = init_adler & 0xffff, b0 = (init_adler >> 16) & 0xffff;
        const uint64_t a = (a0 + acc_a) % P;
        const uint64_t b = (b0 + ((len % P) * a0) % P + acc_b) % P;
        *adler = (uint32_t)(a | (b << 16));
    }
}

}  // namespace zb
