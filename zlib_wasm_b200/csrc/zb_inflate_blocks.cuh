// zb_inflate_blocks.cuh — the deflate blocks of ONE member, found without decoding it, and the pieces that put a
// member decoded chunk by chunk back together.
//
// A member without flush points is one serial chain to inflate.c / inffast.c, and was one to this library too (a
// lone member ran through one team of warps at 80-155 MB/s: slower than one CPU core).  What breaks the chain
// (SURVEY.md §8 f4 "sub-member parallel inflate: speculative block starts"):
//
//   1. CANDIDATES  A dynamic block header (inflate.c:898-1022) validates itself: BTYPE = 2, HLIT <= 29, HDIST <= 29,
//      a COMPLETE code-length code (inftrees.c:126-133 rejects anything else for CODES), then 258..316 code lengths that
//      must decode without a bad repeat, contain an end-of-block code and form a complete (or single-code) literal/length
//      set and a complete / single-code / empty distance set.  Every BIT position of the member is tested
//      (blk_scan_kernel): the first 17 + 3*HCLEN bits reject all but 1 in ~1100 positions with two 64-bit windows,
//      the survivors read their code lengths (blk_validate_kernel).  What passes is, but for a rare accident, a real block start — and the chain
//      check below does not depend on it being one.
//   2. COUNT       Every candidate (and the member's first block) starts a CHUNK: the blocks from there to the first later
//      block boundary that is itself a candidate.  All chunks are decoded at once with nothing written (INF_COUNT in
//      zb_inflate.cu): each reports where it ended, how many bytes and how many matches it holds.
//   3. CHAIN       From the member's first block the host follows end -> start: chunks on the chain are real, the others
//      (false candidates) are dropped.  Prefix sums place every chunk in the output and in the match list.
//   4. LIST        The chain's chunks are decoded again, each at its place (INF_LIST): literals are stored, matches are
//      only listed — what lies before a chunk is not there yet.
//   5. RESOLVE     Every output byte gets a source pointer (itself a literal, or the byte a match copies it from:
//      blk_src_build_kernel); pointer jumping (blk_jump_kernel: src[p] <- src[src[p]], all bytes at once, until nothing
//      moves: log2 of the longest copy chain passes) leaves every byte pointing at a literal, and one gather
//      (blk_gather_kernel) fills the matches in.  inffast.c:196-260 does the same copies one after the other.
//
// Stored and fixed blocks have no header to recognise: they stay inside the chunk of the dynamic block before them.
//
// The header test is __host__ __device__: tests/emul/inf_emul.cpp replays it over reference-made streams.
#pragma once
#include "zb_inflate.cuh"
#include "zb_inflate_round.cuh"

namespace zb {

// 64 stream bits from bit b of an LSB-first bit string held in aligned 32-bit words (words at or past nwords read as 0).
ZB_HD uint64_t blk_bits64(const uint32_t *__restrict__ w, uint64_t nwords, uint64_t b) {
    const uint64_t i = b >> 5;
    const uint32_t s = (uint32_t)b & 31u;
    const uint32_t w0 = i < nwords ? w[i] : 0u, w1 = i + 1 < nwords ? w[i + 1] : 0u, w2 = i + 2 < nwords ? w[i + 2] : 0u;
    return (uint64_t)funnel_r(w0, w1, s) | ((uint64_t)funnel_r(w1, w2, s) << 32);
}

// The fixed fields of a dynamic block header in the low 17 bits of x: BFINAL (any), BTYPE = 2, HLIT, HDIST within
// range (inflate.c:904-909 "too many length or distance symbols").
ZB_HD bool blk_quick(uint32_t x) {
    return ((x >> 1) & 3u) == 2u && ((x >> 3) & 31u) <= 29u && ((x >> 8) & 31u) <= 29u;
}
// ncode 3-bit code-length-code lengths in y: do they form a complete prefix code?  (Kraft sum in units of 2^-7.)
ZB_HD bool blk_cl_complete(uint64_t y, uint32_t ncode) {
    if (ncode < 19) y &= ((uint64_t)1 << (3 * ncode)) - 1;
    uint32_t k = 0;
#pragma unroll
    for (int i = 0; i < 19; ++i) { const uint32_t l = (uint32_t)(y >> (3 * i)) & 7u; k += (128u >> l) & 127u; }
    return k == 128u;
}

// A stored block (inflate.c:863-897) whose LEN field sits at bit P8 (a byte boundary; LEN and NLEN complement each other):
// its 3 header bits — BFINAL, BTYPE = 0 — lie at bit e = P8 - 3 - k for some k in 0..7, with zero padding between them and
// P8 (deflate pads with zeros, trees.c:166-193 bi_windup).  before16: the 16 stream bits in front of P8, bit P8 - 1 on top.
// Returns how many k = 0, 1, ... qualify (they form a run: one more zero bit is needed for each).
ZB_HD uint32_t blk_stored_starts(uint32_t before16, uint64_t P8, uint64_t bit_lo) {
    uint32_t ns = 0;
    for (uint32_t k = 0; k < 8; ++k) {
        if ((before16 >> (14 - k)) != 0u) break;             // bits e + 1 .. P8 - 1 (k + 2 of them) must be zero
        if (P8 - 3 - k < bit_lo) break;
        ++ns;
    }
    return ns;
}

// The rest of the test: read the HLIT + HDIST + 258 code lengths with the code-length code (inflate.c:930-995) and
// apply the acceptance rules of inflate.c:997-1019 / inftrees.c:126-133.  tab: 128 bytes of scratch (the 7-bit decode
// table of the code-length code: symbol | length << 5).  The bit position just behind the header is returned in *end.
ZB_HD bool blk_header_valid(const uint32_t *__restrict__ w, uint64_t nwords, uint64_t total_bits, uint64_t b,
                            uint8_t *tab, const uint8_t *cl_order, uint64_t *end) {
    const uint32_t x = (uint32_t)blk_bits64(w, nwords, b);
    const uint32_t nlen = ((x >> 3) & 31u) + 257u, ndist = ((x >> 8) & 31u) + 1u, ncode = ((x >> 13) & 15u) + 4u;
    const uint64_t y = blk_bits64(w, nwords, b + 17);
    uint64_t cls = 0;                                         // code-length-code lengths by symbol, 3 bits each
    for (uint32_t i = 0; i < ncode; ++i) cls |= ((y >> (3 * i)) & 7u) << (3 * cl_order[i]);
    uint32_t next[8], cnt[8];
    for (int l = 0; l < 8; ++l) cnt[l] = 0;
    for (int sym = 0; sym < 19; ++sym) cnt[(cls >> (3 * sym)) & 7u]++;
    cnt[0] = 0;
    { uint32_t code = 0; for (int l = 1; l < 8; ++l) { code = (code + cnt[l - 1]) << 1; next[l] = code; } }
    for (int sym = 0; sym < 19; ++sym) {
        const uint32_t l = (uint32_t)(cls >> (3 * sym)) & 7u;
        if (!l) continue;
        const uint32_t code = next[l]++;
        uint32_t rc = 0;
        for (uint32_t k = 0; k < l; ++k) rc |= ((code >> k) & 1u) << (l - 1 - k);
        for (uint32_t i = rc; i < 128u; i += 1u << l) tab[i] = (uint8_t)(sym | (l << 5));
    }
    uint64_t p = b + 17 + 3 * (uint64_t)ncode;
    const uint32_t total = nlen + ndist;
    uint32_t have = 0, prev = 0, kl = 0, kd = 0, nzl = 0, nzd = 0, eob = 0;   // Kraft sums in units of 2^-15
    // a 64-bit window over the stream, topped up a word at a time (a code and its extra bits take at most 14)
    uint64_t hold = blk_bits64(w, nwords, p);
    uint32_t nb = 64;                                                         // valid bits in hold; the next unread bit is p + nb
    while (have < total) {
        if (p >= total_bits) return false;
        if (nb < 32) {
            const uint64_t q = p + nb, i = q >> 5;                            // 32 more bits from bit q on (two words unless q is aligned)
            const uint32_t sh = (uint32_t)q & 31u;
            const uint32_t w0 = i < nwords ? w[i] : 0u, w1 = sh && i + 1 < nwords ? w[i + 1] : 0u;
            hold |= (uint64_t)funnel_r(w0, w1, sh) << nb;
            nb += 32;
        }
        const uint32_t v = (uint32_t)hold;
        const uint32_t e = tab[v & 127u], sym = e & 31u, l = e >> 5;          // (the code is complete: every pattern decodes)
        uint32_t rep = 1, val = sym, used = l;
        if (sym >= 16) {
            const uint32_t xb = v >> l;
            if (sym == 16) { if (have == 0) return false; val = prev; rep = 3 + (xb & 3u); used += 2; }
            else if (sym == 17) { val = 0; rep = 3 + (xb & 7u); used += 3; }
            else { val = 0; rep = 11 + (xb & 127u); used += 7; }
            if (have + rep > total) return false;                             // inflate.c:979-983 "invalid bit length repeat"
        }
        p += used; hold >>= used; nb -= used;
        prev = val;
        if (val) {
            const uint32_t unit = 32768u >> val;
            for (uint32_t k = 0; k < rep; ++k, ++have) {
                if (have < nlen) { kl += unit; ++nzl; if (have == 256) eob = 1; }
                else { kd += unit; ++nzd; }
            }
            if (kl > 32768u || kd > 32768u) return false;                     // over-subscribed
        } else have += rep;
    }
    if (p > total_bits) return false;
    if (!eob) return false;                                                   // inflate.c:997-1001 "missing end-of-block"
    if (kl != 32768u && !(nzl == 1 && kl == 16384u)) return false;            // incomplete only as one 1-bit code
    if (kd != 32768u && nzd != 0 && !(nzd == 1 && kd == 16384u)) return false;
    *end = p;
    return true;
}

// Positions are output positions of the member counted like the decoder counts them (from the first byte of the
// history / dictionary in front of it).  kSrcLiteral: the byte is there already.
constexpr uint32_t kSrcLiteral = 0xffffffffu;

#ifdef __CUDACC__
// Stage 1: every bit position in [bit_lo, bit_hi) of the member (words: the aligned words that hold it, bit 0 = bit 0 of
// words[0]) whose fixed header fields and code-length code pass, appended to list[] in no particular order.  The fixed
// fields of all 32 positions of a word are tested at once on the 64-bit window (BTYPE = 2: bit o+1 clear, bit o+2 set;
// HLIT / HDIST >= 30: their four high bits all set); the Kraft sum of the code-length code comes out of a table of four
// 3-bit lengths at a time.  One position in ~1100 survives.
// STORED blocks (inflate.c:863-897) are recognised here too, straight into the candidate list cand[]: a byte position
// whose LEN / NLEN words complement each other, behind 3 header bits (BTYPE = 0) and zero padding — the block may start
// at any of up to eight bit positions before it, each is a candidate (the chain picks the one the previous block ends
// on).  Without them a run of stored blocks (incompressible data) is one chunk, copied by one warp at 1 GB/s.
__global__ void __launch_bounds__(256)
blk_scan_kernel(const uint32_t *__restrict__ words, uint64_t nwords, uint64_t bit_lo, uint64_t bit_hi,
                uint64_t *__restrict__ list, uint32_t cap, uint32_t *__restrict__ count,
                uint64_t *__restrict__ cand, uint32_t cand_cap, uint32_t *__restrict__ cand_count) {
    __shared__ uint16_t kraft4[4096];                         // Kraft sum (units of 2^-7) of four 3-bit code lengths
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) {
        uint32_t k = 0;
        for (int f = 0; f < 4; ++f) { const uint32_t l = (i >> (3 * f)) & 7u; k += (128u >> l) & 127u; }
        kraft4[i] = (uint16_t)k;
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t wi = (bit_lo >> 5) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wi * 32 < bit_hi; wi += stride) {
        const uint32_t w0 = words[wi], w1 = wi + 1 < nwords ? words[wi + 1] : 0u;
        const uint64_t W = (uint64_t)w0 | ((uint64_t)w1 << 32);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {                    // LEN / NLEN at byte 4 wi + j ?
            const uint32_t v = (uint32_t)(W >> (8 * j));
            if (((v ^ (v >> 16)) & 0xffffu) != 0xffffu) continue;
            const uint64_t P8 = wi * 32 + 8 * j;             // bit position of the LEN field
            if (P8 < bit_lo + 3 || P8 + 32 > bit_hi || wi == 0) continue;
            const uint32_t before = __funnelshift_r(words[wi - 1], w0, 8 * j) >> 16;   // the 16 bits in front of it, bit P8 - 1 on top
            const uint32_t ns = blk_stored_starts(before, P8, bit_lo);
            for (uint32_t k = 0; k < ns; ++k) {
                const uint32_t at = atomicAdd(cand_count, 1u);
                if (at < cand_cap) cand[at] = P8 - 3 - k;
            }
        }
        uint32_t mask = (uint32_t)(~(W >> 1) & (W >> 2) & ~((W >> 4) & (W >> 5) & (W >> 6) & (W >> 7)) & ~((W >> 9) & (W >> 10) & (W >> 11) & (W >> 12)));
        if (!mask) continue;
        const uint32_t w2 = wi + 2 < nwords ? words[wi + 2] : 0u, w3 = wi + 3 < nwords ? words[wi + 3] : 0u;
        while (mask) {
            const uint32_t o = (uint32_t)__ffs(mask) - 1u;
            mask &= mask - 1;
            const uint64_t b = wi * 32 + o;
            if (b < bit_lo || b + 20 > bit_hi) continue;
            const uint32_t ncode = ((__funnelshift_r(w0, w1, o) >> 13) & 15u) + 4u;
            const uint32_t sft = o + 17;                      // the code-length-code lengths: up to 57 bits from bit o + 17
            const uint32_t a0 = sft < 32 ? w0 : w1, a1 = sft < 32 ? w1 : w2, a2 = sft < 32 ? w2 : w3;
            uint64_t y = (uint64_t)__funnelshift_r(a0, a1, sft) | ((uint64_t)__funnelshift_r(a1, a2, sft) << 32);
            if (ncode < 19) y &= ((uint64_t)1 << (3 * ncode)) - 1;
            const uint32_t ylo = (uint32_t)y, yhi = (uint32_t)(y >> 32);
            const uint32_t k = kraft4[ylo & 4095u] + kraft4[(ylo >> 12) & 4095u] + kraft4[(uint32_t)(y >> 24) & 4095u] +
                               kraft4[(yhi >> 4) & 4095u] + kraft4[(yhi >> 16) & 4095u];
            if (k != 128u) continue;
            const uint32_t at = atomicAdd(count, 1u);
            if (at < cap) list[at] = b;
        }
    }
}

// Stage 2: the survivors read their code lengths (blk_header_valid), one thread each; what passes is a candidate.
__global__ void __launch_bounds__(128)
blk_validate_kernel(const uint32_t *__restrict__ words, uint64_t nwords, uint64_t bit_hi, const uint64_t *__restrict__ in_list,
                    const uint32_t *__restrict__ in_count, uint32_t in_cap, uint64_t *__restrict__ list, uint32_t cap,
                    uint32_t *__restrict__ count, const FormatTables *__restrict__ fmt) {
    __shared__ uint8_t tabs[128 * 128];
    __shared__ uint8_t order[20];
    if (threadIdx.x < 20) order[threadIdx.x] = fmt->cl_order[threadIdx.x];
    __syncthreads();
    const uint32_t n = *in_count < in_cap ? *in_count : in_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t b = in_list[i];
        uint64_t end;
        if (!blk_header_valid(words, nwords, bit_hi, b, tabs + threadIdx.x * 128, order, &end)) continue;
        const uint32_t k = atomicAdd(count, 1u);
        if (k < cap) list[k] = b;
    }
}

// ---- resolve ------------------------------------------------------------------------------------------------
// src[] covers the output positions [lo, lo + n) of one group of chunks (n <= 256 MiB).  An entry is kSrcLiteral, or the
// position its byte is copied from, counted from `base` = lo - 32768 (0 for the member's first 32 KiB: no source lies
// further back than a window) — 29 bits — with bit 31 set once that position is known to hold a final byte (a literal,
// or a byte before lo).
constexpr uint32_t kSrcFinal = 0x80000000u;

// One warp per 32 matches: byte i of a match comes from dst - dist + i (dist >= len), or from the first period of
// its own output (inffast.c:249-260 byte-serial semantics: dst - dist + i mod dist).
__global__ void __launch_bounds__(256)
blk_src_build_kernel(const QueuedMatch *__restrict__ ml, uint64_t n_matches, uint32_t *__restrict__ src, uint32_t lo, uint32_t base) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t at = ((((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5) * 32; at < n_matches; at += warps * 32) {
        QueuedMatch e;
        e.dst = 0; e.packed = 0;
        if (at + lane < n_matches) e = ml[at + lane];
        const uint32_t cnt = n_matches - at < 32 ? (uint32_t)(n_matches - at) : 32u;
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t dst = __shfl_sync(full, e.dst, j), pk = __shfl_sync(full, e.packed, j);
            const uint32_t len = qm_len(pk), dist = qm_dist(pk);
            const uint32_t from = dst - dist;
            for (uint32_t i = lane; i < len; i += 32) {
                const uint32_t sp = from + (i < dist ? i : i % dist);
                src[dst - lo + i] = (sp - base) | (sp < lo ? kSrcFinal : 0u);
            }
        }
    }
}

// One pass of pointer jumping over src[0, n): an entry that is not final yet follows its source's entry — up to
// kJumpHops times in a row — and is final as soon as it meets a literal, a final entry or a byte before lo.  Sources
// lie behind their bytes and CTAs run in ascending order, so most entries meet a final one in the first pass; entries
// that are final cost later passes one streamed read.  Updates are made in place: any value read is a valid ancestor.
#ifndef ZB_JUMP_HOPS
#define ZB_JUMP_HOPS 4
#endif
constexpr int kJumpHops = ZB_JUMP_HOPS;
__device__ __forceinline__ uint32_t blk_follow(const uint32_t *__restrict__ src, uint32_t v, uint32_t lo, uint32_t base, bool &open) {
    if (v & kSrcFinal) return v;                             // (kSrcLiteral included)
    uint32_t cur = v;
#pragma unroll
    for (int h = 0; h < kJumpHops; ++h) {
        if (cur + base < lo) return cur | kSrcFinal;
        const uint32_t t = src[cur + base - lo];
        if (t == kSrcLiteral) return cur | kSrcFinal;
        if (t & kSrcFinal) return t;
        cur = t;
    }
    open = true;
    return cur;
}
__global__ void __launch_bounds__(256)
blk_jump_kernel(uint32_t *__restrict__ src, uint64_t n, uint32_t lo, uint32_t base, uint32_t *__restrict__ changed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    bool any = false;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            uint4 v = *reinterpret_cast<const uint4 *>(src + i);
            if ((v.x & v.y & v.z & v.w) & kSrcFinal) continue;
            bool open = false;
            v.x = blk_follow(src, v.x, lo, base, open); v.y = blk_follow(src, v.y, lo, base, open);
            v.z = blk_follow(src, v.z, lo, base, open); v.w = blk_follow(src, v.w, lo, base, open);
            *reinterpret_cast<uint4 *>(src + i) = v;
            any |= open;
        } else {
            for (uint64_t k = i; k < n; ++k) {
                const uint32_t v = src[k];
                if (v & kSrcFinal) continue;
                bool open = false;
                src[k] = blk_follow(src, v, lo, base, open);
                any |= open;
            }
        }
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) *changed = 1u;
}

// out[p] = out[source of p] for every copied byte of [lo, lo + n): all entries are final by now.
__global__ void __launch_bounds__(256)
blk_gather_kernel(uint8_t *out, const uint32_t *__restrict__ src, uint64_t n, uint32_t lo, uint32_t base) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + i);
            uint8_t *o = out + lo + i;
            uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
            if (v.x != kSrcLiteral) b0 = out[(v.x & ~kSrcFinal) + base];
            if (v.y != kSrcLiteral) b1 = out[(v.y & ~kSrcFinal) + base];
            if (v.z != kSrcLiteral) b2 = out[(v.z & ~kSrcFinal) + base];
            if (v.w != kSrcLiteral) b3 = out[(v.w & ~kSrcFinal) + base];
            if (v.x != kSrcLiteral) o[0] = (uint8_t)b0;
            if (v.y != kSrcLiteral) o[1] = (uint8_t)b1;
            if (v.z != kSrcLiteral) o[2] = (uint8_t)b2;
            if (v.w != kSrcLiteral) o[3] = (uint8_t)b3;
        } else {
            for (uint64_t k = i; k < n; ++k) { const uint32_t v = src[k]; if (v != kSrcLiteral) out[lo + k] = out[(v & ~kSrcFinal) + base]; }
        }
    }
}
#endif  // __CUDACC__

}  // namespace zb
