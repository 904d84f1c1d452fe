// zb_inflate_round.cuh — warp-parallel decoding of an open Huffman block.
//
// The symbol chain of a DEFLATE block is serial: the start of unit k+1 (a unit is a
// literal, an end-of-block, or a length/distance pair with its extra bits) is known
// only once unit k is decoded (inffast.c:100-287).  One lane decoding alone is what
// bounds a member at ~12 MB/s.  A ROUND breaks the chain with the self-synchronising
// property of prefix codes: the next 32*S bits of the block are cut into 32
// subsequences of S bits; lane i starts decoding at bit i*S — usually in the middle
// of a unit — and after a few (wrong) units its unit starts coincide with the true
// ones and stay so (measured on zlib streams: 96.5 % of the subsequences synchronise
// within S = 512 bits, 99 % within 1024; median ~100 bits).
//
//   P1  speculate  every lane decodes its subsequence from the guessed start and
//                  records, per 64-bit double word, the first unit start it visited
//                  there and the output / match counts accumulated before it.
//   P2  fix up     lane 0 started at a true unit start.  Lane i re-decodes from the
//                  true start handed over by lane i-1 (its end position) only until
//                  its first unit start in a double word equals the recorded one —
//                  from there on the two paths are identical, so the speculative
//                  counts are spliced in.  A lane that never meets its speculative
//                  path simply keeps its own decode.  This repeats until no lane's
//                  start moves (usually 1-2 short passes).
//   scan           exclusive sums of the per-lane output bytes and match counts.
//   P3  emit       every lane decodes its (now true) subsequence once more, storing
//                  literals at their final positions and parking matches in the
//                  round's queue, in stream order.
//   P4  copy       the queue is executed in waves of 32 matches (zb_inflate.cu).
//
// Anything unusual in a round — an invalid code on the true path, a distance
// reaching before the start of the output, output overflow, truncated input — makes
// the warp abandon the round; lane 0 then decodes from the round's start on the
// careful serial path (zb_inflate.cuh), which reports the reference's exact status.
//
// Measured and dropped (round 2): the three loops in a uniform HALF-step form — one table look-up per trip, a match
// taking two trips, so that lanes on literals never wait for lanes in the distance half (ncu had shown 13 of 32 lanes
// active in the unit decoder).  Same bytes, 6 % slower (8.79 -> 9.30 ms per GiB in 4096 members): the kernel is bound
// by each lane's dependent chain (window fetch -> table look-up -> bit arithmetic), not by issue slots, and the half
// steps lengthen that chain by a loop trip per match.
//
// The per-lane phases are __host__ __device__: tests/emul/inf_emul.cpp replays them
// with plain loops over the 32 lanes.
#pragma once
#include "zb_inflate.cuh"

namespace zb {

#ifndef ZB_ROUND_LG_MAX
#define ZB_ROUND_LG_MAX 5
#endif
constexpr int kRoundLgMin = 2, kRoundLgMax = ZB_ROUND_LG_MAX;              // words per lane = 1 << lg  (S = 128 .. 1024 bits)
constexpr int kRoundWordsMax = 1 << kRoundLgMax;
constexpr int kRowExtra = 3;                                 // words of the following subsequence repeated at the end of a row
// Matches a round may park.  The worst case is 32 * S / 2 (a match costs at least 2 bits); text
// parks ~1500.  A round that would park more is left to the serial path like any other oddity.
constexpr uint32_t kRoundQueueCap = 6144;

// Per-warp working memory of a round (shared memory on the device).
// stage: lane i owns row i = the W words of its subsequence followed by the next 3
// words of the stream (a unit may start on the row's last bit and is up to 48 bits
// long).  The row stride W + 3 is odd, so the 32 lanes reading the same column of
// their rows hit 32 different banks.
// fs / cm: what the speculative pass leaves behind per 64-bit double word of a
// subsequence: the bit offset of the first unit starting there (kNoUnit: none) and the
// output / match counts accumulated before that unit.
constexpr int kRoundDwordsMax = kRoundWordsMax / 2;
constexpr uint8_t kNoUnit = 0xff;
// NL = lanes that share a round: 32 (one warp per member) or 128 (a team of four warps per member).
template <int NL>
struct RoundSharedT {
    static constexpr int kLanes = NL;
    uint32_t stage[NL * (kRoundWordsMax + kRowExtra)];
    uint32_t cm[kRoundDwordsMax * NL];           // [dword][lane] output bytes << 12 | matches
    uint8_t  fs[kRoundDwordsMax * NL];           // [dword][lane]
};
using RoundShared = RoundSharedT<32>;
ZB_HD uint32_t stage_row_stride(int lg) { return (1u << lg) + kRowExtra; }

ZB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}

// LSB-first view of one lane's row: at(p) = the 32 stream bits that start at bit p of
// the lane's subsequence.  No running state: a unit re-derives its window from the bit
// position, which costs two shared-memory loads and one funnel shift.
struct LaneWin {
    const uint32_t *row;
    ZB_HD uint32_t at(uint32_t p) const { const uint32_t r = p >> 5; return funnel_r(row[r], row[r + 1], p); }
};

enum : uint32_t { U_LIT = 0, U_MATCH = 1, U_EOB = 2, U_BAD = 3 };
struct Unit { uint32_t kind, used, val, dist; };             // val: literal byte or match length

// Decode the unit at bit p (the work of one trip of inffast.c:100-287).  A literal /
// length code with its extra bits is at most 20 bits, a distance code with its extra
// bits at most 28: each half fits one 32-bit window.
template <bool kWantDist>
ZB_HD Unit decode_unit(const LaneWin &lw, uint32_t p, const uint32_t *__restrict__ L, const uint32_t *__restrict__ D) {
    Unit u;
    u.dist = 0; u.val = 0;
    uint32_t w = lw.at(p), used = 0;
    uint32_t e = L[w & ((1u << kLitRoot) - 1u)];
    if (ZB_E_OP(e) == OP_SUB) {
        e = L[ZB_E_VAL(e) + ((w >> kLitRoot) & ((1u << ZB_E_EXTRA(e)) - 1u))];
        w >>= kLitRoot; used = kLitRoot;
    }
    { const uint32_t k = ZB_E_BITS(e); w >>= k; used += k; }
    const uint32_t op = ZB_E_OP(e);
    if (op == OP_LIT) { u.kind = U_LIT; u.val = e >> 16; u.used = used; return u; }
    if (op != OP_BASE) { u.kind = op == OP_EOB ? U_EOB : U_BAD; u.used = used; return u; }
    {
        const uint32_t x = ZB_E_EXTRA(e);
        u.val = ZB_E_VAL(e) + (w & ((1u << x) - 1u));
        used += x;
    }
    w = lw.at(p + used);
    uint32_t d = D[w & ((1u << kDistRoot) - 1u)];
    if (ZB_E_OP(d) == OP_SUB) {
        d = D[ZB_E_VAL(d) + ((w >> kDistRoot) & ((1u << ZB_E_EXTRA(d)) - 1u))];
        w >>= kDistRoot; used += kDistRoot;
    }
    { const uint32_t k = ZB_E_BITS(d); w >>= k; used += k; }
    if (ZB_E_OP(d) != OP_BASE) { u.kind = U_BAD; u.used = used; return u; }
    {
        const uint32_t dx = ZB_E_EXTRA(d);
        if (kWantDist) u.dist = ZB_E_VAL(d) + (w & ((1u << dx) - 1u));
        used += dx;
    }
    u.kind = U_MATCH; u.used = used;
    return u;
}

// What a lane knows about its subsequence.  Positions are bits relative to the START
// OF THE LANE'S OWN subsequence (lane i's bit 0 = bit i*S of the round); a path ends at
// the first unit start >= S, i.e. at bit (end - S) of the next lane.
struct RoundLane {
    uint32_t spec_end, spec_stop, spec_out, spec_m;          // the speculative path (never changes after P1)
    uint32_t start, end, stop, out, m;                       // the current path (true once P2 has settled)
};
enum : uint32_t { STOP_NONE = 0, STOP_EOB = 1, STOP_BAD = 2 };

// P1.  `start` is 0, except for lane 0 which starts at the block's true position.
template <class RS>
ZB_HD void round_speculate(RoundLane &r, uint32_t lane, int lg, uint32_t start, const uint32_t *stage, RS &rs,
                           const uint32_t *__restrict__ L, const uint32_t *__restrict__ D) {
    const uint32_t DW = 1u << (lg - 1), S = 32u << lg;
    for (uint32_t w = 0; w < DW; ++w) rs.fs[w * RS::kLanes + lane] = kNoUnit;
    LaneWin lw; lw.row = stage + lane * stage_row_stride(lg);
    uint32_t p = start;
    uint32_t curdw = 0xffffffffu, out = 0, m = 0, stop = STOP_NONE;
    while (p < S) {
        const uint32_t dw = p >> 6;
        if (dw != curdw) {                                   // the first unit starting in this double word
            curdw = dw;
            rs.fs[dw * RS::kLanes + lane] = (uint8_t)(p & 63u);
            rs.cm[dw * RS::kLanes + lane] = (out << 12) | m;
        }
        const Unit u = decode_unit<false>(lw, p, L, D);
        p += u.used;
        if (u.kind >= U_EOB) { stop = u.kind == U_EOB ? STOP_EOB : STOP_BAD; break; }
        out += u.kind == U_LIT ? 1u : u.val;
        m += u.kind;
    }
    r.spec_end = r.end = p; r.spec_stop = r.stop = stop; r.spec_out = r.out = out; r.spec_m = r.m = m;
    r.start = start;
}

// P2.  Re-decode from the true start `t` until the path joins the speculative one: two
// paths that have a unit start on the same bit are identical from there on, and once
// joined they share the first unit start of every later double word — which is where
// the join is looked for and the speculative counts are spliced in.
template <class RS>
ZB_HD void round_fix(RoundLane &r, uint32_t lane, int lg, uint32_t t, const uint32_t *stage, const RS &rs,
                     const uint32_t *__restrict__ L, const uint32_t *__restrict__ D) {
    const uint32_t S = 32u << lg;
    LaneWin lw; lw.row = stage + lane * stage_row_stride(lg);
    uint32_t p = t;
    uint32_t out = 0, m = 0, stop = STOP_NONE, curdw = 0xffffffffu;
    r.start = t;
    while (p < S) {
        const uint32_t dw = p >> 6;
        if (dw != curdw) {
            curdw = dw;
            if (rs.fs[dw * RS::kLanes + lane] == (uint8_t)(p & 63u)) {
                const uint32_t c = rs.cm[dw * RS::kLanes + lane];
                r.out = out + (r.spec_out - (c >> 12));
                r.m = m + (r.spec_m - (c & 0xfffu));
                r.end = r.spec_end; r.stop = r.spec_stop;
                return;
            }
        }
        const Unit u = decode_unit<false>(lw, p, L, D);
        p += u.used;
        if (u.kind >= U_EOB) { stop = u.kind == U_EOB ? STOP_EOB : STOP_BAD; break; }
        out += u.kind == U_LIT ? 1u : u.val;
        m += u.kind;
    }
    r.end = p; r.stop = stop; r.out = out; r.m = m;
}

// P3.  Decode the settled subsequence, store literals, park matches (stream order).
// `o` is the lane's first output position inside the member, `qi` its first queue slot.
// Returns 0, or a nonzero reason to abandon the round.
ZB_HD int round_emit(const RoundLane &r, uint32_t lane, int lg, const uint32_t *stage,
                     const uint32_t *__restrict__ L, const uint32_t *__restrict__ D,
                     uint8_t *__restrict__ out, uint32_t o, QueuedMatch *__restrict__ q, uint32_t qi) {
    const uint32_t S = 32u << lg;
    LaneWin lw; lw.row = stage + lane * stage_row_stride(lg);
    uint32_t p = r.start;
    while (p < S) {
        const Unit u = decode_unit<true>(lw, p, L, D);
        p += u.used;
        if (u.kind == U_LIT) { out[o++] = (uint8_t)u.val; continue; }
        if (u.kind != U_MATCH) return u.kind == U_EOB ? 0 : 1;
        if (u.dist > o) return 2;                            // inffast.c:152-161 "invalid distance too far back"
        q[qi].dst = o; q[qi].packed = qm_pack(u.val, u.dist, false);
        ++qi;
        o += u.val;
    }
    return 0;
}

// Subsequence size for a round: the largest S with 32*S bits still inside the member, or -1
// when too little input is left for a round to pay off (the serial path finishes the member).
ZB_HD int round_pick_lg(uint64_t remaining_bits, uint32_t lanes = 32) {
    for (int lg = kRoundLgMax; lg >= kRoundLgMin; --lg)
        if (remaining_bits >= (uint64_t)(32u << lg) * lanes) return lg;
    return -1;
}

}  // namespace zb
