// zb_deflate.cuh — per-thread cores of the deflate pipeline.
//
// B200 re-design of deflate.c (fill_window / longest_match / deflate_fast /
// deflate_slow) and trees.c (build_tree / gen_bitlen / gen_codes / scan_tree /
// send_tree / send_all_trees / compress_block / _tr_flush_block) for inputs cut
// into independent chunks (one Z_FULL_FLUSH-bounded block run each,
// deflate.c:1211-1226).  The reference's single sequential loop is split into
// data-parallel phases (kernels in zb_deflate.cu):
//
//   1 chains   prev_dist[p] = distance to the most recent earlier position with
//              the same 15-bit hash of 3 bytes (deflate.c:141,160-163).  Because
//              deflate_slow inserts EVERY position (deflate.c:1947,1994-2000),
//              the chains do not depend on the parse and can be built up front.
//   2 matches  one thread per position walks its chain exactly like
//              longest_match (deflate.c:1356-1497: max_chain candidates, stop at
//              nice_match, first-longest wins, MAX_DIST limit) and records the
//              result for the full chain budget and for the quartered budget
//              (deflate.c:1390-1392 good_match).
//   3 parse    picks matches with the reference's lazy (deflate.c:1923-2043) or
//              greedy (deflate.c:1824-1915) rule from the precomputed tables, emits
//              symbols and cuts blocks every 16383 symbols (deflate.c:455,512); the
//              32 lanes of a warp parse 32 segments of the chunk concurrently and
//              are stitched where their states coincide (seg_* below).
//   4 trees    per block: histogram, the reference's exact heap Huffman
//              construction + length limiting + header RLE, block type choice
//              (trees.c:997-1089).
//   5 pack     per block: symbol -> code bits, prefix sum of bit lengths,
//              bit-packing into the output stream (trees.c:900-951).
//
// For levels 4..9 the result is byte-identical to the reference on the same
// chunking (tests pin this).  Levels 1..3 use the same machinery with a greedy
// parse over full-insertion chains: the reference's deflate_fast skips hash
// insertions inside matches (deflate.c:1873-1897), which makes its chains depend
// on the parse; ours see strictly more candidates, so the ratio is equal or
// better, but the bytes differ.
#pragma once
#include "zb_format.h"

namespace zb {

constexpr int kMinMatch = 3, kMaxMatch = 258;
constexpr int kWSize = 32768;                                 // the largest window (windowBits 15); staging buffers are sized for it
constexpr int kMinLookahead = kMaxMatch + kMinMatch + 1;      // deflate.h:293
constexpr int kTooFar = 4096;                                 // deflate.c:88-90
constexpr int kHashBitsMax = 16;                              // memLevel 9 (deflate.c:444: hash_bits = memLevel + 7)

enum { MODE_FAST = 0, MODE_SLOW = 1, MODE_HUFF = 2, MODE_RLE = 3 };
enum { STRAT_DEFAULT = 0, STRAT_FILTERED = 1, STRAT_HUFFMAN = 2, STRAT_RLE = 3, STRAT_FIXED = 4 };

struct DeflateParams {
    int level, strategy, mode;
    int good, lazy, nice, chain;      // deflate.c:112-124 configuration_table
    int need_quarter;                 // the quartered-chain result can be asked for by the parse
    uint32_t sym_limit;               // symbols that fill a block: lit_bufsize - 1 = (1 << (memLevel + 6)) - 1 (deflate.c:455,512);
                                      // level 0 cuts its stored blocks at MAX_STORED bytes instead (deflate.c:1615,1664)
    uint32_t w_size, max_dist;        // 1 << windowBits (deflate.c:440-443), MAX_DIST = w_size - MIN_LOOKAHEAD (deflate.h:298)
    uint32_t hash_bits, hash_shift, hash_mask;   // memLevel + 7, (hash_bits + MIN_MATCH - 1) / MIN_MATCH (deflate.c:444-447)
};

ZB_HD DeflateParams deflate_params(int level, int strategy, int window_bits = 15, int mem_level = 8) {
    //                      good lazy nice chain
    const int cfg[10][4] = {{0, 0, 0, 0}, {4, 4, 8, 4}, {4, 5, 16, 8}, {4, 6, 32, 32}, {4, 4, 16, 16},
                            {8, 16, 32, 32}, {8, 16, 128, 128}, {8, 32, 128, 256}, {32, 128, 258, 1024},
                            {32, 258, 258, 4096}};
    DeflateParams p;
    p.level = level; p.strategy = strategy;
    p.good = cfg[level][0]; p.lazy = cfg[level][1]; p.nice = cfg[level][2]; p.chain = cfg[level][3];
    p.mode = (strategy == STRAT_HUFFMAN || level == 0) ? MODE_HUFF : strategy == STRAT_RLE ? MODE_RLE : level >= 4 ? MODE_SLOW : MODE_FAST;
    // a search happens only while prev_length < lazy; it is quartered when prev_length >= good
    p.need_quarter = (p.mode == MODE_SLOW && p.good < p.lazy) ? 1 : 0;
    p.sym_limit = level == 0 ? 65535u : (1u << (mem_level + 6)) - 1u;
    p.w_size = 1u << window_bits; p.max_dist = p.w_size - (uint32_t)kMinLookahead;
    p.hash_bits = (uint32_t)mem_level + 7u; p.hash_shift = (p.hash_bits + kMinMatch - 1) / kMinMatch; p.hash_mask = (1u << p.hash_bits) - 1u;
    return p;
}

ZB_HD uint32_t hash3(const uint8_t *s, const DeflateParams &prm) {   // deflate.c:141 UPDATE_HASH over MIN_MATCH bytes
    return (((uint32_t)s[0] << (2 * prm.hash_shift)) ^ ((uint32_t)s[1] << prm.hash_shift) ^ s[2]) & prm.hash_mask;
}

// ---- phase 2: longest match at one position -------------------------------------
// Result encoding: 0 = no match of length >= 3, else (len << 16) | dist.
struct MatchPair { uint32_t full, quarter; };

// Four bytes at an arbitrary address, little-endian, from two aligned words.
ZB_HD uint32_t load4(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8);
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}
ZB_HD uint32_t ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}

// The walk over one position's hash chain: longest_match (deflate.c:1356-1497) for the
// full chain budget, with a snapshot of the best match when a quarter of the budget has
// been spent (what the search would return when entered with prev_length >= good_match,
// deflate.c:1390-1392).  Per candidate: the quick rejects of deflate.c:1449-1452, the
// common prefix length of a candidate that passed, the next link (deflate.c:1481-1482).
// The loop is written for the GPU's issue slots: the byte scan[best] lives in a register
// (it changes only when the best match does), the next link is requested before the
// candidate is looked at (the two reads are independent, so their latencies overlap), and
// the quarter snapshot is a second loop bound instead of per-step bookkeeping.
// `Mem` supplies the operands by chunk position: byte(pos), word(pos) (four bytes,
// little-endian, any alignment) and link(pos) — plain arrays here, a window staged in
// shared memory in the depth-sorted kernel (32-bit shared addresses instead of generic
// 64-bit pointer arithmetic: a fifth of that kernel's instructions).
struct PlainMem {
    static constexpr bool kNilIsFar = false;              // a chain's end is link 0
    const uint8_t *data; const uint16_t *prev;
    ZB_HD uint32_t byte(uint32_t pos) const { return data[pos]; }
    ZB_HD uint32_t word(uint32_t pos) const { return load4(data + pos); }
    ZB_HD uint32_t link(uint32_t pos) const { return prev[pos]; }
};

template <class Mem>
ZB_HD MatchPair match_walk(const Mem &mem, uint32_t n, uint32_t p, const DeflateParams &prm) {
    MatchPair r;
    r.full = r.quarter = 0;
    if (p + kMinMatch > n) return r;                      // lookahead < MIN_MATCH: no insertion, no search
    const uint32_t d = mem.link(p);
    if (d == 0 || d > prm.max_dist) return r;             // deflate.c:1857/1958: head must be within MAX_DIST
    const uint32_t look = n - p;
    const uint32_t maxlen = look < (uint32_t)kMaxMatch ? look : (uint32_t)kMaxMatch;
    const uint32_t nice = (uint32_t)prm.nice > look ? look : (uint32_t)prm.nice;      // deflate.c:1396
    // The first 12 bytes of the string live in registers: most candidates that pass the quick
    // rejects differ within them, so measuring costs candidate loads only.  (In a warp, the few
    // lanes that measure hold up all the others: the cheaper this path, the better.)
    const bool regs = look >= 16;                         // else: chunk tail, everything bytewise below
    const uint32_t sw0 = regs ? mem.word(p) : 0, sw4 = regs ? mem.word(p + 4) : 0, sw8 = regs ? mem.word(p + 8) : 0;
    uint32_t best = kMinMatch - 1, best_dist = 0;                                     // best < maxlen whenever a candidate is screened
    uint32_t sb = mem.byte(p + kMinMatch - 1), sb1 = mem.byte(p + kMinMatch - 2);     // scan[best], scan[best - 1] (deflate.c:1449-1450 scan_end, scan_end1)
    uint32_t q = p - d, examined = 0;
    const uint32_t budget = (uint32_t)prm.chain, qbudget = budget >> 2;
    uint32_t stop_at = qbudget ? qbudget : budget;        // the next loop bound: quarter snapshot first, then the full budget
    bool have_q = false;
    for (;;) {
        const uint32_t d2 = mem.link(q);
        if (mem.byte(q + best) == sb && mem.byte(q + best - 1) == sb1) {   // deflate.c:1449-1452: the cheapest rejects first
            uint32_t len;
            bool pass;
            if (regs) {
                uint32_t x = mem.word(q) ^ sw0;
                pass = (x & 0xffffu) == 0;
                if (x) len = ctz32(x) >> 3;
                else {
                    x = mem.word(q + 4) ^ sw4;
                    if (x) len = 4 + (ctz32(x) >> 3);
                    else {
                        x = mem.word(q + 8) ^ sw8;
                        if (x) len = 8 + (ctz32(x) >> 3);
                        else {
                            len = 12;
                            while (len + 8 <= maxlen) {   // four bytes per step while the aligned reads stay inside the chunk
                                const uint32_t y = mem.word(q + len) ^ mem.word(p + len);
                                if (y) { len += ctz32(y) >> 3; goto measured; }
                                len += 4;
                            }
                            while (len < maxlen && mem.byte(q + len) == mem.byte(p + len)) ++len;
                        measured:;
                        }
                    }
                }
            } else {
                pass = mem.byte(q) == mem.byte(p) && mem.byte(q + 1) == mem.byte(p + 1);
                len = 2;                                  // m[2]==scan[2] follows from the equal hash when bytes 0,1 agree
                if (pass) while (len < maxlen && mem.byte(q + len) == mem.byte(p + len)) ++len;
            }
            if (pass && len > best) {
                best = len; best_dist = p - q;
                if (len >= nice) break;
                sb = mem.byte(p + best); sb1 = mem.byte(p + best - 1);
            }
        }
        if (++examined == stop_at) {
            if (examined == budget) break;                // deflate.c:1482 --chain_length
            r.quarter = best >= (uint32_t)kMinMatch ? (best << 16) | best_dist : 0;
            have_q = true;
            stop_at = budget;
        }
        if (!Mem::kNilIsFar && d2 == 0) break;            // (a staged window stores NIL as 65535: the window test below ends the walk)
        q -= d2;
        if (p - q >= prm.max_dist) break;                 // deflate.c:1481: cur_match > limit
    }
    r.full = best >= (uint32_t)kMinMatch ? (best << 16) | best_dist : 0;
    if (!have_q) r.quarter = r.full;
    return r;
}

ZB_HD MatchPair match_at(const uint8_t *data, uint32_t n, const uint16_t *prev_dist, uint32_t p,
                         const DeflateParams &prm) {
    const PlainMem mem{data, prev_dist};
    return match_walk(mem, n, p, prm);
}

// ---- levels 1-2: the same walk without data-dependent control flow -----------------
// With a small chain budget and a small nice_match the walk of longest_match has a
// closed form: let len_k be the common prefix length of candidate k capped at NICE;
// the walk stops at the first k with len_k == NICE, and the winner is the first
// candidate that reaches the maximum of the len_k seen up to there (a candidate is
// taken only when strictly longer than the best so far, deflate.c:1466-1476; the
// quick rejects of :1449-1452 only skip candidates that could not be longer).  So
// every lane runs the same CH steps of "load NICE bytes, compare, select" — no lane
// waits for another's longer walk — and only the winner that reached NICE is then
// extended to its true length.  Needs p + kUniformTail <= n (all reads stay inside the
// chunk, lookahead >= MAX_MATCH so neither nice nor the match is clamped).
constexpr uint32_t kUniformTail = kMaxMatch + 16;

template <int NW>
ZB_HD void load_words(const uint8_t *p, uint32_t (&w)[NW]) {       // NW little-endian words from an arbitrary address
#if defined(__CUDA_ARCH__)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t raw[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; ++i) raw[i] = src[i];
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
#else
    for (int i = 0; i < NW; ++i) w[i] = load4(p + 4 * i);
#endif
}

// Operand access for match_uniform.
struct PlainWin {
    const uint8_t *data; const uint16_t *prev;
    template <int NW> ZB_HD void words(uint32_t pos, uint32_t (&w)[NW]) const { load_words<NW>(data + pos, w); }
    ZB_HD uint32_t dist(uint32_t pos) const { return prev[pos]; }
};
template <int CH, int NICE, class Win>
ZB_HD uint32_t match_uniform(const Win &win, uint32_t p, uint32_t max_dist) {
    constexpr int NW = NICE / 4;
    const uint32_t d = win.dist(p);
    if (d == 0 || d > max_dist) return 0;                      // deflate.c:1857: head must be within MAX_DIST
    uint32_t sw[NW];
    win.template words<NW>(p, sw);
    uint32_t q = p - d, best = kMinMatch - 1, best_q = 0;
    bool open = true;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        if (open) {
            uint32_t mw[NW];
            win.template words<NW>(q, mw);
            uint32_t len = NICE;
#pragma unroll
            for (int j = NW - 1; j >= 0; --j) { const uint32_t x = mw[j] ^ sw[j]; if (x) len = 4u * (uint32_t)j + (ctz32(x) >> 3); }
            if (len > best) { best = len; best_q = q; }
            if (len >= (uint32_t)NICE) open = false;
            else if (k + 1 < CH) {
                const uint32_t d2 = win.dist(q);
                if (d2 == 0) open = false;
                else { q -= d2; if (p - q >= max_dist) open = false; }
            }
        }
    }
    if (best < (uint32_t)kMinMatch) return 0;
    if (best >= (uint32_t)NICE) {                                 // the winner reached nice_match: its true length
        uint32_t len = NICE;
        while (len < (uint32_t)kMaxMatch) {                       // reads up to byte p + 259 (< p + kUniformTail)
            uint32_t a[1], c[1];
            win.template words<1>(best_q + len, a);
            win.template words<1>(p + len, c);
            const uint32_t x = a[0] ^ c[0];
            if (x) { len += ctz32(x) >> 3; break; }
            len += 4;
        }
        best = len < (uint32_t)kMaxMatch ? len : (uint32_t)kMaxMatch;
    }
    return (best << 16) | (p - best_q);
}

// Z_RLE (deflate.c:2051-2115): run of the previous byte, distance 1 only.
ZB_HD uint32_t rle_at(const uint8_t *data, uint32_t n, uint32_t p) {
    if (p == 0 || p + kMinMatch > n) return 0;
    const uint32_t look = n - p;
    const uint32_t maxlen = look < (uint32_t)kMaxMatch ? look : (uint32_t)kMaxMatch;
    const uint8_t prev = data[p - 1];
    uint32_t len = 0;
    while (len < maxlen && data[p + len] == prev) ++len;
    return len >= (uint32_t)kMinMatch ? (len << 16) | 1u : 0;
}

// ---- phase 3: parse ------------------------------------------------------------
struct BlockInfo {
    uint32_t sym_start, sym_count;    // symbols of this block inside the chunk's symbol array
    uint32_t byte_start, byte_len;    // input bytes the block covers (stored_len)
    uint32_t flags;                   // bit 0: last block of the stream; bit 1: stored form allowed
    uint32_t bit_start_lo, bit_start_hi;  // filled by the layout step: bit offset inside the output stream
    uint32_t pad;
};
constexpr uint32_t BLK_LAST = 1, BLK_STORED_OK = 2;

ZB_HD uint32_t max_blocks_for(uint32_t chunk_bytes, uint32_t sym_limit) { return chunk_bytes / sym_limit + 2; }

// The parse is a resumable state machine so that the kernel can run it tile by
// tile over operands staged in shared memory.  `Acc` supplies the operands:
//   mf(p), mq(p)  match table entries at position p
//   byte(p)       input byte at position p
//   put(sym)      append one symbol ((dist << 16) | (len-3), or a literal byte)
struct ParseState {
    uint32_t p;                       // next position to examine (strstart)
    uint32_t match_length, cur_dist;  // deflate_slow carry: match found at p-1
    uint32_t match_available;
    uint32_t base;                    // window origin: advances 32 KiB per slide (deflate.c:277-287)
    uint32_t slide_at;                // first loop-top position at which the next slide happens
    uint32_t nsyms, nblocks;          // totals so far for the chunk
    uint32_t blk_sym0, blk_byte0;     // start of the open block
    BlockInfo *blocks;
    uint32_t sym_limit;               // prm.sym_limit (parse_close_block has no prm at hand)
    uint32_t block_mode;              // PB_SERIAL: complete BlockInfo as the parse goes; PB_DEFERRED: a parse that covers
                                      // only a segment leaves (end offset, window base) for seg_finish; PB_NONE: counting pass
};
enum : uint32_t { PB_SERIAL = 0, PB_DEFERRED = 1, PB_NONE = 2 };

ZB_HD uint32_t parse_next_slide(uint32_t base, uint32_t n, const DeflateParams &prm);

ZB_HD uint32_t parse_base_at(uint32_t P, uint32_t n, const DeflateParams &prm);
ZB_HD void parse_init(ParseState &s, BlockInfo *blocks, uint32_t n, const DeflateParams &prm, uint32_t lo = 0) {
    s.base = parse_base_at(lo, n, prm);
    s.slide_at = parse_next_slide(s.base, n, prm);
    s.p = lo; s.match_length = kMinMatch - 1; s.cur_dist = 0; s.match_available = 0;
    s.nsyms = 0; s.nblocks = 0; s.blk_sym0 = 0; s.blk_byte0 = lo; s.blocks = blocks; s.block_mode = PB_SERIAL; s.sym_limit = prm.sym_limit;
}

ZB_HD void parse_close_block(ParseState &s, uint32_t cover_end, bool last) {
    if (s.block_mode != PB_SERIAL) {
        if (s.block_mode == PB_DEFERRED) {             // block index = symbols so far / sym_limit - 1
            BlockInfo &d = s.blocks[s.nsyms / s.sym_limit - 1];
            d.byte_len = cover_end;                    // provisional: end offset of the block
            d.pad = s.base;                            // provisional: window base when the block was closed
        }
        s.blk_sym0 = s.nsyms;
        return;
    }
    BlockInfo b;
    b.sym_start = s.blk_sym0; b.sym_count = s.nsyms - s.blk_sym0;
    b.byte_start = s.blk_byte0; b.byte_len = cover_end - s.blk_byte0;
    // deflate.c:1597-1600: the stored form needs the block start still inside the window
    b.flags = (last ? BLK_LAST : 0) | (s.blk_byte0 >= s.base ? BLK_STORED_OK : 0);
    b.bit_start_lo = b.bit_start_hi = 0; b.pad = 0;
    s.blocks[s.nblocks++] = b;
    s.blk_sym0 = s.nsyms; s.blk_byte0 = cover_end;
}

// fill_window is entered at a loop top when lookahead drops under `trigger` bytes,
// and slides the window when strstart has reached wsize + MAX_DIST (deflate.c:277).
// With the whole chunk available the window is always full (or holds the tail), so
// the first loop top that slides is the first with p >= slide_at:
ZB_HD uint32_t parse_next_slide(uint32_t base, uint32_t n, const DeflateParams &prm) {
    const uint32_t trigger = prm.mode == MODE_HUFF ? 1 : prm.mode == MODE_RLE ? (uint32_t)kMaxMatch + 1 : (uint32_t)kMinLookahead;
    uint64_t fill_end = (uint64_t)base + 2 * (uint64_t)prm.w_size;
    if (fill_end > n) fill_end = n;
    // lookahead = fill_end - p < trigger  <=>  p > fill_end - trigger ;  and  p - base >= wsize + MAX_DIST
    const uint64_t a = fill_end >= trigger ? fill_end - trigger + 1 : 0;
    const uint64_t b = (uint64_t)base + prm.w_size + prm.max_dist;
    const uint64_t at = a > b ? a : b;
    return at > 0xffffffffull ? 0xffffffffu : (uint32_t)at;
}
ZB_HD void parse_slide_check(ParseState &s, uint32_t p, uint32_t n, const DeflateParams &prm) {
    if (p >= s.slide_at) { s.base += prm.w_size; s.slide_at = parse_next_slide(s.base, n, prm); }
}

// Examine positions while p < limit (limit <= n).  A match may carry p past limit.
template <class Acc>
ZB_HD void parse_steps(ParseState &s, uint32_t limit, uint32_t n, const DeflateParams &prm, Acc &acc) {
    auto emit = [&](uint32_t sym, uint32_t cover_end) {
        acc.put(sym);
        ++s.nsyms;
        if (s.nsyms - s.blk_sym0 == prm.sym_limit) parse_close_block(s, cover_end, false);
    };
    uint32_t p = s.p;
    if (prm.mode != MODE_SLOW) {                       // greedy: deflate_fast / deflate_rle / deflate_huff
        while (p < limit) {
            parse_slide_check(s, p, n, prm);
            const uint32_t m = prm.mode == MODE_HUFF ? 0 : acc.mf(p);
            if (m) { const uint32_t len = m >> 16; emit(((m & 0xffff) << 16) | (len - kMinMatch), p + len); p += len; }
            else { emit(acc.byte(p), p + 1); ++p; }
        }
    } else {                                           // lazy: deflate_slow
        uint32_t match_length = s.match_length, cur_dist = s.cur_dist;
        bool match_available = s.match_available != 0;
        while (p < limit) {
            parse_slide_check(s, p, n, prm);
            const uint32_t prev_length = match_length, prev_dist_v = cur_dist;
            match_length = kMinMatch - 1;
            if (prev_length < (uint32_t)prm.lazy) {
                const uint32_t m = (prev_length >= (uint32_t)prm.good) ? acc.mq(p) : acc.mf(p);
                const uint32_t len = m >> 16;
                if (m && len > prev_length) {
                    match_length = len; cur_dist = m & 0xffff;
                    if (match_length <= 5 && (prm.strategy == STRAT_FILTERED ||
                                              (match_length == (uint32_t)kMinMatch && cur_dist > (uint32_t)kTooFar)))
                        match_length = kMinMatch - 1;  // deflate.c:1964-1975
                }
            }
            if (prev_length >= (uint32_t)kMinMatch && match_length <= prev_length) {
                const uint32_t start = p - 1;
                p = start + prev_length;
                match_available = false; match_length = kMinMatch - 1;
                emit((prev_dist_v << 16) | (prev_length - kMinMatch), p);
            } else if (match_available) {
                emit(acc.byte(p - 1), p);
                ++p;
            } else {
                match_available = true; ++p;
            }
        }
        s.match_length = match_length; s.cur_dist = cur_dist; s.match_available = match_available ? 1 : 0;
    }
    s.p = p;
}

// End of the chunk (p == n).  `final_chunk`: flush == Z_FINISH (last block gets
// BFINAL, an empty final block is emitted if nothing is pending,
// deflate.c:1908-1913); otherwise flush == Z_FULL_FLUSH.
template <class Acc>
ZB_HD void parse_finish(ParseState &s, uint32_t n, const DeflateParams &prm, bool final_chunk, Acc &acc) {
    // deflate.c:2026-2030: the pending literal is tallied but its flush flag is
    // ignored (the block is closed by the flush below, never cut here)
    if (prm.mode == MODE_SLOW && s.match_available) { acc.put(acc.byte(n - 1)); ++s.nsyms; s.match_available = 0; }
    parse_slide_check(s, n, n, prm);                   // the loop-top fill_window call that finds lookahead == 0
    if (final_chunk) parse_close_block(s, n, true);
    else if (s.nsyms != s.blk_sym0) parse_close_block(s, n, false);
}

// Whole-chunk convenience over plain arrays (host replay, small inputs).
struct ParseArrays {
    const uint8_t *data; const uint32_t *mfull, *mquarter; uint32_t *syms; uint32_t count;
    ZB_HD uint32_t mf(uint32_t p) const { return mfull[p]; }
    ZB_HD uint32_t mq(uint32_t p) const { return mquarter[p]; }
    ZB_HD uint32_t byte(uint32_t p) const { return data[p]; }
    ZB_HD void put(uint32_t sym) { syms[count++] = sym; }
};

ZB_HD void parse_chunk(const uint8_t *data, uint32_t n, const uint32_t *mfull, const uint32_t *mquarter,
                       const DeflateParams &prm, bool final_chunk, uint32_t *syms, BlockInfo *blocks,
                       uint32_t &nsyms, uint32_t &nblocks, uint32_t lo = 0) {
    ParseState s;
    parse_init(s, blocks, n, prm, lo);
    ParseArrays acc{data, mfull, mquarter, syms, 0};
    parse_steps(s, n, n, prm, acc);
    parse_finish(s, n, prm, final_chunk, acc);
    nsyms = s.nsyms; nblocks = s.nblocks;
}

// ---- deflate_fast with the reference's OWN chains (levels 1-3 byte for byte; opt-in) ----------
// deflate_fast (deflate.c:1824-1915) inserts a position into the hash chains only where its loop stands, plus the inside
// of a match no longer than max_insert_length (= max_lazy: 4 / 5 / 6 at levels 1-3; deflate.c:1873-1897) — so what a search
// finds depends on every parse decision of the 32 KiB before it, and two parses that differ early need not meet again:
// unlike the lazy parse there is nothing to speculate on, a chunk is one serial walk.  The walk itself is the greedy rule of
// parse_steps with an accessor that keeps head / link tables as it goes: mf(p) = INSERT_STRING(p), longest_match from the
// head it displaced (match_walk: prev_length = 2, no TOO_FAR rule, nice_match clamped to the lookahead), then the
// insertions deflate_fast does behind a short match.  Position 0 is never a candidate (a head of 0 is NIL, deflate.c:1366);
// heads that a slide would have cut are further back than MAX_DIST and end the walk either way.
// head: 1 << hash_bits entries, zeroed; link: one u16 distance per position (what the chain kernel writes for the other levels).
struct FastExactAcc {
    const uint8_t *data; uint32_t n; const DeflateParams *prm; uint32_t *head; uint16_t *link; uint32_t *syms; uint32_t count;
    ZB_HD void insert(uint32_t p) {
        const uint32_t h = hash3(data + p, *prm), q = head[h];
        link[p] = (q && p - q <= 65535u) ? (uint16_t)(p - q) : (uint16_t)0;
        head[h] = p;
    }
    ZB_HD uint32_t mf(uint32_t p) {
        if (p + kMinMatch > n) return 0;                       // lookahead < MIN_MATCH: neither inserted nor searched (deflate.c:1847)
        insert(p);
        const PlainMem mem{data, link};
        const uint32_t m = match_walk(mem, n, p, *prm).full;
        const uint32_t len = m >> 16;
        if (len >= (uint32_t)kMinMatch && len <= (uint32_t)prm->lazy && n - (p + len) >= (uint32_t)kMinMatch)
            for (uint32_t q = p + 1; q < p + len; ++q) insert(q);   // deflate.c:1879-1888
        return m;
    }
    ZB_HD uint32_t mq(uint32_t p) { return mf(p); }
    ZB_HD uint32_t byte(uint32_t p) const { return data[p]; }
    ZB_HD void put(uint32_t sym) { syms[count++] = sym; }
};

ZB_HD void fast_exact_chunk(const uint8_t *data, uint32_t n, const DeflateParams &prm, bool final_chunk, uint32_t *head,
                            uint16_t *link, uint32_t *syms, BlockInfo *blocks, uint32_t &nsyms, uint32_t &nblocks) {
    ParseState s;
    parse_init(s, blocks, n, prm, 0);
    FastExactAcc acc{data, n, &prm, head, link, syms, 0};
    parse_steps(s, n, n, prm, acc);
    parse_finish(s, n, prm, final_chunk, acc);
    nsyms = s.nsyms; nblocks = s.nblocks;
}

// ---- the parse, cut into segments that are parsed concurrently ---------------------
// The parse is a serial state machine, but like a prefix-code decoder it forgets its
// past quickly: two parses that are at the same position in the same state (held match
// length / distance, pending literal) are identical from there on, and a parse started
// cold in the middle of the chunk falls into step with the true one within a few
// symbols.  So the chunk is cut into kSegLanes segments, one per thread of a CTA:
//   seg_speculate  every lane parses its segment from a cold state, counting symbols,
//                  and records its state at the first loop top at or after each of up
//                  to kSegRecs evenly spaced boundaries;
//   seg_fix        lane i+1 re-parses from the state lane i really ended in, until its
//                  state at a boundary equals the recorded one; the speculative symbol
//                  count of the rest is then spliced in (repeated until no start moves);
//   seg_emit       after an exclusive scan of the counts every lane parses once more
//                  from its true start state and writes its symbols at their final
//                  indices; a symbol that fills a block (deflate.h:354-372) leaves the
//                  block's end offset and the window base in its BlockInfo slot;
//   seg_finish     once per chunk: pending literal, byte ranges / flags, last block.
// The window base at a loop top P is a closed form of P (every slide trigger <= P has
// fired), so a segment needs no history for "stored form allowed" either.
constexpr uint32_t kSegLanes = 128;                   // segments per chunk (threads of the parse CTA)
constexpr uint32_t kSegRecs = 16;                     // boundaries per segment (the segment end is the last one)

ZB_HD uint32_t parse_base_at(uint32_t P, uint32_t n, const DeflateParams &prm) {
    uint32_t base = 0;
    while (P >= parse_next_slide(base, n, prm)) base += prm.w_size;
    return base;
}

// `lo`: the chunk's first lo bytes are history only — a preset dictionary (deflate.c:550-632:
// hashed and searched like any window content, strstart = block_start = lo when the first
// byte is compressed) — so the segments tile [lo, n).
struct SegGeom { uint32_t blk, seg, nact, lo; };      // boundary spacing, segment length, lanes with work, first parsed position
ZB_HD SegGeom seg_geometry(uint32_t n, uint32_t lo = 0) {
    SegGeom g;
    const uint32_t m = n - lo;
    g.blk = (m + kSegLanes * kSegRecs - 1) / (kSegLanes * kSegRecs);
    if (g.blk < 32) g.blk = 32;
    g.seg = g.blk * kSegRecs;
    g.nact = m ? (m + g.seg - 1) / g.seg : 1;
    g.lo = lo;
    return g;
}

// A parse state at a loop top, comparable: w0 = held match (distance | length << 16) |
// pending-literal flag << 25 — the distance only counts while a match is held.
struct SegState { uint32_t p, w0; };
ZB_HD SegState seg_state_of(const ParseState &s) {
    SegState t;
    t.p = s.p;
    t.w0 = (s.match_length >= (uint32_t)kMinMatch ? s.cur_dist : 0u) | (s.match_length << 16) | (s.match_available ? 1u << 25 : 0u);
    return t;
}
ZB_HD void seg_state_load(ParseState &s, const SegState &t, uint32_t n, const DeflateParams &prm) {
    s.p = t.p; s.cur_dist = t.w0 & 0xffffu; s.match_length = (t.w0 >> 16) & 0x1ffu; s.match_available = (t.w0 >> 25) & 1u;
    s.base = parse_base_at(t.p, n, prm);
    s.slide_at = parse_next_slide(s.base, n, prm);
    s.nsyms = 0; s.nblocks = 0; s.blk_sym0 = 0; s.blk_byte0 = 0; s.blocks = nullptr; s.block_mode = PB_NONE; s.sym_limit = prm.sym_limit;
}
ZB_HD SegState seg_cold(uint32_t p) { SegState t; t.p = p; t.w0 = (uint32_t)(kMinMatch - 1) << 16; return t; }

struct SegRec { uint32_t w0, w1; };                   // state at a boundary: w0 as above, w1 = (p - boundary) | symbols before << 9
struct SegLane {
    SegState start, end; uint32_t count;              // the current path (true once seg_fix has settled)
    SegState spec_end; uint32_t spec_count;           // the speculative path
};
// The operand accessor also paces the walk: a thread's operands may be staged a WINDOW of
// positions at a time (the device stages them in shared memory, all threads of a warp
// together, so that no thread's fetch stalls the others mid-step).  windows(): how many
// windows to go through (the same for every thread); open(w, s0): make window w of the
// segment starting at s0 readable and return its end position; any(b): does any thread
// that is running this pass with me still have b set.
struct NoPut {                                        // counting passes: operands from plain arrays, symbols dropped
    const uint8_t *data; const uint32_t *mfull, *mquarter;
    ZB_HD uint32_t mf(uint32_t p) const { return mfull[p]; }
    ZB_HD uint32_t mq(uint32_t p) const { return mquarter[p]; }
    ZB_HD uint32_t byte(uint32_t p) const { return data[p]; }
    ZB_HD void put(uint32_t) {}
    ZB_HD uint32_t windows(uint32_t) const { return 1; }
    ZB_HD uint32_t open(uint32_t, uint32_t) { return 0xffffffffu; }
    ZB_HD bool any(bool b) const { return b; }
};

// rec[(k - 1) * kSegLanes + lane] belongs to boundary k (k = 1 .. kSegRecs - 1) of lane's segment.
template <class Acc>
ZB_HD void seg_speculate(SegLane &r, uint32_t lane, const SegGeom &g, uint32_t n, const DeflateParams &prm, Acc &acc, SegRec *rec) {
    const uint32_t s0 = g.lo + lane * g.seg, bound = s0 + g.seg < n ? s0 + g.seg : n;
    ParseState st;
    r.start = seg_cold(s0);
    seg_state_load(st, r.start, n, prm);
    uint32_t k = 1;
    const uint32_t nw = acc.windows(g.seg);
    for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t we = acc.open(w, s0), wend = we < bound ? we : bound;
        while (st.p < wend) {
            const uint32_t nextb = s0 + k * g.blk;
            const bool more = k < kSegRecs && nextb < bound;
            parse_steps(st, more && nextb < wend ? nextb : wend, n, prm, acc);
            while (k < kSegRecs && s0 + k * g.blk < bound && st.p >= s0 + k * g.blk) {   // boundaries reached (a match may pass several)
                SegRec e;
                e.w0 = seg_state_of(st).w0; e.w1 = (st.p - (s0 + k * g.blk)) | (st.nsyms << 9);
                rec[(k - 1) * kSegLanes + lane] = e;
                ++k;
            }
        }
    }
    r.end = r.spec_end = seg_state_of(st);
    r.count = r.spec_count = st.nsyms;
}

template <class Acc>
ZB_HD void seg_fix(SegLane &r, uint32_t lane, const SegGeom &g, uint32_t n, const DeflateParams &prm, Acc &acc,
                   const SegRec *rec, const SegState &t) {
    const uint32_t s0 = g.lo + lane * g.seg, bound = s0 + g.seg < n ? s0 + g.seg : n;
    ParseState st;
    r.start = t;
    seg_state_load(st, t, n, prm);
    uint32_t k = 1;
    while (k < kSegRecs && s0 + k * g.blk < bound && s0 + k * g.blk < t.p) ++k;      // boundaries before the true start are not compared
    bool joined = false;
    const uint32_t nw = acc.windows(g.seg);
    for (uint32_t w = 0; w < nw; ++w) {
        if (!acc.any(!joined && st.p < bound)) break;
        const uint32_t we = acc.open(w, s0), wend = we < bound ? we : bound;
        while (!joined && st.p < wend) {
            const uint32_t nextb = s0 + k * g.blk;
            const bool more = k < kSegRecs && nextb < bound;
            parse_steps(st, more && nextb < wend ? nextb : wend, n, prm, acc);
            while (!joined && k < kSegRecs && s0 + k * g.blk < bound && st.p >= s0 + k * g.blk) {
                const SegRec e = rec[(k - 1) * kSegLanes + lane];
                if (e.w0 == seg_state_of(st).w0 && (e.w1 & 0x1ffu) == st.p - (s0 + k * g.blk)) {   // joined the speculative path
                    r.count = st.nsyms + (r.spec_count - (e.w1 >> 9));
                    r.end = r.spec_end;
                    joined = true;
                }
                ++k;
            }
        }
    }
    if (!joined) { r.end = seg_state_of(st); r.count = st.nsyms; }
}

// `acc.put` must store symbol number (first + k) of the lane at index first + k.
template <class Acc>
ZB_HD void seg_emit(const SegLane &r, uint32_t lane, const SegGeom &g, uint32_t n, const DeflateParams &prm, Acc &acc,
                    BlockInfo *blocks, uint32_t first) {
    const uint32_t s0 = g.lo + lane * g.seg, bound = s0 + g.seg < n ? s0 + g.seg : n;
    ParseState st;
    seg_state_load(st, r.start, n, prm);
    st.nsyms = first; st.blk_sym0 = first - first % prm.sym_limit; st.blocks = blocks; st.block_mode = PB_DEFERRED;
    const uint32_t nw = acc.windows(g.seg);
    for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t we = acc.open(w, s0), wend = we < bound ? we : bound;
        if (st.p < wend) parse_steps(st, wend, n, prm, acc);
    }
}

// nsyms: symbols emitted by the segments; pending: the last segment ended holding a
// literal (deflate.c:2026-2030: tallied at the end, its flush flag ignored — the caller
// has stored it at index nsyms).  Returns the number of blocks.
ZB_HD uint32_t seg_finish(BlockInfo *blocks, uint32_t nsyms, bool pending, uint32_t n, const DeflateParams &prm, bool final_chunk,
                          uint32_t lo = 0) {
    const uint32_t lim = prm.sym_limit;
    const uint32_t nfull = nsyms / lim, rem = nsyms % lim + (pending ? 1u : 0u);
    uint32_t start = lo;
    for (uint32_t k = 0; k < nfull; ++k) {
        BlockInfo &b = blocks[k];
        const uint32_t end = b.byte_len, base = b.pad;
        b.sym_start = k * lim; b.sym_count = lim;
        b.byte_start = start; b.byte_len = end - start;
        b.flags = start >= base ? BLK_STORED_OK : 0;  // deflate.c:1597-1600
        b.bit_start_lo = b.bit_start_hi = 0; b.pad = 0;
        start = end;
    }
    uint32_t nblocks = nfull;
    if (final_chunk || rem) {
        BlockInfo &b = blocks[nblocks++];
        b.sym_start = nfull * lim; b.sym_count = rem;
        b.byte_start = start; b.byte_len = n - start;
        b.flags = (final_chunk ? BLK_LAST : 0) | (start >= parse_base_at(n, n, prm) ? BLK_STORED_OK : 0);
        b.bit_start_lo = b.bit_start_hi = 0; b.pad = 0;
    }
    return nblocks;
}

// ---- greedy parse, split for the GPU ----------------------------------------------
// With a greedy rule (deflate_fast / deflate_rle / deflate_huff) the only serial part
// is the chase p -> p + len(p); everything else is a pure function of the visited
// position and of the symbol's index, so it is done by all lanes afterwards:
//   greedy_chase   serial: records the visited positions of one tile
//   greedy_symbol  parallel: symbol word of one visited position; the symbol that
//                  closes a block (every 16383rd) leaves the block's end offset and
//                  loop-top position in its BlockInfo slot
//   seg_finish     once per chunk (shared with the segmented parse below): turns those
//                  into byte ranges / flags, adds the last block.
template <class Acc>
ZB_HD uint32_t greedy_chase(uint32_t &p, uint32_t limit, bool use_m, Acc &acc, uint32_t *visited) {
    uint32_t cnt = 0, q = p;
    while (q < limit) {
        const uint32_t m = use_m ? acc.mf(q) : 0u;
        visited[cnt++] = q;
        q += m ? (m >> 16) : 1u;
    }
    p = q;
    return cnt;
}

template <class Acc>
ZB_HD uint32_t greedy_symbol(uint32_t pos, bool use_m, Acc &acc, uint32_t g, BlockInfo *blocks, uint32_t n, const DeflateParams &prm) {
    const uint32_t m = use_m ? acc.mf(pos) : 0u;
    const uint32_t len = m ? (m >> 16) : 1u;
    if ((g + 1) % prm.sym_limit == 0) {                // this symbol fills a block (deflate.h:354-372 flush flag)
        BlockInfo &b = blocks[g / prm.sym_limit];
        b.byte_len = pos + len;                        // provisional: end offset of the block
        b.pad = parse_base_at(pos, n, prm);            // provisional: window base at the flush (seg_finish completes the slot)
    }
    return m ? (((m & 0xffff) << 16) | (len - kMinMatch)) : acc.byte(pos);
}

// ---- greedy parse with exit tables ---------------------------------------------------
// The chase p -> p + len(p) is a functional graph: whatever position the parse enters a
// short segment at, where it leaves the segment and how many symbols it emits on the way
// depend on the segment alone.  A tile of kGtTile positions is cut into 32 segments of
// kGtSeg; lane i fills, for every position of segment i (backwards), the position at which
// a parse entering there leaves the segment and the symbols it emits until then
// (gt_fill).  With the tables in place the true parse crosses the tile in one dependent
// table read per segment (gt_hop) instead of one per symbol, and each lane then forms the
// symbols of its own segment from its entry point (gt_emit).  Only gt_hop is ordered
// between tiles, so the four warps of a CTA work on four tiles at once and pass (entry
// position, symbol count) from tile to tile.
constexpr uint32_t kGtSeg = 16, kGtTile = 32 * kGtSeg;
ZB_HD uint32_t gt_slot(uint32_t r) { return (r % kGtSeg) * 33u + r / kGtSeg; }   // tile-relative position -> table slot (bank = lane + offset)
constexpr uint32_t kGtSlots = kGtSeg * 33u;

// mfv[gt_slot(r)] = match entry of tile position r (0 beyond the chunk).  Fills lc[] for the lane's
// segment: tile-relative position at which the parse leaves the segment (may exceed the tile by up
// to MAX_MATCH: 10 bits) | symbols emitted until then << 10.  `limit`: tile-relative end of the
// chunk (a parse leaves the tile there at the latest).
ZB_HD void gt_fill(uint32_t lane, uint32_t limit, const uint32_t *mfv, uint16_t *lc) {
    const uint32_t s0 = lane * kGtSeg, s1 = s0 + kGtSeg < limit ? s0 + kGtSeg : limit;
    for (int j = (int)kGtSeg - 1; j >= 0; --j) {
        const uint32_t r = s0 + (uint32_t)j, m = mfv[j * 33 + lane];
        if (r >= limit) continue;
        const uint32_t nx = r + (m ? (m >> 16) : 1u);
        lc[j * 33 + lane] = (uint16_t)(nx >= s1 ? (nx | (1u << 10)) : (lc[(nx - s0) * 33 + lane] + (1u << 10)));
    }
}

// Cross the tile from tile-relative position `e` (may lie beyond the tile: a match carried over it).
// Every lane runs this redundantly; it learns where the parse enters ITS segment (my_entry, or
// 0xffffffff) and how many symbols the segments before it emit (my_first).  Returns the tile-relative
// exit position; `total` = symbols of the tile.  `limit`: tile-relative end of the chunk.
ZB_HD uint32_t gt_hop(uint32_t lane, uint32_t e, uint32_t limit, const uint16_t *lc,
                      uint32_t &my_entry, uint32_t &my_first, uint32_t &total) {
    uint32_t cur = e, sum = 0;
    my_entry = 0xffffffffu; my_first = 0;
    const uint32_t end = limit < kGtTile ? limit : kGtTile;
    while (cur < end) {
        const uint32_t seg = cur / kGtSeg;
        if (seg == lane) { my_entry = cur; my_first = sum; }
        const uint32_t v = lc[(cur % kGtSeg) * 33u + seg];
        sum += v >> 10;
        cur = v & 1023u;
    }
    total = sum;
    return cur;
}

// ---- phase 4: per-block Huffman construction ---------------------------------------
struct BlockCode {
    uint16_t lcode[288]; uint8_t llen[288];   // bit-reversed codes / lengths, literal-length alphabet
    uint16_t dcode[32];  uint8_t dlen[32];    // distance alphabet
    uint32_t type;                            // 0 stored, 1 fixed, 2 dynamic
    uint32_t hdr_bits;                        // bits of the dynamic header that follow the 3 block-header bits
    uint32_t body_bits;                       // 3 + (static_len | opt_len): bits of a fixed/dynamic block
    uint32_t pad;
    uint32_t hdr[160];                        // dynamic header, LSB-first bit string
};

struct StaticTrees {                          // trees.h static_ltree / static_dtree, generated at start-up
    uint16_t lcode[288]; uint8_t llen[288];
    uint16_t dcode[32];  uint8_t dlen[32];
};

struct TreeNode { uint16_t fc; uint16_t dl; };   // Freq|Code, Dad|Len   (deflate.h:75-90)

struct TreeWork {                             // one per CTA (shared memory)
    TreeNode lt[2 * 286 + 1], dt[2 * 30 + 1], bt[2 * 19 + 1];
    int16_t heap[2 * 286 + 1];
    uint8_t depth[2 * 286 + 1];
    uint16_t bl_count[16];
    int heap_len, heap_max;
    uint32_t opt_len, static_len;
    uint32_t hkey[286 + 2];                   // tree_build_fast: (Freq << 8) | depth of heap[1 .. heap_len]; the kernel also
                                              // counts literal/length frequencies here and stages the dynamic header here
    uint32_t dhist[32];                       // ... and distance-code frequencies here
};

ZB_HD uint32_t bit_reverse(uint32_t code, int len) {   // trees.c:154
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1); code >>= 1; }
    return r;
}

ZB_HD void tree_gen_codes(TreeNode *t, int max_code, const uint16_t *blc) {   // trees.c:203-232
    uint16_t next[16];
    uint32_t code = 0;
    for (int b = 1; b <= 15; ++b) { code = (code + blc[b - 1]) << 1; next[b] = (uint16_t)code; }
    for (int n = 0; n <= max_code; ++n) {
        const int len = t[n].dl;
        if (len) t[n].fc = (uint16_t)bit_reverse(next[len]++, len);
    }
}

inline void static_trees_fill(StaticTrees &st) {       // trees.c:374-390
    TreeNode t[288];
    uint16_t blc[16] = {0};
    int n = 0;
    for (; n <= 143; ++n) { t[n].dl = 8; blc[8]++; }
    for (; n <= 255; ++n) { t[n].dl = 9; blc[9]++; }
    for (; n <= 279; ++n) { t[n].dl = 7; blc[7]++; }
    for (; n <= 287; ++n) { t[n].dl = 8; blc[8]++; }
    tree_gen_codes(t, 287, blc);
    for (n = 0; n < 288; ++n) { st.lcode[n] = t[n].fc; st.llen[n] = (uint8_t)t[n].dl; }
    for (n = 0; n < 32; ++n) { st.dcode[n] = (uint16_t)bit_reverse((uint32_t)n, 5); st.dlen[n] = 5; }
}

#define ZB_SMALLER(t, n, m) ((t)[n].fc < (t)[m].fc || ((t)[n].fc == (t)[m].fc && w.depth[n] <= w.depth[m]))   // trees.c:499-501

ZB_HD void tree_sift(TreeWork &w, TreeNode *t, int k) {  // trees.c:509 pqdownheap
    const int v = w.heap[k];
    for (int j = k << 1; j <= w.heap_len; j <<= 1) {
        if (j < w.heap_len && ZB_SMALLER(t, w.heap[j + 1], w.heap[j])) ++j;
        if (ZB_SMALLER(t, v, w.heap[j])) break;
        w.heap[k] = w.heap[j]; k = j;
    }
    w.heap[k] = (int16_t)v;
}

// trees.c:540-613 gen_bitlen.  st_len: static code lengths (nullptr for the bit-length tree).
ZB_HD void tree_gen_bitlen(TreeWork &w, TreeNode *t, int max_code, const uint8_t *st_len,
                           const uint8_t *extra, int base, int max_length) {
    constexpr int HEAP_SIZE = 2 * 286 + 1;
    int h, overflow = 0;
    for (int b = 0; b <= 15; ++b) w.bl_count[b] = 0;
    t[w.heap[w.heap_max]].dl = 0;
    for (h = w.heap_max + 1; h < HEAP_SIZE; ++h) {
        const int n = w.heap[h];
        int bits = t[t[n].dl].dl + 1;
        if (bits > max_length) { bits = max_length; ++overflow; }
        t[n].dl = (uint16_t)bits;
        if (n > max_code) continue;
        w.bl_count[bits]++;
        const int xb = (n >= base) ? extra[n - base] : 0;
        const uint32_t f = t[n].fc;
        w.opt_len += f * (uint32_t)(bits + xb);
        if (st_len) w.static_len += f * (uint32_t)(st_len[n] + xb);
    }
    if (!overflow) return;
    do {
        int bits = max_length - 1;
        while (w.bl_count[bits] == 0) --bits;
        w.bl_count[bits]--; w.bl_count[bits + 1] += 2; w.bl_count[max_length]--;
        overflow -= 2;
    } while (overflow > 0);
    for (int bits = max_length; bits != 0; --bits) {
        int n = w.bl_count[bits];
        while (n != 0) {
            const int m = w.heap[--h];
            if (m > max_code) continue;
            if (t[m].dl != (uint32_t)bits) {
                w.opt_len += ((uint32_t)bits - t[m].dl) * t[m].fc;
                t[m].dl = (uint16_t)bits;
            }
            --n;
        }
    }
}

// trees.c:627-706 build_tree; returns max_code.
ZB_HD int tree_build(TreeWork &w, TreeNode *t, int elems, const uint8_t *st_len,
                     const uint8_t *extra, int base, int max_length) {
    constexpr int HEAP_SIZE = 2 * 286 + 1;
    int max_code = -1, node;
    w.heap_len = 0; w.heap_max = HEAP_SIZE;
    for (int n = 0; n < elems; ++n) {
        if (t[n].fc) { w.heap[++w.heap_len] = (int16_t)(max_code = n); w.depth[n] = 0; }
        else t[n].dl = 0;
    }
    while (w.heap_len < 2) {                           // trees.c:655-661: force two codes
        node = w.heap[++w.heap_len] = (int16_t)(max_code < 2 ? ++max_code : 0);
        t[node].fc = 1; w.depth[node] = 0;
        w.opt_len--; if (st_len) w.static_len -= st_len[node];
    }
    for (int n = w.heap_len / 2; n >= 1; --n) tree_sift(w, t, n);
    node = elems;
    do {
        const int n = w.heap[1];
        w.heap[1] = w.heap[w.heap_len--];
        tree_sift(w, t, 1);
        const int m = w.heap[1];
        w.heap[--w.heap_max] = (int16_t)n;
        w.heap[--w.heap_max] = (int16_t)m;
        t[node].fc = (uint16_t)(t[n].fc + t[m].fc);
        w.depth[node] = (uint8_t)((w.depth[n] >= w.depth[m] ? w.depth[n] : w.depth[m]) + 1);
        t[n].dl = t[m].dl = (uint16_t)node;
        w.heap[1] = (int16_t)node++;
        tree_sift(w, t, 1);
    } while (w.heap_len >= 2);
    w.heap[--w.heap_max] = w.heap[1];
    tree_gen_bitlen(w, t, max_code, st_len, extra, base, max_length);
    tree_gen_codes(t, max_code, w.bl_count);
    return max_code;
}

// ---- the same construction, arranged for one GPU thread ------------------------------------
// pqdownheap's comparison (trees.c:499-501: smaller by frequency, ties by subtree depth, `<=`) reads heap[j], then the
// node's Freq, then its depth: three dependent shared-memory round trips per heap level.  Here the active heap carries
// its sort key beside the node number — (Freq << 8) | depth, compared with one `<=` — so a level costs one round trip
// (both children are requested together).  Merge order, heap contents and therefore every code length are those of
// tree_build (the host replay compares the two over random and adversarial histograms).
ZB_HD void tree_sift_fast(TreeWork &w, int k) {
    const uint32_t vk = w.hkey[k];
    const int16_t vn = w.heap[k];
    const int len = w.heap_len;
    int j = k << 1;
    while (j <= len) {
        uint32_t kj = w.hkey[j];
        if (j < len) { const uint32_t kr = w.hkey[j + 1]; if (kr <= kj) { kj = kr; ++j; } }
        if (vk <= kj) break;
        w.hkey[k] = kj; w.heap[k] = w.heap[j];
        k = j; j <<= 1;
    }
    w.hkey[k] = vk; w.heap[k] = vn;
}

ZB_HD uint32_t bit_reverse_fast(uint32_t code, int len) {
#if defined(__CUDA_ARCH__)
    return __brev(code) >> (32 - len);
#else
    return bit_reverse(code, len);
#endif
}

ZB_HD int tree_build_fast(TreeWork &w, TreeNode *t, int elems, const uint8_t *st_len,
                          const uint8_t *extra, int base, int max_length) {
    constexpr int HEAP_SIZE = 2 * 286 + 1;
    int max_code = -1, node;
    w.heap_len = 0; w.heap_max = HEAP_SIZE;
    for (int n = 0; n < elems; ++n) {
        const uint32_t f = t[n].fc;
        if (f) { w.heap[++w.heap_len] = (int16_t)(max_code = n); w.hkey[w.heap_len] = f << 8; w.depth[n] = 0; }
        else t[n].dl = 0;
    }
    while (w.heap_len < 2) {                           // trees.c:655-661: force two codes
        node = w.heap[++w.heap_len] = (int16_t)(max_code < 2 ? ++max_code : 0);
        t[node].fc = 1; w.depth[node] = 0; w.hkey[w.heap_len] = 1u << 8;
        w.opt_len--; if (st_len) w.static_len -= st_len[node];
    }
    for (int n = w.heap_len / 2; n >= 1; --n) tree_sift_fast(w, n);
    node = elems;
    do {
        const int n = w.heap[1];
        const uint32_t kn = w.hkey[1];
        w.heap[1] = w.heap[w.heap_len]; w.hkey[1] = w.hkey[w.heap_len]; --w.heap_len;
        tree_sift_fast(w, 1);
        const int m = w.heap[1];
        const uint32_t km = w.hkey[1];
        w.heap[--w.heap_max] = (int16_t)n;
        w.heap[--w.heap_max] = (int16_t)m;
        const uint32_t f = ((kn >> 8) + (km >> 8)) & 0xffffu;                       // ush arithmetic, as deflate.h:75-90
        const uint32_t dn = kn & 0xffu, dm = km & 0xffu, d = ((dn >= dm ? dn : dm) + 1u) & 0xffu;   // uch depth
        t[node].fc = (uint16_t)f; w.depth[node] = (uint8_t)d;
        t[n].dl = t[m].dl = (uint16_t)node;
        w.heap[1] = (int16_t)node++; w.hkey[1] = (f << 8) | d;
        tree_sift_fast(w, 1);
    } while (w.heap_len >= 2);
    w.heap[--w.heap_max] = w.heap[1];
    tree_gen_bitlen(w, t, max_code, st_len, extra, base, max_length);
    {                                                  // trees.c:203-232 gen_codes
        uint16_t next[16];
        uint32_t code = 0;
        for (int b = 1; b <= 15; ++b) { code = (code + w.bl_count[b - 1]) << 1; next[b] = (uint16_t)code; }
        for (int n = 0; n <= max_code; ++n) {
            const int len = t[n].dl;
            if (len) t[n].fc = (uint16_t)bit_reverse_fast(next[len]++, len);
        }
    }
    return max_code;
}

struct HdrWriter {                                    // LSB-first appender for the dynamic header
    uint32_t *words; uint32_t nbits;
    ZB_HD void put(uint32_t v, int len) {
        const uint32_t w = nbits >> 5, sh = nbits & 31;
        words[w] |= v << sh;
        if (sh + (uint32_t)len > 32) words[w + 1] |= v >> (32 - sh);
        nbits += (uint32_t)len;
    }
};

// trees.c:712-745 scan_tree (out == nullptr) / :753-795 send_tree share the run-length walk.
ZB_HD void tree_walk(TreeWork &w, TreeNode *t, int max_code, HdrWriter *out) {
    int prevlen = -1, nextlen = t[0].dl, count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    if (!out) t[max_code + 1].dl = 0xffff;            // guard
    TreeNode *bt = w.bt;
    for (int n = 0; n <= max_code; ++n) {
        const int curlen = nextlen;
        nextlen = t[n + 1].dl;
        if (++count < max_count && curlen == nextlen) continue;
        if (count < min_count) {
            if (out) { do out->put(bt[curlen].fc, bt[curlen].dl); while (--count); }
            else bt[curlen].fc += (uint16_t)count;
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (out) { out->put(bt[curlen].fc, bt[curlen].dl); --count; }
                else bt[curlen].fc++;
            }
            if (out) { out->put(bt[16].fc, bt[16].dl); out->put((uint32_t)count - 3, 2); }
            else bt[16].fc++;
        } else if (count <= 10) {
            if (out) { out->put(bt[17].fc, bt[17].dl); out->put((uint32_t)count - 3, 3); }
            else bt[17].fc++;
        } else {
            if (out) { out->put(bt[18].fc, bt[18].dl); out->put((uint32_t)count - 11, 7); }
            else bt[18].fc++;
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

// trees.c:997-1089 _tr_flush_block for one block whose symbol frequencies are
// already in w.lt[0..285].fc / w.dt[0..29].fc (END_BLOCK counted).  Decides
// stored / fixed / dynamic and fills `out`.
// kFast: the keyed heap (tree_build_fast).  kCopy: also copy the chosen code tables into `out` (the GPU kernel lets
// the whole warp do that afterwards: out.pad = l_max | d_max << 16 tells it how far the dynamic lengths reach).
// hdr: where the dynamic header's bit string goes (160 words; out.hdr, or a staging area the caller copies out).
template <bool kFast, bool kCopy>
ZB_HD void block_build_t(TreeWork &w, const BlockInfo &blk, int strategy, const StaticTrees &st,
                         const FormatTables &fmt, BlockCode &out, uint32_t *hdr) {
    const uint8_t bl_extra[19] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 3, 7};
    w.opt_len = 0; w.static_len = 0;
    for (int n = 0; n < 19; ++n) w.bt[n].fc = 0;
    const int l_max = kFast ? tree_build_fast(w, w.lt, 286, st.llen, fmt.len_extra, 257, 15) : tree_build(w, w.lt, 286, st.llen, fmt.len_extra, 257, 15);
    const int d_max = kFast ? tree_build_fast(w, w.dt, 30, st.dlen, fmt.dist_extra, 0, 15) : tree_build(w, w.dt, 30, st.dlen, fmt.dist_extra, 0, 15);
    tree_walk(w, w.lt, l_max, nullptr);                // trees.c:800-829 build_bl_tree
    tree_walk(w, w.dt, d_max, nullptr);
    if (kFast) tree_build_fast(w, w.bt, 19, nullptr, bl_extra, 0, 7); else tree_build(w, w.bt, 19, nullptr, bl_extra, 0, 7);
    int mb;
    for (mb = 18; mb >= 3; --mb) if (w.bt[fmt.cl_order[mb]].dl) break;
    w.opt_len += 3 * ((uint32_t)mb + 1) + 5 + 5 + 4;
    uint32_t opt_lenb = (w.opt_len + 3 + 7) >> 3;
    const uint32_t static_lenb = (w.static_len + 3 + 7) >> 3;
    const bool force_stored = (strategy & 0x100) != 0;   // level 0: trees.c:1041-1043 forces stored blocks
    strategy &= 0xff;
    if (static_lenb <= opt_lenb || strategy == STRAT_FIXED) opt_lenb = static_lenb;
    out.hdr_bits = 0; out.pad = (uint32_t)l_max | ((uint32_t)d_max << 16);
    if (force_stored || (blk.byte_len + 4 <= opt_lenb && (blk.flags & BLK_STORED_OK))) {
        out.type = 0; out.body_bits = 0;
    } else if (static_lenb == opt_lenb) {
        out.type = 1; out.body_bits = 3 + w.static_len;
        if (kCopy) {
            for (int n = 0; n < 288; ++n) { out.lcode[n] = st.lcode[n]; out.llen[n] = st.llen[n]; }
            for (int n = 0; n < 32; ++n) { out.dcode[n] = st.dcode[n]; out.dlen[n] = st.dlen[n]; }
        }
    } else {
        out.type = 2; out.body_bits = 3 + w.opt_len;
        if (kCopy) {
            for (int n = 0; n < 288; ++n) { out.lcode[n] = n < 286 ? w.lt[n].fc : 0; out.llen[n] = n <= l_max ? (uint8_t)w.lt[n].dl : 0; }
            for (int n = 0; n < 32; ++n) { out.dcode[n] = n < 30 ? w.dt[n].fc : 0; out.dlen[n] = n <= d_max ? (uint8_t)w.dt[n].dl : 0; }
        }
        for (int i = 0; i < 160; ++i) hdr[i] = 0;
        HdrWriter hw{hdr, 0};
        hw.put((uint32_t)l_max + 1 - 257, 5);         // trees.c:833-855 send_all_trees
        hw.put((uint32_t)d_max + 1 - 1, 5);
        hw.put((uint32_t)mb + 1 - 4, 4);
        for (int r = 0; r <= mb; ++r) hw.put(w.bt[fmt.cl_order[r]].dl, 3);
        tree_walk(w, w.lt, l_max, &hw);
        tree_walk(w, w.dt, d_max, &hw);
        out.hdr_bits = hw.nbits;
    }
}
ZB_HD void block_build_fast(TreeWork &w, const BlockInfo &blk, int strategy, const StaticTrees &st,
                            const FormatTables &fmt, BlockCode &out) {   // host replay of the kernel's construction
    block_build_t<true, true>(w, blk, strategy, st, fmt, out, out.hdr);
    out.pad = 0;
}
ZB_HD void block_build(TreeWork &w, const BlockInfo &blk, int strategy, const StaticTrees &st,
                       const FormatTables &fmt, BlockCode &out) {
    block_build_t<false, true>(w, blk, strategy, st, fmt, out, out.hdr);
    out.pad = 0;
}

// ---- phase 5: symbol -> bits (trees.c:900-951 compress_block) ----------------------
ZB_HD uint64_t symbol_bits(uint32_t sym, const uint16_t *lcode, const uint8_t *llen, const uint16_t *dcode,
                           const uint8_t *dlen, const FormatTables &fmt, uint32_t &nbits) {
    const uint32_t dist = sym >> 16, lc = sym & 0xffff;
    if (dist == 0) { nbits = llen[lc]; return lcode[lc]; }
    uint32_t c = fmt.len_code[lc];
    uint64_t v = lcode[257 + c];
    uint32_t nb = llen[257 + c];
    const uint32_t lx = fmt.len_extra[c];
    if (lx) { v |= (uint64_t)(lc + kMinMatch - fmt.len_base[c]) << nb; nb += lx; }
    c = dist_to_code(fmt, dist);
    v |= (uint64_t)dcode[c] << nb; nb += dlen[c];
    const uint32_t dx = fmt.dist_extra[c];
    if (dx) { v |= (uint64_t)(dist - fmt.dist_base[c]) << nb; nb += dx; }
    nbits = nb;
    return v;
}

// Bit layout of a chunk: given each block's type and size, where does it start?
// Returns the chunk's size in bytes.  `marker`: append the 00 00 FF FF sync
// marker (non-final chunks, deflate.c:1214-1215 / trees.c:860-875).
ZB_HD uint64_t block_end_bit(const BlockInfo &b, const BlockCode &c, uint64_t start) {
    uint64_t end;
    if (c.type == 0) end = ((start + 3 + 7) & ~7ull) + 32 + 8ull * b.byte_len;
    else end = start + c.body_bits;
    if (b.flags & BLK_LAST) end = (end + 7) & ~7ull;  // bi_windup
    return end;
}

}  // namespace zb
