// zb_inflate.cuh — the per-member inflate state machine.
//
// B200 replacement for inflate.c:590-1264 (inflate), inftrees.c:32-299
// (inflate_table) and inffast.c:50-304 (inflate_fast) for members decoded into
// one contiguous output buffer: all distances resolve inside the output, so the
// 32 KiB side window of inflate.c:368-412 (updatewindow) is not needed
// (inflate.c:362-366).
//
// Division of labour inside the warp that owns a member (zb_inflate.cu):
//   lane 0   runs this state machine: wrapper header, block headers, dynamic
//            table construction, Huffman decode.  Literals are stored as they
//            are decoded.  It stops at the next EVENT.
//   all 32   lanes execute the event: an LZ77 match copy or a stored-block
//            copy, striped over lanes, then lane 0 resumes.
// Everything here is __host__ __device__: tests/emul replays the identical code
// on the CPU (events executed by a plain loop) before any GPU time is spent.
#pragma once
#include "zb_format.h"
#include "../../include/zb200.h"

namespace zb {

// ---- decode tables -----------------------------------------------------------
// entry: [31:16] value   [15:12] extra-bit count / sub-table index bits
//        [11:8]  op      [7:0]   code bits consumed at this level
enum : uint32_t { OP_LIT = 0, OP_BASE = 1, OP_EOB = 2, OP_SUB = 3, OP_BAD = 4 };
ZB_HD uint32_t mk_entry(uint32_t val, uint32_t extra, uint32_t op, uint32_t bits) {
    return (val << 16) | (extra << 12) | (op << 8) | bits;
}
#define ZB_E_BITS(e)  ((e) & 0xffu)
#define ZB_E_OP(e)    (((e) >> 8) & 0xfu)
#define ZB_E_EXTRA(e) (((e) >> 12) & 0xfu)
#define ZB_E_VAL(e)   ((e) >> 16)

constexpr int kLitRoot = 9, kDistRoot = 6, kClRoot = 7;        // inflate.c:927,1002,1011
constexpr int kLitEntries = 852, kDistEntries = 592;           // inftrees.h:49-51 ENOUGH_LENS / ENOUGH_DISTS
constexpr int kClEntries = 128;
enum { TBL_CODELEN = 0, TBL_LITLEN = 1, TBL_DIST = 2 };

// Per-warp working memory (shared memory on the device).
struct InflateScratch {
    uint32_t lit[kLitEntries];
    uint32_t dist[kDistEntries];       // the first kClEntries double as the code-length table while a header is read
    uint16_t work[320];                // symbols sorted by code length
    uint8_t  lens[320];                // code lengths of the block being set up
};

// Decode-ahead queue: lane 0 keeps decoding past a match, parking up to kQueue
// matches here; the warp then executes the whole batch.  A match whose source
// cannot overlap the destination of an earlier parked match ("independent") is
// copied by its own lane, all such lanes in parallel, so their L2 round trips
// overlap; the others run afterwards, in order, striped over the warp.
constexpr uint32_t kQueue = 64;
struct QueuedMatch {
    uint32_t dst;                      // output position (members are limited to < 4 GiB of output)
    uint32_t packed;                   // len (bits 0..8) | dist << 9 (bits 9..24) | dependent << 31
};
ZB_HD uint32_t qm_pack(uint32_t len, uint32_t dist, bool dep) { return len | (dist << 9) | (dep ? 0x80000000u : 0u); }
ZB_HD uint32_t qm_len(uint32_t pk) { return pk & 0x1ffu; }
ZB_HD uint32_t qm_dist(uint32_t pk) { return (pk >> 9) & 0xffffu; }
ZB_HD bool qm_dep(uint32_t pk) { return (pk >> 31) != 0; }

// Build a two-level decode table from code lengths (the job of inftrees.c:32-299).
// Returns 0 = complete code, 1 = incomplete, 2 = no codes at all, -1 = over-subscribed
// or out of table space.  The caller applies the reference's acceptance rules.
ZB_HD int build_decode_table(int type, const uint8_t *lens, int n, uint32_t *table, int cap, int root,
                             uint16_t *work, const FormatTables &fmt) {
    uint16_t count[16], offs[16];
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    int max = 15;
    while (max >= 1 && count[max] == 0) --max;
    const int root_size = 1 << root;
    if (max == 0) {                                   // inftrees.c:111-119: decode-time error markers
        for (int i = 0; i < root_size; ++i) table[i] = mk_entry(0, 0, OP_BAD, 1);
        return 2;
    }
    int left = 1;
    for (int len = 1; len <= 15; ++len) {             // inftrees.c:125-131
        left <<= 1;
        left -= count[len];
        if (left < 0) return -1;
    }
    offs[1] = 0;
    for (int len = 1; len < 15; ++len) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
    for (int i = 0; i < n; ++i) if (lens[i]) work[offs[lens[i]]++] = (uint16_t)i;
    if (left > 0)                                     // incomplete: unassigned patterns must fail at decode time
        for (int i = 0; i < root_size; ++i) table[i] = mk_entry(0, 0, OP_BAD, 1);

    auto entry_for = [&](int sym, int bits) -> uint32_t {
        if (type == TBL_CODELEN) return mk_entry((uint32_t)sym, 0, OP_LIT, (uint32_t)bits);
        if (type == TBL_LITLEN) {
            if (sym < 256) return mk_entry((uint32_t)sym, 0, OP_LIT, (uint32_t)bits);
            if (sym == 256) return mk_entry(0, 0, OP_EOB, (uint32_t)bits);
            if (sym > 285) return mk_entry(0, 0, OP_BAD, (uint32_t)bits);      // inftrees.c:57-60: 286,287 invalid
            return mk_entry(fmt.len_base[sym - 257], fmt.len_extra[sym - 257], OP_BASE, (uint32_t)bits);
        }
        if (sym > 29) return mk_entry(0, 0, OP_BAD, (uint32_t)bits);           // inftrees.c:65-68: 30,31 invalid
        return mk_entry(fmt.dist_base[sym], fmt.dist_extra[sym], OP_BASE, (uint32_t)bits);
    };

    int used = root_size;                              // next free entry (sub-tables follow the root table)
    uint32_t code = 0;                                 // canonical code, MSB-first, of the current symbol
    int idx = 0;                                       // position in work[]
    uint32_t cur_prefix = 0xffffffffu;                 // root index owning the open sub-table
    int sub_base = 0, sub_bits = 0;
    for (int len = 1; len <= max; ++len) {
        for (int k = count[len]; k > 0; --k, ++idx, ++code) {
            const int sym = work[idx];
            // bit-reverse the len-bit code: DEFLATE packs Huffman codes MSB first into an LSB-first stream
#if defined(__CUDA_ARCH__)
            const uint32_t rc = __brev(code) >> (32 - len);
#else
            uint32_t rc = 0;
            for (int b = 0; b < len; ++b) rc |= ((code >> b) & 1u) << (len - 1 - b);
#endif
            if (len <= root) {
                const uint32_t e = entry_for(sym, len);
                for (uint32_t i = rc; i < (uint32_t)root_size; i += 1u << len) table[i] = e;
            } else {
                const uint32_t prefix = rc & (uint32_t)(root_size - 1);
                if (prefix != cur_prefix) {
                    // size the sub-table: grow until the remaining codes under this prefix fit
                    // (inftrees.c:243-252), counting the codes not yet placed
                    int curr = len - root;
                    int room = 1 << curr;
                    int remaining = k;                 // codes of this length still to place, incl. this one
                    int l2 = len;
                    while (curr + root < max) {
                        room -= remaining;
                        if (room <= 0) break;
                        ++curr; room <<= 1; ++l2;
                        remaining = count[l2];
                    }
                    if (used + (1 << curr) > cap) return -1;
                    sub_base = used; sub_bits = curr; used += 1 << curr;
                    cur_prefix = prefix;
                    for (int i = 0; i < (1 << curr); ++i) table[sub_base + i] = mk_entry(0, 0, OP_BAD, 1);
                    table[prefix] = mk_entry((uint32_t)sub_base, (uint32_t)curr, OP_SUB, (uint32_t)root);
                }
                const uint32_t e = entry_for(sym, len - root);
                for (uint32_t i = rc >> root; i < (1u << sub_bits); i += 1u << (len - root)) table[sub_base + i] = e;
            }
        }
        code <<= 1;
    }
    return left > 0 ? 1 : 0;
}

// ---- events handed from lane 0 to the whole warp ---------------------------------
enum : uint32_t { EV_MATCH = 1, EV_STORED = 2, EV_DONE = 3, EV_BATCH = 4, EV_HUFF = 5, EV_TABLES = 6 };
struct InflateEvent {
    uint32_t kind;
    uint32_t len;        // match length / stored byte count
    uint32_t dist;       // match distance
    uint64_t src;        // stored: byte offset inside the member's input
    uint64_t dst;        // output position the copy starts at
};

// ---- the state machine -------------------------------------------------------------
struct InflateState {
    // input
    const uint8_t *in; uint64_t in_len; uint64_t next;     // next = bytes of input already loaded into hold
    uint64_t hold; int bits;                               // LSB-first accumulator; bits < 0 => ran past the end
    // output
    uint8_t *out; uint64_t out_cap; uint64_t pos;
    // block state
    int last; int in_block;                                // in_block: 1 = Huffman block open
    const uint32_t *lt, *dt;                               // tables of the open block
    int wrap_kind;                                         // 0 raw, 1 zlib, 2 gzip (resolved)
    int dyn_nlen, dyn_ndist;                               // symbol counts of the dynamic header just read
    int have_dict;                                         // a preset dictionary lies before the output (pos starts behind it)
    int status;
    uint64_t ck_bit, ck_out;                               // last block boundary (resume point)
    uint32_t stored_check, stored_isize;                   // trailer values
    uint64_t in_used;
    uint32_t *tlit, *tdist;                                // decode tables of a dynamic block (kLitEntries / kDistEntries)
    uint16_t *work; uint8_t *lens;                         // table-construction scratch (320 entries each)
    const uint32_t *fixed_lit, *fixed_dist;
    const FormatTables *fmt;
    InflateEvent parked; int has_parked;                   // event held back until the queue has been executed
    int huff_external;                                     // 1: hand Huffman blocks to the caller (EV_HUFF) instead of decoding them here
    // Optional log of every deflate-block boundary passed (inflate(Z_BLOCK), zlib.h:540-560): blog[0] = entries, entry k at
    // blog[2 + 2k]: bit offset inside the member, output position | (BFINAL of the block that ended there) << 63.  When the log
    // is full the member stops at that boundary as if its input ended there (the caller resumes from it).
    uint64_t *blog; uint32_t blog_cap, blog_n;
    int tables_external;                                   // 1: a dynamic header's code lengths are read here, its two decode tables
                                                           //    are built by the caller (EV_TABLES: len = nlen, dist = ndist), who
                                                           //    reports back through tables_done()
    // A CHUNK of a member decoded on its own (zb_inflate_blocks.cuh): the member's blocks from a known block start up to
    // the first later boundary that is one of the `cand_n` sorted candidate block starts in cand[] (bit offsets inside
    // the member) — the chunk stops there like a member whose input ends (ZB200_INF_TRUNCATED, ck_bit = the boundary).
    const uint64_t *cand; uint32_t cand_n; uint64_t start_bit;
    int count_only;                                        // 1: nothing is written — pos counts the output, the caller counts
                                                           //    the matches; what lies before the chunk is unknown, so no
                                                           //    distance is "too far back" yet

    ZB_HD void refill() {
        while (bits <= 32) {
            if (next >= in_len) break;
            const uint8_t *p = in + next;
            if ((((uintptr_t)p) & 3) == 0 && next + 4 <= in_len) {
                hold |= (uint64_t)(*reinterpret_cast<const uint32_t *>(p)) << bits;
                next += 4; bits += 32;
            } else {
                hold |= (uint64_t)(*p) << bits;
                next += 1; bits += 8;
            }
        }
    }
    ZB_HD uint32_t peek(int k) const { return (uint32_t)hold & ((1u << k) - 1u); }
    ZB_HD void drop(int k) { hold >>= k; bits -= k; }
    ZB_HD uint32_t take(int k) { uint32_t v = peek(k); drop(k); return v; }
    ZB_HD uint32_t take32() { uint32_t v = (uint32_t)hold; drop(32); return v; }
    ZB_HD uint64_t bitpos() const { return next * 8 - (uint64_t)(int64_t)bits; }
    // bytes really available as bits (negative bits means we consumed zero padding)
    ZB_HD bool overrun() const { return bits < 0; }

    // Rebind the working memory (the device keeps the construction scratch in memory it
    // shares with the rounds, which never run while a header is being read).
    ZB_HD void bind(uint32_t *lit, uint32_t *dist, uint16_t *w, uint8_t *l) { tlit = lit; tdist = dist; work = w; lens = l; }

    ZB_HD void init(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, InflateScratch *s,
                    const uint32_t *flit, const uint32_t *fdist, const FormatTables *f) {
        in = src; in_len = n; next = 0; hold = 0; bits = 0;
        out = dst; out_cap = cap; pos = 0; last = 0; in_block = 0; lt = dt = nullptr;
        wrap_kind = 0; status = ZB200_INF_OK; ck_bit = 0; ck_out = 0; stored_check = 0; stored_isize = 0;
        in_used = 0; fixed_lit = flit; fixed_dist = fdist; fmt = f; has_parked = 0; huff_external = 0; tables_external = 0; have_dict = 0;
        blog = nullptr; blog_cap = 0; blog_n = 0;
        cand = nullptr; cand_n = 0; start_bit = 0; count_only = 0;
        if (s) { tlit = s->lit; tdist = s->dist; work = s->work; lens = s->lens; }
        else { tlit = tdist = nullptr; work = nullptr; lens = nullptr; }
    }

    // A preset dictionary of n bytes occupies out[0..n): the member's output follows it
    // (inflate.c:1278-1312 copies it into the window instead).  Call after init().
    ZB_HD void preset(uint64_t n) { if (n) { pos = n; ck_out = n; have_dict = 1; } }

    // Continue a member at a block boundary reported by an earlier, truncated run
    // (the streaming inflate() of the host API feeds input piecewise).
    ZB_HD void resume(uint64_t bit_off, uint64_t out_pos, int kind) {
        next = bit_off >> 3; hold = 0; bits = 0;
        refill();
        drop((int)(bit_off & 7));
        pos = out_pos; wrap_kind = kind; ck_bit = bit_off; ck_out = out_pos; start_bit = bit_off;
    }

    // Is bit offset b one of the candidate block starts?  (binary search; a few probes per deflate block)
    ZB_HD bool is_candidate(uint64_t b) const {
        uint32_t lo = 0, hi = cand_n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint64_t v = cand[mid];
            if (v == b) return true;
            if (v < b) lo = mid + 1; else hi = mid;
        }
        return false;
    }

    // Re-seed the bit reader at an absolute bit offset: after an externally decoded
    // block (still_open = 0), or inside one whose remainder is decoded here.
    ZB_HD void seek(uint64_t bit_off, uint64_t out_pos, int still_open = 0) {
        next = bit_off >> 3; hold = 0; bits = 0;
        refill();
        drop((int)(bit_off & 7));
        pos = out_pos; in_block = still_open;
    }

    // Wrapper header: inflate.c:622-669 (zlib), :671-808 (gzip).  Returns status.
    ZB_HD int parse_header(int wrap) {
        if (wrap == ZB200_WRAP_RAW) { wrap_kind = 0; return ZB200_INF_OK; }
        if (in_len < 2) return ZB200_INF_TRUNCATED;
        const uint32_t h0 = in[0], h1 = in[1];
        if ((wrap & ZB200_WRAP_GZIP) && h0 == 0x1f && h1 == 0x8b) {
            wrap_kind = 2;
            if (in_len < 10) return ZB200_INF_TRUNCATED;
            if (in[2] != 8) return ZB200_INF_METHOD;
            const uint32_t flg = in[3];
            if (flg & 0xe0) return ZB200_INF_GZ_FLAGS;
            uint64_t p = 10;
            if (flg & 4) {
                if (p + 2 > in_len) return ZB200_INF_TRUNCATED;
                p += 2 + (uint64_t)(in[p] | (in[p + 1] << 8));
            }
            for (int f = 8; f <= 16; f <<= 1)
                if (flg & f) {
                    for (;;) { if (p >= in_len) return ZB200_INF_TRUNCATED; if (in[p++] == 0) break; }
                }
            if (flg & 2) {
                if (p + 2 > in_len) return ZB200_INF_TRUNCATED;
                uint32_t c = 0xffffffffu;
                for (uint64_t i = 0; i < p; ++i) c = crc_byte_bitwise(c, in[i]);
                if (((~c) & 0xffff) != (uint32_t)(in[p] | (in[p + 1] << 8))) return ZB200_INF_GZ_HCRC;
                p += 2;
            }
            if (p > in_len) return ZB200_INF_TRUNCATED;
            next = p;
            return ZB200_INF_OK;
        }
        if (!(wrap & ZB200_WRAP_ZLIB)) return ZB200_INF_HEADER_CHECK;
        if (((h0 << 8) + h1) % 31) return ZB200_INF_HEADER_CHECK;
        if ((h0 & 0xf) != 8) return ZB200_INF_METHOD;
        if ((h0 >> 4) + 8 > 15) return ZB200_INF_WINDOW;
        wrap_kind = 1; next = 2;
        if (h1 & 0x20) {                               // FDICT: inflate.c:660-669 reads the DICTID, then wants the dictionary
            if (in_len < 6) return ZB200_INF_TRUNCATED;
            stored_check = ((uint32_t)in[2] << 24) | ((uint32_t)in[3] << 16) | ((uint32_t)in[4] << 8) | (uint32_t)in[5];
            if (!have_dict) return ZB200_INF_NEED_DICT;
            next = 6;
        }
        return ZB200_INF_OK;
    }

    // Dynamic block header: inflate.c:898-1022.
    ZB_HD int read_dynamic() {
        refill();
        const int nlen = (int)take(5) + 257, ndist = (int)take(5) + 1, ncode = (int)take(4) + 4;
        if (overrun()) return ZB200_INF_TRUNCATED;
        if (nlen > 286 || ndist > 30) return ZB200_INF_TOO_MANY_SYMS;
        for (int i = 0; i < 19; ++i) lens[i] = 0;
        for (int i = 0; i < ncode; ++i) {
            refill();
            lens[fmt->cl_order[i]] = (uint8_t)take(3);
        }
        if (overrun()) return ZB200_INF_TRUNCATED;
        uint32_t *cl = tdist;                          // reuse: the distance table is built afterwards
        int r = build_decode_table(TBL_CODELEN, lens, 19, cl, kClEntries, kClRoot, work, *fmt);
        if (r < 0 || r == 1) return ZB200_INF_CODE_LENGTHS;       // inftrees.c:131: incomplete CODES set is an error
        const int total = nlen + ndist;
        int have = 0;
        // lens[] is about to be overwritten with the real code lengths
        while (have < total) {
            refill();
            uint32_t e = cl[peek(kClRoot)];
            uint32_t sym, b;
            if (ZB_E_OP(e) == OP_BAD) { sym = 0; b = 1; }          // empty code-length code: inflate.c:940-947 reads val 0, 1 bit
            else { sym = ZB_E_VAL(e); b = ZB_E_BITS(e); }
            if (sym < 16) { drop((int)b); if (overrun()) return ZB200_INF_TRUNCATED; lens[have++] = (uint8_t)sym; continue; }
            int rep; uint8_t val = 0;
            if (sym == 16) {
                if (bits < (int)b + 2) return ZB200_INF_TRUNCATED;
                drop((int)b);
                if (have == 0) return ZB200_INF_BIT_REPEAT;
                val = lens[have - 1];
                rep = 3 + (int)take(2);
            } else if (sym == 17) {
                if (bits < (int)b + 3) return ZB200_INF_TRUNCATED;
                drop((int)b); rep = 3 + (int)take(3);
            } else {
                if (bits < (int)b + 7) return ZB200_INF_TRUNCATED;
                drop((int)b); rep = 11 + (int)take(7);
            }
            if (have + rep > total) return ZB200_INF_BIT_REPEAT;
            while (rep--) lens[have++] = val;
        }
        if (lens[256] == 0) return ZB200_INF_NO_EOB;
        if (tables_external) { dyn_nlen = nlen; dyn_ndist = ndist; return ZB200_INF_OK; }
        r = build_decode_table(TBL_LITLEN, lens, nlen, tlit, kLitEntries, kLitRoot, work, *fmt);
        if (r < 0) return ZB200_INF_LITLEN_SET;
        if (r == 1) {                                  // inftrees.c:131-132: incomplete only if a single 1-bit code
            int nz = 0, ones = 0;
            for (int i = 0; i < nlen; ++i) { nz += lens[i] != 0; ones += lens[i] == 1; }
            if (!(nz == 1 && ones == 1)) return ZB200_INF_LITLEN_SET;
        }
        r = build_decode_table(TBL_DIST, lens + nlen, ndist, tdist, kDistEntries, kDistRoot, work, *fmt);
        if (r < 0) return ZB200_INF_DIST_SET;
        if (r == 1) {
            int nz = 0, ones = 0;
            for (int i = 0; i < ndist; ++i) { nz += lens[nlen + i] != 0; ones += lens[nlen + i] == 1; }
            if (!(nz == 1 && ones == 1)) return ZB200_INF_DIST_SET;
        }
        lt = tlit; dt = tdist;
        return ZB200_INF_OK;
    }

    // The caller has built the tables of the dynamic block announced by EV_TABLES (st = the
    // status the serial construction would have returned): the block opens, or the member ends.
    ZB_HD InflateEvent tables_done(int st) {
        if (st) return done(st);
        lt = tlit; dt = tdist;
        in_block = 1;
        InflateEvent ev; ev.kind = EV_HUFF; ev.len = 0; ev.dist = 0; ev.src = bitpos(); ev.dst = pos;
        return ev;
    }

    ZB_HD InflateEvent done(int st) {
        status = st;
        InflateEvent ev; ev.kind = EV_DONE; ev.len = 0; ev.dist = 0; ev.src = 0; ev.dst = pos;
        return ev;
    }

    // Trailer: inflate.c:1183-1219.  Values are recorded; the comparison with the
    // computed checksum happens in the verify kernel (zb_inflate.cu).
    ZB_HD InflateEvent finish() {
        drop(bits & 7);
        uint64_t p = next - (uint64_t)(bits >> 3);     // first unread byte
        if (wrap_kind == 2) {
            if (p + 8 > in_len) { in_used = p; return done(ZB200_INF_TRUNCATED); }
            stored_check = (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
            stored_isize = (uint32_t)in[p + 4] | ((uint32_t)in[p + 5] << 8) | ((uint32_t)in[p + 6] << 16) | ((uint32_t)in[p + 7] << 24);
            p += 8;
        } else if (wrap_kind == 1) {
            if (p + 4 > in_len) { in_used = p; return done(ZB200_INF_TRUNCATED); }
            stored_check = ((uint32_t)in[p] << 24) | ((uint32_t)in[p + 1] << 16) | ((uint32_t)in[p + 2] << 8) | (uint32_t)in[p + 3];
            p += 4;
        }
        in_used = p;
        ck_bit = p * 8; ck_out = pos;
        return done(ZB200_INF_OK);
    }

    // Run until something needs the whole warp (or the member ends).
    ZB_HD InflateEvent step() {
        for (;;) {
            if (!in_block) {
                if (blog) {                            // a block boundary (the one after the last block included)
                    if (blog_n >= blog_cap) { ck_bit = bitpos(); ck_out = pos; return done(ZB200_INF_TRUNCATED); }
                    blog[2 + 2 * blog_n] = bitpos();
                    blog[3 + 2 * blog_n] = pos | ((uint64_t)(last ? 1 : 0) << 63);
                    blog[0] = ++blog_n;
                }
                if (last) return finish();
                ck_bit = bitpos(); ck_out = pos;       // a block boundary: safe resume point
                if (cand_n && ck_bit != start_bit && is_candidate(ck_bit)) return done(ZB200_INF_TRUNCATED);   // the next chunk starts here
                refill();
                last = (int)take(1);
                const uint32_t type = take(2);         // inflate.c:827-862
                if (overrun()) return done(ZB200_INF_TRUNCATED);
                if (type == 0) {                       // stored: inflate.c:863-897
                    drop(bits & 7);
                    refill();
                    if (bits < 32) return done(ZB200_INF_TRUNCATED);
                    const uint32_t v = take32();
                    const uint32_t len = v & 0xffff;
                    if (len != ((v >> 16) ^ 0xffff)) return done(ZB200_INF_STORED_LEN);
                    const uint64_t src = next - (uint64_t)(bits >> 3);
                    if (len > in_len - src) return done(ZB200_INF_TRUNCATED);
                    if (len > out_cap - pos) return done(ZB200_INF_OUTPUT_FULL);
                    InflateEvent ev; ev.kind = EV_STORED; ev.len = len; ev.dist = 0; ev.src = src; ev.dst = pos;
                    next = src + len; hold = 0; bits = 0; pos += len;
                    if (len == 0) continue;
                    return ev;
                } else if (type == 1) {
                    lt = fixed_lit; dt = fixed_dist;
                } else if (type == 2) {
                    const int st = read_dynamic();
                    if (st) return done(st);
                    if (tables_external) {
                        InflateEvent ev; ev.kind = EV_TABLES; ev.len = (uint32_t)dyn_nlen; ev.dist = (uint32_t)dyn_ndist; ev.src = 0; ev.dst = pos;
                        return ev;
                    }
                } else {
                    return done(ZB200_INF_BLOCK_TYPE);
                }
                in_block = 1;
                if (huff_external) {                   // the warp-parallel decoder takes the symbols of this block
                    InflateEvent ev; ev.kind = EV_HUFF; ev.len = (lt == fixed_lit) ? 1u : 0u; ev.dist = 0;
                    ev.src = bitpos(); ev.dst = pos;
                    return ev;
                }
            }
            // ---- symbol loop: inffast.c:100-287 ----
            for (;;) {
                refill();
                uint32_t e = lt[peek(kLitRoot)];
                if (ZB_E_OP(e) == OP_SUB) {
                    const uint32_t sub = ZB_E_VAL(e) + (((uint32_t)(hold >> kLitRoot)) & ((1u << ZB_E_EXTRA(e)) - 1u));
                    drop(kLitRoot);
                    e = lt[sub];
                }
                drop((int)ZB_E_BITS(e));
                const uint32_t op = ZB_E_OP(e);
                if (op == OP_LIT) {
                    if (overrun()) return done(ZB200_INF_TRUNCATED);
                    if (pos >= out_cap) return done(ZB200_INF_OUTPUT_FULL);
                    if (!count_only) out[pos] = (uint8_t)ZB_E_VAL(e);
                    ++pos;
                    continue;
                }
                if (op == OP_EOB) {
                    if (overrun()) return done(ZB200_INF_TRUNCATED);
                    in_block = 0;
                    break;
                }
                if (op != OP_BASE) return done(overrun() ? ZB200_INF_TRUNCATED : ZB200_INF_LITLEN_CODE);
                const uint32_t len = ZB_E_VAL(e) + take((int)ZB_E_EXTRA(e));
                refill();
                uint32_t d = dt[peek(kDistRoot)];
                if (ZB_E_OP(d) == OP_SUB) {
                    const uint32_t sub = ZB_E_VAL(d) + (((uint32_t)(hold >> kDistRoot)) & ((1u << ZB_E_EXTRA(d)) - 1u));
                    drop(kDistRoot);
                    d = dt[sub];
                }
                drop((int)ZB_E_BITS(d));
                if (ZB_E_OP(d) != OP_BASE) return done(overrun() ? ZB200_INF_TRUNCATED : ZB200_INF_DIST_CODE);
                const uint32_t dist = ZB_E_VAL(d) + take((int)ZB_E_EXTRA(d));
                if (overrun()) return done(ZB200_INF_TRUNCATED);
                if (dist > pos && !count_only) return done(ZB200_INF_DIST_FAR);   // inffast.c:152-161
                if (len > out_cap - pos) return done(ZB200_INF_OUTPUT_FULL);
                InflateEvent ev; ev.kind = EV_MATCH; ev.len = len; ev.dist = dist; ev.src = 0; ev.dst = pos;
                pos += len;
                return ev;
            }
        }
    }

    // The lean inner loop (the role of inflate_fast, inffast.c:50-304): while at least
    // 16 input bytes and 264 output bytes remain, symbols are decoded with no
    // per-symbol bounds or overrun checks, refills are single aligned 32-bit loads,
    // literals are stored directly and matches are parked in q[].  Anything unusual
    // (invalid code, distance too far, leaving the safe zone) rewinds to the start of
    // the offending symbol and returns to the careful path (step()), which reports it.
    // Returns 0: continue on the careful path, 1: queue full, 2: end of block consumed.
    ZB_HD int fast_symbols(QueuedMatch *q, uint32_t &count, uint64_t &first_dst) {
        if (in_len < 16 || out_cap < 264 || count_only) return 0;
        uint64_t h = hold; int b = bits; uint64_t nx = next;
        // byte-align the loader to 4 so every refill is one aligned word
        while (((reinterpret_cast<uintptr_t>(in) + nx) & 3) != 0) {
            if (b > 56 || nx >= in_len) return 0;
            h |= (uint64_t)in[nx++] << b; b += 8;
        }
        const uint64_t room = out_cap - 264;
        if (pos > room || nx > in_len - 16) { hold = h; bits = b; next = nx; return 0; }
        uint64_t span = room - pos;
        if (span > 0x40000000ull) span = 0x40000000ull;
        const uint32_t lim = (uint32_t)span;            // offsets below are 32-bit, relative to `o`
        uint8_t *o = out + pos;
        const uint32_t base = (uint32_t)pos;            // low word of the absolute position (members < 4 GiB)
        const uint32_t near0 = pos < 65536 ? (uint32_t)pos : 65536u;   // for the "too far back" test
        // first parked destination relative to `o` (may be negative when parked by the careful path)
        int32_t first_rel = count ? (int32_t)((int64_t)first_dst - (int64_t)pos) : 0;
        const uint8_t *ip = in + nx, *ip_end = in + (in_len - 16);
        const uint32_t *L = lt, *D = dt;
        uint32_t rel = 0;
        int reason = 0;
        for (;;) {
            // Leaving the safe zone: the careful path takes over.  Checked before EVERY symbol, not only when a refill
            // is due: the refill in the middle of a match below is unconditional, and a run of symbols that each
            // reach it with 32 or more bits left at their start would otherwise walk past the end of the input
            // (found by tools/fuzz_inflate_slices.py: a truncated stream decoded bytes that were not there).
            if (ip > ip_end) break;
            if (b < 32) { h |= (uint64_t)(*reinterpret_cast<const uint32_t *>(ip)) << b; ip += 4; b += 32; }
            if (rel > lim) break;
            uint32_t e = L[(uint32_t)h & ((1u << kLitRoot) - 1)];
            if ((e & 0xf00u) == 0) {                     // OP_LIT == 0: a literal straight from the root table
                const int k = (int)(e & 0xffu);
                h >>= k; b -= k;
                o[rel++] = (uint8_t)(e >> 16);
                continue;
            }
            if (ZB_E_OP(e) == OP_SUB) {
                const uint32_t sub = ZB_E_VAL(e) + (((uint32_t)(h >> kLitRoot)) & ((1u << ZB_E_EXTRA(e)) - 1u));
                h >>= kLitRoot; b -= kLitRoot;
                e = L[sub];
            }
            { const int k = (int)ZB_E_BITS(e); h >>= k; b -= k; }
            const uint32_t op = ZB_E_OP(e);
            if (op == OP_LIT) { o[rel++] = (uint8_t)(e >> 16); continue; }
            if (op == OP_EOB) { in_block = 0; reason = 2; break; }
            // In the safe zone the whole symbol is present, so a bad code is final (no rewind needed).
            if (op != OP_BASE) { status = ZB200_INF_LITLEN_CODE; reason = 3; break; }
            const uint32_t x = ZB_E_EXTRA(e);
            const uint32_t len = ZB_E_VAL(e) + ((uint32_t)h & ((1u << x) - 1u));
            h >>= x; b -= (int)x;
            if (b < 32) { h |= (uint64_t)(*reinterpret_cast<const uint32_t *>(ip)) << b; ip += 4; b += 32; }   // ip <= ip_end + 4: still 12 bytes of slack
            uint32_t d = D[(uint32_t)h & ((1u << kDistRoot) - 1)];
            if (ZB_E_OP(d) == OP_SUB) {
                const uint32_t sub = ZB_E_VAL(d) + (((uint32_t)(h >> kDistRoot)) & ((1u << ZB_E_EXTRA(d)) - 1u));
                h >>= kDistRoot; b -= kDistRoot;
                d = D[sub];
            }
            { const int k = (int)ZB_E_BITS(d); h >>= k; b -= k; }
            if (ZB_E_OP(d) != OP_BASE) { status = ZB200_INF_DIST_CODE; reason = 3; break; }
            const uint32_t dx = ZB_E_EXTRA(d);
            const uint32_t dist = ZB_E_VAL(d) + ((uint32_t)h & ((1u << dx) - 1u));
            h >>= dx; b -= (int)dx;
            if (dist > rel + near0) { status = ZB200_INF_DIST_FAR; reason = 3; break; }   // inffast.c:152-161
            const bool dep = dist < len || (count && (int32_t)(rel + len - dist) > first_rel);
            if (!count) { first_rel = (int32_t)rel; first_dst = pos + rel; }
            q[count].dst = base + rel; q[count].packed = qm_pack(len, dist, dep);
            rel += len;
            if (++count == kQueue) { reason = 1; break; }
        }
        hold = h; bits = b; next = (uint64_t)(ip - in); pos += rel;
        return reason;
    }

    // Decode until the queue is full or something other than a match happens.
    // Returns EV_BATCH with ev.len parked matches in q[], or a STORED / DONE event
    // (only once the queue has been handed over).
    ZB_HD InflateEvent run_batch(QueuedMatch *q) {
        InflateEvent ev;
        if (has_parked) { has_parked = 0; return parked; }
        uint32_t count = 0;
        uint64_t first_dst = 0;
        for (;;) {
            if (in_block) {
                const int r = fast_symbols(q, count, first_dst);
                if (r == 1) break;
                if (r == 2) continue;                  // block done: step() reads the next header
                if (r == 3) {                          // a final decoding error met in the fast loop
                    ev = done(status);
                    if (!count) return ev;
                    parked = ev; has_parked = 1;
                    break;
                }
            }
            ev = step();
            if (ev.kind == EV_MATCH) {
                // dependent: overlaps itself, or may read what an earlier parked match will write
                const bool dep = ev.dist < ev.len || (count && ev.dst - ev.dist + ev.len > first_dst);
                if (!count) first_dst = ev.dst;
                q[count].dst = (uint32_t)ev.dst; q[count].packed = qm_pack(ev.len, ev.dist, dep);
                if (++count == kQueue) break;
                continue;
            }
            if (!count) return ev;
            parked = ev; has_parked = 1;
            break;
        }
        ev.kind = EV_BATCH; ev.len = count; ev.dist = 0; ev.src = 0; ev.dst = 0;
        return ev;
    }
};

}  // namespace zb
