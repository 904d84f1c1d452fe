// zb_inflate_tables.cuh — the two decode tables of a dynamic block, built by a whole warp.
//
// inftrees.c:32-299 (inflate_table) walks the symbols in code order and replicates each
// one's entry through the table: a serial loop of ~26 k instructions per block header on one
// lane, 13 % of a member's instructions and — with the rounds sharing the symbols among 32
// or 128 lanes — most of a member's latency.  Here the construction is turned round: every
// TABLE ENTRY finds its symbol.  The lengths give the canonical code (RFC 1951 3.2.2):
//   first[L]  the first code of length L (MSB first),  limit[L] = first[L] + count[L],
//   offs[L]   how many symbols have shorter codes,
// and the symbols ordered by (length, symbol) (inftrees.c:137-144 `work`) are placed by an
// ordered histogram (match.any over 32 symbols at a time).  An entry's index is the bit-
// reversed prefix of the codes it serves; the smallest L whose L-bit prefix is below
// limit[L] is the length of the code that owns it (the canonical decode of puff.c:230-260),
// prefix - first[L] + offs[L] the symbol's place.  Root prefixes above every code of at most
// `root` bits own a sub-table as wide as the longest code below them — for a complete code
// the very size inftrees.c:243-252 arrives at; sizes are summed by a warp scan in prefix
// order (the order the reference allots them), and a sub-table entry finds its prefix by
// bisection.  Entries no code reaches keep the reference's decode-time error marker
// (inftrees.c:111-119, :287-296).  Same result codes as build_decode_table (zb_inflate.cuh),
// which stays in use for the 19-symbol code-length code, the host replay and the parity test
// of this file (tests/test_gpu_inflate.py::test_warp_table_builder).
#pragma once
#include "zb_inflate.cuh"

namespace zb {

struct TableScratch {                       // shared memory, one per warp
    uint16_t first[16], limit[16], offs[16], cnt[16];
    uint16_t sub[512];                      // per long prefix, in prefix order: first entry | index bits << 12
};

__device__ __forceinline__ uint32_t tbl_entry_for(int type, uint32_t sym, uint32_t bits, const FormatTables &fmt) {
    if (type == TBL_LITLEN) {
        if (sym < 256) return mk_entry(sym, 0, OP_LIT, bits);
        if (sym == 256) return mk_entry(0, 0, OP_EOB, bits);
        if (sym > 285) return mk_entry(0, 0, OP_BAD, bits);                    // inftrees.c:57-60
        return mk_entry(fmt.len_base[sym - 257], fmt.len_extra[sym - 257], OP_BASE, bits);
    }
    if (sym > 29) return mk_entry(0, 0, OP_BAD, bits);                         // inftrees.c:65-68
    return mk_entry(fmt.dist_base[sym], fmt.dist_extra[sym], OP_BASE, bits);
}

// All 32 lanes call.  lens[0..n) (shared memory), n <= 320.  Returns (same in every lane)
// 0 complete, 1 incomplete, 2 no codes, -1 over-subscribed / out of table space.
// `single` is set when the code is exactly one symbol of one bit (inftrees.c:131-132).
template <int ROOT>
__device__ __forceinline__ int build_decode_table_warp(int type, const uint8_t *lens, int n, uint32_t *table, int cap,
                                                       uint16_t *work, TableScratch &ws, const FormatTables &fmt,
                                                       uint32_t lane, bool &single) {
    const unsigned full = 0xffffffffu;
    constexpr uint32_t root_size = 1u << ROOT;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (lane < 16) ws.cnt[lane] = 0;
    __syncwarp(full);
    // ordered histogram: rank of every symbol among those of its length
    uint32_t rk[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        rk[c] = 0;
        if (c * 32 < n) {
            const int i = c * 32 + (int)lane;
            const uint32_t l = i < n ? lens[i] : 0u;
            const uint32_t m = __match_any_sync(full, l);
            const uint32_t r = __popc(m & lt_mask);
            const uint32_t before = ws.cnt[l];
            __syncwarp(full);
            if (r == 0) ws.cnt[l] = (uint16_t)(before + __popc(m));
            __syncwarp(full);
            rk[c] = before + r;
        }
    }
    // canonical code: one lane, 15 steps
    if (lane == 0) {
        uint32_t code = 0, off = 0;
        for (int L = 1; L <= 15; ++L) {
            const uint32_t c = ws.cnt[L];
            ws.first[L] = (uint16_t)code; ws.offs[L] = (uint16_t)off;
            ws.limit[L] = (uint16_t)(code + c);        // <= 2^L: fits (an over-subscribed set is rejected below)
            code = (code + c) << 1; off += c;
        }
    }
    int max = 0, left = 1, nz = 0;
    bool over = false;
#pragma unroll
    for (int L = 1; L <= 15; ++L) {
        const int c = ws.cnt[L];
        if (c) max = L;
        nz += c;
        left = (left << 1) - c;
        if (left < 0) over = true;                     // inftrees.c:125-131
    }
    single = nz == 1 && ws.cnt[1] == 1;
    __syncwarp(full);
    if (max == 0) {
        for (uint32_t i = lane; i < root_size; i += 32) table[i] = mk_entry(0, 0, OP_BAD, 1);
        __syncwarp(full);
        return 2;
    }
    if (over) return -1;
    // symbols in (length, symbol) order
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        const int i = c * 32 + (int)lane;
        if (i < n) {
            const uint32_t l = lens[i];
            if (l) work[ws.offs[l] + rk[c]] = (uint16_t)i;
        }
    }
    // long prefixes: [p_lo, p_hi] in MSB-first order
    uint32_t p_lo = 0, np = 0, used = root_size;
    if (max > ROOT) {
        p_lo = ws.limit[ROOT];
        const uint32_t p_hi = ((uint32_t)ws.limit[max] - 1u) >> (max - ROOT);
        np = p_hi - p_lo + 1u;
        for (uint32_t base = 0; base < np; base += 32) {
            const uint32_t k = base + lane, p = p_lo + k;
            uint32_t size = 0, curr = 0;
            if (k < np) {
                int ml = ROOT + 1;
                for (int L = ROOT + 1; L <= max; ++L)
                    if (ws.cnt[L] && ((uint32_t)ws.first[L] >> (L - ROOT)) <= p) ml = L;
                curr = (uint32_t)(ml - ROOT);
                size = 1u << curr;
            }
            uint32_t inc = size;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(full, inc, d);
                if (lane >= (uint32_t)d) inc += y;
            }
            const uint32_t mybase = used + inc - size;
            used += __shfl_sync(full, inc, 31);
            if (k < np && mybase + size <= (uint32_t)cap) {
                ws.sub[k] = (uint16_t)(mybase | (curr << 12));
                table[__brev(p) >> (32 - ROOT)] = mk_entry(mybase, curr, OP_SUB, ROOT);
            }
        }
        if (used > (uint32_t)cap) return -1;
    }
    __syncwarp(full);
    // root entries
    for (uint32_t i = lane; i < root_size; i += 32) {
        const uint32_t c = __brev(i) >> (32 - ROOT);
        if (np && c >= p_lo) {
            if (c - p_lo >= np) table[i] = mk_entry(0, 0, OP_BAD, 1);
            continue;
        }
        uint32_t e = mk_entry(0, 0, OP_BAD, 1);
        const int top = max < ROOT ? max : ROOT;
        for (int L = 1; L <= top; ++L) {
            const uint32_t pre = c >> (ROOT - L);
            if (pre < ws.limit[L]) { e = tbl_entry_for(type, work[ws.offs[L] + pre - ws.first[L]], (uint32_t)L, fmt); break; }
        }
        table[i] = e;
    }
    // sub-table entries
    const uint32_t nsub = used - root_size;
    for (uint32_t t = lane; t < nsub; t += 32) {
        const uint32_t at = root_size + t;
        uint32_t lo = 0, hi = np - 1;                  // the last prefix whose sub-table starts at or before `at`
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (((uint32_t)ws.sub[mid] & 0xfffu) <= at) lo = mid; else hi = mid - 1;
        }
        const uint32_t sb = ws.sub[lo], base = sb & 0xfffu, curr = sb >> 12;
        const uint32_t j = at - base, p = p_lo + lo;
        const uint32_t W = ROOT + curr;
        const uint32_t c = (p << curr) | (__brev(j) >> (32 - curr));
        uint32_t e = mk_entry(0, 0, OP_BAD, 1);
        for (uint32_t L = ROOT + 1; L <= W; ++L) {
            const uint32_t pre = c >> (W - L);
            if (pre < ws.limit[L]) { e = tbl_entry_for(type, work[ws.offs[L] + pre - ws.first[L]], L - ROOT, fmt); break; }
        }
        table[at] = e;
    }
    __syncwarp(full);
    return left > 0 ? 1 : 0;
}

// Both tables of a dynamic block (the second half of read_dynamic, zb_inflate.cuh), with the
// reference's acceptance rules (inflate.c:1002-1018, inftrees.c:131-132).  Returns the status.
__device__ __forceinline__ int build_dynamic_tables_warp(const uint8_t *lens, int nlen, int ndist, uint32_t *tlit, uint32_t *tdist,
                                                         uint16_t *work, TableScratch &ws, const FormatTables &fmt, uint32_t lane) {
    bool single;
    int r = build_decode_table_warp<kLitRoot>(TBL_LITLEN, lens, nlen, tlit, kLitEntries, work, ws, fmt, lane, single);
    if (r < 0 || (r == 1 && !single)) return ZB200_INF_LITLEN_SET;
    __syncwarp(0xffffffffu);
    r = build_decode_table_warp<kDistRoot>(TBL_DIST, lens + nlen, ndist, tdist, kDistEntries, work, ws, fmt, lane, single);
    if (r < 0 || (r == 1 && !single)) return ZB200_INF_DIST_SET;
    return ZB200_INF_OK;
}

}  // namespace zb
