// zb_inflate.cu — one warp per member inflate (north_star: "inffast.c/inflate.c
// as one-warp-per-member inflate_fast") plus trailer verification.
//
// Kernel shape: CTAs of kInfWarps warps; every warp pulls member indices from a
// global counter (members are 64 KiB..1 MiB, so static striping would leave
// the tail unbalanced).  Per warp: 6.7 KiB of shared memory holds the decode
// tables of the open block (zb_inflate.cuh InflateScratch).  Lane 0 runs the
// state machine; match / stored copies are executed by all 32 lanes.
//
// Algorithmic bytes per member: C_i read + U_i written; the LZ
// back-references re-read recently written output from L1/L2.
#include "zb_internal.h"
#include "zb_inflate.cuh"
#include "zb_inflate_round.cuh"
#include "zb_inflate_tables.cuh"
#include "zb_inflate_blocks.cuh"
#include <string.h>
#include <stdio.h>
#include <algorithm>
#include <vector>
#include <stdlib.h>

namespace zb {

constexpr int kInfWarps = 4;
#ifndef ZB_INF_CTAS_PER_SM
#define ZB_INF_CTAS_PER_SM 4
#endif
constexpr int kInfCtasPerSm = ZB_INF_CTAS_PER_SM;          // 4 CTAs x (4 warps x 12.6 KiB + 3.4 KiB) of shared memory, 128 registers: 16 warps per SM
constexpr size_t kMaxGridWarps = 4096;    // upper bound on resident warps (sizes the per-warp match queues: 48 KiB each)
// Two symbol decoders exist: the serial loop on lane 0 (InflateState::fast_symbols,
// ~79 warp instructions per symbol with one lane active) and the warp-parallel rounds
// (huff_rounds_warp, all lanes active).  The rounds are the default; the serial loop
// finishes member tails and reports errors.
constexpr bool kWarpParallelHuffman = true;

// What a launch of the decode kernels does with the members it is given.
//   INF_MEMBER  members decoded into their output (the batch path: one warp / team per member)
//   INF_COUNT   CHUNKS of one member (zb_inflate_blocks.cuh): every chunk's output bytes and matches are counted,
//               nothing is written — the chunk's place in the output is not known yet
//   INF_LIST    chunks again, each at its place: literals are stored, matches are appended to the member's match
//               list instead of being copied (what lies before a chunk may not be there yet)
enum : int { INF_MEMBER = 0, INF_COUNT = 1, INF_LIST = 2 };
struct ChunkArgs {
    const uint64_t *cand; uint32_t n_cand;     // sorted candidate block starts (bit offsets inside the member)
    const uint64_t *mbase;                     // INF_LIST: per chunk, index of its first entry in mlist
    QueuedMatch *mlist;                        // INF_LIST: the matches, chunk after chunk, in stream order inside a chunk
    uint64_t *nmatch;                          // INF_COUNT: per chunk, matches counted
};

struct InflateDeviceTables {
    uint32_t fixed_lit[512];
    uint32_t fixed_dist[64];
    FormatTables fmt;
};

// Per-warp shared memory: the two decode tables of the open dynamic block, and an area
// that serves the table construction + the serial path's queue while lane 0 reads a
// header, and the rounds while the block's symbols are decoded (never both at once).
struct InflateWarpShared {
    uint32_t lit[kLitEntries];
    uint32_t dist[kDistEntries];
    union {
        struct { uint16_t work[320]; uint8_t lens[320]; QueuedMatch q[kQueue]; TableScratch tb; } serial;
        RoundShared rnd;
    };
};
struct InflateShared {
    InflateWarpShared w[kInfWarps];
    uint32_t fixed_lit[512];
    uint32_t fixed_dist[64];
    FormatTables fmt;
};



// Execute `count` parked matches: independent ones one per lane (their L2 round
// trips overlap), then the dependent ones in order, striped over the warp.
__device__ __forceinline__ void exec_queue(uint8_t *dst, const QueuedMatch *q, uint32_t count) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = 0; base < count; base += 32) {
        const bool mine = base + lane < count;
        QueuedMatch e;
        e.dst = 0; e.packed = 0;
        if (mine) e = q[base + lane];
        if (mine && !qm_dep(e.packed)) {
            const uint32_t len = qm_len(e.packed);
            uint8_t *d = dst + e.dst;
            const uint8_t *s = d - qm_dist(e.packed);          // dist >= len: source and destination do not overlap
            uint32_t i = 0;
            for (; i + 8 <= len; i += 8) {
                uint8_t t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) t[k] = s[i + k];
#pragma unroll
                for (int k = 0; k < 8; ++k) d[i + k] = t[k];
            }
            for (; i < len; ++i) d[i] = s[i];
        }
    }
    __syncwarp(full);
    for (uint32_t base = 0; base < count; base += 32) {
        const bool mine = base + lane < count;
        uint32_t depmask = __ballot_sync(full, mine && qm_dep(q[base + lane].packed));
        while (depmask) {
            const int j = __ffs(depmask) - 1;
            depmask &= depmask - 1;
            const QueuedMatch m = q[base + j];
            const uint32_t dist = qm_dist(m.packed), len = qm_len(m.packed);
            uint8_t *d = dst + m.dst;
            const uint8_t *s = d - dist;
            if (dist >= len) { for (uint32_t i = lane; i < len; i += 32) d[i] = s[i]; }
            else { for (uint32_t i = lane; i < len; i += 32) d[i] = s[i % dist]; }   // byte-serial semantics (inffast.c:249-260)
            __syncwarp(full);
        }
    }
}

// A stored block (inflate.c:863-897): len bytes from the member's input to its output, by the nt threads that share the
// member.  Destination words are aligned, the source is read as aligned words and funnel-shifted into place, four words
// in flight per thread (a byte-by-byte loop whose stores may alias its loads ran at one memory round trip per 32 bytes:
// 20 ms per MiB of incompressible data).
__device__ __forceinline__ void copy_stored(uint8_t *__restrict__ d, const uint8_t *__restrict__ s, uint32_t len, uint32_t tid, uint32_t nt) {
    const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(d) & 3)) & 3);
    if (len < 64 + head) { for (uint32_t i = tid; i < len; i += nt) d[i] = s[i]; return; }
    if (tid < head) d[tid] = s[tid];
    const uint32_t nw = (len - head) >> 2;
    uint32_t *__restrict__ dw = reinterpret_cast<uint32_t *>(d + head);
    const uintptr_t sa = reinterpret_cast<uintptr_t>(s + head);
    const uint32_t *__restrict__ sw = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(sa & 3) * 8;
    if (sh) {
#pragma unroll 4
        for (uint32_t w = tid; w < nw; w += nt) dw[w] = __funnelshift_r(sw[w], sw[w + 1], sh);   // (word w + 1 holds a byte of the block)
    } else {
#pragma unroll 4
        for (uint32_t w = tid; w < nw; w += nt) dw[w] = sw[w];
    }
    for (uint32_t i = head + nw * 4 + tid; i < len; i += nt) d[i] = s[i];
}

// ---- warp-parallel Huffman block decode (rounds: zb_inflate_round.cuh) ---------------
// Copy one match owned by this lane (source and destination do not overlap): all loads
// of a 16-byte step are issued before its stores, so a step costs one memory round
// trip instead of one per byte.  Written as predicated PTX (one setp + one ld/st per
// byte slot, immediate offsets): the compiler's version rebuilt a 64-bit address and
// spilled predicates for every slot (14 instructions per byte).
#define ZB_LDB(t, p, n, k) asm volatile("{ .reg .pred q; setp.gt.u32 q, %2, " #k "; @q ld.global.u8 %0, [%1+" #k "]; }" : "+r"(t) : "l"(p), "r"(n))
#define ZB_STB(t, p, n, k) asm volatile("{ .reg .pred q; setp.gt.u32 q, %2, " #k "; @q st.global.u8 [%1+" #k "], %0; }" :: "r"(t), "l"(p), "r"(n) : "memory")
__device__ __forceinline__ void copy_own(uint8_t *d, const uint8_t *s, uint32_t len) {
    for (;;) {
        uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0, t8 = 0, t9 = 0, ta = 0, tb = 0, tc = 0, td = 0, te = 0, tf = 0;
        ZB_LDB(t0, s, len, 0); ZB_LDB(t1, s, len, 1); ZB_LDB(t2, s, len, 2); ZB_LDB(t3, s, len, 3);
        ZB_LDB(t4, s, len, 4); ZB_LDB(t5, s, len, 5); ZB_LDB(t6, s, len, 6); ZB_LDB(t7, s, len, 7);
        ZB_LDB(t8, s, len, 8); ZB_LDB(t9, s, len, 9); ZB_LDB(ta, s, len, 10); ZB_LDB(tb, s, len, 11);
        ZB_LDB(tc, s, len, 12); ZB_LDB(td, s, len, 13); ZB_LDB(te, s, len, 14); ZB_LDB(tf, s, len, 15);
        ZB_STB(t0, d, len, 0); ZB_STB(t1, d, len, 1); ZB_STB(t2, d, len, 2); ZB_STB(t3, d, len, 3);
        ZB_STB(t4, d, len, 4); ZB_STB(t5, d, len, 5); ZB_STB(t6, d, len, 6); ZB_STB(t7, d, len, 7);
        ZB_STB(t8, d, len, 8); ZB_STB(t9, d, len, 9); ZB_STB(ta, d, len, 10); ZB_STB(tb, d, len, 11);
        ZB_STB(tc, d, len, 12); ZB_STB(td, d, len, 13); ZB_STB(te, d, len, 14); ZB_STB(tf, d, len, 15);
        if (len <= 16) break;
        len -= 16; s += 16; d += 16;
    }
}

// Execute the `count` matches of a round, parked in stream order, in waves of 32.  Only
// match destinations are holes (the literals are stored already), and they ascend: a match
// whose source — [dst - dist, dst - dist + min(len, dist)) — ends before the wave's first
// destination cannot depend on a wave-mate and is copied by its own lane at once (all such
// copies overlap).  The rest (2-8 %) know which earlier wave-mates' destinations their source
// touches (a bit set, built against the dependent ones only: the others are done by then)
// and go in passes: whoever has no unfinished wave-mate in its set copies, so a pass costs
// one L2 round trip however many matches it carries; chains (runs) take a pass per link.
// A lane copies its match alone when 16-byte load-then-store steps reproduce the byte-serial
// semantics of inffast.c:249-260 (dist >= len, or dist >= 16); short periods are striped
// over the warp, one match at a time.
__device__ __forceinline__ void copy_striped(uint8_t *d, uint32_t mdist, uint32_t mlen, uint32_t lane) {
    const uint8_t *s = d - mdist;
    for (uint32_t i = lane; i < mlen; i += 32) d[i] = s[i % mdist];
}
__device__ __forceinline__ void copy_ready(uint8_t *dst, bool ready, uint32_t edst, uint32_t len, uint32_t dist, uint32_t lane) {
    const unsigned full = 0xffffffffu;
    const bool simple = dist >= len || dist >= 16;
    if (ready && simple) copy_own(dst + edst, dst + edst - dist, len);
    uint32_t pm = __ballot_sync(full, ready && !simple);
    while (pm) {
        const int j = __ffs(pm) - 1;
        pm &= pm - 1;
        copy_striped(dst + __shfl_sync(full, edst, j), __shfl_sync(full, dist, j), __shfl_sync(full, len, j), lane);
    }
}
__device__ __forceinline__ void exec_round_queue(uint8_t *dst, const QueuedMatch *__restrict__ q, uint32_t count) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    QueuedMatch nxt;
    nxt.dst = 0; nxt.packed = 0;
    if (lane < count) nxt = q[lane];
    for (uint32_t base = 0; base < count; base += 32) {
        const bool mine = base + lane < count;
        const QueuedMatch e = nxt;
        if (base + 32 + lane < count) nxt = q[base + 32 + lane];   // the next wave's entries travel while this one is copied
        const uint32_t first = __shfl_sync(full, e.dst, 0);
        const uint32_t len = qm_len(e.packed), dist = qm_dist(e.packed);
        const uint32_t srcb = e.dst - dist, srce = srcb + (len < dist ? len : dist);
        const bool dep = mine && srce > first;
        copy_ready(dst, mine && !dep, e.dst, len, dist, lane);
        uint32_t undone = __ballot_sync(full, dep);
        uint32_t waits = 0;                                          // dependent wave-mates whose destination my source touches
        for (uint32_t m = undone; m; m &= m - 1) {
            const uint32_t j = (uint32_t)__ffs(m) - 1u;
            const uint32_t dj = __shfl_sync(full, e.dst, j), lj = __shfl_sync(full, len, j);
            if (dep && j < lane && dj < srce && dj + lj > srcb) waits |= 1u << j;
        }
        __syncwarp(full);
        while (undone) {
            const bool ready = ((undone >> lane) & 1u) && !(waits & undone);
            copy_ready(dst, ready, e.dst, len, dist, lane);
            undone &= ~__ballot_sync(full, ready);
            __syncwarp(full);
        }
    }
}

// Fetch a round's rows into shared memory with cp.async (4-byte copies, zero-filled
// past the end of the member's input): row i = words [W0 + i*W, W0 + i*W + W + 3).
template <int LG, int NL>
__device__ __forceinline__ void stage_fetch_lg(uint32_t *stage, const uint32_t *__restrict__ words, uint64_t W0, uint64_t nwords, uint32_t lane) {
    constexpr uint32_t W = 1u << LG, stride = W + kRowExtra, total = NL * stride;
    for (uint32_t k = lane; k < total; k += NL) {
        const uint32_t row = k / stride, col = k - row * stride;
        const uint64_t w = W0 + row * W + col;
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(stage + k);
        const uint32_t *g = words + (w < nwords ? w : nwords - 1);
        const int bytes = w < nwords ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(g), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int NL>                                       // NL threads fetch the rows of NL lanes
__device__ __forceinline__ void stage_fetch(uint32_t *stage, const uint32_t *words, uint64_t W0, uint64_t nwords, int lg, uint32_t lane) {
    switch (lg) {
        case 2: stage_fetch_lg<2, NL>(stage, words, W0, nwords, lane); break;
        case 3: stage_fetch_lg<3, NL>(stage, words, W0, nwords, lane); break;
        case 4: stage_fetch_lg<4, NL>(stage, words, W0, nwords, lane); break;
        default: stage_fetch_lg<5, NL>(stage, words, W0, nwords, lane); break;
    }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// status 0: the block's end-of-block code was consumed; 1: lane 0 continues inside
// the block on the serial path from (bitpos, pos) — end of the member's input, or a
// condition the careful path has to report.
struct HuffResult { int status; uint64_t bitpos; uint64_t pos; uint64_t nm; };   // nm: matches passed (chunk modes)

template <int MODE>
__device__ __forceinline__ HuffResult
huff_rounds_warp(const uint8_t *src, uint64_t in_len, uint8_t *dst, uint64_t out_cap, uint64_t bitpos, uint64_t pos,
                 const uint32_t *__restrict__ lt, const uint32_t *__restrict__ dt, RoundShared &rs, QueuedMatch *gq) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3;
    const uint32_t bias = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    const uint32_t *words = reinterpret_cast<const uint32_t *>(a0);
    const uint64_t total_bits = bias + in_len * 8;                  // valid bits are [bias, total_bits)
    const uint64_t nwords = (total_bits + 31) >> 5;
    HuffResult res;
    res.status = 1; res.bitpos = bitpos; res.pos = pos; res.nm = 0;
    uint64_t B = bias + bitpos;
    if (B >= total_bits) return res;
    int lg = round_pick_lg(total_bits - B), lg_cap = kRoundLgMax;
    if (lg < 0) return res;
    for (;;) {
        res.status = 1; res.bitpos = bitpos; res.pos = pos;         // where the serial path would take over
        const uint32_t S = 32u << lg;
        const uint64_t W0 = B >> 5;
        const uint32_t *stage = rs.stage;
        __syncwarp(full);                                            // every lane is done with the previous round's rows
        stage_fetch<32>(rs.stage, words, W0, nwords, lg, lane);
        stage_wait();
        __syncwarp(full);
        RoundLane r;
        round_speculate(r, lane, lg, lane ? 0u : (uint32_t)(B & 31u), stage, rs, lt, dt);
        for (;;) {                                                   // P2: until no start moves
            const uint32_t stopmask = __ballot_sync(full, r.stop != STOP_NONE);
            const uint32_t nvalid = stopmask ? (uint32_t)__ffs(stopmask) : 32u;
            const uint32_t t = __shfl_up_sync(full, r.end, 1) - S;   // lane i-1's end, in lane i's coordinates
            const bool need = lane > 0 && lane < nvalid && t != r.start;
            if (!__ballot_sync(full, need)) break;
            if (need) round_fix(r, lane, lg, t, stage, rs, lt, dt);
        }
        const uint32_t stopmask = __ballot_sync(full, r.stop != STOP_NONE);
        const uint32_t last = stopmask ? (uint32_t)__ffs(stopmask) - 1u : 31u;
        const uint32_t stop_l = __shfl_sync(full, r.stop, last);
        const uint64_t end_abs = W0 * 32 + (uint64_t)last * S + __shfl_sync(full, r.end, last);
        if (stop_l == STOP_BAD) return res;
        if (end_abs > total_bits) return res;                        // ran past the input: the serial path reports it
        const int next_lg = stop_l == STOP_NONE && end_abs < total_bits ? round_pick_lg(total_bits - end_abs) : -1;
        const bool valid = lane <= last;
        const uint32_t myout = valid ? r.out : 0u, mym = valid ? r.m : 0u;
        uint32_t inc_o = myout, inc_m = mym;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t yo = __shfl_up_sync(full, inc_o, dlt), ym = __shfl_up_sync(full, inc_m, dlt);
            if (lane >= (uint32_t)dlt) { inc_o += yo; inc_m += ym; }
        }
        const uint32_t tot_o = __shfl_sync(full, inc_o, 31), tot_m = __shfl_sync(full, inc_m, 31);
        if (MODE == INF_MEMBER && tot_m > kRoundQueueCap) {          // runs: matches of 2-3 bits.  Redo the round with shorter subsequences
            lg_cap = lg >= kRoundLgMin + 2 ? lg - 2 : kRoundLgMin;   // (at S = 128 a round holds at most 2048 matches), and keep
            lg = lg_cap;                                             // them short for the rest of the block
            continue;
        }
        int err = tot_o > out_cap - pos;
        if (MODE != INF_COUNT) {                                     // (INF_LIST: gq is the chunk's part of the match list, filled round after round)
            QueuedMatch *rq = MODE == INF_LIST ? gq + res.nm : gq;
            if (!err && valid) err = round_emit(r, lane, lg, stage, lt, dt, dst, (uint32_t)pos + inc_o - myout, rq, inc_m - mym);
        }
        if (__ballot_sync(full, err != 0)) return res;
        __syncwarp(full);                                            // literals and queue entries -> visible to every lane
        if (MODE == INF_MEMBER) exec_round_queue(dst, gq, tot_m);
        pos += tot_o;
        B = end_abs;
        bitpos = B - bias;
        res.bitpos = bitpos; res.pos = pos; res.nm += tot_m;
        if (stop_l == STOP_EOB) { res.status = 0; return res; }
        if (next_lg < 0) { res.status = 1; return res; }
        lg = next_lg < lg_cap ? next_lg : lg_cap;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kInfWarps * 32, kInfCtasPerSm)
inflate_kernel(const uint8_t *__restrict__ in, uint8_t *out, const zb200_member *__restrict__ members,
               uint32_t n_members, int wrap, zb200_member_result *__restrict__ results,
               const InflateDeviceTables *__restrict__ tabs, unsigned int *__restrict__ counter,
               const uint32_t *__restrict__ order, uint64_t *__restrict__ seg_off, uint64_t *__restrict__ seg_len,
               QueuedMatch *__restrict__ round_queues, uint64_t *blog, uint32_t blog_cap, const ChunkArgs ca) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateShared &sh = *reinterpret_cast<InflateShared *>(smem_raw);
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tabs);
        uint32_t *dst = sh.fixed_lit;
        constexpr int words = (int)(sizeof(InflateDeviceTables) / 4);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;

    for (;;) {
        unsigned int m = 0;
        if (lane == 0) { m = atomicAdd(counter, 1u); if (MODE == INF_MEMBER && m < n_members) m = order[m]; }   // largest members first
        m = __shfl_sync(full, m, 0);
        if (m >= n_members) break;
        const zb200_member mb = members[m];
        const uint8_t *src = in + mb.in_off;
        const uint64_t dict = mb.dict_len <= 32768 && mb.dict_len <= mb.out_off ? mb.dict_len : 0;
        uint8_t *dst = out + mb.out_off - dict;              // positions count from the dictionary's first byte
        const uint64_t cap = (mb.out_cap < 0xfffffff0ull - dict ? mb.out_cap : 0xfffffff0ull - dict) + dict;   // queue entries hold 32-bit output offsets
        uint64_t nm = 0;                                     // chunk modes: matches of this chunk so far
        QueuedMatch *const ml = MODE == INF_LIST ? ca.mlist + ca.mbase[m] : nullptr;

        InflateState st;
        int hs = ZB200_INF_OK;
        if (lane == 0) {
            st.init(src, mb.in_len, dst, cap, nullptr, sh.fixed_lit, sh.fixed_dist, &sh.fmt);
            st.bind(sh.w[warp].lit, sh.w[warp].dist, sh.w[warp].serial.work, sh.w[warp].serial.lens);
            st.huff_external = kWarpParallelHuffman ? 1 : 0;
            st.tables_external = 1;
            if (MODE == INF_MEMBER && blog && n_members == 1) { st.blog = blog; st.blog_cap = blog_cap; }
            if (MODE != INF_MEMBER) { st.cand = ca.cand; st.cand_n = ca.n_cand; st.count_only = MODE == INF_COUNT; }
            st.preset(dict);
            if (mb.resume_bit) st.resume(mb.resume_bit, mb.resume_out + dict, wrap);
            else { hs = st.parse_header(wrap); st.start_bit = st.bitpos(); }
            if (hs) st.status = hs;
        }
        hs = __shfl_sync(full, hs, 0);
        QueuedMatch *q = sh.w[warp].serial.q;
        while (hs == ZB200_INF_OK) {
            InflateEvent ev;
            ev.kind = EV_DONE; ev.len = 0; ev.dist = 0; ev.src = 0; ev.dst = 0;
            if (lane == 0) ev = st.run_batch(q);
            uint32_t kind = __shfl_sync(full, ev.kind, 0);
            if (kind == EV_DONE) break;
            uint32_t len = __shfl_sync(full, ev.len, 0);
            __syncwarp(full);                                  // lane 0's literal stores and queue writes -> visible
            if (kind == EV_TABLES) {                           // a dynamic header's lengths are in: the warp builds both tables
                const int ts = build_dynamic_tables_warp(sh.w[warp].serial.lens, (int)len, (int)__shfl_sync(full, ev.dist, 0),
                                                         sh.w[warp].lit, sh.w[warp].dist, sh.w[warp].serial.work,
                                                         sh.w[warp].serial.tb, sh.fmt, (uint32_t)lane);
                if (lane == 0) ev = st.tables_done(ts);
                kind = __shfl_sync(full, ev.kind, 0);
                if (kind == EV_DONE) break;
                len = 0;
            }
            if (kWarpParallelHuffman && kind == EV_HUFF) {     // a Huffman block: all 32 lanes decode it
                const uint64_t bp = __shfl_sync(full, ev.src, 0), op = __shfl_sync(full, ev.dst, 0);
                const uint32_t *lt = len ? sh.fixed_lit : sh.w[warp].lit;
                const uint32_t *dt = len ? sh.fixed_dist : sh.w[warp].dist;
                const HuffResult hr = huff_rounds_warp<MODE>(src, mb.in_len, dst, cap, bp, op, lt, dt, sh.w[warp].rnd,
                    MODE == INF_MEMBER ? round_queues + (size_t)(blockIdx.x * kInfWarps + warp) * kRoundQueueCap : ml + nm);
                nm += hr.nm;
                if (lane == 0) st.seek(hr.bitpos, hr.pos, hr.status);   // status 1: the serial path finishes the block
            } else if (kind == EV_BATCH) {
                if (MODE == INF_MEMBER) exec_queue(dst, q, len);
                if (MODE == INF_LIST) { for (uint32_t i = lane; i < len; i += 32) ml[nm + i] = q[i]; __syncwarp(full); }
                nm += len;
            } else if (MODE != INF_COUNT) {                    // stored block: input -> output
                const uint64_t to = __shfl_sync(full, ev.dst, 0);
                const uint64_t from = __shfl_sync(full, ev.src, 0);
                copy_stored(dst + to, src + from, len, lane, 32);
                __syncwarp(full);
            }
        }
        if (lane == 0) {
            zb200_member_result r;
            r.status = st.status;
            r.wrap_kind = (uint32_t)st.wrap_kind;
            r.check = st.stored_check;                         // replaced by the computed value in inflate_verify_kernel
            r.isize = st.stored_isize;
            r.out_len = st.pos - dict;
            r.in_used = st.in_used;
            r.resume_bit = st.ck_bit;
            r.resume_out = st.ck_out - dict;
            results[m] = r;
            if (MODE == INF_MEMBER) {
                seg_off[m] = mb.out_off;
                seg_len[m] = st.status == ZB200_INF_OK ? st.pos - dict : 0;
            }
            if (MODE == INF_COUNT) ca.nmatch[m] = nm;
        }
    }
}

// ---- one member across a TEAM of four warps ---------------------------------------------
// With few members in the batch a member's own latency is what counts (72 MB/s per warp: 14 ms
// per MiB).  The rounds do not care how many lanes share them: a team runs them with 128
// subsequences — the fix-up chain and the scans cross the warps through shared memory, the
// copy waves are 128 matches wide — while thread 0 remains the serial skeleton.  Tails too
// short for a team round go to warp 0's 32-lane rounds, oddities to the serial path as ever.
#ifndef ZB_TEAM_WARPS
#define ZB_TEAM_WARPS 4
#endif
constexpr int kTeamWarps = ZB_TEAM_WARPS, kTeamLanes = kTeamWarps * 32, kTeamCtasPerSm = 16 / kTeamWarps, kTeamLgDefault = 4;
struct TeamShared {
    uint32_t lit[kLitEntries];
    uint32_t dist[kDistEntries];
    union {
        struct { uint16_t work[320]; uint8_t lens[320]; QueuedMatch q[kQueue]; TableScratch tb; } serial;
        RoundSharedT<kTeamLanes> rnd;
        RoundShared rnd1;                                    // warp 0's rounds over a tail
    };
    uint32_t fixed_lit[512];
    uint32_t fixed_dist[64];
    FormatTables fmt;
    uint32_t x_end[kTeamLanes], x_stop[kTeamLanes];          // the fix-up chain across warps
    uint32_t x_wfirst[kTeamWarps], x_wsum_o[kTeamWarps], x_wsum_m[kTeamWarps], x_dep[2][kTeamWarps];
    QueuedMatch x_q[kTeamLanes];                             // the copy wave in flight
    InflateEvent ev;
    int hs;
    unsigned int member;
};

// The copy waves of exec_round_queue, 128 matches wide: the wave and the sets of unfinished
// matches live in shared memory, a pass ends at a block barrier.
__device__ __forceinline__ void exec_team_queue(uint8_t *dst, const QueuedMatch *__restrict__ q, uint32_t count, TeamShared &ts) {
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t base = 0; base < count; base += kTeamLanes) {
        const bool mine = base + tid < count;
        QueuedMatch e;
        e.dst = 0; e.packed = 0;
        if (mine) e = q[base + tid];
        const uint32_t first = q[base].dst;                  // same address in every thread: one transaction
        const uint32_t len = qm_len(e.packed), dist = qm_dist(e.packed);
        const uint32_t srcb = e.dst - dist, srce = srcb + (len < dist ? len : dist);
        const bool dep = mine && srce > first;
        ts.x_q[tid] = e;
        const uint32_t dm = __ballot_sync(full, dep);
        if (lane == 0) ts.x_dep[0][warp] = dm;
        copy_ready(dst, mine && !dep, e.dst, len, dist, lane);
        __syncthreads();                                     // independent copies done and visible, wave published
        uint32_t waits[kTeamWarps];
#pragma unroll
        for (int w = 0; w < kTeamWarps; ++w) waits[w] = 0;
        if (dep) {
#pragma unroll
            for (int w = 0; w < kTeamWarps; ++w) {
                uint32_t m = ts.x_dep[0][w];
                if ((uint32_t)w > warp) m = 0;
                else if ((uint32_t)w == warp) m &= (1u << lane) - 1u;
                for (; m; m &= m - 1) {
                    const uint32_t j = (uint32_t)__ffs(m) - 1u;
                    const QueuedMatch o = ts.x_q[w * 32 + j];
                    if (o.dst < srce && o.dst + qm_len(o.packed) > srcb) waits[w] |= 1u << j;
                }
            }
        }
        uint32_t mine_undone = dm;                           // my warp's share of the unfinished set
        for (int p = 0; __syncthreads_or(mine_undone != 0); p ^= 1) {
            uint32_t blocked = 0;
#pragma unroll
            for (int w = 0; w < kTeamWarps; ++w) blocked |= waits[w] & ts.x_dep[p][w];
            const bool ready = ((mine_undone >> lane) & 1u) && !blocked;
            copy_ready(dst, ready, e.dst, len, dist, lane);
            mine_undone &= ~__ballot_sync(full, ready);
            if (lane == 0) ts.x_dep[p ^ 1][warp] = mine_undone;
        }
    }
}

template <int MODE>
__device__ __forceinline__ HuffResult
huff_rounds_team(const uint8_t *src, uint64_t in_len, uint8_t *dst, uint64_t out_cap, uint64_t bitpos, uint64_t pos,
                 const uint32_t *__restrict__ lt, const uint32_t *__restrict__ dt, TeamShared &ts, QueuedMatch *gq, int team_lg_cap) {
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    RoundSharedT<kTeamLanes> &rs = ts.rnd;
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3;
    const uint32_t bias = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    const uint32_t *words = reinterpret_cast<const uint32_t *>(a0);
    const uint64_t total_bits = bias + in_len * 8;
    const uint64_t nwords = (total_bits + 31) >> 5;
    HuffResult res;
    res.status = 1; res.bitpos = bitpos; res.pos = pos; res.nm = 0;
    uint64_t B = bias + bitpos;
    if (B >= total_bits) return res;
    int lg = round_pick_lg(total_bits - B, kTeamLanes), lg_cap = team_lg_cap;
    if (lg < 0) return res;
    if (lg > lg_cap) lg = lg_cap;
    for (;;) {
        res.status = 1; res.bitpos = bitpos; res.pos = pos;
        const uint32_t S = 32u << lg;
        const uint64_t W0 = B >> 5;
        __syncthreads();                                             // every thread is done with the previous round's rows
        stage_fetch<kTeamLanes>(rs.stage, words, W0, nwords, lg, tid);
        stage_wait();
        __syncthreads();
        RoundLane r;
        round_speculate(r, tid, lg, tid ? 0u : (uint32_t)(B & 31u), rs.stage, rs, lt, dt);
        uint32_t firststop;
        for (;;) {                                                   // P2 across the team: until no start moves
            ts.x_end[tid] = r.end; ts.x_stop[tid] = r.stop;
            const uint32_t sm = __ballot_sync(full, r.stop != STOP_NONE);
            if (lane == 0) ts.x_wfirst[warp] = sm ? warp * 32 + (uint32_t)__ffs(sm) - 1u : (uint32_t)kTeamLanes;
            __syncthreads();
            firststop = ts.x_wfirst[0];
#pragma unroll
            for (int w = 1; w < kTeamWarps; ++w) firststop = ts.x_wfirst[w] < firststop ? ts.x_wfirst[w] : firststop;
            const uint32_t nvalid = firststop < (uint32_t)kTeamLanes ? firststop + 1u : (uint32_t)kTeamLanes;
            const uint32_t t = tid ? ts.x_end[tid - 1] - S : 0u;
            const bool need = tid > 0 && tid < nvalid && t != r.start;
            if (!__syncthreads_or(need)) break;                      // (the barrier also orders these reads before the next writes)
            if (need) round_fix(r, tid, lg, t, rs.stage, rs, lt, dt);
        }
        const uint32_t last = firststop < (uint32_t)kTeamLanes ? firststop : (uint32_t)kTeamLanes - 1u;
        const uint32_t stop_l = ts.x_stop[last];
        const uint64_t end_abs = W0 * 32 + (uint64_t)last * S + ts.x_end[last];
        if (stop_l == STOP_BAD) return res;
        if (end_abs > total_bits) return res;
        const int next_lg = stop_l == STOP_NONE && end_abs < total_bits ? round_pick_lg(total_bits - end_abs, kTeamLanes) : -1;
        const bool valid = tid <= last;
        const uint32_t myout = valid ? r.out : 0u, mym = valid ? r.m : 0u;
        uint32_t inc_o = myout, inc_m = mym;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            const uint32_t yo = __shfl_up_sync(full, inc_o, dlt), ym = __shfl_up_sync(full, inc_m, dlt);
            if (lane >= (uint32_t)dlt) { inc_o += yo; inc_m += ym; }
        }
        if (lane == 31) { ts.x_wsum_o[warp] = inc_o; ts.x_wsum_m[warp] = inc_m; }
        __syncthreads();
        uint32_t before_o = 0, before_m = 0, tot_o = 0, tot_m = 0;
#pragma unroll
        for (uint32_t w = 0; w < (uint32_t)kTeamWarps; ++w) {
            const uint32_t xo = ts.x_wsum_o[w], xm = ts.x_wsum_m[w];
            if (w < warp) { before_o += xo; before_m += xm; }
            tot_o += xo; tot_m += xm;
        }
        if (MODE == INF_MEMBER && tot_m > kRoundQueueCap * kTeamWarps) {   // runs: redo the round with shorter subsequences
            lg_cap = lg >= kRoundLgMin + 2 ? lg - 2 : kRoundLgMin;
            lg = lg_cap;
            continue;
        }
        int err = tot_o > out_cap - pos;
        if (MODE != INF_COUNT) {
            QueuedMatch *rq = MODE == INF_LIST ? gq + res.nm : gq;
            if (!err && valid) err = round_emit(r, tid, lg, rs.stage, lt, dt, dst, (uint32_t)pos + before_o + inc_o - myout, rq, before_m + inc_m - mym);
        }
        if (__syncthreads_or(err != 0)) return res;                  // (the barrier also publishes literals and queue entries)
        if (MODE == INF_MEMBER) exec_team_queue(dst, gq, tot_m, ts);
        pos += tot_o;
        B = end_abs;
        bitpos = B - bias;
        res.bitpos = bitpos; res.pos = pos; res.nm += tot_m;
        if (stop_l == STOP_EOB) { res.status = 0; return res; }
        if (next_lg < 0) { res.status = 1; return res; }
        lg = next_lg < lg_cap ? next_lg : lg_cap;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kTeamLanes, kTeamCtasPerSm)
inflate_team_kernel(const uint8_t *__restrict__ in, uint8_t *out, const zb200_member *__restrict__ members,
                    uint32_t n_members, int wrap, zb200_member_result *__restrict__ results,
                    const InflateDeviceTables *__restrict__ tabs, unsigned int *__restrict__ counter,
                    const uint32_t *__restrict__ order, uint64_t *__restrict__ seg_off, uint64_t *__restrict__ seg_len,
                    QueuedMatch *__restrict__ round_queues, int team_lg_cap, uint64_t *blog, uint32_t blog_cap, const ChunkArgs ca) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TeamShared &ts = *reinterpret_cast<TeamShared *>(smem_raw);
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(tabs);
        uint32_t *dst = ts.fixed_lit;
        constexpr int words = (int)(sizeof(InflateDeviceTables) / 4);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    QueuedMatch *gq = round_queues + (size_t)blockIdx.x * kTeamWarps * kRoundQueueCap;
    for (;;) {
        __syncthreads();
        if (tid == 0) { unsigned int m = atomicAdd(counter, 1u); ts.member = m < n_members ? (MODE == INF_MEMBER ? order[m] : m) : 0xffffffffu; }
        __syncthreads();
        const unsigned int m = ts.member;
        if (m == 0xffffffffu) break;
        const zb200_member mb = members[m];
        const uint8_t *src = in + mb.in_off;
        const uint64_t dict = mb.dict_len <= 32768 && mb.dict_len <= mb.out_off ? mb.dict_len : 0;
        uint8_t *dst = out + mb.out_off - dict;
        const uint64_t cap = (mb.out_cap < 0xfffffff0ull - dict ? mb.out_cap : 0xfffffff0ull - dict) + dict;
        uint64_t nm = 0;                                         // chunk modes: matches of this chunk so far
        QueuedMatch *const ml = MODE == INF_LIST ? ca.mlist + ca.mbase[m] : nullptr;
        InflateState st;
        if (tid == 0) {
            int hs = ZB200_INF_OK;
            st.init(src, mb.in_len, dst, cap, nullptr, ts.fixed_lit, ts.fixed_dist, &ts.fmt);
            st.bind(ts.lit, ts.dist, ts.serial.work, ts.serial.lens);
            st.huff_external = 1;
            st.tables_external = 1;
            if (MODE == INF_MEMBER && blog && n_members == 1) { st.blog = blog; st.blog_cap = blog_cap; }
            if (MODE != INF_MEMBER) { st.cand = ca.cand; st.cand_n = ca.n_cand; st.count_only = MODE == INF_COUNT; }
            st.preset(dict);
            if (mb.resume_bit) st.resume(mb.resume_bit, mb.resume_out + dict, wrap);
            else { hs = st.parse_header(wrap); st.start_bit = st.bitpos(); }
            if (hs) st.status = hs;
            ts.hs = hs;
        }
        __syncthreads();
        const int hs = ts.hs;
        while (hs == ZB200_INF_OK) {
            if (tid == 0) ts.ev = st.run_batch(ts.serial.q);
            __syncthreads();
            if (ts.ev.kind == EV_TABLES) {                         // a dynamic header's lengths are in: warp 0 builds both tables
                __syncthreads();                                   // (everybody has seen the event's kind)
                if (warp == 0) {
                    const int bs = build_dynamic_tables_warp(ts.serial.lens, (int)ts.ev.len, (int)ts.ev.dist, ts.lit, ts.dist,
                                                             ts.serial.work, ts.serial.tb, ts.fmt, tid);
                    __syncwarp(0xffffffffu);                       // every lane has read the event
                    if (tid == 0) ts.ev = st.tables_done(bs);
                }
                __syncthreads();
            }
            const InflateEvent ev = ts.ev;
            __syncthreads();                                       // everybody has the event before thread 0 writes the next
            if (ev.kind == EV_DONE) break;
            if (ev.kind == EV_HUFF) {
                const uint32_t *lt = ev.len ? ts.fixed_lit : ts.lit;
                const uint32_t *dt = ev.len ? ts.fixed_dist : ts.dist;
                HuffResult hr = huff_rounds_team<MODE>(src, mb.in_len, dst, cap, ev.src, ev.dst, lt, dt, ts, MODE == INF_MEMBER ? gq : ml + nm, team_lg_cap);
                nm += hr.nm;
                __syncthreads();
                if (hr.status == 1) {                              // a tail too short for a team round: warp 0's rounds
                    if (warp == 0) {
                        hr = huff_rounds_warp<MODE>(src, mb.in_len, dst, cap, hr.bitpos, hr.pos, lt, dt, ts.rnd1, MODE == INF_MEMBER ? gq : ml + nm);
                        if (tid == 0) ts.x_wsum_m[0] = (uint32_t)hr.nm;
                    }
                    if (MODE != INF_MEMBER) { __syncthreads(); nm += ts.x_wsum_m[0]; }
                }
                if (tid == 0) st.seek(hr.bitpos, hr.pos, hr.status);   // status 1: the serial path finishes the block
                __syncthreads();
            } else if (ev.kind == EV_BATCH) {
                if (MODE == INF_MEMBER && warp == 0) exec_queue(dst, ts.serial.q, ev.len);
                if (MODE == INF_LIST) for (uint32_t i = tid; i < ev.len; i += kTeamLanes) ml[nm + i] = ts.serial.q[i];
                nm += ev.len;
                __syncthreads();
            } else if (MODE == INF_COUNT) {                        // (a stored block: nothing is written)
            } else {                                               // stored block: input -> output
                copy_stored(dst + ev.dst, src + ev.src, ev.len, tid, kTeamLanes);
                __syncthreads();
            }
        }
        if (tid == 0) {
            zb200_member_result r;
            r.status = st.status;
            r.wrap_kind = (uint32_t)st.wrap_kind;
            r.check = st.stored_check;
            r.isize = st.stored_isize;
            r.out_len = st.pos - dict;
            r.in_used = st.in_used;
            r.resume_bit = st.ck_bit;
            r.resume_out = st.ck_out - dict;
            results[m] = r;
            if (MODE == INF_MEMBER) {
                seg_off[m] = mb.out_off;
                seg_len[m] = st.status == ZB200_INF_OK ? st.pos - dict : 0;
            }
            if (MODE == INF_COUNT) ca.nmatch[m] = nm;
        }
    }
}

// ---- self-test of the warp table builder (zb_inflate_tables.cuh) --------------------------
// One warp per case: lane 0 builds both tables of a set of code lengths the serial way
// (build_decode_table + read_dynamic's acceptance rules), the warp builds them again, and
// the two are compared entry by entry over tables pre-filled with one sentinel.
// verdict: 0 equal, 1 status differs, 2 literal/length table differs, 3 distance table differs
// (| first differing index << 8).
__global__ void __launch_bounds__(32)
tables_selftest_kernel(const uint8_t *__restrict__ lens_all, const uint32_t *__restrict__ counts, uint32_t ncases,
                       const InflateDeviceTables *__restrict__ tabs, uint32_t *__restrict__ verdict) {
    __shared__ uint32_t a_lit[kLitEntries], a_dist[kDistEntries], b_lit[kLitEntries], b_dist[kDistEntries];
    __shared__ uint16_t work[320];
    __shared__ uint8_t lens[320];
    __shared__ TableScratch tb;
    __shared__ FormatTables fmt;
    __shared__ int st_a;
    const uint32_t lane = threadIdx.x;
    const unsigned full = 0xffffffffu;
    for (uint32_t i = lane; i < sizeof(FormatTables) / 4; i += 32)
        reinterpret_cast<uint32_t *>(&fmt)[i] = reinterpret_cast<const uint32_t *>(&tabs->fmt)[i];
    for (uint32_t c = blockIdx.x; c < ncases; c += gridDim.x) {
        __syncwarp(full);
        const int nlen = (int)counts[2 * c], ndist = (int)counts[2 * c + 1];
        for (uint32_t i = lane; i < 320; i += 32) lens[i] = lens_all[(size_t)c * 320 + i];
        for (uint32_t i = lane; i < (uint32_t)kLitEntries; i += 32) a_lit[i] = b_lit[i] = 0xdeadbeefu;
        for (uint32_t i = lane; i < (uint32_t)kDistEntries; i += 32) a_dist[i] = b_dist[i] = 0xdeadbeefu;
        __syncwarp(full);
        if (lane == 0) {
            int st = ZB200_INF_OK;
            int r = build_decode_table(TBL_LITLEN, lens, nlen, a_lit, kLitEntries, kLitRoot, work, fmt);
            if (r == 1) { int nz = 0, ones = 0; for (int i = 0; i < nlen; ++i) { nz += lens[i] != 0; ones += lens[i] == 1; } if (!(nz == 1 && ones == 1)) r = -1; }
            if (r < 0) st = ZB200_INF_LITLEN_SET;
            else {
                r = build_decode_table(TBL_DIST, lens + nlen, ndist, a_dist, kDistEntries, kDistRoot, work, fmt);
                if (r == 1) { int nz = 0, ones = 0; for (int i = 0; i < ndist; ++i) { nz += lens[nlen + i] != 0; ones += lens[nlen + i] == 1; } if (!(nz == 1 && ones == 1)) r = -1; }
                if (r < 0) st = ZB200_INF_DIST_SET;
            }
            st_a = st;
        }
        __syncwarp(full);
        const int st_b = build_dynamic_tables_warp(lens, nlen, ndist, b_lit, b_dist, work, tb, fmt, lane);
        __syncwarp(full);
        uint32_t v = 0;
        if (st_a != st_b) v = 1;
        else if (st_a == ZB200_INF_OK) {
            uint32_t bad = 0xffffffffu;
            for (uint32_t i = lane; i < (uint32_t)kLitEntries; i += 32) if (a_lit[i] != b_lit[i] && i < bad) bad = i;
            for (int d = 16; d; d >>= 1) { const uint32_t o = __shfl_xor_sync(full, bad, d); bad = o < bad ? o : bad; }
            if (bad != 0xffffffffu) v = 2 | (bad << 8);
            else {
                for (uint32_t i = lane; i < (uint32_t)kDistEntries; i += 32) if (a_dist[i] != b_dist[i] && i < bad) bad = i;
                for (int d = 16; d; d >>= 1) { const uint32_t o = __shfl_xor_sync(full, bad, d); bad = o < bad ? o : bad; }
                if (bad != 0xffffffffu) v = 3 | (bad << 8);
            }
        }
        if (lane == 0) verdict[c] = v;
    }
}

// ---- members of a gzip file, discovered (gzread.c:76-234 walks them one after another) -----
// A gzip file has no index: a member's length is known once it has been decoded.  Every byte
// position that looks like a member header (1f 8b 08, no reserved flag, a known XFL) is a
// candidate; the host takes each candidate's extent to be "up to the next candidate" and its
// output size from the ISIZE field in front of the next candidate, all of them are decoded in one
// batch, and the chain is then checked from the front.  A candidate inside another member's
// data makes that member come out truncated or over-full: it is dropped and the batch redone.
__global__ void __launch_bounds__(256)
gz_candidates_kernel(const uint8_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ list, uint32_t cap, uint32_t *__restrict__ count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 10 <= n; i += stride) {
        if (in[i] != 0x1f) continue;
        if (in[i + 1] != 0x8b || in[i + 2] != 8 || (in[i + 3] & 0xe0)) continue;
        const uint32_t xfl = in[i + 8];
        if (xfl != 0 && xfl != 2 && xfl != 4) continue;         // deflate.c:1049-1052 writes nothing else
        const uint32_t k = atomicAdd(count, 1u);
        if (k < cap) list[k] = i;
    }
}

// ---- one stream, decoded in parallel at its flush points -------------------------------------
// A deflate stream written with Z_FULL_FLUSH points (deflate.c:1211-1226: byte-aligned empty
// stored block 00 00 FF FF, history reset) — everything this library's deflate() emits, pigz -i
// output, zlib's own full flushes — falls into runs of blocks that decode on their own.  Nothing
// in the stream marks those points, and a member decoded by one warp runs at 80-155 MB/s, so:
// every 00 00 FF FF is a candidate boundary, the runs between candidates are inflated as one batch
// of raw members into slots of a scratch buffer (a run's output size is not known beforehand),
// and the chain is checked from the front — a run is good when it ends exactly on the next
// candidate, at a block boundary.  A run that stops mid-block had a false successor (the pattern
// inside compressed or stored data): merged with it.  A run that reaches behind its own start
// ("invalid distance too far back") follows a sync flush, not a full one: merged into its
// predecessor.  The good runs are then packed, the check value is computed over the whole.
__global__ void __launch_bounds__(256)
flush_candidates_kernel(const uint8_t *__restrict__ in, uint64_t lo, uint64_t n, uint64_t *__restrict__ list, uint32_t cap, uint32_t *__restrict__ count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 4 <= n; i += stride) {
        if (in[i + 2] != 0xff) continue;
        if (in[i + 3] != 0xff || in[i] != 0 || in[i + 1] != 0) continue;
        const uint32_t k = atomicAdd(count, 1u);
        if (k < cap) list[k] = i + 4;                        // the next run starts behind the marker
    }
}

int inflate_stream_blocks(zb200_ctx *ctx, const uint8_t *src, size_t n, int wrap, uint8_t *out, size_t out_cap,
                          size_t *out_len, int *status, size_t *in_used, uint32_t *check, int *applicable,
                          const StreamContinuation *cont, bool uploaded, bool prefix_ok, uint32_t *end_bit, StreamOutAlt *alt);

struct GatherSeg { uint64_t src, dst, len; };
__global__ void __launch_bounds__(256)
gather_segments_kernel(const uint8_t *__restrict__ from, uint8_t *__restrict__ to, const GatherSeg *__restrict__ segs) {
    const GatherSeg g = segs[blockIdx.y];
    const uint8_t *s = from + g.src;
    uint8_t *d = to + g.dst;
    const uint64_t per = (g.len + gridDim.x - 1) / gridDim.x;
    uint64_t a = (uint64_t)blockIdx.x * per, e = a + per < g.len ? a + per : g.len;
    if (a >= e) return;
    // destination words; the source is read as aligned words and funnel-shifted into place
    const uint64_t head = (4 - ((reinterpret_cast<uintptr_t>(d) + a) & 3)) & 3;
    for (uint64_t i = a + threadIdx.x; i < a + head && i < e; i += 256) d[i] = s[i];
    const uint64_t w0 = a + head;
    if (w0 < e) {
        const uint64_t nw = (e - w0) >> 2;
        uint32_t *dw = reinterpret_cast<uint32_t *>(d + w0);
        const uintptr_t sa = reinterpret_cast<uintptr_t>(s + w0);
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(sa & 3) * 8;
        for (uint64_t w = threadIdx.x; w < nw; w += 256) {
            const uint32_t lo32 = sw[w];
            const uint32_t hi32 = sh ? sw[w + 1] : 0u;       // (never reads past the source run: sh != 0 => bytes of word w+1 belong to it)
            dw[w] = __funnelshift_r(lo32, hi32, sh);
        }
        for (uint64_t i = w0 + nw * 4 + threadIdx.x; i < e; i += 256) d[i] = s[i];
    }
}

// Members are handed out by an atomic counter.  A member is decoded serially by its
// warp, so the launch ends when the slowest member ends: hand out the largest first
// (bucketed by log2 of the compressed size; one CTA builds the order).
__global__ void __launch_bounds__(1024)
inflate_order_kernel(const zb200_member *__restrict__ members, uint32_t n, uint32_t *__restrict__ order) {
    __shared__ unsigned int hist[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[63 - __clzll((long long)(members[i].in_len | 1))], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int bkt = 63; bkt >= 0; --bkt) { const unsigned int c = hist[bkt]; hist[bkt] = run; run += c; }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        order[atomicAdd(&hist[63 - __clzll((long long)(members[i].in_len | 1))], 1u)] = i;
}

// Trailer check: inflate.c:1183-1219 ("incorrect data check" / "incorrect length check").
__global__ void inflate_verify_kernel(zb200_member_result *__restrict__ results, uint32_t n, int verify,
                                      const uint32_t *__restrict__ crc, const uint32_t *__restrict__ adler) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    zb200_member_result r = results[i];
    if (r.status != ZB200_INF_OK) return;
    const uint32_t stored = r.check;
    if (r.wrap_kind == 1) {
        r.check = adler[i];
        if (verify && stored != r.check) r.status = ZB200_INF_DATA_CHECK;
    } else {
        r.check = crc[i];
        if (verify && r.wrap_kind == 2) {
            if (stored != r.check) r.status = ZB200_INF_DATA_CHECK;
            else if (r.isize != (uint32_t)r.out_len) r.status = ZB200_INF_LENGTH_CHECK;
        }
    }
    results[i] = r;
}

int inflate_init(zb200_ctx *ctx) {
    static InflateDeviceTables h;          // host image, identical for every context
    static std::once_flag once;
    std::call_once(once, [] {
        memset(&h, 0, sizeof h);
        format_fill(h.fmt);
        uint8_t lens[288]; uint16_t work[320];
        int i = 0;                         // inflate.c:252-290 fixedtables: 8/9/7/8-bit literal/length code, 5-bit distances
        for (; i < 144; ++i) lens[i] = 8;
        for (; i < 256; ++i) lens[i] = 9;
        for (; i < 280; ++i) lens[i] = 7;
        for (; i < 288; ++i) lens[i] = 8;
        build_decode_table(TBL_LITLEN, lens, 288, h.fixed_lit, 512, kLitRoot, work, h.fmt);
        for (i = 0; i < 32; ++i) lens[i] = 5;
        build_decode_table(TBL_DIST, lens, 32, h.fixed_dist, 64, kDistRoot, work, h.fmt);
    });
    void *d = nullptr;
    ZB_CUDA(cudaMalloc(&d, sizeof h));
    ZB_CUDA(cudaMemcpy(d, &h, sizeof h, cudaMemcpyHostToDevice));
    ctx->d_inflate_tables = d;
    ZB_CUDA(cudaFuncSetAttribute(inflate_kernel<INF_MEMBER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InflateShared)));
    ZB_CUDA(cudaFuncSetAttribute(inflate_kernel<INF_COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InflateShared)));
    ZB_CUDA(cudaFuncSetAttribute(inflate_kernel<INF_LIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InflateShared)));
    ZB_CUDA(cudaFuncSetAttribute(inflate_team_kernel<INF_MEMBER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TeamShared)));
    ZB_CUDA(cudaFuncSetAttribute(inflate_team_kernel<INF_COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TeamShared)));
    ZB_CUDA(cudaFuncSetAttribute(inflate_team_kernel<INF_LIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TeamShared)));
    return ZB200_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// scratch layout for one call
struct InflateWork {
    unsigned int *counter; uint64_t *seg_off, *seg_len; uint32_t *crc, *adler, *order; CkAccum *acc; QueuedMatch *queues;
    // warps that may own a queue: one per member, or four when few members run as teams
    static size_t grid_warps(size_t n) { return 4 * n < kMaxGridWarps ? 4 * n : kMaxGridWarps; }
    static size_t bytes(size_t n) {
        return 256 + align_up(n * 8, 256) * 2 + align_up(n * 4, 256) * 3 + align_up(n * sizeof(CkAccum), 256) +
               grid_warps(n) * kRoundQueueCap * sizeof(QueuedMatch);
    }
    void carve(void *base, size_t n) {
        uint8_t *p = (uint8_t *)base;
        counter = (unsigned int *)p; p += 256;
        seg_off = (uint64_t *)p; p += align_up(n * 8, 256);
        seg_len = (uint64_t *)p; p += align_up(n * 8, 256);
        crc = (uint32_t *)p; p += align_up(n * 4, 256);
        adler = (uint32_t *)p; p += align_up(n * 4, 256);
        order = (uint32_t *)p; p += align_up(n * 4, 256);
        acc = (CkAccum *)p; p += align_up(n * sizeof(CkAccum), 256);
        queues = (QueuedMatch *)p;
    }
};

size_t inflate_work_bytes(size_t n_members) { return InflateWork::bytes(n_members); }

int inflate_launch(zb200_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, const zb200_member *d_members,
                   size_t n, int wrap, int verify, zb200_member_result *d_results, void *d_work, cudaStream_t s,
                   uint64_t *d_blog, uint32_t blog_cap) {
    InflateWork w;
    w.carve(d_work, n);
    ZB_CUDA(cudaMemsetAsync(w.counter, 0, 256, s));
    if (d_blog) ZB_CUDA(cudaMemsetAsync(d_blog, 0, 16, s));
    prof_mark(ctx, s, "inflate_order_kernel");
    inflate_order_kernel<<<1, 1024, 0, s>>>(d_members, (uint32_t)n, w.order);
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    // Few members: a team of four warps per member (a member's own latency is what counts);
    // many: one warp per member (throughput).  $ZB200_INF_TEAM = 0 / 1 forces either.
    static const int team_knob = [] { const char *e = getenv("ZB200_INF_TEAM"); return e ? atoi(e) : -1; }();
    const bool team = team_knob >= 0 ? team_knob != 0 : n <= (size_t)ctx->sm_count * kTeamCtasPerSm;   // every member a resident team
    if (team) {
        size_t ctas = n;
        const size_t cap = (size_t)ctx->sm_count * kTeamCtasPerSm;
        if (ctas > cap) ctas = cap;
        if (ctas * kTeamWarps > kMaxGridWarps) ctas = kMaxGridWarps / kTeamWarps;
        static const int lg_knob = [] { const char *e = getenv("ZB200_INF_TEAM_LG"); return e ? atoi(e) : kTeamLgDefault; }();
        const int team_lg = lg_knob < kRoundLgMin ? kRoundLgMin : lg_knob > kRoundLgMax ? kRoundLgMax : lg_knob;
        prof_mark(ctx, s, "inflate_team_kernel");
        inflate_team_kernel<INF_MEMBER><<<(unsigned)ctas, kTeamLanes, sizeof(TeamShared), s>>>(
            d_in, d_out, d_members, (uint32_t)n, wrap, d_results,
            (const InflateDeviceTables *)ctx->d_inflate_tables, w.counter, w.order, w.seg_off, w.seg_len, w.queues, team_lg, d_blog, blog_cap, ChunkArgs{});
    } else {
        size_t ctas = (n + kInfWarps - 1) / kInfWarps;
        static const int per_sm_knob = [] { const char *e = getenv("ZB200_INF_CTAS_PER_SM"); return e ? atoi(e) : 0; }();   // profiling knob
        size_t cap = (size_t)ctx->sm_count * (per_sm_knob > 0 && per_sm_knob < kInfCtasPerSm ? per_sm_knob : kInfCtasPerSm);
        if (cap * kInfWarps > kMaxGridWarps) cap = kMaxGridWarps / kInfWarps;
        if (ctas > cap) ctas = cap;
        prof_mark(ctx, s, "inflate_kernel");
        inflate_kernel<INF_MEMBER><<<(unsigned)ctas, kInfWarps * 32, sizeof(InflateShared), s>>>(
            d_in, d_out, d_members, (uint32_t)n, wrap, d_results,
            (const InflateDeviceTables *)ctx->d_inflate_tables, w.counter, w.order, w.seg_off, w.seg_len, w.queues, d_blog, blog_cap, ChunkArgs{});
    }
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    const int which = wrap == ZB200_WRAP_ZLIB ? ZB200_ADLER32 : wrap == ZB200_WRAP_AUTO ? (ZB200_CRC32 | ZB200_ADLER32) : ZB200_CRC32;
    int r = checksum_launch(ctx, d_out, w.seg_off, w.seg_len, 0, n, which, 0, 1, w.crc, w.adler, w.acc, s);
    if (r) return r;
    prof_mark(ctx, s, "inflate_verify_kernel");
    inflate_verify_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_results, (uint32_t)n, verify, w.crc, w.adler);
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    prof_mark(ctx, s, nullptr);
    return ZB200_OK;
}

// One raw / zlib / gzip stream (host pointers), decoded run by run in parallel (see flush_candidates_kernel).
// *applicable = 0: the stream offers nothing to split on (no candidate, a preset dictionary, a header that does
// not parse) and nothing has been written — the caller takes the one-member path.  Otherwise *status is the
// ZB200_INF_* outcome, *out_len the valid output bytes (also when they exceed out_cap: ZB200_INF_OUTPUT_FULL and
// nothing copied), *in_used the stream's length, *check its computed check value.
int inflate_stream_parallel(zb200_ctx *ctx, const uint8_t *src, size_t n, int wrap, uint8_t *out, size_t out_cap,
                            size_t *out_len, int *status, size_t *in_used, uint32_t *check, int *applicable,
                            const StreamContinuation *cont, int blocks_mode, uint32_t *end_bit, StreamOutAlt *alt) {
    *applicable = 0; *out_len = 0; *status = ZB200_INF_OK; *in_used = 0; *check = 0;
    if (end_bit) *end_bit = 0;
    if (n < 64) return ZB200_OK;
    // nothing (or too little) to split on at flush points: the stream's dynamic block headers are the next thing to try
    auto blocks = [&](bool uploaded) {
        if (!blocks_mode) return (int)ZB200_OK;
        return inflate_stream_blocks(ctx, src, n, wrap, out, out_cap, out_len, status, in_used, check, applicable, cont, uploaded,
                                     blocks_mode == 2, end_bit, alt);
    };
    const size_t hist_len = cont ? cont->hist_len : 0;
    if (cont && (wrap != ZB200_WRAP_RAW || hist_len > 32768 || cont->bit0 > 7)) return ZB200_ERR_PARAM;
    InflateState hs;
    hs.init(src, n, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    if (hs.parse_header(wrap) != ZB200_INF_OK) return ZB200_OK;
    const uint64_t hdr = hs.next;
    const int kind = hs.wrap_kind;
    cudaStream_t s = ctx->stream;
    int r = ensure_io(ctx, n + 16, 16);
    if (r) return r;
    if ((r = h2d_auto(ctx, ctx->d_io_in, src, n, s))) return r;
    const uint32_t cand_cap = (uint32_t)(n / 5 + 16 < (1u << 22) ? n / 5 + 16 : (1u << 22));
    if ((r = ensure_scratch(ctx, 256 + (size_t)cand_cap * 8))) return r;
    uint32_t *d_count = (uint32_t *)ctx->d_scratch;
    uint64_t *d_list = (uint64_t *)((uint8_t *)ctx->d_scratch + 256);
    ZB_CUDA(cudaMemsetAsync(d_count, 0, 256, s));
    prof_mark(ctx, s, "flush_candidates_kernel");
    flush_candidates_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(ctx->d_io_in, hdr, n, d_list, cand_cap, d_count);
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_count, 8, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint32_t nc = *(const uint32_t *)ctx->h_small;
    if (nc > cand_cap) return ZB200_OK;
    if ((uint64_t)(nc + 1) * ((uint64_t)1 << 20) < n) return blocks(true);   // runs of more than 1 MiB on average: too few to fill the GPU
    std::vector<uint64_t> start(nc + 1);
    ZB_CUDA(cudaMemcpy(start.data() + 1, d_list, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    start[0] = hdr;
    std::sort(start.begin() + 1, start.end());
    while (!start.empty() && start.back() >= n) start.pop_back();          // a marker at the very end starts nothing
    if (start.size() < 2) return blocks(true);
    *applicable = 1;                                         // (2 once the deflate data is seen to end: BFINAL reached)
    struct Run { uint64_t in_off, in_len, cap, slot; zb200_member_result res; bool done; };   // done: decoded into its slot, res valid
    std::vector<Run> runs(start.size());
    for (size_t k = 0; k < runs.size(); ++k) {
        runs[k].in_off = start[k];
        runs[k].in_len = (k + 1 < runs.size() ? start[k + 1] : n) - start[k];
        runs[k].cap = 0; runs[k].done = false;
    }
    auto first_cap = [](uint64_t len) { const uint64_t a = len * 8 + 65536, b = len * 1032 + 65536; return a < b ? a : b; };
    std::vector<zb200_member> tab;
    std::vector<uint32_t> which;
    size_t good = 0;                                         // runs [0, good) are verified
    int final_status = -1;                                   // set when the chain ends (stream end or an error)
    uint64_t total = 0;                                      // scratch bytes handed out so far
    for (int pass = 0; final_status < 0; ++pass) {
        if (pass > 24) { *applicable = 0; return ZB200_OK; } // a damaged or pattern-ridden stream: the one-member path reports it
        // a run that has to be (re)decoded gets a fresh slot behind all others: the bytes of the rest stay where they are
        for (size_t k = 0; k < runs.size(); ++k) {
            Run &u = runs[k];
            if (u.done) continue;
            if (!u.cap) u.cap = first_cap(u.in_len);
            u.slot = total;
            total += (u.cap + (k == 0 ? hist_len : 0) + 15) & ~(uint64_t)15;
        }
        // every candidate costs a slot with 64 KiB of slack: a stream riddled with false markers (stored or binary data)
        // would ask for far more memory than its output can be — leave it to the one-member path
        if (total > (uint64_t)n * 16 + 2 * (uint64_t)out_cap + ((uint64_t)256 << 20)) { *applicable = 0; return ZB200_OK; }
        const size_t before = ctx->io_out_bytes;
        if ((r = ensure_io(ctx, n + 16, total + 16))) return r;
        if (ctx->io_out_bytes != before && pass) {           // the buffer was replaced: every run's bytes are gone, lay them out afresh
            total = 0;
            for (size_t k = 0; k < runs.size(); ++k) { Run &u = runs[k]; u.done = false; u.slot = total; total += (u.cap + (k == 0 ? hist_len : 0) + 15) & ~(uint64_t)15; }
            if ((r = ensure_io(ctx, n + 16, total + 16))) return r;
        }
        tab.clear(); which.clear();
        for (size_t k = 0; k < runs.size(); ++k) {
            const Run &u = runs[k];
            if (u.done) continue;
            zb200_member m;
            m.in_off = u.in_off; m.in_len = u.in_len; m.out_off = u.slot; m.out_cap = u.cap; m.resume_bit = m.resume_out = m.dict_len = 0;
            if (k == 0 && cont) {                            // a continued stream: what came before lies ahead of run 0's output
                m.out_off = u.slot + hist_len; m.dict_len = hist_len; m.resume_bit = cont->bit0;
                if (hist_len) ZB_CUDA(cudaMemcpyAsync(ctx->d_io_out + u.slot, cont->hist, hist_len, cudaMemcpyHostToDevice, s));
            }
            tab.push_back(m); which.push_back((uint32_t)k);
        }
        const size_t m = tab.size();
        if (m) {
            const size_t tbl = align_up(m * sizeof(zb200_member), 256), rsl = align_up(m * sizeof(zb200_member_result), 256);
            if ((r = ensure_scratch(ctx, tbl + rsl + InflateWork::bytes(m)))) return r;
            uint8_t *base = (uint8_t *)ctx->d_scratch;
            zb200_member *d_members = (zb200_member *)base;
            zb200_member_result *d_results = (zb200_member_result *)(base + tbl);
            ZB_CUDA(cudaMemcpyAsync(d_members, tab.data(), m * sizeof(zb200_member), cudaMemcpyHostToDevice, s));
            if ((r = inflate_launch(ctx, ctx->d_io_in, ctx->d_io_out, d_members, m, ZB200_WRAP_RAW, 0, d_results, base + tbl + rsl, s))) return r;
            std::vector<zb200_member_result> res(m);
            ZB_CUDA(cudaMemcpyAsync(res.data(), d_results, m * sizeof(zb200_member_result), cudaMemcpyDeviceToHost, s));
            ZB_CUDA(cudaStreamSynchronize(s));
            for (size_t j = 0; j < m; ++j) { runs[which[j]].res = res[j]; runs[which[j]].done = true; }
        }
        // the verified prefix grows over runs that end exactly on the next candidate
        auto clean = [](const Run &u) { return u.res.status == ZB200_INF_TRUNCATED && u.res.resume_bit == u.in_len * 8 && u.res.resume_out == u.res.out_len; };
        while (good < runs.size() && clean(runs[good])) ++good;
        if (good == runs.size()) { final_status = ZB200_INF_TRUNCATED; break; }                    // the input ends on a boundary
        if (runs[good].res.status == ZB200_INF_OK) { ++good; final_status = ZB200_INF_OK; break; } // BFINAL: the stream ends here
        // Many runs that reach behind their start: the markers are SYNC flushes (pigz, deflate(Z_SYNC_FLUSH), this library's
        // ZB200_CHUNK_CARRY) — merging would leave one long run for one team of warps (242 MB of level-1 text: 4.2 s).  Such a
        // stream is one run of blocks: its block headers are what to split on.  If that path declines, the merges go ahead.
        if (pass == 0 && blocks_mode) {
            size_t far = 0;
            for (size_t k = good; k < runs.size(); ++k) far += runs[k].res.status == ZB200_INF_DIST_FAR;
            if (far >= 2 && far * 8 >= runs.size() - good) {
                const int rb = blocks(true);
                if (rb != ZB200_OK || *applicable) return rb;
                *applicable = 1;                             // (its buffers are gone: every run is decoded again, then repaired as below)
                for (size_t k = 0; k < runs.size(); ++k) { runs[k].done = false; }
                total = 0; good = 0;
                continue;
            }
        }
        // repairs, for every run from there on (a merge never harms: it only removes a split)
        bool changed = false;
        for (size_t k = good; k < runs.size(); ++k) {
            Run &u = runs[k];
            const int st = u.res.status;
            const uint64_t most = u.in_len * 1032 + 65536;
            if (st < 0 || st == ZB200_INF_OK || clean(u)) continue;                                // (st < 0: merged in this pass, decoded in the next)
            if (st == ZB200_INF_TRUNCATED && k + 1 == runs.size()) continue;                       // the input ends inside the last run
            if (st == ZB200_INF_OUTPUT_FULL && u.cap < most) { u.cap = u.cap * 8 < most ? u.cap * 8 : most; u.done = false; changed = true; }
            else if (st == ZB200_INF_TRUNCATED && k + 1 < runs.size()) {                           // stopped mid-block: the successor was none
                u.in_len += runs[k + 1].in_len; u.cap = 0; u.res.status = -1; u.done = false;
                runs.erase(runs.begin() + (long)k + 1);
                changed = true;
            } else if (k > good || (st == ZB200_INF_DIST_FAR && k > 0)) {                          // needs what came before (a sync flush point), or started on a false boundary
                Run &pv = runs[k - 1];
                pv.in_len += u.in_len; pv.cap = 0; pv.res.status = -1; pv.done = false;
                runs.erase(runs.begin() + (long)k);
                if (k == good) --good;
                --k;
                changed = true;
            }
        }
        if (!changed) { final_status = runs[good].res.status; break; }                             // the stream's own error (or its end of input)
    }
    // pack the verified runs
    std::vector<GatherSeg> gs(good);
    uint64_t produced = 0;
    for (size_t k = 0; k < good; ++k) { gs[k].src = runs[k].slot + (k == 0 ? hist_len : 0); gs[k].dst = produced; gs[k].len = runs[k].res.out_len; produced += runs[k].res.out_len; }
    *out_len = (size_t)produced;
    *status = final_status;
    if (final_status == ZB200_INF_OK) *applicable = 2;
    uint64_t end = good ? runs[good - 1].in_off + (final_status == ZB200_INF_OK ? runs[good - 1].res.in_used : runs[good - 1].in_len) : hdr;
    uint32_t stored = 0, isize = 0;
    if (final_status == ZB200_INF_OK) {                      // trailer: inflate.c:1183-1219
        if (kind == 2) {
            if (end + 8 > n) *status = ZB200_INF_TRUNCATED;
            else { stored = (uint32_t)src[end] | ((uint32_t)src[end + 1] << 8) | ((uint32_t)src[end + 2] << 16) | ((uint32_t)src[end + 3] << 24);
                   isize = (uint32_t)src[end + 4] | ((uint32_t)src[end + 5] << 8) | ((uint32_t)src[end + 6] << 16) | ((uint32_t)src[end + 7] << 24); end += 8; }
        } else if (kind == 1) {
            if (end + 4 > n) *status = ZB200_INF_TRUNCATED;
            else { stored = ((uint32_t)src[end] << 24) | ((uint32_t)src[end + 1] << 16) | ((uint32_t)src[end + 2] << 8) | (uint32_t)src[end + 3]; end += 4; }
        }
    }
    *in_used = (size_t)end;
    if (produced > out_cap) {                                // the caller's second chance: a buffer it grows to the size needed
        uint8_t *p = alt ? alt->grow(alt->self, (size_t)produced) : nullptr;
        if (!p) { *status = ZB200_INF_OUTPUT_FULL; return ZB200_OK; }
        out = p; out_cap = (size_t)produced; alt->used = true;
    }
    const size_t gtab = align_up(good * sizeof(GatherSeg) + 16, 256);
    if ((r = ensure_scratch(ctx, gtab + 1024 + produced + 16))) return r;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    GatherSeg *d_gs = (GatherSeg *)base;
    uint32_t *d_sum = (uint32_t *)(base + gtab);
    CkAccum *d_acc = (CkAccum *)(base + gtab + 256);
    uint8_t *d_final = base + gtab + 1024;
    if (good) {
        ZB_CUDA(cudaMemcpyAsync(d_gs, gs.data(), good * sizeof(GatherSeg), cudaMemcpyHostToDevice, s));
        prof_mark(ctx, s, "gather_segments_kernel");
        gather_segments_kernel<<<dim3(16, (unsigned)good), 256, 0, s>>>(ctx->d_io_out, d_final, d_gs);
        ZB_LAUNCHED();
        ZB_CHECK_LAUNCH();
    }
    const int ck = cont ? cont->check_kind : kind;           // a continued stream: the caller says which check its wrapper wants
    if ((r = checksum_launch(ctx, d_final, nullptr, nullptr, produced, 1, ck == 1 ? ZB200_ADLER32 : ZB200_CRC32, 0, 1, d_sum, d_sum + 1, d_acc, s))) return r;
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_sum, 8, cudaMemcpyDeviceToHost, s));
    if (produced && (r = d2h_auto(ctx, out, d_final, (size_t)produced, s))) return r;
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint32_t *hsum = (const uint32_t *)ctx->h_small;
    *check = ck == 1 ? hsum[1] : hsum[0];
    if (*status == ZB200_INF_OK) {
        if (kind != 0 && stored != *check) *status = ZB200_INF_DATA_CHECK;
        else if (kind == 2 && isize != (uint32_t)produced) *status = ZB200_INF_LENGTH_CHECK;
    }
    return ZB200_OK;
}

// ---- one member, decoded chunk by chunk in parallel (zb_inflate_blocks.cuh) -------------------------------------
// Launch the decode kernels over chunks of ONE member (mode INF_COUNT / INF_LIST): a team of four warps per chunk while
// every chunk can have one, a warp per chunk beyond.  `kind`: the member's resolved wrapper (0 raw, 1 zlib, 2 gzip).
static int inflate_chunks_launch(zb200_ctx *ctx, int mode, const uint8_t *d_in, uint8_t *d_out, const zb200_member *d_members, size_t n,
                                 int kind, zb200_member_result *d_results, unsigned int *d_counter, const ChunkArgs &ca, cudaStream_t s) {
    ZB_CUDA(cudaMemsetAsync(d_counter, 0, 256, s));
    const InflateDeviceTables *tabs = (const InflateDeviceTables *)ctx->d_inflate_tables;
    static const int team_knob = [] { const char *e = getenv("ZB200_INF_TEAM"); return e ? atoi(e) : -1; }();
    const bool team = team_knob >= 0 ? team_knob != 0 : n <= (size_t)ctx->sm_count * kTeamCtasPerSm;
    if (team) {
        size_t ctas = n;
        const size_t cap = (size_t)ctx->sm_count * kTeamCtasPerSm;
        if (ctas > cap) ctas = cap;
        prof_mark(ctx, s, mode == INF_COUNT ? "inflate_team_count_kernel" : "inflate_team_list_kernel");
        if (mode == INF_COUNT)
            inflate_team_kernel<INF_COUNT><<<(unsigned)ctas, kTeamLanes, sizeof(TeamShared), s>>>(
                d_in, d_out, d_members, (uint32_t)n, kind, d_results, tabs, d_counter, nullptr, nullptr, nullptr, nullptr, kTeamLgDefault, nullptr, 0, ca);
        else
            inflate_team_kernel<INF_LIST><<<(unsigned)ctas, kTeamLanes, sizeof(TeamShared), s>>>(
                d_in, d_out, d_members, (uint32_t)n, kind, d_results, tabs, d_counter, nullptr, nullptr, nullptr, nullptr, kTeamLgDefault, nullptr, 0, ca);
    } else {
        size_t ctas = (n + kInfWarps - 1) / kInfWarps;
        const size_t cap = (size_t)ctx->sm_count * kInfCtasPerSm;
        if (ctas > cap) ctas = cap;
        prof_mark(ctx, s, mode == INF_COUNT ? "inflate_count_kernel" : "inflate_list_kernel");
        if (mode == INF_COUNT)
            inflate_kernel<INF_COUNT><<<(unsigned)ctas, kInfWarps * 32, sizeof(InflateShared), s>>>(
                d_in, d_out, d_members, (uint32_t)n, kind, d_results, tabs, d_counter, nullptr, nullptr, nullptr, nullptr, nullptr, 0, ca);
        else
            inflate_kernel<INF_LIST><<<(unsigned)ctas, kInfWarps * 32, sizeof(InflateShared), s>>>(
                d_in, d_out, d_members, (uint32_t)n, kind, d_results, tabs, d_counter, nullptr, nullptr, nullptr, nullptr, nullptr, 0, ca);
    }
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    return ZB200_OK;
}

// One raw / zlib / gzip stream (host pointers, or already on the device in ctx->d_io_in when `uploaded`), decoded in
// parallel at its dynamic block headers.  Same contract as inflate_stream_parallel; in addition *end_bit receives the
// bit (0..7) inside src[*in_used] at which a prefix delivered with ZB200_INF_TRUNCATED ends (a block boundary).
// Streams this has nothing to offer to (fewer than two chunks, a damaged chain, output of 4 GiB or more) come back
// with *applicable = 0 and nothing written: the one-member path decodes them and reports their exact status.
constexpr uint64_t kBlkGroupOut = (uint64_t)256 << 20;       // output bytes resolved together (4 bytes of scratch per byte)
int inflate_stream_blocks(zb200_ctx *ctx, const uint8_t *src, size_t n, int wrap, uint8_t *out, size_t out_cap,
                          size_t *out_len, int *status, size_t *in_used, uint32_t *check, int *applicable,
                          const StreamContinuation *cont, bool uploaded, bool prefix_ok, uint32_t *end_bit, StreamOutAlt *alt) {
    *applicable = 0; *out_len = 0; *status = ZB200_INF_OK; *in_used = 0; *check = 0;
    if (end_bit) *end_bit = 0;
    static const int knob = [] { const char *e = getenv("ZB200_INF_BLOCKS"); return e ? atoi(e) : 1; }();
    if (!knob || n < 4096) return ZB200_OK;
    const size_t hist_len = cont ? cont->hist_len : 0;
    if (cont && (wrap != ZB200_WRAP_RAW || hist_len > 32768 || cont->bit0 > 7)) return ZB200_ERR_PARAM;
    InflateState hs;
    hs.init(src, n, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    if (hs.parse_header(wrap) != ZB200_INF_OK) return ZB200_OK;
    const uint64_t hdr = hs.next;
    const int kind = hs.wrap_kind;
    const uint64_t bit0 = hdr * 8 + (cont ? cont->bit0 : 0);  // the member's first block
    cudaStream_t s = ctx->stream;
    int r = ensure_io(ctx, n + 16, 16);
    if (r) return r;
    if (!uploaded && (r = h2d_auto(ctx, ctx->d_io_in, src, n, s))) return r;
    // 1. candidates: positions that pass the cheap part of the test (one in ~1100), then those of them that pass all of it
    const uint32_t cand_cap = (uint32_t)(n / 64 + 1024 < (1u << 24) ? n / 64 + 1024 : (1u << 24));
    const uint32_t surv_cap = (uint32_t)(n / 16 + 4096 < (1u << 26) ? n / 16 + 4096 : (1u << 26));
    if ((r = ensure_scratch(ctx, 512 + (size_t)cand_cap * 8 + (size_t)surv_cap * 8))) return r;
    uint32_t *d_count = (uint32_t *)ctx->d_scratch;           // [0] candidates, [1] survivors of stage 1
    uint64_t *d_list = (uint64_t *)((uint8_t *)ctx->d_scratch + 512);
    uint64_t *d_surv = d_list + cand_cap;
    ZB_CUDA(cudaMemsetAsync(d_count, 0, 512, s));
    {
        const uint64_t nwords = ((uint64_t)n + 3) / 4;       // (the buffer holds n + 16 bytes; what follows the stream is never taken for part of it)
        const uint64_t want = (nwords + 255) / 256;
        const unsigned grid = (unsigned)(want < (uint64_t)ctx->sm_count * 16 ? (want ? want : 1) : (uint64_t)ctx->sm_count * 16);
        prof_mark(ctx, s, "blk_scan_kernel");
        blk_scan_kernel<<<grid, 256, 0, s>>>((const uint32_t *)ctx->d_io_in, nwords, bit0, (uint64_t)n * 8, d_surv, surv_cap, d_count + 1, d_list, cand_cap, d_count);
        ZB_LAUNCHED();
        ZB_CHECK_LAUNCH();
        const uint64_t wantv = n / (128 * 137) + 1;          // (about one survivor per 137 bytes)
        const unsigned gridv = (unsigned)(wantv < (uint64_t)ctx->sm_count * 8 ? wantv : (uint64_t)ctx->sm_count * 8);
        prof_mark(ctx, s, "blk_validate_kernel");
        blk_validate_kernel<<<gridv, 128, 0, s>>>((const uint32_t *)ctx->d_io_in, nwords, (uint64_t)n * 8, d_surv, d_count + 1, surv_cap, d_list, cand_cap,
                                                  d_count, &((const InflateDeviceTables *)ctx->d_inflate_tables)->fmt);
        ZB_LAUNCHED();
        ZB_CHECK_LAUNCH();
    }
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_count, 8, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint32_t nc = *(const uint32_t *)ctx->h_small;
    if (nc < 2 || nc > cand_cap || ((const uint32_t *)ctx->h_small)[1] > surv_cap) return ZB200_OK;
    std::vector<uint64_t> cand(nc);
    ZB_CUDA(cudaMemcpy(cand.data(), d_list, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    std::sort(cand.begin(), cand.end());
    // 2. count: chunk 0 = the member's first block, chunk k = candidate k - 1
    const size_t nA = (size_t)nc + 1;
    std::vector<zb200_member> tab(nA);
    for (size_t k = 0; k < nA; ++k) {
        zb200_member &m = tab[k];
        m.in_off = 0; m.in_len = n; m.out_off = 0; m.out_cap = 0xfffffff0ull; m.resume_out = 0; m.dict_len = 0;
        m.resume_bit = k == 0 ? (cont && cont->bit0 ? bit0 : 0) : cand[k - 1];
    }
    const size_t cand_b = align_up((size_t)nc * 8, 256), tab_b = align_up(nA * sizeof(zb200_member), 256),
                 res_b = align_up(nA * sizeof(zb200_member_result), 256), u64_b = align_up(nA * 8, 256);
    if ((r = ensure_scratch(ctx, 512 + cand_b + tab_b + res_b + u64_b))) return r;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    unsigned int *d_counter = (unsigned int *)base;
    uint64_t *d_cand = (uint64_t *)(base + 512);
    zb200_member *d_members = (zb200_member *)(base + 512 + cand_b);
    zb200_member_result *d_results = (zb200_member_result *)(base + 512 + cand_b + tab_b);
    uint64_t *d_nmatch = (uint64_t *)(base + 512 + cand_b + tab_b + res_b);
    ZB_CUDA(cudaMemcpyAsync(d_cand, cand.data(), (size_t)nc * 8, cudaMemcpyHostToDevice, s));
    ZB_CUDA(cudaMemcpyAsync(d_members, tab.data(), nA * sizeof(zb200_member), cudaMemcpyHostToDevice, s));
    ChunkArgs ca;
    ca.cand = d_cand; ca.n_cand = nc; ca.mbase = nullptr; ca.mlist = nullptr; ca.nmatch = d_nmatch;
    if ((r = inflate_chunks_launch(ctx, INF_COUNT, ctx->d_io_in, nullptr, d_members, nA, kind, d_results, d_counter, ca, s))) return r;
    std::vector<zb200_member_result> res(nA);
    std::vector<uint64_t> nmatch(nA);
    ZB_CUDA(cudaMemcpyAsync(res.data(), d_results, nA * sizeof(zb200_member_result), cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaMemcpyAsync(nmatch.data(), d_nmatch, nA * 8, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    // 3. chain
    struct Link { size_t idx; uint64_t out_off, m_off; };
    std::vector<Link> chain;
    uint64_t total = 0, total_m = 0;
    int final_status = ZB200_INF_TRUNCATED;                  // a prefix, unless the chain reaches the end of the deflate data
    uint64_t stop_bit = bit0;                                // where the delivered prefix ends
    for (size_t cur = 0;;) {
        const zb200_member_result &q = res[cur];
        const uint64_t begin = cur == 0 ? bit0 : cand[cur - 1];
        if (q.status == ZB200_INF_OK) {                      // BFINAL and the trailer: the member ends inside this chunk
            chain.push_back({cur, total, total_m});
            total += q.out_len; total_m += nmatch[cur];
            final_status = ZB200_INF_OK;
            break;
        }
        if (q.status != ZB200_INF_TRUNCATED || q.resume_out != q.out_len || q.resume_bit <= begin) break;
        const auto it = std::lower_bound(cand.begin(), cand.end(), q.resume_bit);
        if (it == cand.end() || *it != q.resume_bit) break;  // stopped for another reason (the input ends inside the chunk)
        chain.push_back({cur, total, total_m});
        total += q.out_len; total_m += nmatch[cur];
        stop_bit = q.resume_bit;
        cur = (size_t)(it - cand.begin()) + 1;
    }
    static const int debug = [] { const char *e = getenv("ZB200_BLOCKS_DEBUG"); return e ? atoi(e) : 0; }();
    if (debug) {
        uint64_t max_out = 0, max_in = 0;
        for (size_t k = 0; k < chain.size(); ++k) {
            const zb200_member_result &q = res[chain[k].idx];
            const uint64_t b0 = chain[k].idx ? cand[chain[k].idx - 1] : bit0;
            const uint64_t in_bits = (q.status == ZB200_INF_OK ? q.in_used * 8 : q.resume_bit) - b0;
            if (q.out_len > max_out) max_out = q.out_len;
            if (in_bits > max_in) max_in = in_bits;
        }
        fprintf(stderr, "[blocks] n=%zu candidates=%u chain=%zu out=%llu matches=%llu largest chunk: %llu bytes out, %llu bytes in\n", n, nc, chain.size(),
                (unsigned long long)total, (unsigned long long)total_m, (unsigned long long)max_out, (unsigned long long)(max_in / 8));
    }
    if (chain.size() < 2 || hist_len + total >= 0xfffffff0ull) return ZB200_OK;
    if (final_status != ZB200_INF_OK && !prefix_ok) return ZB200_OK;
    *applicable = final_status == ZB200_INF_OK ? 2 : 1;
    *out_len = (size_t)total;
    *status = final_status;
    const zb200_member_result &fin = res[chain.back().idx];
    uint64_t end = final_status == ZB200_INF_OK ? fin.in_used : stop_bit >> 3;
    *in_used = (size_t)end;
    if (end_bit) *end_bit = final_status == ZB200_INF_OK ? 0u : (uint32_t)(stop_bit & 7);
    // (an output buffer that is too small is reported only once the second decode has confirmed the chain: the counting pass
    //  does not know what lies before a chunk, so it cannot see a distance that reaches too far back — found by
    //  tools/fuzz_blocks.py: a spliced stream came back "output buffer full" where the reference says "invalid distance")
    // 4. list + 5. resolve, group after group of chunks
    if ((r = ensure_io(ctx, n + 16, hist_len + total + 16))) return r;
    uint8_t *d_out = ctx->d_io_out;
    if (hist_len) ZB_CUDA(cudaMemcpyAsync(d_out, cont->hist, hist_len, cudaMemcpyHostToDevice, s));
    const size_t nB = chain.size();
    uint64_t grp_m_max = 0, grp_o_max = 0;
    std::vector<size_t> grp_first;                           // groups of consecutive chunks with <= kBlkGroupOut bytes of output
    for (size_t k = 0; k < nB;) {
        grp_first.push_back(k);
        const uint64_t o0 = chain[k].out_off, m0 = chain[k].m_off;
        size_t e = k + 1;
        auto o_end = [&](size_t j) { return j < nB ? chain[j].out_off : total; };
        while (e < nB && o_end(e + 1) - o0 <= kBlkGroupOut) ++e;
        const uint64_t go = o_end(e) - o0, gm = (e < nB ? chain[e].m_off : total_m) - m0;
        if (go >= 0x7fff0000ull) { *applicable = 0; *out_len = 0; *status = ZB200_INF_OK; *in_used = 0; return ZB200_OK; }   // (one chunk of 2 GiB: source offsets hold 31 bits)
        if (go > grp_o_max) grp_o_max = go;
        if (gm > grp_m_max) grp_m_max = gm;
        k = e;
    }
    grp_first.push_back(nB);
    const size_t tabB_b = align_up(nB * sizeof(zb200_member), 256), resB_b = align_up(nB * sizeof(zb200_member_result), 256), mb_b = align_up(nB * 8, 256),
                 ml_b = align_up((size_t)grp_m_max * sizeof(QueuedMatch) + 16, 256), src_b = align_up((size_t)grp_o_max * 4 + 16, 256);
    if ((r = ensure_scratch(ctx, 1024 + cand_b + tabB_b + resB_b + mb_b + ml_b + src_b))) return r;
    base = (uint8_t *)ctx->d_scratch;                        // (the block may have been replaced: everything is laid out afresh)
    d_counter = (unsigned int *)base;
    uint32_t *d_flag = (uint32_t *)(base + 512);
    uint32_t *d_sum = (uint32_t *)(base + 640);
    CkAccum *d_acc = (CkAccum *)(base + 768);
    d_cand = (uint64_t *)(base + 1024);
    d_members = (zb200_member *)(base + 1024 + cand_b);
    d_results = (zb200_member_result *)(base + 1024 + cand_b + tabB_b);
    uint64_t *d_mbase = (uint64_t *)(base + 1024 + cand_b + tabB_b + resB_b);
    QueuedMatch *d_ml = (QueuedMatch *)(base + 1024 + cand_b + tabB_b + resB_b + mb_b);
    uint32_t *d_src = (uint32_t *)(base + 1024 + cand_b + tabB_b + resB_b + mb_b + ml_b);
    ZB_CUDA(cudaMemcpyAsync(d_cand, cand.data(), (size_t)nc * 8, cudaMemcpyHostToDevice, s));
    std::vector<zb200_member> tabB(nB);
    std::vector<uint64_t> mbase(nB);
    std::vector<zb200_member_result> resB(nB);
    for (size_t g = 0; g + 1 < grp_first.size(); ++g) {
        const size_t k0 = grp_first[g], k1 = grp_first[g + 1], ng = k1 - k0;
        const uint64_t o0 = chain[k0].out_off, o1 = k1 < nB ? chain[k1].out_off : total;
        const uint64_t m0 = chain[k0].m_off, m1 = k1 < nB ? chain[k1].m_off : total_m;
        for (size_t k = k0; k < k1; ++k) {
            zb200_member &m = tabB[k];
            m = tab[chain[k].idx];
            m.out_off = hist_len; m.dict_len = hist_len; m.out_cap = total; m.resume_out = chain[k].out_off;
            mbase[k] = chain[k].m_off - m0;
        }
        ZB_CUDA(cudaMemcpyAsync(d_members + k0, tabB.data() + k0, ng * sizeof(zb200_member), cudaMemcpyHostToDevice, s));
        ZB_CUDA(cudaMemcpyAsync(d_mbase + k0, mbase.data() + k0, ng * 8, cudaMemcpyHostToDevice, s));
        ZB_CUDA(cudaMemsetAsync(d_src, 0xff, (size_t)(o1 - o0) * 4, s));
        ca.cand = d_cand; ca.n_cand = nc; ca.mbase = d_mbase + k0; ca.mlist = d_ml; ca.nmatch = nullptr;
        if ((r = inflate_chunks_launch(ctx, INF_LIST, ctx->d_io_in, d_out, d_members + k0, ng, kind, d_results + k0, d_counter, ca, s))) return r;
        ZB_CUDA(cudaMemcpyAsync(resB.data() + k0, d_results + k0, ng * sizeof(zb200_member_result), cudaMemcpyDeviceToHost, s));
        // The second decode must tell the first one's story (it adds the "too far back" test) BEFORE its match list is used:
        // a chunk that stopped early leaves the rest of its part of the list unwritten (found by tools/fuzz_blocks.py).
        ZB_CUDA(cudaStreamSynchronize(s));
        for (size_t k = k0; k < k1; ++k) {
            const zb200_member_result &a = res[chain[k].idx], &b = resB[k];
            if (a.status != b.status || b.out_len != chain[k].out_off + a.out_len || a.resume_bit != b.resume_bit) {
                *applicable = 0; *out_len = 0; *status = ZB200_INF_OK; *in_used = 0;
                return ZB200_OK;
            }
        }
        const uint32_t lo = (uint32_t)(hist_len + o0), sbase = lo > 32768u ? lo - 32768u : 0u;
        const uint64_t gn = o1 - o0, gm = m1 - m0;
        if (gm) {
            const uint64_t want = (gm + 255) / 256;          // 8 warps x 32 matches per CTA
            const unsigned grid = (unsigned)(want < (uint64_t)ctx->sm_count * 32 ? want : (uint64_t)ctx->sm_count * 32);
            prof_mark(ctx, s, "blk_src_build_kernel");
            blk_src_build_kernel<<<grid, 256, 0, s>>>(d_ml, gm, d_src, lo, sbase);
            ZB_LAUNCHED();
            ZB_CHECK_LAUNCH();
            const uint64_t wantj = (gn / 4 + 255) / 256 + 1;
            const unsigned gridj = (unsigned)(wantj < (uint64_t)ctx->sm_count * 32 ? wantj : (uint64_t)ctx->sm_count * 32);
            for (int pass = 0; pass < 64; pass += 2) {       // the flag is looked at every other pass
                ZB_CUDA(cudaMemsetAsync(d_flag, 0, 4, s));
                for (int k = 0; k < 2; ++k) {
                    prof_mark(ctx, s, "blk_jump_kernel");
                    blk_jump_kernel<<<gridj, 256, 0, s>>>(d_src, gn, lo, sbase, d_flag + (k == 1 ? 0 : 1));
                    ZB_LAUNCHED();
                }
                ZB_CHECK_LAUNCH();
                ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_flag, 4, cudaMemcpyDeviceToHost, s));
                ZB_CUDA(cudaStreamSynchronize(s));
                if (*(const uint32_t *)ctx->h_small == 0) break;
            }
            prof_mark(ctx, s, "blk_gather_kernel");
            blk_gather_kernel<<<gridj, 256, 0, s>>>(d_out, d_src, gn, lo, sbase);
            ZB_LAUNCHED();
            ZB_CHECK_LAUNCH();
        }
    }
    prof_mark(ctx, s, nullptr);
    if (total > out_cap) {                                   // (*out_len says how much room is needed; the caller may offer a buffer it grows)
        uint8_t *p = alt ? alt->grow(alt->self, (size_t)total) : nullptr;
        if (!p) { *status = ZB200_INF_OUTPUT_FULL; return ZB200_OK; }
        out = p; out_cap = (size_t)total; alt->used = true;
    }
    // the check value over the whole, the bytes back, the trailer (inflate.c:1183-1219)
    const int ck = cont ? cont->check_kind : kind;
    uint8_t *d_final = d_out + hist_len;
    if ((r = checksum_launch(ctx, d_final, nullptr, nullptr, total, 1, ck == 1 ? ZB200_ADLER32 : ZB200_CRC32, 0, 1, d_sum, d_sum + 1, d_acc, s))) return r;
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_sum, 8, cudaMemcpyDeviceToHost, s));
    if (total && (r = d2h_auto(ctx, out, d_final, (size_t)total, s))) return r;
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint32_t *hsum = (const uint32_t *)ctx->h_small;
    *check = ck == 1 ? hsum[1] : hsum[0];
    if (final_status == ZB200_INF_OK) {
        if (kind == 1 && fin.check != *check) *status = ZB200_INF_DATA_CHECK;
        else if (kind == 2 && fin.check != *check) *status = ZB200_INF_DATA_CHECK;
        else if (kind == 2 && fin.isize != (uint32_t)total) *status = ZB200_INF_LENGTH_CHECK;
    }
    return ZB200_OK;
}

}  // namespace zb

using namespace zb;

extern "C" {

const char *zb200_inflate_msg(int status) {
    static const char *const msgs[ZB200_INF_COUNT] = {
        "", "incorrect header check", "unknown compression method", "invalid window size",
        "unknown header flags set", "header crc mismatch", "invalid block type",
        "invalid stored block lengths", "too many length or distance symbols",
        "invalid code lengths set", "invalid bit length repeat",
        "invalid code -- missing end-of-block", "invalid literal/lengths set",
        "invalid distances set", "invalid literal/length code", "invalid distance code",
        "invalid distance too far back", "incorrect data check", "incorrect length check",
        "truncated input", "output buffer full", "need dictionary"};
    return (status >= 0 && status < ZB200_INF_COUNT) ? msgs[status] : "unknown status";
}

int zb200_inflate_dev(zb200_ctx *ctx, const void *d_in, void *d_out, const zb200_member *d_members,
                      size_t n_members, int wrap, int verify, zb200_member_result *d_results, void *stream) {
    if (!ctx || !d_members || !d_results || wrap < 0 || wrap > 3 || n_members > 0xfffffff0ull) return ZB200_ERR_PARAM;
    if (n_members == 0) return ZB200_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, pick_stream(ctx, stream));
    int r = ensure_scratch(ctx, InflateWork::bytes(n_members));
    if (r) return r;
    return inflate_launch(ctx, (const uint8_t *)d_in, (uint8_t *)d_out, d_members, n_members, wrap, verify,
                          d_results, ctx->d_scratch, pick_stream(ctx, stream));
}

// Many members, pinned buffers, members laid out in order: the member list is cut into up to
// 8 pieces of similar output size (1 GiB or more); piece k+1's input travels to the device and piece k-1's
// output travels back while piece k is decoded (three streams).
static int inflate_host_pipelined(zb200_ctx *ctx, const uint8_t *in, uint8_t *out, const zb200_member *members, size_t n,
                                  int wrap, int verify, zb200_member_result *results, size_t in_bytes, size_t out_bytes) {
    const size_t tbl = align_up(n * sizeof(zb200_member), 256), rsl = align_up(n * sizeof(zb200_member_result), 256);
    int r = ensure_io(ctx, in_bytes + 16, out_bytes + 16);
    if (r) return r;
    if ((r = ensure_scratch(ctx, tbl + rsl + InflateWork::bytes(n)))) return r;
    cudaStream_t s = ctx->stream;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    zb200_member *d_members = (zb200_member *)base;
    zb200_member_result *d_results = (zb200_member_result *)(base + tbl);
    void *d_work = base + tbl + rsl;
    // pieces of >= 1 GiB of output: a piece's kernel cannot be shorter than its largest member's own
    // decode (14 ms for 1 MiB), so small pieces would only queue those latencies one after the other
    size_t target = out_bytes / 8 > ((size_t)1 << 30) ? out_bytes / 8 : ((size_t)1 << 30);
    size_t first[10], np = 0;
    first[0] = 0;
    {
        size_t acc = 0;
        for (size_t i = 0; i < n; ++i) {
            acc += members[i].out_cap;
            if (acc >= target && i + 1 < n && np + 1 < 8) { first[++np] = i + 1; acc = 0; }
        }
        first[++np] = n;
    }
    cudaEvent_t ev_in[8], ev_out[8];
    for (size_t k = 0; k < np; ++k) {
        ZB_CUDA(cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming));
        ZB_CUDA(cudaEventCreateWithFlags(&ev_out[k], cudaEventDisableTiming));
    }
    auto cleanup = [&]() { for (size_t k = 0; k < np; ++k) { cudaEventDestroy(ev_in[k]); cudaEventDestroy(ev_out[k]); } };
    int rc = ZB200_OK;
    if (cudaMemcpyAsync(d_members, members, n * sizeof(zb200_member), cudaMemcpyHostToDevice, s) != cudaSuccess) rc = ZB200_ERR_CUDA;
    for (size_t k = 0; k < np && rc == ZB200_OK; ++k) {
        const size_t lo = members[first[k]].in_off, hi = members[first[k + 1] - 1].in_off + members[first[k + 1] - 1].in_len;
        if (cudaMemcpyAsync(ctx->d_io_in + lo, in + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess ||
            cudaEventRecord(ev_in[k], ctx->copy_stream) != cudaSuccess) rc = ZB200_ERR_CUDA;
    }
    for (size_t k = 0; k < np && rc == ZB200_OK; ++k) {
        cudaStreamWaitEvent(s, ev_in[k], 0);
        rc = inflate_launch(ctx, ctx->d_io_in, ctx->d_io_out, d_members + first[k], first[k + 1] - first[k], wrap, verify,
                            d_results + first[k], d_work, s);
        if (rc == ZB200_OK && cudaEventRecord(ev_out[k], s) != cudaSuccess) rc = ZB200_ERR_CUDA;
    }
    for (size_t k = 0; k < np && rc == ZB200_OK; ++k) {
        if (cudaEventSynchronize(ev_out[k]) != cudaSuccess) { rc = ZB200_ERR_CUDA; break; }
        const size_t lo = members[first[k]].out_off, hi = members[first[k + 1] - 1].out_off + members[first[k + 1] - 1].out_cap;
        if (cudaMemcpyAsync(out + lo, ctx->d_io_out + lo, hi - lo, cudaMemcpyDeviceToHost, ctx->back_stream) != cudaSuccess) rc = ZB200_ERR_CUDA;
    }
    if (rc == ZB200_OK && cudaMemcpyAsync(results, d_results, n * sizeof(zb200_member_result), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = ZB200_ERR_CUDA;
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(s);
    cudaStreamSynchronize(ctx->back_stream);
    cleanup();
    if (rc == ZB200_ERR_CUDA) set_error("inflate: pipelined transfer failed");
    return rc;
}

int zb200_selftest_tables(zb200_ctx *ctx, const uint8_t *lens, const uint32_t *counts, size_t n_cases, uint32_t *verdict) {
    if (!ctx || !lens || !counts || !verdict) return ZB200_ERR_PARAM;
    if (n_cases == 0) return ZB200_OK;
    for (size_t i = 0; i < n_cases; ++i)
        if (counts[2 * i] < 257 || counts[2 * i] > 288 || counts[2 * i + 1] < 1 || counts[2 * i + 1] > 32) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    const size_t lb = align_up(n_cases * 320, 256), cb = align_up(n_cases * 8, 256), vb = align_up(n_cases * 4, 256);
    int r = ensure_scratch(ctx, lb + cb + vb);
    if (r) return r;
    cudaStream_t s = ctx->stream;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    ZB_CUDA(cudaMemcpyAsync(base, lens, n_cases * 320, cudaMemcpyHostToDevice, s));
    ZB_CUDA(cudaMemcpyAsync(base + lb, counts, n_cases * 8, cudaMemcpyHostToDevice, s));
    const unsigned grid = (unsigned)(n_cases < (size_t)ctx->sm_count * 16 ? n_cases : (size_t)ctx->sm_count * 16);
    prof_mark(ctx, s, "tables_selftest_kernel");
    tables_selftest_kernel<<<grid, 32, 0, s>>>(base, (const uint32_t *)(base + lb), (uint32_t)n_cases,
                                               (const InflateDeviceTables *)ctx->d_inflate_tables, (uint32_t *)(base + lb + cb));
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    ZB_CUDA(cudaMemcpyAsync(verdict, base + lb + cb, n_cases * 4, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    return ZB200_OK;
}

int zb200_gunzip_host(zb200_ctx *ctx, const void *in, size_t n, void *out, size_t out_cap, size_t *out_len,
                      int *inf_status, zb200_member *members, size_t max_members, size_t *n_members) {
    if (!ctx || (!in && n) || (!out && out_cap) || !out_len || !inf_status) return ZB200_ERR_PARAM;
    *out_len = 0; *inf_status = ZB200_INF_OK;
    if (n_members) *n_members = 0;
    const uint8_t *src = (const uint8_t *)in;
    if (n < 18 || src[0] != 0x1f || src[1] != 0x8b) { *inf_status = n < 2 ? ZB200_INF_TRUNCATED : ZB200_INF_HEADER_CHECK; return ZB200_OK; }
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    cudaStream_t s = ctx->stream;
    int r = ensure_io(ctx, n + 16, 16);
    if (r) return r;
    if ((r = h2d_auto(ctx, ctx->d_io_in, in, n, s))) return r;
    // 1. candidates
    const uint32_t cand_cap = (uint32_t)(n / 18 + 16 < (1u << 24) ? n / 18 + 16 : (1u << 24));
    if ((r = ensure_scratch(ctx, 256 + (size_t)cand_cap * 8))) return r;
    uint32_t *d_count = (uint32_t *)ctx->d_scratch;
    uint64_t *d_list = (uint64_t *)((uint8_t *)ctx->d_scratch + 256);
    ZB_CUDA(cudaMemsetAsync(d_count, 0, 256, s));
    prof_mark(ctx, s, "gz_candidates_kernel");
    gz_candidates_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(ctx->d_io_in, n, d_list, cand_cap, d_count);
    ZB_LAUNCHED();
    ZB_CHECK_LAUNCH();
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_count, 8, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint32_t nc = *(const uint32_t *)ctx->h_small;
    if (nc > cand_cap) { set_error("gunzip: more than %u member candidates", cand_cap); return ZB200_ERR_PARAM; }
    std::vector<uint64_t> cand(nc);
    if (nc) ZB_CUDA(cudaMemcpy(cand.data(), d_list, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    std::sort(cand.begin(), cand.end());
    if (cand.empty() || cand[0] != 0) { *inf_status = ZB200_INF_HEADER_CHECK; return ZB200_OK; }
    auto le32 = [&](uint64_t at) { return (uint32_t)src[at] | ((uint32_t)src[at + 1] << 8) | ((uint32_t)src[at + 2] << 16) | ((uint32_t)src[at + 3] << 24); };
    // 1b. FEW, LARGE members (the usual .gz file is one): a member decoded by one team runs at 150 MB/s, so the members are
    // taken one after the other instead, each through the single-stream decoders (flush-point runs, else block-header
    // chunks).  Whatever does not come back clean from there is left to the batch below, which starts over.
    if (n >= ((size_t)4 << 20) && cand.size() <= 16) {
        size_t pos = 0, produced = 0, nm = 0;
        std::vector<zb200_member> found;
        bool clean = true;
        while (pos + 18 <= n && src[pos] == 0x1f && src[pos + 1] == 0x8b) {
            size_t out_len1 = 0, used = 0;
            int st1 = 0, appl = 0;
            uint32_t ck = 0;
            r = inflate_stream_parallel(ctx, src + pos, n - pos, ZB200_WRAP_GZIP, (uint8_t *)out + produced, out_cap - produced, &out_len1, &st1, &used,
                                        &ck, &appl, nullptr, 1);
            if (r != ZB200_OK || !appl || st1 != ZB200_INF_OK || used == 0) { clean = false; break; }
            zb200_member m;
            m.in_off = pos; m.in_len = used; m.out_off = produced; m.out_cap = out_len1; m.resume_bit = m.resume_out = m.dict_len = 0;
            found.push_back(m);
            produced += out_len1; pos += used; ++nm;
        }
        if (clean && nm) {                                   // (bytes that are no member may follow: gzread.c:gz_look ignores them)
            if (n_members) *n_members = nm;
            if (members) for (size_t k = 0; k < nm && k < max_members; ++k) members[k] = found[k];
            *out_len = produced;
            return ZB200_OK;
        }
        if ((r = ensure_io(ctx, n + 16, 16))) return r;      // (the single-stream decoders put their own input at the buffer's start)
        if ((r = h2d_auto(ctx, ctx->d_io_in, in, n, s))) return r;
    }
    // 2. hypotheses -> batch -> chain check; a false candidate is dropped and the batch redone
    std::vector<zb200_member> tab;
    std::vector<zb200_member_result> res;
    size_t good = 0, total = 0;
    bool last_at_max = false;                            // bytes after the last member hid its ISIZE: size it by the format's maximum
    for (int pass = 0;; ++pass) {
        if (pass > 64) { set_error("gunzip: too many false member candidates"); return ZB200_ERR_PARAM; }
        // a candidate whose predecessor's implied output size is impossible lies inside that predecessor
        for (size_t i = 0; i + 1 < cand.size();) {
            const uint64_t len = cand[i + 1] - cand[i];
            if (len < 18 || (uint64_t)le32(cand[i + 1] - 4) > len * 1032 + 65536) cand.erase(cand.begin() + (long)i + 1);
            else ++i;
        }
        const size_t m = cand.size();
        tab.assign(m, zb200_member());
        total = 0;
        for (size_t i = 0; i < m; ++i) {
            const uint64_t end = i + 1 < m ? cand[i + 1] : n, most = (end - cand[i]) * 1032 + 65536;
            uint64_t isz = le32(end - 4);
            if (i + 1 == m && (last_at_max || isz > most)) isz = most;
            tab[i].in_off = cand[i]; tab[i].in_len = end - cand[i];
            tab[i].out_off = total; tab[i].out_cap = isz;
            tab[i].resume_bit = tab[i].resume_out = tab[i].dict_len = 0;
            total += isz;
        }
        const size_t tbl = align_up(m * sizeof(zb200_member), 256), rsl = align_up(m * sizeof(zb200_member_result), 256);
        if ((r = ensure_io(ctx, n + 16, total + 16))) return r;
        if ((r = ensure_scratch(ctx, tbl + rsl + InflateWork::bytes(m)))) return r;
        uint8_t *base = (uint8_t *)ctx->d_scratch;
        zb200_member *d_members = (zb200_member *)base;
        zb200_member_result *d_results = (zb200_member_result *)(base + tbl);
        ZB_CUDA(cudaMemcpyAsync(d_members, tab.data(), m * sizeof(zb200_member), cudaMemcpyHostToDevice, s));
        if ((r = inflate_launch(ctx, ctx->d_io_in, ctx->d_io_out, d_members, m, ZB200_WRAP_GZIP, 1, d_results, base + tbl + rsl, s))) return r;
        res.resize(m);
        ZB_CUDA(cudaMemcpyAsync(res.data(), d_results, m * sizeof(zb200_member_result), cudaMemcpyDeviceToHost, s));
        ZB_CUDA(cudaStreamSynchronize(s));
        bool redo = false;
        good = 0;
        *inf_status = ZB200_INF_OK;
        for (size_t i = 0; i < m; ++i) {
            const zb200_member_result &q = res[i];
            const bool last = i + 1 == m;
            if (q.status == ZB200_INF_OK) {
                ++good;
                if (last || q.in_used == tab[i].in_len) continue;
                break;                                   // bytes that are no member follow: gzread.c:gz_look ignores them
            }
            if (!last && (q.status == ZB200_INF_TRUNCATED || q.status == ZB200_INF_OUTPUT_FULL)) {
                cand.erase(cand.begin() + (long)i + 1);  // the next candidate was none
                redo = true;
            } else if (last && q.status == ZB200_INF_OUTPUT_FULL && !last_at_max) {
                last_at_max = true;
                redo = true;
            } else *inf_status = q.status;
            break;
        }
        if (!redo) break;
    }
    // 3. the valid prefix: members' outputs lie back to back (ISIZE was right for each of them)
    size_t produced = 0;
    for (size_t i = 0; i < good; ++i) produced += (size_t)res[i].out_len;
    if (n_members) *n_members = good;
    if (members)
        for (size_t k = 0, at = 0; k < good && k < max_members; ++k) {
            members[k] = tab[k];
            members[k].in_len = res[k].in_used; members[k].out_off = at; members[k].out_cap = res[k].out_len;
            at += (size_t)res[k].out_len;
        }
    *out_len = produced;
    if (produced > out_cap) { set_error("gunzip: %zu bytes do not fit the output buffer", produced); return ZB200_ERR_OUTPUT; }
    if (produced && (r = d2h_auto(ctx, out, ctx->d_io_out, produced, s))) return r;
    ZB_CUDA(cudaStreamSynchronize(s));
    return ZB200_OK;
}

int zb200_inflate_host(zb200_ctx *ctx, const void *in, void *out, const zb200_member *members,
                       size_t n_members, int wrap, int verify, zb200_member_result *results) {
    if (!ctx || !members || !results || wrap < 0 || wrap > 3) return ZB200_ERR_PARAM;
    if (n_members == 0) return ZB200_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    size_t in_bytes = 0, out_bytes = 0;
    for (size_t i = 0; i < n_members; ++i) {
        if (members[i].in_off + members[i].in_len > in_bytes) in_bytes = members[i].in_off + members[i].in_len;
        if (members[i].out_off + members[i].out_cap > out_bytes) out_bytes = members[i].out_off + members[i].out_cap;
        if (members[i].dict_len > 32768 || members[i].dict_len > members[i].out_off) {
            set_error("inflate: a preset dictionary is at most 32768 bytes and lies before out_off");
            return ZB200_ERR_PARAM;
        }
    }
    // the pipelined path: enough work, pinned data buffers, members in order, nothing to resume
    if (n_members >= 64 && out_bytes >= ((size_t)1536 << 20) && is_pinned(in) && is_pinned(out)) {
        bool ordered = true;
        for (size_t i = 0; i < n_members && ordered; ++i) {
            if (members[i].resume_bit || members[i].dict_len) ordered = false;
            if (i && (members[i].in_off < members[i - 1].in_off + members[i - 1].in_len ||
                      members[i].out_off < members[i - 1].out_off + members[i - 1].out_cap)) ordered = false;
        }
        if (ordered) return inflate_host_pipelined(ctx, (const uint8_t *)in, (uint8_t *)out, members, n_members, wrap, verify, results, in_bytes, out_bytes);
    }
    const size_t tbl = align_up(n_members * sizeof(zb200_member), 256);
    const size_t rsl = align_up(n_members * sizeof(zb200_member_result), 256);
    int r = ensure_io(ctx, in_bytes + 16, out_bytes + 16);
    if (r) return r;
    r = ensure_scratch(ctx, tbl + rsl + InflateWork::bytes(n_members));
    if (r) return r;
    cudaStream_t s = ctx->stream;
    uint8_t *base = (uint8_t *)ctx->d_scratch;
    zb200_member *d_members = (zb200_member *)base;
    zb200_member_result *d_results = (zb200_member_result *)(base + tbl);
    void *d_work = base + tbl + rsl;
    if ((r = h2d_auto(ctx, ctx->d_io_in, in, in_bytes, s))) return r;
    // members that resume need the output produced so far (back-references reach into it)
    for (size_t i = 0; i < n_members; ++i)
        if (members[i].resume_bit && members[i].resume_out)
            if ((r = h2d_auto(ctx, ctx->d_io_out + members[i].out_off, (const uint8_t *)out + members[i].out_off,
                              members[i].resume_out, s))) return r;
    for (size_t i = 0; i < n_members; ++i)                      // preset dictionaries travel with their members
        if (members[i].dict_len)
            if ((r = h2d_auto(ctx, ctx->d_io_out + members[i].out_off - members[i].dict_len,
                              (const uint8_t *)out + members[i].out_off - members[i].dict_len, members[i].dict_len, s))) return r;
    if ((r = h2d_auto(ctx, d_members, members, n_members * sizeof(zb200_member), s))) return r;
    r = inflate_launch(ctx, ctx->d_io_in, ctx->d_io_out, d_members, n_members, wrap, verify, d_results, d_work, s);
    if (r) return r;
    if ((r = d2h_auto(ctx, results, d_results, n_members * sizeof(zb200_member_result), s))) return r;
    ZB_CUDA(cudaStreamSynchronize(s));
    // bring back only what was produced: per member for small batches, one
    // covering range for large ones (members are normally laid out back to back)
    if (n_members <= 64) {
        for (size_t i = 0; i < n_members; ++i) {
            const size_t from = members[i].resume_bit ? members[i].resume_out : 0;
            if (results[i].out_len > from)
                if ((r = d2h_auto(ctx, (uint8_t *)out + members[i].out_off + from, ctx->d_io_out + members[i].out_off + from,
                                  results[i].out_len - from, s))) return r;
        }
    } else {
        size_t lo = (size_t)-1, hi = 0;
        for (size_t i = 0; i < n_members; ++i) {
            if (!results[i].out_len) continue;
            if (members[i].out_off < lo) lo = members[i].out_off;
            if (members[i].out_off + results[i].out_len > hi) hi = members[i].out_off + results[i].out_len;
        }
        if (hi > lo && (r = d2h_auto(ctx, (uint8_t *)out + lo, ctx->d_io_out + lo, hi - lo, s))) return r;
    }
    ZB_CUDA(cudaStreamSynchronize(s));
    return ZB200_OK;
}

int zb200_inflate_stream_host(zb200_ctx *ctx, const void *in, size_t n, int wrap, void *out, size_t out_cap,
                              zb200_member_result *result) {
    if (!ctx || (!in && n) || (!out && out_cap) || !result || wrap < 0 || wrap > 3) return ZB200_ERR_PARAM;
    {
        ZB_CUDA(cudaSetDevice(ctx->device));
        CtxUse use(ctx, ctx->stream);
        size_t out_len = 0, in_used = 0;
        int status = 0, applicable = 0;
        uint32_t check = 0;
        const int r = inflate_stream_parallel(ctx, (const uint8_t *)in, n, wrap, (uint8_t *)out, out_cap, &out_len, &status, &in_used, &check, &applicable, nullptr, 1);
        // the run-parallel decode is an optimisation: when it cannot get its memory (a stream of stored / binary data with
        // many false 00 00 FF FF candidates asks for a slot per candidate) the one-member path below decodes the stream
        if (r != ZB200_OK && r != ZB200_ERR_NOMEM) return r;
        if (r == ZB200_OK && applicable) {
            memset(result, 0, sizeof *result);
            result->status = status; result->check = check; result->out_len = out_len; result->in_used = in_used;
            result->isize = (uint32_t)out_len;
            const uint8_t *p = (const uint8_t *)in;
            result->wrap_kind = wrap == ZB200_WRAP_RAW ? 0u : (n >= 2 && p[0] == 0x1f && p[1] == 0x8b && (wrap & ZB200_WRAP_GZIP)) ? 2u : 1u;
            return ZB200_OK;
        }
    }
    zb200_member m;
    m.in_off = 0; m.in_len = n; m.out_off = 0; m.out_cap = out_cap; m.resume_bit = m.resume_out = m.dict_len = 0;
    return zb200_inflate_host(ctx, in, out, &m, 1, wrap, 1, result);
}

}  // extern "C"
