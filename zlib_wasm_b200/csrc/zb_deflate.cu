// zb_deflate.cu — the deflate pipeline kernels and their orchestration.
// Algorithm and reference citations: zb_deflate.cuh.  Data layout in HBM for a
// sub-batch of B chunks of S bytes (scratch, per input byte): prev_dist u16 (2 B),
// match tables 2 x u32 (8 B), symbols u32 (4 B); per block 32 B of BlockInfo and
// 1.6 KiB of BlockCode.  Algorithmic bytes per chunk: U read + C written.
#include "zb_internal.h"
#include "zb_deflate.cuh"
#include <string.h>
#include <stdlib.h>

namespace zb {

// ld.shared through 32-bit shared-space addresses held in registers
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
// Named barriers 1..4 with IMMEDIATE ids.  With the id in a register the compiler has to assume that all 16 barriers of a
// CTA are in use, and the SM's barrier budget then allows four CTAs (ncu: launch__occupancy_limit_barriers = 4) — which
// is what bounded the greedy parse, a latency-bound kernel, at 16 of 64 warps per SM.
__device__ __forceinline__ void bar_sync64(uint32_t id) {
    switch (id) {
        case 1: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 3: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void bar_arrive64(uint32_t id) {
    switch (id) {
        case 1: asm volatile("bar.arrive 1, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.arrive 2, 64;" ::: "memory"); break;
        case 3: asm volatile("bar.arrive 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.arrive 4, 64;" ::: "memory"); break;
    }
}
// Match entries are (length << 16 | distance): bits 25..31 are free.  Bit 31 of the FULL-budget entry says that the
// quartered search (deflate.c:1390-1392) found something else — only then is the second table written and read.  At level 6
// the two agree at all but a few per cent of the positions: the parse stages one table instead of two (60 -> 40 KiB of
// shared memory: five CTAs per SM instead of three) and the match kernel writes 4 bytes per position instead of 8.
constexpr uint32_t kQuarterDiffers = 0x80000000u;
struct DeflateDeviceTables {
    FormatTables fmt;
    StaticTrees st;
};

struct ParseLink;
struct DeflateWork {
    ParseLink *links; unsigned int *mstate;        // one long chunk parsed by many CTAs (dfl_parse_multi_kernel)
    uint16_t *prev; uint32_t *mfull, *mquarter, *syms;
    BlockInfo *blocks; BlockCode *codes;
    uint32_t *nblocks; uint64_t *chunk_bytes, *chunk_off;
    uint32_t *chunk_crc; uint64_t *seg_off, *seg_len; CkAccum *acc;
};

static size_t al(size_t v) { return (v + 255) & ~(size_t)255; }
constexpr size_t kMultiRange = 262144;        // one long chunk: positions per CTA of the chain and parse kernels

static size_t work_bytes(size_t nb, size_t S, uint32_t MB) {
    const size_t np = nb * S;
    return al(np * 2) + 3 * al(np * 4 + 16) + al(nb * MB * sizeof(BlockInfo)) + al(nb * MB * sizeof(BlockCode)) +
           al(nb * 4) * 2 + al(nb * 8) * 4 + al(nb * sizeof(CkAccum)) + al((np / kMultiRange + 2) * 32) + 256 + 4096;
}

static void carve(DeflateWork &w, void *base, size_t nb, size_t S, uint32_t MB) {
    uint8_t *p = (uint8_t *)base;
    const size_t np = nb * S;
    w.prev = (uint16_t *)p; p += al(np * 2);
    w.mfull = (uint32_t *)p; p += al(np * 4 + 16);
    w.mquarter = (uint32_t *)p; p += al(np * 4 + 16);
    w.syms = (uint32_t *)p; p += al(np * 4 + 16);
    w.blocks = (BlockInfo *)p; p += al(nb * MB * sizeof(BlockInfo));
    w.codes = (BlockCode *)p; p += al(nb * MB * sizeof(BlockCode));
    w.nblocks = (uint32_t *)p; p += al(nb * 4);
    w.chunk_crc = (uint32_t *)p; p += al(nb * 4);
    w.chunk_bytes = (uint64_t *)p; p += al(nb * 8);
    w.chunk_off = (uint64_t *)p; p += al(nb * 8);
    w.seg_off = (uint64_t *)p; p += al(nb * 8);
    w.seg_len = (uint64_t *)p; p += al(nb * 8);
    w.acc = (CkAccum *)p; p += al(nb * sizeof(CkAccum));
    w.links = (ParseLink *)p; p += al((np / kMultiRange + 2) * 32);
    w.mstate = (unsigned int *)p;
}

struct Batch {                 // one sub-batch of chunks, passed by value to the kernels
    const uint8_t *in;         // start of the sub-batch's input (chunk 0's history first, if it has any)
    uint64_t bytes;            // input bytes in the sub-batch (history of chunk 0 included)
    uint32_t S;                // scratch stride: the longest a chunk gets with its history (= step + carry)
    uint32_t nb;               // chunks in the sub-batch
    uint32_t MB;               // block slots per chunk
    int last_is_final;         // the last chunk of this sub-batch ends the stream (Z_FINISH)
    int all_final;             // every chunk is its own stream (gzip members)
    uint32_t skip;             // history of chunk 0: its first `skip` bytes are searched, never emitted (a preset dictionary, or
                               // the tail of the previous sub-batch / piece when history is carried)
    uint32_t first_bit;        // deflatePrime: chunk 0 of the call starts at this bit (0..7) of its first output byte
    uint32_t range;            // ONE long chunk (nb = 1) worked on by many CTAs of the ordered kernels: positions per CTA (0: off)
    uint32_t step;             // new input bytes per chunk (= S unless history is carried)
    uint32_t carry;            // history carried from chunk to chunk (pigz-style, deflateSetDictionary of the previous w_size bytes
                               // per chunk, deflate.c:550-632): every chunk after the first sees the `carry` bytes before it; 0 = off
};

// A chunk as its kernels see it: `lo` bytes of history (never emitted) followed by its new bytes, scratch at c * S.
// (chunks shorter than the window: an early chunk's history is what there is — everything from b.in on)
__device__ __forceinline__ uint32_t chunk_lo(const Batch &b, uint32_t c) {
    if (c == 0) return b.skip;
    const uint64_t before = (uint64_t)b.skip + (uint64_t)c * b.step;
    return before < b.carry ? (uint32_t)before : b.carry;
}
__device__ __forceinline__ const uint8_t *chunk_data(const Batch &b, uint32_t c) {
    return b.in + ((uint64_t)b.skip + (uint64_t)c * b.step - chunk_lo(b, c));
}
__device__ __forceinline__ uint32_t chunk_len(const Batch &b, uint32_t c) {
    const uint64_t rem = b.bytes - b.skip - (uint64_t)c * b.step;
    return chunk_lo(b, c) + (uint32_t)(rem < b.step ? rem : b.step);
}
__device__ __forceinline__ bool chunk_final(const Batch &b, uint32_t c) {
    return b.all_final || (b.last_is_final && c + 1 == b.nb);
}

// ---- phase 1: hash chains, one CTA of four warps per chunk ------------------------------
// 32 consecutive positions per step: same-hash positions inside the step are linked
// with match.any, the rest through a 32K-entry u16 head table in shared memory holding
// window-relative positions; every 32 KiB the table slides like deflate.c:187-209 so
// that entries stay 16-bit.  Only the head-table part of a step is inherently ordered;
// reading the input, hashing and storing the links are not.  The 64 KiB table allows
// three CTAs per SM, and a lone warp runs at its own dependent-issue latency (ncu: 46
// instructions per step at ~5.5 cycles each), so four warps share one table: warp w
// takes every fourth TRIP of 128 positions, hashes it while the others are busy, waits
// for its turn (a named barrier shared with the previous trip's warp: a blocked warp
// takes no issue slots; spinning on a shared-memory counter measured 6.7x slower), does
// the four ordered head-table steps, passes the turn on and stores its links.
constexpr int kChainWarps = 4, kChainTrip = 128;
// head table: 2 << hash_bits bytes of dynamic shared memory (64 KiB at memLevel 8, 128 KiB at memLevel 9)
// The body is compiled once per warp of the CTA (W = the warp's index): the ids of the two named barriers a warp uses
// are then immediates, and ptxas counts 5 barriers instead of assuming all 16 — with a register id the SM's barrier budget
// allows 4 CTAs (ncu: launch__occupancy_limit_barriers), which binds as soon as the head table is smaller than 64 KiB.
template <uint32_t W, bool IMM>
__device__ __forceinline__ void chain_body(const Batch &b, const DeflateParams &prm, uint16_t *__restrict__ prev_all, uint16_t *head) {
    const uint32_t lane_warp = threadIdx.x >> 5;             // (IMM: W is this warp's index, known at compile time)
    const uint32_t warp = IMM ? W : lane_warp;
    const uint32_t c = b.range ? 0u : blockIdx.x, lane = threadIdx.x & 31;
    const uint8_t *data = chunk_data(b, c);
    const uint32_t n = chunk_len(b, c);
    uint16_t *prev = prev_all + (uint64_t)c * b.S;
    uint4 *h4 = reinterpret_cast<uint4 *>(head);
    const int nvec = (int)((2u << prm.hash_bits) >> 4);                 // uint4 vectors of the table (>= 32)
    const uint32_t hshift = prm.hash_shift, hmask = prm.hash_mask, wsz = prm.w_size, wlog = 31u - (uint32_t)__clz(prm.w_size);
    const uint32_t wsub = wsz | (wsz << 16);
    constexpr int G = kChainTrip / 32;
    const uintptr_t in_hi = reinterpret_cast<uintptr_t>(b.in) + b.bytes;
    // the three bytes at a position as one little-endian word (bytes past the input read as 0)
    auto bytes3 = [&](uint32_t p) -> uint32_t {
        const uintptr_t a = reinterpret_cast<uintptr_t>(data) + p;
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const uint32_t lo = w[0];
        const uint32_t hi = (reinterpret_cast<uintptr_t>(w) + 4 < in_hi) ? w[1] : 0u;    // read only words that hold at least one input byte
        return __funnelshift_r(lo, hi, (uint32_t)(a & 3) * 8);
    };
    // One long chunk in RANGES (b.range positions per CTA): a CTA first re-inserts the w_size positions before its range —
    // no candidate of its own positions lies further back (MAX_DIST < w_size) — and stores the links of its range only.
    // Links that the one-CTA walk would have drawn to positions further back than that come out as NIL here; both end a
    // match walk (deflate.c:1481 cur_match > limit), so the matches found are the same.
    uint32_t t_first = 0, t_last = (n + kChainTrip - 1) / kChainTrip, store_lo = 0;
    if (b.range) {
        const uint32_t lo_r = blockIdx.x * b.range, hi_r = n - lo_r < b.range ? n : lo_r + b.range;
        t_first = (lo_r > wsz ? lo_r - wsz : 0u) / kChainTrip;
        t_last = (hi_r + kChainTrip - 1) / kChainTrip;
        store_lo = lo_r;
    }
    uint32_t hs[G], nxt[G];
    auto hash_trip = [&](uint32_t t, uint32_t (&h)[G]) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t p = t * kChainTrip + 32 * j + lane;
            if (p + kMinMatch <= n) {
                const uint32_t x = bytes3(p);
                h[j] = (((x & 0xffu) << (2 * hshift)) ^ (((x >> 8) & 0xffu) << hshift) ^ ((x >> 16) & 0xffu)) & hmask;
            } else h[j] = 0x10000u | lane;
        }
    };
    if (t_first + warp < t_last) hash_trip(t_first + warp, nxt);
    for (uint32_t t = t_first + warp; t < t_last; t += kChainWarps) {
#pragma unroll
        for (int j = 0; j < G; ++j) hs[j] = nxt[j];
        if (t + kChainWarps < t_last) hash_trip(t + kChainWarps, nxt);     // the next trip's hashes: off the ordered path
        const uint32_t p0 = t * kChainTrip;
        // window origin after every slide due up to this trip (the first one happens at position 2 * w_size)
        const uint32_t base = p0 < 2u * wsz ? 0u : ((p0 >> wlog) - 1u) << wlog;
        if (t != t_first) {                                        // trip t-1 has left the head table (its warp arrived here)
            if (IMM) asm volatile("bar.sync %0, 64;" ::"n"(W + 1) : "memory");
            else asm volatile("bar.sync %0, 64;" ::"r"(lane_warp + 1) : "memory");
        }
        if (p0 >= 2u * wsz && (p0 & (wsz - 1u)) == 0) {            // slide: subtract w_size, saturating at 0 (= NIL)
            for (int i = lane; i < nvec; i += 32) {
                uint4 v = h4[i];
                v.x = __vsubus2(v.x, wsub); v.y = __vsubus2(v.y, wsub);
                v.z = __vsubus2(v.z, wsub); v.w = __vsubus2(v.w, wsub);
                h4[i] = v;
            }
            __syncwarp();
        }
        uint32_t dists[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t g0 = p0 + 32 * j, p = g0 + lane;
            const bool valid = p + kMinMatch <= n;
            const uint32_t rel = p - base;                         // window-relative position, 1..65535 here (0 = NIL)
            // Optimistic step: read the old head, let every lane write its own position and
            // read back.  If every lane reads back its own value the 32 hashes were distinct
            // and the old heads are the links.  match.any (whose cost grows with the number
            // of distinct values) is only needed when two lanes of the group collide.
            uint32_t old = 0;
            if (valid) old = head[hs[j]];
            __syncwarp();
            if (valid) head[hs[j]] = (uint16_t)rel;
            __syncwarp();
            const bool clash = valid && head[hs[j]] != (uint16_t)rel;
            uint32_t dist = (valid && old) ? rel - old : 0;
            if (__any_sync(0xffffffffu, clash)) {
                const uint32_t mask = __match_any_sync(0xffffffffu, hs[j]);
                const uint32_t lower = mask & ((1u << lane) - 1u);
                if (valid && lower) {
                    const uint32_t q = g0 + (31u - (uint32_t)__clz(lower));
                    dist = q ? p - q : 0;                          // position 0 is never a match target (deflate.c:1366)
                }
                __syncwarp();
                if (valid && (mask >> lane) == 1u) head[hs[j]] = (uint16_t)rel;   // the group's highest lane wins
                __syncwarp();
            }
            dists[j] = dist;
        }
        if (t + 1 < t_last) {                                      // pass the turn on
            if (IMM) asm volatile("bar.arrive %0, 64;" ::"n"(((W + 1) & (kChainWarps - 1)) + 1) : "memory");
            else asm volatile("bar.arrive %0, 64;" ::"r"(((lane_warp + 1) & (kChainWarps - 1)) + 1) : "memory");
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t p = p0 + 32 * j + lane;
            if (p < n && p >= store_lo) prev[p] = (uint16_t)dists[j];
        }
    }
}
// IMM = false: one copy of the body, barrier ids in registers — the form for 64 KiB tables and more (levels 4-9: the table
// bounds the kernel at 3 CTAs per SM, the barrier budget would allow 4, and the four copies measured 7 % slower there).
template <bool IMM>
__global__ void __launch_bounds__(kChainWarps * 32) dfl_chain_kernel(Batch b, DeflateParams prm, uint16_t *__restrict__ prev_all) {
    extern __shared__ __align__(16) uint16_t head[];
    {
        uint4 *h4 = reinterpret_cast<uint4 *>(head);
        const int nvec = (int)((2u << prm.hash_bits) >> 4);
        for (int i = threadIdx.x; i < nvec; i += kChainWarps * 32) h4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (!IMM) { chain_body<0, false>(b, prm, prev_all, head); return; }
    switch (threadIdx.x >> 5) {
        case 0: chain_body<0, IMM>(b, prm, prev_all, head); break;
        case 1: chain_body<1, IMM>(b, prm, prev_all, head); break;
        case 2: chain_body<2, IMM>(b, prm, prev_all, head); break;
        default: chain_body<3, IMM>(b, prm, prev_all, head); break;
    }
}

// ---- phase 2: longest match per position --------------------------------------------
// One position per thread.  The walk itself (MatchWalk: screen / measure / advance) is
// irregular -- ncu: 8 of 32 lanes active on average, issue slots 74 % busy -- and four
// re-schedulings of it inside the warp were measured and dropped:
// they raise the active-lane count but add as many scheduling instructions as they save.
// This plain form (operands from global memory) now serves Z_RLE only; levels 1-2 run
// dfl_match_uniform_kernel and levels 3-9 dfl_match_sorted_kernel, both on a window
// staged in shared memory.
__global__ void __launch_bounds__(256)
dfl_match_kernel(Batch b, DeflateParams prm, const uint16_t *__restrict__ prev_all,
                 uint32_t *__restrict__ mfull, uint32_t *__restrict__ mquarter) {
    const uint32_t c = blockIdx.y;
    const uint32_t n = chunk_len(b, c);
    const uint32_t p = blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    const uint64_t off = (uint64_t)c * b.S;
    const uint8_t *data = chunk_data(b, c);
    if (prm.mode == MODE_RLE) { mfull[off + p] = rle_at(data, n, p); return; }
    const MatchPair r = match_at(data, n, prev_all + off, p, prm);
    const bool diff = prm.need_quarter && r.quarter != r.full;
    mfull[off + p] = r.full | (diff ? kQuarterDiffers : 0u);
    if (diff) mquarter[off + p] = r.quarter;
}

// The same walk with the positions of a tile handed to the warps in order of their
// chain length (links inside the window, up to the budget: counted on the staged links).  A warp costs as many steps as its longest walk; consecutive positions
// have unrelated walk lengths (ncu: 8 of 32 lanes active; host replay: 30 % of the
// lane-steps useful on text at level 6), while positions with the same depth bucket
// have similar ones (75 % useful with quarter-octave buckets, 90 % on mixed data).
// Per tile: bucket of every position (0 = no candidate at all), counting sort in shared
// memory, then thread i of the CTA takes the sorted entries i, i + 256, ...
//
// The walks themselves run on shared memory.  From global memory every lane's candidate
// byte and chain link is a different cache line, and the kernel is bound by the L1's tag
// stage (one line per cycle: 10 G scattered reads per 512 MiB at level 6 ~ 28 ms, the
// time measured with and without the sort).  So the CTA first stages everything the
// walks of its tile can touch — the 32 KiB of history before the tile, the tile and its
// lookahead (input bytes) and the chain links of the same positions — with coalesced
// 16-byte loads; a 16 Ki-position tile makes that 9 bytes of staging per position.
constexpr uint32_t kMsTile = 16384, kMsThreads = 1024, kMsBuckets = 48;
constexpr uint32_t kMsLook = kMaxMatch + 16;                                  // bytes a walk may read past its position
constexpr uint32_t kMsDataBytes = (kWSize + kMsTile + kMsLook + 32 + 15) & ~15u;   // + alignment skew / slack
constexpr uint32_t kMsLinkBytes = ((kWSize + kMsTile) * 2 + 32 + 15) & ~15u;
constexpr uint32_t kMsSmem = kMsDataBytes + kMsLinkBytes + kMsTile * 2;
__device__ __forceinline__ uint32_t depth_bucket(uint32_t e) {   // 0,1,2,3, then four buckets per octave
    if (e < 4u) return e;
    const uint32_t l = 31u - (uint32_t)__clz(e);
    return 4u + (l - 2u) * 4u + ((e >> (l - 2u)) & 3u);          // e <= 4095 -> at most 43
}
extern __shared__ __align__(16) uint8_t ms_smem[];
// Operands from the staged window.  Addresses are 32-bit shared-space addresses held in
// registers and the loads are ld.shared in PTX: written as C++ indexing of ms_smem the
// compiler rebuilt the CTA's shared-window base (S2UR CgaCtaId, ULEA, ...) in every
// iteration of the walk, a quarter of its instructions.
struct StagedMem {
    static constexpr bool kNilIsFar = true;              // stage_window stores the end of a chain (link 0) as 65535: farther than any window
    uint32_t dbase, lbase;                               // shared address of chunk position 0 (wraps), same for link 0
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return lds_u8(dbase + pos); }
    __device__ __forceinline__ uint32_t word(uint32_t pos) const {
        const uint32_t a = dbase + pos, w = a & ~3u;
        return __funnelshift_r(lds_u32(w), lds_u32(w + 4), (a & 3u) * 8u);
    }
    __device__ __forceinline__ uint32_t link(uint32_t pos) const { return lds_u16(lbase + 2u * pos); }
    // the accessors match_uniform uses
    template <int NW> __device__ __forceinline__ void words(uint32_t pos, uint32_t (&w)[NW]) const {
        const uint32_t a = dbase + pos, sh = (a & 3u) * 8u, w0 = a & ~3u;
        uint32_t raw[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; ++i) raw[i] = lds_u32(w0 + 4u * i);
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
    }
    __device__ __forceinline__ uint32_t dist(uint32_t pos) const { return link(pos); }
};

// Bytes from the staged window, chain links straight from global memory: what levels 1-2 use.  Their walk reads two links
// per position — its own (coalesced) and its first candidate's — so staging the 48 Ki links of a tile's window (96 KiB of
// the 180) cost more than it saved: it held the kernel at one CTA per SM, whose staging and walking phases then took turns.
struct StagedDataMem {
    static constexpr bool kNilIsFar = false;             // a chain's end is link 0
    uint32_t dbase; const uint16_t *prev;
    __device__ __forceinline__ uint32_t byte(uint32_t pos) const { return lds_u8(dbase + pos); }
    __device__ __forceinline__ uint32_t word(uint32_t pos) const {
        const uint32_t a = dbase + pos, w = a & ~3u;
        return __funnelshift_r(lds_u32(w), lds_u32(w + 4), (a & 3u) * 8u);
    }
    __device__ __forceinline__ uint32_t link(uint32_t pos) const { return prev[pos]; }
    template <int NW> __device__ __forceinline__ void words(uint32_t pos, uint32_t (&w)[NW]) const {
        const uint32_t a = dbase + pos, sh = (a & 3u) * 8u, w0 = a & ~3u;
        uint32_t raw[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; ++i) raw[i] = lds_u32(w0 + 4u * i);
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = __funnelshift_r(raw[i], raw[i + 1], sh);
    }
    __device__ __forceinline__ uint32_t dist(uint32_t pos) const { return prev[pos]; }
};

// Stage bytes [lo, hi) and (LINKS) links [lo, t1) of a chunk into ms_smem (all threads of the CTA;
// ends with a barrier) and return the accessor.
template <bool LINKS = true>
__device__ __forceinline__ StagedMem stage_window(const Batch &b, const uint8_t *data, const uint16_t *prev,
                                                  uint32_t lo, uint32_t hi, uint32_t t1) {
    uint4 *sdata = reinterpret_cast<uint4 *>(ms_smem);
    uint4 *slink = reinterpret_cast<uint4 *>(ms_smem + kMsDataBytes);
    // input bytes: 16-byte vectors by ADDRESS; vectors that straddle the ends of the input buffer are read bytewise
    const uintptr_t in_lo = reinterpret_cast<uintptr_t>(b.in), in_hi = in_lo + b.bytes;
    const uintptr_t d_first = ((reinterpret_cast<uintptr_t>(data) + lo) & ~(uintptr_t)15) - 16;   // one vector of slack before
    const uint32_t dskew = (uint32_t)(reinterpret_cast<uintptr_t>(data) + lo - d_first);
    const uint32_t dvecs = (uint32_t)((reinterpret_cast<uintptr_t>(data) + hi - d_first + 15) >> 4) + 1;   // and one after
    for (uint32_t v = threadIdx.x; v < dvecs; v += blockDim.x) {
        const uintptr_t a = d_first + 16ull * v;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (a >= in_lo && a + 16 <= in_hi) x = *reinterpret_cast<const uint4 *>(a);
        else if (a + 16 > in_lo && a < in_hi) {
            uint8_t *xb = reinterpret_cast<uint8_t *>(&x);
            for (int k = 0; k < 16; ++k) if (a + k >= in_lo && a + k < in_hi) xb[k] = *reinterpret_cast<const uint8_t *>(a + k);
        }
        sdata[v] = x;
    }
    // chain links of [lo, t1): the array sits inside the engine's scratch (16-byte aligned, followed by other arrays)
    const uintptr_t p_first = reinterpret_cast<uintptr_t>(prev + lo) & ~(uintptr_t)15;
    const uint32_t pskew = (uint32_t)((reinterpret_cast<uintptr_t>(prev + lo) - p_first) >> 1);
    const uint32_t pvecs = LINKS ? (uint32_t)((reinterpret_cast<uintptr_t>(prev + t1) - p_first + 15) >> 4) : 0u;
    for (uint32_t v = threadIdx.x; v < pvecs; v += blockDim.x) {
        uint4 x = *reinterpret_cast<const uint4 *>(p_first + 16ull * v);
        x.x |= __vcmpeq2(x.x, 0u); x.y |= __vcmpeq2(x.y, 0u);      // NIL (0) -> 65535: one test (the window limit) ends a walk
        x.z |= __vcmpeq2(x.z, 0u); x.w |= __vcmpeq2(x.w, 0u);
        slink[v] = x;
    }
    __syncthreads();
    StagedMem mem;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ms_smem);
    mem.dbase = sbase + dskew - lo;                                // chunk position p lives at shared address dbase + p
    mem.lbase = sbase + kMsDataBytes + 2u * pskew - 2u * lo;       // link p at lbase + 2 p
    return mem;
}

// Levels 1-2 on the staged window: the branch-free walk reads 16 scattered words and 4
// links per position; from global memory that is bound by the L1's tag stage (7.6 ms per
// 512 MiB at level 1), from shared memory by bank conflicts (~3.5 wavefronts per request).
// GLINKS: links from global memory (level 1: two per position; measured 2.90 -> 2.54 ms per 444 MiB, two CTAs per SM),
// else from the staged window (level 2 reads up to eight per position: from global memory 8.5 -> 10.4 ms).
template <int CH, int NICE, bool GLINKS>
__global__ void __launch_bounds__(kMsThreads, GLINKS ? 2 : 1)
dfl_match_uniform_kernel(Batch b, DeflateParams prm, const uint16_t *__restrict__ prev_all, uint32_t *__restrict__ mfull) {
    const uint32_t c = blockIdx.y, n = chunk_len(b, c), t0 = blockIdx.x * kMsTile;
    if (t0 >= n || t0 + kMsTile <= chunk_lo(b, c)) return;         // (a tile of history only: no position of it is ever parsed)
    const uint64_t off = (uint64_t)c * b.S;
    const uint8_t *data = chunk_data(b, c);
    const uint16_t *prev = prev_all + off;
    const uint32_t cnt = n - t0 < kMsTile ? n - t0 : kMsTile, t1 = t0 + cnt;
    const uint32_t lo = t0 > (uint32_t)kWSize ? t0 - kWSize : 0;
    const uint32_t hi = t1 + kMsLook < n ? t1 + kMsLook : n;
    const StagedMem smem = stage_window<!GLINKS>(b, data, prev, lo, hi, t1);
    StagedDataMem gmem;
    gmem.dbase = smem.dbase; gmem.prev = prev;
    for (uint32_t p = t0 + threadIdx.x; p < t1; p += kMsThreads) {
        uint32_t r;
        if (GLINKS) {
            if (p + kUniformTail <= n) r = match_uniform<CH, NICE>(gmem, p, prm.max_dist);
            else r = match_walk(gmem, n, p, prm).full;             // chunk tail: lookahead clamps apply
        } else {
            if (p + kUniformTail <= n) r = match_uniform<CH, NICE>(smem, p, prm.max_dist);
            else r = match_walk(smem, n, p, prm).full;
        }
        mfull[off + p] = r;
    }
}

__global__ void __launch_bounds__(kMsThreads)
dfl_match_sorted_kernel(Batch b, DeflateParams prm, const uint16_t *__restrict__ prev_all,
                        uint32_t *__restrict__ mfull, uint32_t *__restrict__ mquarter, uint32_t key_cap) {
    __shared__ uint32_t s_hist[kMsBuckets], s_cur[kMsBuckets];
    uint16_t *s_order = reinterpret_cast<uint16_t *>(ms_smem + kMsDataBytes + kMsLinkBytes);
    const uint32_t c = blockIdx.y, n = chunk_len(b, c), t0 = blockIdx.x * kMsTile;
    if (t0 >= n || t0 + kMsTile <= chunk_lo(b, c)) return;         // (a tile of history only: no position of it is ever parsed)
    const uint64_t off = (uint64_t)c * b.S;
    const uint8_t *data = chunk_data(b, c);
    const uint16_t *prev = prev_all + off;
    const uint32_t cnt = n - t0 < kMsTile ? n - t0 : kMsTile, t1 = t0 + cnt;
    const uint32_t lo = t0 > (uint32_t)kWSize ? t0 - kWSize : 0;             // candidates lie less than 32 KiB back
    const uint32_t hi = t1 + kMsLook < n ? t1 + kMsLook : n;                  // bytes [lo, hi) can be read
    if (threadIdx.x < kMsBuckets) s_hist[threadIdx.x] = 0;
    StagedMem mem = stage_window(b, data, prev, lo, hi, t1);
    // Pin the two bases in registers.  At 64 registers per thread the compiler otherwise REMATERIALISES them inside the
    // walk loop — 21 instructions (S2UR ctaid, LDC, S2UR CgaCtaId, ULEA, ...) per candidate step, half of the step
    // (ncu source page, round 2).  The result of an asm statement cannot be recomputed, so it has to stay live.
    uint32_t n_pin = n;                                            // (the chunk length too: it is read in the measuring path)
    asm volatile("" : "+r"(mem.dbase), "+r"(mem.lbase), "+r"(n_pin));

    uint32_t bk[kMsTile / kMsThreads];
#pragma unroll
    for (uint32_t k = 0; k < kMsTile / kMsThreads; ++k) {
        const uint32_t i = k * kMsThreads + threadIdx.x;
        bk[k] = 0xffffffffu;
        if (i < cnt) {
            const uint32_t p = t0 + i, d = mem.link(p);
            uint32_t e = 0;
            if (p + kMinMatch <= n && d != 0 && d <= prm.max_dist) {
                // length of the walk if no candidate ends it early: links inside the window, up to the budget.
                // Counted in position order (lanes diverge: 12 of 32 active), so it is capped at $ZB200_KEY_CAP links, 8 by
                // default: with the walk's step at its present cost a finer key costs more than it saves (512 MiB at level 6,
                // text / mixed: cap 4 22.0 / 37.4 ms, 8 19.2 / 36.1, 16 19.6 / 36.7, 32 21.3 / 38.4, 64 22.4 / 41.3).
                const uint32_t cap = (uint32_t)prm.chain < key_cap ? (uint32_t)prm.chain : key_cap;
                uint32_t q = p - d;
                e = 1;
                while (e < cap) {
                    const uint32_t d2 = mem.link(q);
                    if (d2 == 0) break;
                    q -= d2;
                    if (p - q >= prm.max_dist) break;
                    ++e;
                }
            }
            bk[k] = depth_bucket(e);
            atomicAdd(&s_hist[bk[k]], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                        // longest walks first
        uint32_t run = 0;
        for (int q = kMsBuckets - 1; q >= 0; --q) { s_cur[q] = run; run += s_hist[q]; }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t k = 0; k < kMsTile / kMsThreads; ++k)
        if (bk[k] != 0xffffffffu) s_order[atomicAdd(&s_cur[bk[k]], 1u)] = (uint16_t)(k * kMsThreads + threadIdx.x);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt; i += kMsThreads) {
        const uint32_t p = t0 + s_order[i];
        const MatchPair r = match_walk(mem, n_pin, p, prm);
        const bool diff = prm.need_quarter && r.quarter != r.full;
        mfull[off + p] = r.full | (diff ? kQuarterDiffers : 0u);
        if (diff) mquarter[off + p] = r.quarter;
    }
}

// ---- levels 1-3 with the reference's own chains (ZB200_EXACT_FAST): one chunk per THREAD ----
// fast_exact_chunk (zb_deflate.cuh) is one serial walk per chunk — deflate_fast's chains depend on its parse — so the only
// parallelism is across chunks: a CTA of one working thread per chunk, every chunk resident at once (each thread is its own
// warp: no lane waits for another chunk's control flow; the walk is bound by the latency of its dependent table reads).
// head table = the first 1 << hash_bits words of the chunk's (zeroed) match-table slot, links = its link array.
__global__ void __launch_bounds__(32)
dfl_fast_exact_kernel(Batch b, DeflateParams prm, uint16_t *__restrict__ prev_all, uint32_t *__restrict__ mfull,
                      uint32_t *__restrict__ syms, BlockInfo *__restrict__ blocks, uint32_t *__restrict__ nblocks) {
    if (threadIdx.x != 0) return;
    const uint32_t c = blockIdx.x;
    const uint64_t off = (uint64_t)c * b.S;
    uint32_t ns = 0, nbk = 0;
    fast_exact_chunk(chunk_data(b, c), chunk_len(b, c), prm, chunk_final(b, c), mfull + off, prev_all + off, syms + off,
                     blocks + (uint64_t)c * b.MB, ns, nbk);
    nblocks[c] = nbk;
}

// ---- phase 3, greedy rules (deflate_fast / deflate_rle / deflate_huff): one chunk per CTA ----
// Exit tables per tile (zb_deflate.cuh gt_*): warp w takes tiles w, w + 4, ...; it loads
// the tile's match entries, fills the tables, waits for the previous tile's exit (named
// barrier, as in the chain kernel), crosses its tile in 32 dependent table reads, passes
// (exit position, symbol count) on and then forms its symbols.
constexpr int kGtWarps = 4;
struct GtWarpShared {
    uint32_t mfv[kGtSlots];
    uint16_t lc[kGtSlots];
};
struct GtAcc {                                         // greedy_symbol's operands: entries from the tile, bytes from global memory
    const uint32_t *mfv; const uint8_t *data; uint32_t t0;
    __device__ __forceinline__ uint32_t mf(uint32_t p) const { return mfv[gt_slot(p - t0)]; }
    __device__ __forceinline__ uint32_t byte(uint32_t p) const { return data[p]; }
};

__global__ void __launch_bounds__(kGtWarps * 32)
dfl_parse_greedy_kernel(Batch b, DeflateParams prm, const uint32_t *__restrict__ mfull,
                        uint32_t *__restrict__ syms, BlockInfo *__restrict__ blocks, uint32_t *__restrict__ nblocks) {
    __shared__ GtWarpShared sh[kGtWarps];
    __shared__ volatile uint32_t s_entry, s_nsyms;     // handed from tile to tile: absolute entry position, symbols so far
    const uint32_t c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t off = (uint64_t)c * b.S;
    const uint8_t *data = chunk_data(b, c);
    const uint32_t n = chunk_len(b, c);
    const bool use_m = prm.mode != MODE_HUFF;
    const uint32_t *mf = mfull + off;
    uint32_t *out = syms + off;
    BlockInfo *blk = blocks + (uint64_t)c * b.MB;
    const uint32_t lo = chunk_lo(b, c), tile0 = lo / kGtTile;   // the parse enters at `lo` (history before it)
    if (threadIdx.x == 0) { s_entry = lo; s_nsyms = 0; }
    __syncthreads();
    GtWarpShared &w = sh[warp];
    const uint32_t ntiles = (n + kGtTile - 1) / kGtTile;
    for (uint32_t t = tile0 + warp; t < ntiles; t += kGtWarps) {
        const uint32_t t0 = t * kGtTile;
        __syncwarp();                                  // the previous tile's tables are no longer read
#pragma unroll
        for (uint32_t k = 0; k < kGtSeg; ++k) {        // coalesced loads, transposed (padded) stores
            const uint32_t r = k * 32 + lane, p = t0 + r;
            w.mfv[gt_slot(r)] = (use_m && p < n) ? mf[p] : 0u;
        }
        __syncwarp();
        gt_fill(lane, n - t0, w.mfv, w.lc);
        __syncwarp();
        if (t != tile0) bar_sync64(warp + 1);          // the previous tile has been crossed
        const uint32_t entry = s_entry, first_sym = s_nsyms;
        uint32_t my_entry, my_first, total;
        const uint32_t exit_rel = gt_hop(lane, entry - t0, n - t0, w.lc, my_entry, my_first, total);
        if (lane == 0) { s_entry = t0 + exit_rel; s_nsyms = first_sym + total; }
        if (t + 1 < ntiles) bar_arrive64(((warp + 1) & (kGtWarps - 1)) + 1);
        if (my_entry != 0xffffffffu) {                 // the symbols of this lane's segment
            GtAcc acc{w.mfv, data, t0};
            const uint32_t seg_end = (lane + 1) * kGtSeg, lim = n - t0;
            uint32_t r = my_entry, g = first_sym + my_first;
            while (r < seg_end && r < lim) {
                const uint32_t m = w.mfv[gt_slot(r)];
                out[g] = greedy_symbol(t0 + r, use_m, acc, g, blk, n, prm);
                ++g;
                r += m ? (m >> 16) : 1u;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) nblocks[c] = seg_finish(blk, s_nsyms, false, n, prm, chunk_final(b, c), lo);
}

// ---- phase 3, lazy rule (deflate_slow): one chunk per CTA, one segment per thread -------------------------
// zb_deflate.cuh seg_*: speculate / fix up / scan / emit.  Every thread walks its own
// 1/128 of the chunk.
// Operands of one thread's walk, a window of kPw positions at a time.  Read straight from
// global memory every 4-byte entry costs a sector through a thrashed L1 (8x traffic, 12.5 ms
// per 512 MiB at level 6); with a per-thread sector buffer refilled on demand the refill is a
// divergent branch whose load stalls the whole warp (58 % of the stall samples, 9.3 ms).  So
// the windows are WARP-SYNCHRONOUS: every thread of a warp steps through the windows of its
// own segment together; opening window w waits for its cp.async copies (issued one window
// earlier) and starts those of window w + 1, all lanes at the same instruction, and inside a
// window every operand is one ld.shared from the thread's own row.
constexpr uint32_t kPw = 16, kPwRow = kPw + 4;            // entries per window / words per row (stride keeps 16-byte alignment)
constexpr uint32_t kPwBytesRow = 8;                       // words per byte row: [start - 1, start + 16) at any alignment
struct WinAcc {
    const uint8_t *data; const uint32_t *mfull, *mquarter; uint32_t *out; uint32_t at;
    uint32_t row_f, row_b;                               // shared addresses of this thread's rows (buffer 0)
    uint64_t abs0;                                       // absolute entry index of chunk position 0
    uintptr_t in_lo, in_hi;
    uint32_t nw;
    uint32_t fadj, badj;                                 // current window: entry p at fadj + 4 p, byte p at badj + p
    uint32_t nb[kPwBytesRow - 2];                        // bytes of the next window, on their way in registers
    __device__ __forceinline__ uint32_t mf(uint32_t p) const { return lds_u32(fadj + 4u * p) & ~kQuarterDiffers; }
    __device__ __forceinline__ uint32_t mq(uint32_t p) const {                   // (the second table only where it differs: rare)
        const uint32_t v = lds_u32(fadj + 4u * p);
        return (v & kQuarterDiffers) ? mquarter[abs0 + p] : v;
    }
    __device__ __forceinline__ uint32_t byte(uint32_t p) const { return lds_u8(badj + p); }
    __device__ __forceinline__ void put(uint32_t sym) { if (out) out[at++] = sym; }
    __device__ __forceinline__ uint32_t windows(uint32_t seg) { nw = seg / kPw + 2; return nw; }
    __device__ __forceinline__ bool any(bool b) const { return __any_sync(__activemask(), b); }
    __device__ __forceinline__ void issue(uint32_t w, uint64_t wabs) {            // cp.async the entries of window w (absolute start wabs)
        const uint32_t bo = (w & 1u) * kPwRow * 4u * kSegLanes;
#pragma unroll
        for (uint32_t v = 0; v < kPw / 4; ++v) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(row_f + bo + 16u * v), "l"(mfull + wabs + 4u * v) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void fetch_bytes(uint64_t wabs) {                  // words covering bytes [start - 1, start + kPw) -> nb[]
        const uintptr_t a = (reinterpret_cast<uintptr_t>(data) + (uintptr_t)(wabs - abs0) - 1) & ~(uintptr_t)3;
#pragma unroll
        for (uint32_t k = 0; k < kPwBytesRow - 2; ++k) {
            const uintptr_t wa = a + 4u * k;
            nb[k] = (wa >= (in_lo & ~(uintptr_t)3) && wa < in_hi) ? *reinterpret_cast<const uint32_t *>(wa) : 0u;
        }
    }
    // Called by all threads of the warp that run the pass, at the same point.
    __device__ __forceinline__ uint32_t open(uint32_t w, uint32_t s0) {
        const uint64_t first = (abs0 + s0) / kPw * kPw;                           // absolute start of window 0
        const uint64_t wabs = first + (uint64_t)w * kPw;
        if (w == 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");                  // nothing of an earlier pass still in flight
            issue(0, wabs);
            fetch_bytes(wabs);
        }
        // bytes of this window: registers -> row
#pragma unroll
        for (uint32_t k = 0; k < kPwBytesRow - 2; ++k) asm volatile("st.shared.u32 [%0], %1;" ::"r"(row_b + 4u * k), "r"(nb[k]) : "memory");
        if (w + 1 < nw) {
            issue(w + 1, wabs + kPw);
            fetch_bytes(wabs + kPw);
            asm volatile("cp.async.wait_group 1;" ::: "memory");                  // window w has landed
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const uint32_t bo = (w & 1u) * kPwRow * 4u * kSegLanes;
        const uint32_t wrel = (uint32_t)(wabs - abs0);                            // chunk-relative window start (wraps below 0 for the first window)
        fadj = row_f + bo - 4u * wrel;
        const uintptr_t a = (reinterpret_cast<uintptr_t>(data) + (uintptr_t)(wabs - abs0) - 1) & ~(uintptr_t)3;
        badj = row_b - (uint32_t)(a - reinterpret_cast<uintptr_t>(data));         // byte p sits at row_b + (addr(p) - a)
        return wrel + kPw;
    }
};

constexpr uint32_t kParseSmem = (2 * kPwRow + kPwBytesRow) * kSegLanes * 4 + (kSegRecs - 1) * kSegLanes * 8 + 2 * kSegLanes * 4 + 64;
__global__ void __launch_bounds__(kSegLanes)
dfl_parse_kernel(Batch b, DeflateParams prm, const uint32_t *__restrict__ mfull,
                 const uint32_t *__restrict__ mquarter, uint32_t *__restrict__ syms,
                 BlockInfo *__restrict__ blocks, uint32_t *__restrict__ nblocks) {
    extern __shared__ __align__(16) uint8_t ps_smem[];             // 60 KiB: over the static limit, carved by hand
    uint32_t *s_f = reinterpret_cast<uint32_t *>(ps_smem);         // [2][kSegLanes][kPwRow]
    uint32_t *s_b = s_f + 2 * kPwRow * kSegLanes;                  // [kSegLanes][kPwBytesRow]
    SegRec *rec = reinterpret_cast<SegRec *>(s_b + kPwBytesRow * kSegLanes);   // [(kSegRecs - 1)][kSegLanes]
    uint32_t *s_p = reinterpret_cast<uint32_t *>(rec + (kSegRecs - 1) * kSegLanes);
    uint32_t *s_w0 = s_p + kSegLanes, *s_wsum = s_w0 + kSegLanes;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = blockIdx.x;
    const uint64_t off = (uint64_t)c * b.S;
    const uint8_t *data = chunk_data(b, c);
    const uint32_t n = chunk_len(b, c);
    uint32_t *out = syms + off;
    BlockInfo *blk = blocks + (uint64_t)c * b.MB;
    const SegGeom g = seg_geometry(n, chunk_lo(b, c));
    const bool active = tid < g.nact;

    SegLane r;
    r.start = r.end = r.spec_end = seg_cold(g.lo); r.count = r.spec_count = 0;
    WinAcc acc;
    acc.data = data; acc.mfull = mfull; acc.mquarter = prm.need_quarter ? mquarter : mfull; acc.out = nullptr; acc.at = 0;
    acc.row_f = (uint32_t)__cvta_generic_to_shared(s_f + tid * kPwRow);
    acc.row_b = (uint32_t)__cvta_generic_to_shared(s_b + tid * kPwBytesRow);
    acc.abs0 = off; acc.nw = 0;
    acc.in_lo = reinterpret_cast<uintptr_t>(b.in); acc.in_hi = acc.in_lo + b.bytes;
    if (active) seg_speculate(r, tid, g, n, prm, acc, rec);
    for (;;) {                                                     // until no start moves
        s_p[tid] = r.end.p; s_w0[tid] = r.end.w0;
        __syncthreads();
        SegState t = r.start;
        if (tid) { t.p = s_p[tid - 1]; t.w0 = s_w0[tid - 1]; }
        const bool need = tid > 0 && active && (t.p != r.start.p || t.w0 != r.start.w0);
        if (!__syncthreads_or(need)) break;
        if (need) seg_fix(r, tid, g, n, prm, acc, rec, t);
    }
    // exclusive scan of the symbol counts over the CTA
    const uint32_t cnt = active ? r.count : 0u;
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(full, inc, d); if (lane >= (uint32_t)d) inc += y; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t k = 0; k < kSegLanes / 32; ++k) { const uint32_t x = s_wsum[k]; if (k < warp) before += x; total += x; }
    const uint32_t first = before + inc - cnt;
    acc.out = out; acc.at = first;
    if (active) seg_emit(r, tid, g, n, prm, acc, blk, first);
    __syncthreads();                                               // provisional BlockInfo of every thread -> visible to thread 0
    if (tid == 0) {
        const bool pending = prm.mode == MODE_SLOW && ((s_w0[g.nact - 1] >> 25) & 1u);
        if (pending) out[total] = data[n - 1];
        nblocks[c] = seg_finish(blk, total, pending, n, prm, chunk_final(b, c), g.lo);
    }
}


// ---- phase 3, lazy rule, ONE long chunk parsed by many CTAs -------------------------------------------------------
// compress2() / deflate(Z_FINISH) emit one run of blocks however long the input is (compress.c:22-59), and the parse of
// a run is one state machine from its first byte to its last.  The segment scheme above does not care how many segments
// there are, only that lane i+1 learns where lane i really ended — so a long chunk gets G CTAs of kSegLanes segments each
// and the hand-over crosses the CTAs as well: CTA k (tickets are drawn in order, so every CTA before it is running or
// done) settles its own segments from a cold start, waits for the LINK of CTA k-1 — its true end state and the number of
// symbols before CTA k — lets its first lane fall into step with that (a re-parse of a few dozen bytes, which as a rule
// changes nothing behind it), publishes its own link, and emits its symbols at their final indices.  The chain of links
// costs a few microseconds per CTA; everything else runs at once.  The last CTA to finish closes the block table.
// Two chains: the end state of a CTA's last segment (`ready`) — the one whose latency counts — and, behind it, the number of
// symbols up to and including the CTA (`counted`).
// The state chain is walked TWICE: a CTA publishes the end state it reaches from its cold start at once (`pready`), its
// successor falls into step with that while the true states are still on their way, and when the true one arrives and
// is the same — the rule: a parse forgets its start within a segment — the successor's own state is true as it stands.
// A link of the chain then costs a flag's round trip instead of a re-parse (6-18 us per CTA before).
struct ParseLink { uint32_t p, w0, ready, syms, counted, pp, pw0, pready; };
__global__ void __launch_bounds__(kSegLanes)
dfl_parse_multi_kernel(Batch b, DeflateParams prm, const uint32_t *__restrict__ mfull,
                       const uint32_t *__restrict__ mquarter, uint32_t *__restrict__ syms,
                       BlockInfo *__restrict__ blocks, uint32_t *__restrict__ nblocks, uint32_t G,
                       unsigned int *__restrict__ mstate, volatile ParseLink *links) {
    extern __shared__ __align__(16) uint8_t ps_smem[];
    uint32_t *s_f = reinterpret_cast<uint32_t *>(ps_smem);         // [2][kSegLanes][kPwRow]
    uint32_t *s_b = s_f + 2 * kPwRow * kSegLanes;                  // [kSegLanes][kPwBytesRow]
    SegRec *rec = reinterpret_cast<SegRec *>(s_b + kPwBytesRow * kSegLanes);   // [(kSegRecs - 1)][kSegLanes]
    uint32_t *s_p = reinterpret_cast<uint32_t *>(rec + (kSegRecs - 1) * kSegLanes);
    uint32_t *s_w0 = s_p + kSegLanes, *s_wsum = s_w0 + kSegLanes;
    __shared__ uint32_t s_k, s_pred[3];
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_k = atomicAdd(&mstate[0], 1u);
    __syncthreads();
    const uint32_t k = s_k;
    const uint8_t *data = b.in;
    const uint32_t n = chunk_len(b, 0), lo = b.skip;
    uint32_t *out = syms;
    SegGeom g;
    {
        const uint32_t m = n - lo;
        const uint64_t lanes = (uint64_t)G * kSegLanes * kSegRecs;
        g.blk = (uint32_t)((m + lanes - 1) / lanes);
        if (g.blk < 32) g.blk = 32;
        g.seg = g.blk * kSegRecs;
        const uint32_t nact_total = m ? (m + g.seg - 1) / g.seg : 1;
        const uint64_t first = (uint64_t)k * kSegLanes;
        g.nact = nact_total > first ? (nact_total - first < kSegLanes ? (uint32_t)(nact_total - first) : kSegLanes) : 0u;
        g.lo = g.nact ? lo + (uint32_t)first * g.seg : n;
    }
    const bool active = tid < g.nact;

    SegLane r;
    r.start = r.end = r.spec_end = seg_cold(g.lo); r.count = r.spec_count = 0;
    WinAcc acc;
    acc.data = data; acc.mfull = mfull; acc.mquarter = prm.need_quarter ? mquarter : mfull; acc.out = nullptr; acc.at = 0;
    acc.row_f = (uint32_t)__cvta_generic_to_shared(s_f + tid * kPwRow);
    acc.row_b = (uint32_t)__cvta_generic_to_shared(s_b + tid * kPwBytesRow);
    acc.abs0 = 0; acc.nw = 0;
    acc.in_lo = reinterpret_cast<uintptr_t>(b.in); acc.in_hi = acc.in_lo + b.bytes;
    if (active) seg_speculate(r, tid, g, n, prm, acc, rec);
    SegState pred0 = seg_cold(g.lo);                               // where the segment before this CTA's first one ended (cold: not known yet)
    auto settle = [&]() {                                          // until no start moves
        for (;;) {
            s_p[tid] = r.end.p; s_w0[tid] = r.end.w0;
            __syncthreads();
            SegState t = pred0;
            if (tid) { t.p = s_p[tid - 1]; t.w0 = s_w0[tid - 1]; }
            const bool need = active && (t.p != r.start.p || t.w0 != r.start.w0);
            if (!__syncthreads_or(need)) break;
            if (need) seg_fix(r, tid, g, n, prm, acc, rec, t);
        }
    };
    settle();
    if (tid == 0) {                                                // what this CTA ends in when started cold
        links[k].pp = g.nact ? s_p[g.nact - 1] : pred0.p;
        links[k].pw0 = g.nact ? s_w0[g.nact - 1] : pred0.w0;
        __threadfence();
        links[k].pready = 1u;
    }
    if (k > 0) {
        if (tid == 0) {
            while (links[k - 1].pready == 0u) __nanosleep(32);
            __threadfence();
            s_pred[0] = links[k - 1].pp; s_pred[1] = links[k - 1].pw0;
        }
        __syncthreads();
        pred0.p = s_pred[0]; pred0.w0 = s_pred[1];
        settle();                                                  // in step with the predecessor's provisional end
        __syncthreads();
        if (tid == 0) {
            while (links[k - 1].ready == 0u) __nanosleep(32);
            __threadfence();
            s_pred[0] = links[k - 1].p; s_pred[1] = links[k - 1].w0;
        }
        __syncthreads();
        if (s_pred[0] != pred0.p || s_pred[1] != pred0.w0) {       // (uniform) the true end is another one: once more
            pred0.p = s_pred[0]; pred0.w0 = s_pred[1];
            settle();
        }
    }
    if (tid == 0) {                                                // the state link: the next CTA may go on
        links[k].p = g.nact ? s_p[g.nact - 1] : pred0.p;
        links[k].w0 = g.nact ? s_w0[g.nact - 1] : pred0.w0;
        __threadfence();
        links[k].ready = 1u;
    }
    // exclusive scan of the symbol counts over the CTA
    const uint32_t cnt = active ? r.count : 0u;
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(full, inc, d); if (lane >= (uint32_t)d) inc += y; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (uint32_t q = 0; q < kSegLanes / 32; ++q) { const uint32_t x = s_wsum[q]; if (q < warp) before += x; total += x; }
    if (tid == 0) {                                                // the count link
        uint32_t sb = 0;
        if (k > 0) {
            while (links[k - 1].counted == 0u) __nanosleep(32);
            __threadfence();
            sb = links[k - 1].syms;
        }
        links[k].syms = sb + total;
        __threadfence();
        links[k].counted = 1u;
        s_pred[2] = sb;
    }
    __syncthreads();
    const uint32_t syms_before = s_pred[2];
    const uint32_t first = syms_before + before + inc - cnt;
    acc.out = out; acc.at = first;
    if (active) seg_emit(r, tid, g, n, prm, acc, blocks, first);
    __syncthreads();
    if (tid == 0) {
        __threadfence();                                           // this CTA's symbols and provisional BlockInfo entries -> visible
        if (atomicAdd(&mstate[1], 1u) == G - 1) {                  // the last CTA to get here closes the chunk
            __threadfence();
            const uint32_t all = links[G - 1].syms;
            const bool pending = prm.mode == MODE_SLOW && ((links[G - 1].w0 >> 25) & 1u);
            if (pending) out[all] = data[n - 1];
            nblocks[0] = seg_finish(blocks, all, pending, n, prm, chunk_final(b, 0), lo);
        }
    }
}

// ---- phase 4: per-block histogram + Huffman construction -------------------------------
// One WARP per block.  The construction itself is the reference's heap algorithm — its tie-breaks (trees.c:499-501) decide
// the code lengths, so it is replayed exactly, and it is serial — which makes this kernel latency-bound on ONE thread per
// block: what counts is how many blocks an SM works on at once (a warp per block: ~40 per SM, against 16 with a
// 128-thread CTA per block) and how short the thread's dependent chain is (keyed heap, zb_deflate.cuh tree_build_fast).
// The warp's other lanes count the frequencies beforehand and copy the tables out afterwards.
constexpr int kTreeWarps = 4;
__global__ void __launch_bounds__(kTreeWarps * 32)
dfl_tree_kernel(Batch b, int strategy, const uint32_t *__restrict__ syms, const BlockInfo *__restrict__ blocks,
                const uint32_t *__restrict__ nblocks, const DeflateDeviceTables *__restrict__ tabs,
                BlockCode *__restrict__ codes) {
    __shared__ TreeWork ws[kTreeWarps];
    __shared__ StaticTrees s_st;                                   // the serial thread's operands: never a global-memory round trip
    __shared__ FormatTables s_fmt;
    const uint32_t c = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < sizeof(StaticTrees) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(&s_st)[i] = reinterpret_cast<const uint32_t *>(&tabs->st)[i];
    for (uint32_t i = threadIdx.x; i < sizeof(FormatTables) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(&s_fmt)[i] = reinterpret_cast<const uint32_t *>(&tabs->fmt)[i];
    __syncthreads();
    const uint32_t bi = blockIdx.x * kTreeWarps + warp;
    if (bi >= nblocks[c]) return;                                  // (warp-uniform; no CTA-wide barrier from here on)
    TreeWork &w = ws[warp];
    const unsigned full = 0xffffffffu;
    for (uint32_t i = lane; i < 288; i += 32) w.hkey[i] = 0;
    w.dhist[lane] = 0;
    __syncwarp(full);
    const BlockInfo blk = blocks[(uint64_t)c * b.MB + bi];
    const uint32_t *s = syms + (uint64_t)c * b.S + blk.sym_start;
    constexpr uint32_t U = 16;                                     // symbols in flight per lane: the loop is bound by load latency
    for (uint32_t base = 0; base < blk.sym_count; base += 32 * U) {
        uint32_t v[U];
#pragma unroll
        for (uint32_t k = 0; k < U; ++k) { const uint32_t i = base + k * 32 + lane; v[k] = i < blk.sym_count ? s[i] : 0xffffffffu; }
#pragma unroll
        for (uint32_t k = 0; k < U; ++k) {
            if (v[k] == 0xffffffffu) continue;                     // (no symbol has distance 65535)
            const uint32_t dist = v[k] >> 16, lc = v[k] & 0xffff;
            if (!dist) atomicAdd(&w.hkey[lc], 1u);
            else {
                atomicAdd(&w.hkey[257 + s_fmt.len_code[lc]], 1u);
                atomicAdd(&w.dhist[dist_to_code(s_fmt, dist)], 1u);
            }
        }
    }
    __syncwarp(full);
    for (uint32_t i = lane; i < 286; i += 32) w.lt[i].fc = (uint16_t)w.hkey[i];
    if (lane < 30) w.dt[lane].fc = (uint16_t)w.dhist[lane];
    __syncwarp(full);
    BlockCode &out = codes[(uint64_t)c * b.MB + bi];
    if (lane == 0) {
        w.lt[256].fc = 1;                                          // END_BLOCK (trees.c:480 init_block)
        block_build_t<true, false>(w, blk, strategy, s_st, s_fmt, out, w.hkey);
    }
    __syncwarp(full);
    const uint32_t type = out.type;
    if (type == 1) {
        for (uint32_t n = lane; n < 288; n += 32) { out.lcode[n] = s_st.lcode[n]; out.llen[n] = s_st.llen[n]; }
        out.dcode[lane] = s_st.dcode[lane]; out.dlen[lane] = s_st.dlen[lane];
    } else if (type == 2) {
        const uint32_t l_max = out.pad & 0xffffu, d_max = out.pad >> 16;
        for (uint32_t n = lane; n < 288; n += 32) {
            out.lcode[n] = n < 286 ? w.lt[n].fc : (uint16_t)0;
            out.llen[n] = n <= l_max ? (uint8_t)w.lt[n].dl : (uint8_t)0;
        }
        out.dcode[lane] = lane < 30 ? w.dt[lane].fc : (uint16_t)0;
        out.dlen[lane] = lane <= d_max ? (uint8_t)w.dt[lane].dl : (uint8_t)0;
        const uint32_t nw = (out.hdr_bits + 31) >> 5;
        for (uint32_t i = lane; i < nw; i += 32) out.hdr[i] = w.hkey[i];
    }
    __syncwarp(full);
    if (lane == 0) out.pad = 0;
}

// ---- layout: bit offsets of blocks inside their chunk, chunk sizes ---------------------
__global__ void dfl_layout_kernel(Batch b, BlockInfo *__restrict__ blocks, const BlockCode *__restrict__ codes,
                                  const uint32_t *__restrict__ nblocks, uint64_t *__restrict__ chunk_bytes,
                                  uint32_t member_overhead, uint32_t *__restrict__ bi_used) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.nb) return;
    uint64_t bit = c == 0 ? b.first_bit : 0;                       // deflatePrime: bits already in the first byte
    const uint32_t nbk = nblocks[c];
    for (uint32_t i = 0; i < nbk; ++i) {
        BlockInfo &bi = blocks[(uint64_t)c * b.MB + i];
        bi.bit_start_lo = (uint32_t)bit; bi.bit_start_hi = (uint32_t)(bit >> 32);
        if (i + 1 == nbk && c + 1 == b.nb) {                       // deflateUsed: bits in use in the stream's last byte (trees.c:187, deflate.c:1758)
            const BlockCode &bc = codes[(uint64_t)c * b.MB + i];
            *bi_used = (bc.type == 0 || !chunk_final(b, c)) ? 8u : (uint32_t)(((bit + bc.body_bits - 1) & 7) + 1);
        }
        bit = block_end_bit(bi, codes[(uint64_t)c * b.MB + i], bit);
    }
    if (!chunk_final(b, c)) bit = ((bit + 3 + 7) & ~7ull) + 32;   // 000 + pad + 00 00 FF FF
    chunk_bytes[c] = ((bit + 7) >> 3) + member_overhead;
}

// The same for ONE long chunk with thousands of blocks: one CTA scans them.  Where a block ends is a function of where it
// starts — start + its bits, or, for a stored block, the next byte boundary behind its 3 header bits + 32 + 8 bytes per
// byte (block_end_bit) — and functions of the form x -> x + d and x -> ((x + c) & ~7) + d are closed under composition, so
// the ends come out of an ordinary prefix scan over (aligning?, c, d) instead of a walk from block to block.
struct EndFn { uint32_t al; uint64_t c, d; };
__device__ __forceinline__ EndFn endfn_of(const BlockInfo &bi, const BlockCode &bc) {
    EndFn f;
    if (bc.type == 0) { f.al = 1; f.c = 10; f.d = 32 + 8ull * bi.byte_len; }
    else { f.al = 0; f.c = 0; f.d = bc.body_bits; }
    return f;
}
__device__ __forceinline__ EndFn endfn_then(const EndFn &f, const EndFn &g) {   // first f, then g
    EndFn h;
    if (!g.al) { h.al = f.al; h.c = f.c; h.d = f.d + g.d; }
    else if (!f.al) { h.al = 1; h.c = f.d + g.c; h.d = g.d; }
    else { h.al = 1; h.c = f.c; h.d = ((f.d + g.c) & ~7ull) + g.d; }
    return h;
}
__device__ __forceinline__ uint64_t endfn_apply(const EndFn &f, uint64_t x) { return f.al ? ((x + f.c) & ~7ull) + f.d : x + f.d; }
__global__ void __launch_bounds__(1024)
dfl_layout_long_kernel(Batch b, BlockInfo *__restrict__ blocks, const BlockCode *__restrict__ codes,
                       const uint32_t *__restrict__ nblocks, uint64_t *__restrict__ chunk_bytes,
                       uint32_t member_overhead, uint32_t *__restrict__ bi_used) {
    __shared__ EndFn s_fn[1024];
    const uint32_t nbk = nblocks[0], tid = threadIdx.x;
    const uint32_t per = (nbk + 1023) / 1024, i0 = tid * per, i1 = i0 + per < nbk ? i0 + per : nbk;
    EndFn mine; mine.al = 0; mine.c = 0; mine.d = 0;                // identity
    for (uint32_t i = i0; i < i1; ++i) mine = endfn_then(mine, endfn_of(blocks[i], codes[i]));
    s_fn[tid] = mine;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {                       // inclusive scan (Hillis-Steele) of the composed functions
        EndFn prev; prev.al = 0; prev.c = 0; prev.d = 0;
        if (tid >= d) prev = s_fn[tid - d];
        __syncthreads();
        if (tid >= d) s_fn[tid] = endfn_then(prev, s_fn[tid]);
        __syncthreads();
    }
    const uint64_t bit0 = b.first_bit;                              // deflatePrime: bits already in the first byte
    uint64_t bit = tid ? endfn_apply(s_fn[tid - 1], bit0) : bit0;
    for (uint32_t i = i0; i < i1; ++i) {
        BlockInfo &bi = blocks[i];
        bi.bit_start_lo = (uint32_t)bit; bi.bit_start_hi = (uint32_t)(bit >> 32);
        if (i + 1 == nbk) {                                         // the chunk's last block: deflateUsed, windup, the chunk's size
            const BlockCode &bc = codes[i];
            *bi_used = (bc.type == 0 || !chunk_final(b, 0)) ? 8u : (uint32_t)(((bit + bc.body_bits - 1) & 7) + 1);
            uint64_t end = block_end_bit(bi, bc, bit);
            if (!chunk_final(b, 0)) end = ((end + 3 + 7) & ~7ull) + 32;
            chunk_bytes[0] = ((end + 7) >> 3) + member_overhead;
        }
        bit = block_end_bit(bi, codes[i], bit);
    }
}

// Exclusive scan of the chunk sizes of a sub-batch onto the running stream length.
__global__ void __launch_bounds__(1024)
dfl_scan_kernel(uint32_t nb, const uint64_t *__restrict__ chunk_bytes, uint64_t *__restrict__ chunk_off,
                uint64_t *running, uint64_t *chunk_end_out) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = *running;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < nb ? chunk_bytes[i] : 0;
        uint64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t s = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint64_t y = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += y; }
            wsum[lane] = s;
        }
        __syncthreads();
        const uint64_t before = carry + (warp ? wsum[warp - 1] : 0) + (x - v);
        if (i < nb) { chunk_off[i] = before; if (chunk_end_out) chunk_end_out[i] = before + v; }
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *running = carry;
}

// ---- phase 5: bit packing ------------------------------------------------------------
// (Round 2 measured the alternative the review suggested — symbols ORed into a shared-memory staging area, complete words
// leaving by plain coalesced stores, global atomics only for the two words a block shares with its neighbours: correct,
// but 1.00 ms against 0.84 ms per 512 MiB at level 1.  The global atomicOr is a fire-and-forget reduction at the L2; the
// staged form pays four CTA barriers per 256 symbols and the same number of shared-memory atomics.  Dropped.)
__device__ __forceinline__ void or_bits(uint32_t *out32, uint64_t bit, uint64_t v, uint32_t nb) {
    if (nb == 0) return;
    const uint64_t w = bit >> 5;
    const uint32_t sh = (uint32_t)bit & 31u;
    const uint64_t lo = v << sh;
    const uint32_t a = (uint32_t)lo, bb = (uint32_t)(lo >> 32);
    if (a) atomicOr(out32 + w, a);
    if (bb) atomicOr(out32 + w + 1, bb);
    if (sh && nb + sh > 64) { const uint32_t cc = (uint32_t)(v >> (64 - sh)); if (cc) atomicOr(out32 + w + 2, cc); }
}

__global__ void __launch_bounds__(256)
dfl_pack_kernel(Batch b, const uint32_t *__restrict__ syms, const BlockInfo *__restrict__ blocks,
                const BlockCode *__restrict__ codes, const uint32_t *__restrict__ nblocks,
                const uint64_t *__restrict__ chunk_off, const uint64_t *__restrict__ chunk_bytes,
                const DeflateDeviceTables *__restrict__ tabs, uint32_t *__restrict__ out32,
                uint32_t member_header) {
    const uint32_t c = blockIdx.y, bi = blockIdx.x;
    const uint32_t nbk = nblocks[c];
    if (bi > nbk) return;
    const uint64_t chunk_bit0 = (chunk_off[c] + member_header) * 8;
    if (bi == nbk) {                                               // sync marker after the last block
        if (chunk_final(b, c) || threadIdx.x != 0) return;
        // the marker's only set bits are the FF FF of its last two bytes
        const uint64_t end_bit = (chunk_off[c] + chunk_bytes[c]) * 8;
        or_bits(out32, end_bit - 16, 0xffffu, 16);
        return;
    }
    __shared__ uint16_t s_lcode[288]; __shared__ uint8_t s_llen[288];
    __shared__ uint16_t s_dcode[32];  __shared__ uint8_t s_dlen[32];
    __shared__ uint32_t s_warp[8];
    const BlockInfo blk = blocks[(uint64_t)c * b.MB + bi];
    const BlockCode &code = codes[(uint64_t)c * b.MB + bi];
    const uint32_t type = code.type;
    const uint64_t start = chunk_bit0 + (((uint64_t)blk.bit_start_hi << 32) | blk.bit_start_lo);
    if (threadIdx.x == 0) or_bits(out32, start, (blk.flags & BLK_LAST) | (type << 1), 3);
    if (type == 0) {                                               // stored: trees.c:860-875
        const uint64_t body = (start + 3 + 7) & ~7ull;
        if (threadIdx.x == 0)
            or_bits(out32, body, (uint64_t)(blk.byte_len & 0xffff) | ((uint64_t)(~blk.byte_len & 0xffff) << 16), 32);
        const uint8_t *src = chunk_data(b, c) + blk.byte_start;
        for (uint32_t i = threadIdx.x; i < blk.byte_len; i += 256) or_bits(out32, body + 32 + 8ull * i, src[i], 8);
        return;
    }
    for (int i = threadIdx.x; i < 288; i += 256) { s_lcode[i] = code.lcode[i]; s_llen[i] = code.llen[i]; }
    if (threadIdx.x < 32) { s_dcode[threadIdx.x] = code.dcode[threadIdx.x]; s_dlen[threadIdx.x] = code.dlen[threadIdx.x]; }
    const uint32_t hdr_bits = code.hdr_bits;
    for (uint32_t wd = threadIdx.x; wd * 32 < hdr_bits; wd += 256) {
        const uint32_t left = hdr_bits - wd * 32;
        or_bits(out32, start + 3 + 32ull * wd, code.hdr[wd], left < 32 ? left : 32);
    }
    __syncthreads();
    uint64_t cursor = start + 3 + hdr_bits;
    const uint32_t *s = syms + (uint64_t)c * b.S + blk.sym_start;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < blk.sym_count; base += 256) {
        const uint32_t i = base + threadIdx.x;
        uint32_t nb = 0;
        uint64_t v = 0;
        if (i < blk.sym_count) v = symbol_bits(s[i], s_lcode, s_llen, s_dcode, s_dlen, tabs->fmt, nb);
        uint32_t x = nb;                                           // inclusive scan of bit lengths
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const uint32_t t = s_warp[k]; if (k < warp) before += t; total += t; }
        or_bits(out32, cursor + before + (x - nb), v, nb);
        cursor += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) or_bits(out32, cursor, s_lcode[256], s_llen[256]);   // END_BLOCK
}

// ---- framing ---------------------------------------------------------------------------
// Stream header at offset 0 and trailer after the last chunk (deflate.c:1004-1054,1239-1256).
__global__ void dfl_frame_kernel(uint8_t *out, int frame, int level, int strategy, int finish, int write_header,
                                 uint64_t *running, const uint32_t *sums /* crc, adler */, uint64_t n,
                                 uint64_t *total_out, int window_bits) {
    if (threadIdx.x || blockIdx.x) return;
    if (write_header) {
        if (frame == ZB200_FRAME_ZLIB) {
            const uint32_t lf = (strategy >= STRAT_HUFFMAN || level < 2) ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3;
            uint32_t hdr = ((8u + ((uint32_t)(window_bits - 8) << 4)) << 8) | (lf << 6);   // deflate.c:1006: Z_DEFLATED + (w_bits - 8) << 4
            hdr += 31 - hdr % 31;
            out[0] = (uint8_t)(hdr >> 8); out[1] = (uint8_t)hdr;
        } else if (frame == ZB200_FRAME_GZIP) {
            const uint8_t g[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0,
                                   (uint8_t)(level == 9 ? 2 : (strategy >= STRAT_HUFFMAN || level < 2) ? 4 : 0), 3};
            for (int i = 0; i < 10; ++i) out[i] = g[i];
        }
    }
    uint64_t pos = *running;
    if (finish) {
        if (frame == ZB200_FRAME_ZLIB) {
            const uint32_t a = sums[1];
            out[pos++] = (uint8_t)(a >> 24); out[pos++] = (uint8_t)(a >> 16); out[pos++] = (uint8_t)(a >> 8); out[pos++] = (uint8_t)a;
        } else if (frame == ZB200_FRAME_GZIP) {
            const uint32_t cr = sums[0], l = (uint32_t)n;
            for (int i = 0; i < 4; ++i) out[pos++] = (uint8_t)(cr >> (8 * i));
            for (int i = 0; i < 4; ++i) out[pos++] = (uint8_t)(l >> (8 * i));
        }
    }
    *running = pos;
    if (total_out) *total_out = pos;
}

// One gzip member per chunk: header before, CRC-32 + ISIZE after each chunk's deflate data.
__global__ void dfl_member_frame_kernel(Batch b, uint8_t *out, int level, int strategy,
                                        const uint64_t *__restrict__ chunk_off, const uint64_t *__restrict__ chunk_bytes,
                                        const uint32_t *__restrict__ crc) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.nb) return;
    uint8_t *h = out + chunk_off[c];
    const uint8_t g[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0,
                           (uint8_t)(level == 9 ? 2 : (strategy >= STRAT_HUFFMAN || level < 2) ? 4 : 0), 3};
    for (int i = 0; i < 10; ++i) h[i] = g[i];
    uint8_t *t = out + chunk_off[c] + chunk_bytes[c] - 8;
    const uint32_t cr = crc[c], l = chunk_len(b, c);
    for (int i = 0; i < 4; ++i) t[i] = (uint8_t)(cr >> (8 * i));
    for (int i = 0; i < 4; ++i) t[4 + i] = (uint8_t)(l >> (8 * i));
}

__global__ void dfl_segments_kernel(Batch b, uint64_t first_off, uint64_t *seg_off, uint64_t *seg_len) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= b.nb) return;
    seg_off[c] = first_off + (uint64_t)c * b.S;
    seg_len[c] = chunk_len(b, c);
}

// ---------------------------------------------------------------------------
int deflate_init(zb200_ctx *ctx) {
    static DeflateDeviceTables h;
    static std::once_flag once;
    std::call_once(once, [] { memset(&h, 0, sizeof h); format_fill(h.fmt); static_trees_fill(h.st); });
    void *d = nullptr;
    ZB_CUDA(cudaMalloc(&d, sizeof h));
    ZB_CUDA(cudaMemcpy(d, &h, sizeof h, cudaMemcpyHostToDevice));
    ctx->d_deflate_tables = d;
    ZB_CUDA(cudaFuncSetAttribute(dfl_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 << kHashBitsMax));
    ZB_CUDA(cudaFuncSetAttribute(dfl_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 << kHashBitsMax));
    ZB_CUDA(cudaFuncSetAttribute(dfl_chain_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ZB_CUDA(cudaFuncSetAttribute(dfl_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kParseSmem));
    ZB_CUDA(cudaFuncSetAttribute(dfl_parse_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kParseSmem));
    ZB_CUDA(cudaFuncSetAttribute(dfl_parse_greedy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ZB_CUDA(cudaFuncSetAttribute(dfl_tree_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsSmem));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_uniform_kernel<4, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsDataBytes));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_uniform_kernel<4, 8, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_uniform_kernel<2, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsDataBytes));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_uniform_kernel<2, 8, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_uniform_kernel<8, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMsSmem));
    ZB_CUDA(cudaFuncSetAttribute(dfl_match_uniform_kernel<8, 16, false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    return ZB200_OK;
}

static size_t chunk_bound(size_t len) {
    // fixed-Huffman worst case (deflate.c:852-855) + per-block and marker slack (blocks of 127 symbols at memLevel 1:
    // a stored block costs 5 bytes per 127, well inside the len/8 term)
    return len + (len >> 3) + (len >> 8) + (len >> 9) + 4 + 8 * (len / 16383 + 2) + 16;
}

static size_t frame_overhead(int frame) { return frame == ZB200_FRAME_ZLIB ? 6 : frame == ZB200_FRAME_GZIP ? 18 : 0; }

constexpr size_t kBatchBytes = 512u << 20;    // input bytes per sub-batch (bounds scratch at ~14x this)

static size_t batch_chunks(size_t n, size_t S, uint32_t MB) {
    size_t nch = (n + S - 1) / S;
    if (nch == 0) nch = 1;
    size_t nb = kBatchBytes / S;
    // small memLevels cut a chunk into many blocks (127 symbols each at memLevel 1): keep the block tables (1.6 KiB per
    // block slot) of a sub-batch under 2 GiB
    const size_t per_chunk = (size_t)MB * (sizeof(BlockInfo) + sizeof(BlockCode));
    const size_t cap = ((size_t)2 << 30) / per_chunk;
    if (nb > cap) nb = cap;
    if (nb == 0) nb = 1;
    if (nb > 32768) nb = 32768;                 // gridDim.y limit of the per-chunk kernels
    return nb < nch ? nb : nch;
}

size_t deflate_wave_chunks(zb200_ctx *ctx, int mem_level) {
    int per_sm = 0;
    const int hb = mem_level + 7;
    const cudaError_t e = hb < 15 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dfl_chain_kernel<true>, kChainWarps * 32, 2u << hb)
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dfl_chain_kernel<false>, kChainWarps * 32, 2u << hb);
    if (e != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
    return (size_t)(per_sm > 0 ? per_sm : 1) * (size_t)ctx->sm_count;
}
size_t deflate_piece_bytes(zb200_ctx *ctx, size_t want, size_t S, int mem_level, int level) {
    size_t chunks = (want + S - 1) / S;
    if (level < 1) return (chunks ? chunks : 1) * S;
    const size_t wave = deflate_wave_chunks(ctx, level <= 2 && mem_level > 7 ? 7 : mem_level), batch = kBatchBytes / S;   // (levels 1-2 hash into 14 bits at most)
    if (chunks >= wave) {
        chunks = chunks / wave * wave;
        if (chunks > batch && batch >= wave) chunks = chunks / (batch / wave * wave) * (batch / wave * wave);   // whole sub-batches
    }
    return (chunks ? chunks : 1) * S;
}

static thread_local const int *tl_deflate_tune = nullptr;
const int *deflate_tune_override() { return tl_deflate_tune; }
void deflate_tune_set(const int *four) { tl_deflate_tune = four; }

int deflate_launch_opts(zb200_ctx *ctx, const uint8_t *d_in, size_t n, size_t S, const DeflateOpts &o, int frame,
                        int finish, uint8_t *d_out, size_t out_cap, uint64_t *d_chunk_end, uint64_t *d_total,
                        uint32_t *d_sums_out, cudaStream_t s) {
    const int level = o.level, strategy = o.strategy;
    const size_t skip = o.skip;
    if (o.window_bits < 9 || o.window_bits > 15 || o.mem_level < 1 || o.mem_level > 9 || o.first_bit > 7 ||
        (o.first_bit && frame != ZB200_FRAME_RAW)) {
        set_error("deflate: windowBits 9..15, memLevel 1..9, first bit 0..7 (raw streams only)");
        return ZB200_ERR_PARAM;
    }
    // history carried from chunk to chunk: levels 1-9 of one stream (stored blocks have no use for it, members are independent)
    const bool carry = o.carry && level >= 1 && frame != ZB200_FRAME_GZIP_MEMBERS;
    if (skip && (skip > kWSize || skip > n || (!carry && n > S) || frame != ZB200_FRAME_RAW)) {
        set_error("deflate: a preset dictionary is at most 32768 bytes at the head of a single raw chunk");
        return ZB200_ERR_PARAM;
    }
    if (level < 0 || level > 9 || strategy < 0 || strategy > 4 || frame < 0 || frame > 3 || S < 1 ||
        S > 0x40000000ull || ((uintptr_t)d_out & 3)) {
        set_error("deflate: bad parameter (level 0..9, strategy 0..4, frame 0..3, chunk 1..2^30, 4-byte aligned output)");
        return ZB200_ERR_PARAM;
    }
    const bool members = frame == ZB200_FRAME_GZIP_MEMBERS;
    const size_t n_new = n - skip;                                 // bytes to compress (a dictionary alone may exceed S when not carried: one chunk)
    size_t nch = carry ? (n_new + S - 1) / S : (n + S - 1) / S;
    if (nch == 0 && (finish || members)) nch = 1;
    if (out_cap < zb200_deflate_bound(n_new, S, frame)) { set_error("deflate: output capacity below zb200_deflate_bound()"); return ZB200_ERR_OUTPUT; }
    DeflateParams prm = deflate_params(level, strategy, o.window_bits, o.mem_level);
    const int *tune = deflate_tune_override();                     // deflateTune (deflate.c:805-816): this thread's next calls
    if (tune && (prm.mode == MODE_SLOW || prm.mode == MODE_FAST)) {
        prm.good = tune[0]; prm.lazy = tune[1]; prm.nice = tune[2]; prm.chain = tune[3];
        prm.need_quarter = (prm.mode == MODE_SLOW && prm.good < prm.lazy) ? 1 : 0;
    }
    // Levels 1 and 2 hash into 14 bits, whatever memLevel says (deflate.c:444 hash_bits = memLevel + 7 = 15 by default):
    // a 32 KiB head table lets six chunks share an SM where 64 KiB allows three, and the chain kernel — bound by the latency
    // of its ordered steps, not by issue slots — goes from 3.33 to 1.99 ms per 444 MiB.  Their streams are not the
    // reference's byte for byte anyway (every position is inserted); the shorter hash costs 1.4 % of size (level 1 on
    // markov text: 0.983 -> 0.997 x the reference's size; 13 bits: 1.46 ms but 1.020 x).  Level 3 walks 32 links: the
    // fuller chains of a shorter hash cost its match kernel more (15.1 -> 18.9 ms) than the chain kernel gains, it keeps
    // memLevel's table.  $ZB200_FAST_HASH_BITS overrides.
    // ZB200_EXACT_FAST: levels 1-3 as the reference's own deflate_fast (its table's hash width, its chains); needs the head
    // table to fit the chunk's match-table slot, no history in front of the call
    const bool exact_fast = o.exact_fast && prm.mode == MODE_FAST && skip == 0 && !o.carry && S >= ((size_t)1 << prm.hash_bits);
    if (prm.mode == MODE_FAST && level <= 2 && !tune && !exact_fast) {
        static const int fast_hash = [] { const char *e = getenv("ZB200_FAST_HASH_BITS"); return e ? atoi(e) : 14; }();
        const uint32_t hb = (uint32_t)(fast_hash >= 9 && fast_hash <= 16 ? fast_hash : 14);
        if (prm.hash_bits > hb) { prm.hash_bits = hb; prm.hash_shift = (hb + kMinMatch - 1) / kMinMatch; prm.hash_mask = (1u << hb) - 1u; }
    }
    // levels 3..9: chain walks of very different lengths -> depth-sorted scheduling (levels 1-2 use the branch-free walk,
    // which is compiled for their table values)
    const bool sorted_walks = (prm.mode == MODE_SLOW || prm.mode == MODE_FAST) && (prm.level >= 3 || tune);
    // (Measured and dropped: blocks of 32767 symbols at levels 1-2 — half as many Huffman constructions, tree kernel 0.91 ->
    //  0.66 ms per 444 MiB, the same size on text — cost 1-2 % of size where text and incompressible data alternate: with the
    //  two-candidate search that crossed the 3 % bound on one generator.)
    const uint32_t MB = max_blocks_for((uint32_t)S, prm.sym_limit);
    // Carried history: chunk c > 0 is worked on as [w_size bytes before it | its own bytes] — deflateSetDictionary of the
    // previous w_size bytes per chunk (what pigz does; deflate.c:550-632 keeps the last w_size bytes of a dictionary) — so
    // its working arrays hold E = S + w_size positions.  The kernels see the same picture as for a preset dictionary.
    // (a dictionary longer than a small window is chunk 0's history all the same: its slot is as long as that takes)
    const size_t W = carry ? prm.w_size : 0, E = S + (carry && skip > W ? skip : W);
    if (E > 0x40000000ull) { set_error("deflate: carried history: chunk + window at most 2^30 bytes"); return ZB200_ERR_PARAM; }
    const size_t nb_max = carry ? batch_chunks((nch ? nch : 1) * E, E, MB) : batch_chunks(n, S, MB);
    // Two sub-batches in flight (a call of at least two): sub-batch j runs on stream j & 1 with its own working arrays.  The
    // ordered kernels (chain, parse) and the tree kernel are bound by latency, not by issue slots (ncu: 43-52 % issue
    // active), and the match / pack kernels by issue slots: kernels of neighbouring sub-batches fill each other's gaps.
    // Only the running stream length orders them: layout + scan of sub-batch j wait for the scan of j - 1.
    // Off while kernels are being timed one by one (zb200_profile_enable) and for gzip members; $ZB200_DUAL_STREAM=0 turns it off.
    // (Measured: three and four in flight, or sub-batches of one wave, give nothing over two: 60.8 / 60.3 / 60.8 GB/s at level 1.)
    // (Also measured: an explicit software pipeline — stage X of sub-batch j gated on stage X of j - 1 by events, so that unlike
    //  stages are paired by construction — 59.4-59.7 GB/s against 60.8 at level 1, 11.16 against 11.33 at level 6: the block
    //  scheduler's own drift pairs them as well, the gates only add bubbles.)
    // A caller that keeps two CALLS in flight itself (the host pipeline: piece k on flight slot k & 1) says so in o.slot: slot 1
    // has its own stream-length / checksum cells, its own part of the scratch and its own second stream.
    static const int dual_knob = [] { const char *e = getenv("ZB200_DUAL_STREAM"); const int v = e ? atoi(e) : 2; return v >= 1 ? 2 : 0; }();
    constexpr int kMaxFlight = 2;
    const size_t so = o.slot ? 32 : 0;                             // u64 cells of ctx->d_small per slot
    cudaStream_t const s_other = ctx->aux_stream[o.slot ? 2 : 0];
    const size_t nch_eff = nch ? nch : 1;
    const bool may_fly = dual_knob >= 2 && !ctx->prof_on && !members && s_other && level >= 1;
    // Chunks per sub-batch: a whole number of WAVES of the chain kernel (its CTAs per SM x the SM count: 444 chunks at
    // memLevel 8).  The ordered kernels take the same time for 1 chunk or a full wave, so 2048 chunks (4.6 waves) cost
    // five waves: 1776 chunks per sub-batch measured 44.2 GB/s at level 1 against 42.0 for 2048.
    // (Levels 4-9 spend their time in the match kernel, whose grid has a CTA per 16 Ki positions: no waves to speak of,
    // and one more launch per 2 GiB measured 2 % slower.)
    size_t nb_run = nb_max;
    if (prm.mode == MODE_FAST) {
        const size_t wave = deflate_wave_chunks(ctx, (int)prm.hash_bits - 7);
        if (nb_run > wave) nb_run = nb_run / wave * wave;
        static const int sub_waves = [] { const char *e = getenv("ZB200_SUB_WAVES"); return e ? atoi(e) : 0; }();
        if (sub_waves > 0 && nb_run > wave * (size_t)sub_waves) nb_run = wave * (size_t)sub_waves;
    } else if (prm.mode == MODE_SLOW && may_fly && nch_eff <= nb_run) {
        // the lazy levels know no waves: a call (or a piece of the host pipeline) that would be ONE sub-batch goes as two halves in flight
        static const size_t split_min = [] { const char *e = getenv("ZB200_SPLIT_MIN"); const long v = e ? atol(e) : 256; return (size_t)(v > 0 ? v : 0); }();
        if (split_min && nch_eff >= 2 * split_min) nb_run = (nch_eff + 1) / 2;
    }
    const bool dual = may_fly && nch_eff > nb_run;
    const int K = dual ? 2 : 1;                                    // sub-batches in flight
    const size_t wb = work_bytes(nb_run, E, MB);
    const size_t slot_base = o.slot ? o.slot_bytes : 0;
    // The caller's estimate of a slot's scratch is for memLevel 8; a call that needs more (small memLevels: many more blocks
    // per chunk) would reach into the other slot's part — it then runs alone: after everything enqueued so far, and to its end
    // before the caller enqueues the next piece.
    const bool alone = o.slot_bytes && (size_t)K * wb > o.slot_bytes;
    if (alone) ZB_CUDA(cudaDeviceSynchronize());
    int r = ensure_scratch(ctx, slot_base + (size_t)K * wb);
    if (r) return r;
    DeflateWork w_set[kMaxFlight];
    for (int k = 0; k < K; ++k) carve(w_set[k], (uint8_t *)ctx->d_scratch + slot_base + (size_t)k * wb, nb_run, E, MB);
    struct Events {                                                // [k]: scan of the latest sub-batch on stream k done; [kMaxFlight]: start / end
        cudaEvent_t e[kMaxFlight + 1] = {};
        ~Events() { for (auto x : e) if (x) cudaEventDestroy(x); }
    } evs;
    if (dual) for (auto &x : evs.e) ZB_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    cudaStream_t const s_main = s;
    auto flight_stream = [&](size_t j) { return (j % (size_t)K) ? s_other : s_main; };
    const DeflateDeviceTables *tabs = (const DeflateDeviceTables *)ctx->d_deflate_tables;
    uint64_t *running = ctx->d_small + so + 16;                    // stream length so far (device)
    uint32_t *sums = (uint32_t *)(ctx->d_small + so + 20);         // crc, adler of the whole input
    CkAccum *acc1 = (CkAccum *)(ctx->d_small + so + 24);
    uint32_t *bi_used = (uint32_t *)(ctx->d_small + so + 19);      // deflateUsed of the call's last byte
    const size_t zero_bytes = zb200_deflate_bound(n_new, S, frame);
    prof_mark(ctx, s, "memset_output");
    ZB_CUDA(cudaMemsetAsync(d_out, 0, (zero_bytes + 3) & ~(size_t)3, s));
    const uint64_t hdr = frame == ZB200_FRAME_ZLIB ? 2 : frame == ZB200_FRAME_GZIP ? 10 : 0;
    ZB_CUDA(cudaMemcpyAsync(running, &ctx->h_small[32 + (hdr == 2 ? 1 : hdr == 10 ? 2 : 0)], 8, cudaMemcpyHostToDevice, s));
    if (frame == ZB200_FRAME_ZLIB || frame == ZB200_FRAME_GZIP || d_sums_out) {
        r = checksum_launch(ctx, d_in + skip, nullptr, nullptr, n - skip, 1, ZB200_CRC32 | ZB200_ADLER32, 0, 1, sums, sums + 1, acc1, s);
        if (r) return r;
        if (d_sums_out) ZB_CUDA(cudaMemcpyAsync(d_sums_out, sums, 8, cudaMemcpyDeviceToDevice, s));
    }
    if (dual) {                                                    // the other streams start behind the memset and the first offset
        ZB_CUDA(cudaEventRecord(evs.e[kMaxFlight], s_main));
        ZB_CUDA(cudaStreamWaitEvent(s_other, evs.e[kMaxFlight], 0));
    }
    for (size_t c0 = 0; c0 < nch; c0 += nb_run) {
        const size_t j = c0 / nb_run;
        cudaStream_t const s = flight_stream(j);                   // (shadows the call's stream inside the loop)
        DeflateWork &w = w_set[j % (size_t)K];
        Batch b;
        b.nb = (uint32_t)(nch - c0 < nb_run ? nch - c0 : nb_run);
        const size_t off = c0 * S;                                 // the sub-batch's first new byte, counted from the call's
        const size_t span = (size_t)b.nb * S;
        b.skip = (uint32_t)(c0 ? (skip + off < W ? skip + off : W) : skip);   // history in front of it: carried, or the call's own
        b.in = d_in + skip + off - b.skip;
        b.bytes = b.skip + (n_new - off < span ? n_new - off : span);
        b.S = (uint32_t)E; b.step = (uint32_t)S; b.carry = (uint32_t)W; b.MB = MB;
        b.last_is_final = (finish && c0 + b.nb == nch) ? 1 : 0;
        b.all_final = members ? 1 : 0;
        b.first_bit = c0 == 0 ? o.first_bit : 0u;
        // ONE long chunk (a one-shot call's single run of blocks): many CTAs share its ordered phases
        static const int multi_knob = [] { const char *e = getenv("ZB200_MULTI_CTA_RUN"); return e ? atoi(e) : 1; }();
        const bool one_long = multi_knob && b.nb == 1 && b.bytes >= 2 * kMultiRange;
        const uint32_t n_ranges = one_long ? (uint32_t)((b.bytes + kMultiRange - 1) / kMultiRange) : 1u;
        b.range = one_long ? (uint32_t)kMultiRange : 0u;
        if (exact_fast) {
            prof_mark(ctx, s, "dfl_fast_exact_kernel");
            ZB_CUDA(cudaMemsetAsync(w.mfull, 0, (size_t)b.nb * E * 4, s));   // the head tables
            dfl_fast_exact_kernel<<<b.nb, 32, 0, s>>>(b, prm, w.prev, w.mfull, w.syms, w.blocks, w.nblocks);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        } else if (prm.mode == MODE_FAST || prm.mode == MODE_SLOW) {
            prof_mark(ctx, s, "dfl_chain_kernel");
            const unsigned chain_grid = one_long ? n_ranges : b.nb;
            if (prm.hash_bits < 15) dfl_chain_kernel<true><<<chain_grid, kChainWarps * 32, 2u << prm.hash_bits, s>>>(b, prm, w.prev);
            else dfl_chain_kernel<false><<<chain_grid, kChainWarps * 32, 2u << prm.hash_bits, s>>>(b, prm, w.prev);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        }
        if (exact_fast) {
        } else if (sorted_walks) {
            dim3 g((unsigned)((E + kMsTile - 1) / kMsTile), b.nb);
            prof_mark(ctx, s, "dfl_match_sorted_kernel");
            static const int key_knob = [] { const char *e = getenv("ZB200_KEY_CAP"); return e ? atoi(e) : 0; }();
            uint32_t key_cap = (uint32_t)(key_knob > 0 ? key_knob : prm.chain / 16);   // (level 6: 8, level 9: 256)
            key_cap = key_cap < 8 ? 8 : key_cap > 4095 ? 4095 : key_cap;
            dfl_match_sorted_kernel<<<g, kMsThreads, kMsSmem, s>>>(b, prm, w.prev, w.mfull, w.mquarter, key_cap);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        } else if (prm.mode == MODE_FAST) {                        // levels 1-2
            dim3 g((unsigned)((E + kMsTile - 1) / kMsTile), b.nb);
            prof_mark(ctx, s, "dfl_match_uniform_kernel");
            // Level 1 looks at TWO candidates per position, not the table's four (deflate.c:114 max_chain 4).  Every position is
            // inserted here (deflate_fast skips the inside of matches longer than max_insert, deflate.c:1873-1897), which alone
            // makes the stream 3.6 % smaller than the reference's; two candidates keep 1.7 % of that and cost 3.3 ms per 512 MiB
            // instead of 4.6 (measured, markov text: 4 / 3 / 2 / 1 candidates -> 0.964 / 0.970 / 0.983 / 1.021 x the reference's
            // size).  $ZB200_L1_CHAIN=4 restores the table value.
            static const int l1_chain = [] { const char *e = getenv("ZB200_L1_CHAIN"); return e ? atoi(e) : 2; }();
            // (a small memLevel's short hash fills the chains with other strings: there all four are looked at)
            if (prm.level == 1 && l1_chain != 4 && prm.hash_bits >= 14) dfl_match_uniform_kernel<2, 8, true><<<g, kMsThreads, kMsDataBytes, s>>>(b, prm, w.prev, w.mfull);
            else if (prm.level == 1) dfl_match_uniform_kernel<4, 8, true><<<g, kMsThreads, kMsDataBytes, s>>>(b, prm, w.prev, w.mfull);
            else dfl_match_uniform_kernel<8, 16, false><<<g, kMsThreads, kMsSmem, s>>>(b, prm, w.prev, w.mfull);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        } else if (prm.mode != MODE_HUFF) {
            dim3 g((unsigned)((E + 255) / 256), b.nb);
            prof_mark(ctx, s, "dfl_match_kernel");
            dfl_match_kernel<<<g, 256, 0, s>>>(b, prm, w.prev, w.mfull, w.mquarter);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        }
        if (!exact_fast) prof_mark(ctx, s, "dfl_parse_kernel");
        if (exact_fast) {
        } else if (prm.mode == MODE_SLOW && one_long) {
            ZB_CUDA(cudaMemsetAsync(w.links, 0, (size_t)n_ranges * sizeof(ParseLink), s));
            ZB_CUDA(cudaMemsetAsync(w.mstate, 0, 16, s));
            dfl_parse_multi_kernel<<<n_ranges, kSegLanes, kParseSmem, s>>>(b, prm, w.mfull, w.mquarter, w.syms, w.blocks, w.nblocks, n_ranges, w.mstate, w.links);
        } else if (prm.mode == MODE_SLOW) dfl_parse_kernel<<<b.nb, kSegLanes, kParseSmem, s>>>(b, prm, w.mfull, w.mquarter, w.syms, w.blocks, w.nblocks);
        else dfl_parse_greedy_kernel<<<b.nb, kGtWarps * 32, 0, s>>>(b, prm, w.mfull, w.syms, w.blocks, w.nblocks);
        if (!exact_fast) { ZB_LAUNCHED(); ZB_CHECK_LAUNCH(); }
        prof_mark(ctx, s, "dfl_tree_kernel");
        dfl_tree_kernel<<<dim3((MB + kTreeWarps - 1) / kTreeWarps, b.nb), kTreeWarps * 32, 0, s>>>(b, strategy | (level == 0 ? 0x100 : 0), w.syms, w.blocks, w.nblocks, tabs, w.codes);
        ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        if (dual && j > 0) ZB_CUDA(cudaStreamWaitEvent(s, evs.e[(j - 1) % (size_t)K], 0));   // the stream length up to here (and deflateUsed's order)
        prof_mark(ctx, s, "dfl_layout_kernel");
        if (one_long) dfl_layout_long_kernel<<<1, 1024, 0, s>>>(b, w.blocks, w.codes, w.nblocks, w.chunk_bytes, members ? 18 : 0, bi_used);
        else dfl_layout_kernel<<<(b.nb + 127) / 128, 128, 0, s>>>(b, w.blocks, w.codes, w.nblocks, w.chunk_bytes, members ? 18 : 0, bi_used);
        ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        prof_mark(ctx, s, "dfl_scan_kernel");
        dfl_scan_kernel<<<1, 1024, 0, s>>>(b.nb, w.chunk_bytes, w.chunk_off, running, d_chunk_end ? d_chunk_end + c0 : nullptr);
        ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        if (dual) ZB_CUDA(cudaEventRecord(evs.e[j % (size_t)K], s));
        prof_mark(ctx, s, "dfl_pack_kernel");
        dfl_pack_kernel<<<dim3(MB + 1, b.nb), 256, 0, s>>>(b, w.syms, w.blocks, w.codes, w.nblocks, w.chunk_off, w.chunk_bytes,
                                                          tabs, (uint32_t *)d_out, members ? 10 : 0);
        ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        if (members) {
            prof_mark(ctx, s, "dfl_segments_kernel");
            dfl_segments_kernel<<<(b.nb + 127) / 128, 128, 0, s>>>(b, off, w.seg_off, w.seg_len);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
            r = checksum_launch(ctx, d_in, w.seg_off, w.seg_len, 0, b.nb, ZB200_CRC32, 0, 1, w.chunk_crc, nullptr, w.acc, s);
            if (r) return r;
            prof_mark(ctx, s, "dfl_member_frame_kernel");
            dfl_member_frame_kernel<<<(b.nb + 63) / 64, 64, 0, s>>>(b, d_out, level, strategy, w.chunk_off, w.chunk_bytes, w.chunk_crc);
            ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
        }
    }
    if (dual) {                                                    // the call's stream ends behind everything the other one did
        ZB_CUDA(cudaEventRecord(evs.e[kMaxFlight], s_other));
        ZB_CUDA(cudaStreamWaitEvent(s_main, evs.e[kMaxFlight], 0));
    }
    prof_mark(ctx, s, "dfl_frame_kernel");
    dfl_frame_kernel<<<1, 32, 0, s>>>(d_out, frame, level, strategy, finish, 1, running, sums, n - skip, d_total, o.window_bits);
    ZB_LAUNCHED(); ZB_CHECK_LAUNCH();
    prof_mark(ctx, s, nullptr);
    if (alone) ZB_CUDA(cudaStreamSynchronize(s));
    return ZB200_OK;
}

}  // namespace zb

using namespace zb;

extern "C" {

size_t zb200_deflate_bound(size_t n, size_t chunk_size, int frame) {
    frame &= 0xff;                                                 // (the flags change what a chunk's matches see, not its worst case)
    if (chunk_size == 0) chunk_size = 1;
    size_t nch = (n + chunk_size - 1) / chunk_size;
    if (nch == 0) nch = 1;
    const size_t full = n / chunk_size, rem = n % chunk_size;
    size_t per_member = frame == ZB200_FRAME_GZIP_MEMBERS ? 18 : 0;
    size_t total = full * (chunk_bound(chunk_size) + per_member);
    if (rem || full == 0) total += chunk_bound(rem) + per_member;
    return (total + frame_overhead(frame) + 64 + 15) & ~(size_t)15;
}

size_t zb200_deflate_scratch_bytes(size_t n, size_t chunk_size) {
    if (chunk_size == 0 || chunk_size > 0x40000000ull) return 0;
    const uint32_t MB = max_blocks_for((uint32_t)chunk_size, 16383u);
    return work_bytes(batch_chunks(n, chunk_size, MB), chunk_size, MB);
}

int zb200_deflate_dev(zb200_ctx *ctx, const void *d_in, size_t n, size_t chunk_size, int level, int strategy,
                      int frame, int finish, void *d_out, size_t out_cap, uint64_t *d_chunk_end,
                      uint64_t *d_total, void *stream) {
    if (!ctx || (!d_in && n) || !d_out) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, pick_stream(ctx, stream));
    DeflateOpts o;
    o.level = level; o.strategy = strategy; o.carry = (frame & ZB200_CHUNK_CARRY) != 0; o.exact_fast = (frame & ZB200_EXACT_FAST) != 0;
    return deflate_launch_opts(ctx, (const uint8_t *)d_in, n, chunk_size, o, frame & 0xff, finish,
                               (uint8_t *)d_out, out_cap, d_chunk_end, d_total, nullptr, pick_stream(ctx, stream));
}

// Large inputs in pinned memory: the input is cut into up to 9 pieces (whole chunks); piece
// k+1 travels to the device and piece k-1 travels back while piece k is compressed (three
// streams, PCIe is full duplex).  Pieces are raw deflate runs — exactly the chunks a single
// call would emit — so their concatenation is the same byte stream; the stream header and
// trailer (deflate.c:1004-1054,1239-1256) are written here on the host, the input checksums
// of the pieces merged with crc32_combine / adler32_combine.
static int deflate_host_pipelined(zb200_ctx *ctx, const uint8_t *in, size_t n, size_t S, const DeflateOpts &opts,
                                  int frame, int finish, uint8_t *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc,
                                  uint32_t *bits_used) {
    const int level = opts.level, strategy = opts.strategy;
    const int pframe = frame == ZB200_FRAME_GZIP_MEMBERS ? ZB200_FRAME_GZIP_MEMBERS : ZB200_FRAME_RAW;
    // Pieces of 128 .. 512 MiB: smaller ones leave the GPU underfilled (the chain kernel needs ~450 chunks in flight),
    // larger ones lengthen the two ends of the pipeline that nothing overlaps — the first piece's way in and the last
    // piece's way out (8 GiB in 1 GiB pieces: 21 + 9 ms of 269; in 512 MiB pieces half of that).
    // Levels 3-9 take long pieces: on data of mixed kinds a chunk's cost in the ordered kernels varies several-fold, and a
    // piece of one wave (444 chunks) then lasts as long as its slowest chunk — 2 GiB of the mixed generator at level 6 in
    // 19 pieces: chain 44.6 + parse 42.8 ms against 27.9 + 27.3 in four sub-batches of 2048 chunks.
    size_t piece = level >= 3 ? n / 4 : n / 16;
    if (piece < ((size_t)128 << 20)) piece = (size_t)128 << 20;
    // (levels 1-2 run at the speed of the link: there the ends weigh more — 8 GiB at level 1 in pieces of 444 / 222 / 111 MiB:
    //  47.8 / 49.5 / 38.8 GB/s end to end, the last being half a wave of the chain kernel)
    static const size_t piece_knob = [] { const char *e = getenv("ZB200_PIPE_PIECE_MIB"); const long v = e ? atol(e) : 0; return (size_t)(v >= 16 ? v : 0) << 20; }();
    const size_t piece_max = piece_knob ? piece_knob : (size_t)(level >= 1 && level <= 2 ? 256 : 512) << 20;
    if (piece > piece_max) piece = piece_max;
    {   // never less than one wave of the chain kernel while the input holds two (a piece of 0.6 waves costs a whole one)
        const size_t wave_bytes = deflate_wave_chunks(ctx, level <= 2 && opts.mem_level > 7 ? 7 : opts.mem_level) * S;
        if (level >= 1 && piece < wave_bytes && n >= 2 * wave_bytes) piece = wave_bytes;
    }
    piece = deflate_piece_bytes(ctx, piece, S, opts.mem_level, level);
    constexpr size_t kMaxPieces = 64;
    while ((n + piece - 1) / piece > kMaxPieces) piece += (piece + S - 1) / S * S;   // (inputs beyond 32 GiB: longer pieces)
    const size_t np = (n + piece - 1) / piece;
    size_t ooff[kMaxPieces + 1], bound[kMaxPieces], total_bound = 0;
    for (size_t k = 0; k < np; ++k) {
        const size_t len = k + 1 < np ? piece : n - k * piece;
        bound[k] = zb200_deflate_bound(len, S, pframe);
        ooff[k] = total_bound; total_bound += bound[k];
    }
    int r = ensure_io(ctx, n + 16, total_bound + 16);
    if (r) return r;
    // carried history: piece k > 0 is compressed behind the last W bytes of piece k - 1, which lie just before it on the device
    const size_t W = (opts.carry && level >= 1 && pframe == ZB200_FRAME_RAW) ? (size_t)1 << opts.window_bits : 0;
    // Two PIECES in flight: piece k is compressed on flight slot k & 1 (its own stream, stream-length cells and part of the
    // scratch, DeflateOpts::slot).  A piece is short — one wave of the chain kernel at levels 1-2 — and its kernels leave the
    // GPU half idle in turn (the ordered ones wait on latency, the others on issue slots); with 37 pieces of 222 MiB the kernels
    // of an 8 GiB call summed to 176 ms of its 180: the pipeline was bound by them, not by the link.
    static const int fly_knob = [] { const char *e = getenv("ZB200_DUAL_STREAM"); return e ? atoi(e) : 2; }();
    const bool fly = fly_knob >= 1 && !ctx->prof_on && level >= 1 && ctx->aux_stream[1] && np >= 2;
    const size_t per_slot = 2 * (W ? zb200_deflate_scratch_bytes((piece / S + 1) * (S + W), S + W) : zb200_deflate_scratch_bytes(piece, S)) + 4096;
    if ((r = ensure_scratch(ctx, fly ? 2 * per_slot : per_slot))) return r;   // no reallocation (= implicit sync) mid-pipeline
    cudaEvent_t ev_in[kMaxPieces], ev_out[kMaxPieces];
    for (size_t k = 0; k < np; ++k) {
        ZB_CUDA(cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming));
        ZB_CUDA(cudaEventCreateWithFlags(&ev_out[k], cudaEventDisableTiming));
    }
    auto cleanup = [&]() { for (size_t k = 0; k < np; ++k) { cudaEventDestroy(ev_in[k]); cudaEventDestroy(ev_out[k]); } };
    cudaStream_t const s0 = ctx->stream;
    for (size_t k = 0; k < np; ++k) {
        const size_t len = k + 1 < np ? piece : n - k * piece;
        if (cudaMemcpyAsync(ctx->d_io_in + k * piece, in + k * piece, len, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess ||
            cudaEventRecord(ev_in[k], ctx->copy_stream) != cudaSuccess) { cleanup(); set_error("deflate: H2D of piece %zu failed", k); return ZB200_ERR_CUDA; }
    }
    for (size_t k = 0; k < np; ++k) {
        const size_t len = k + 1 < np ? piece : n - k * piece;
        cudaStream_t const s = fly && (k & 1) ? ctx->aux_stream[1] : s0;
        cudaStreamWaitEvent(s, ev_in[k], 0);
        DeflateOpts po = opts;
        po.skip = k ? W : 0; po.first_bit = k == 0 ? opts.first_bit : 0u;
        po.slot = fly ? (int)(k & 1) : 0; po.slot_bytes = fly ? per_slot : 0;
        r = deflate_launch_opts(ctx, ctx->d_io_in + k * piece - po.skip, len + po.skip, S, po, pframe, (finish && k + 1 == np) ? 1 : 0,
                                ctx->d_io_out + ooff[k], bound[k], nullptr, ctx->d_pipe + 2 * k, (uint32_t *)(ctx->d_pipe + 2 * k + 1), s);
        if (r) { cudaStreamSynchronize(s0); for (auto a : ctx->aux_stream) cudaStreamSynchronize(a); cudaStreamSynchronize(ctx->copy_stream); cleanup(); return r; }
        cudaMemcpyAsync(ctx->h_pipe + 2 * k, ctx->d_pipe + 2 * k, 16, cudaMemcpyDeviceToHost, s);
        cudaEventRecord(ev_out[k], s);
    }
    // header
    size_t pos = 0;
    uint8_t hdr[10];
    size_t hlen = 0;
    if (frame == ZB200_FRAME_ZLIB) {
        const unsigned lf = (strategy >= STRAT_HUFFMAN || level < 2) ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3;
        unsigned h = ((8u + ((unsigned)(opts.window_bits - 8) << 4)) << 8) | (lf << 6);
        h += 31 - h % 31;
        hdr[0] = (uint8_t)(h >> 8); hdr[1] = (uint8_t)h; hlen = 2;
    } else if (frame == ZB200_FRAME_GZIP) {
        const uint8_t g[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(level == 9 ? 2 : (strategy >= STRAT_HUFFMAN || level < 2) ? 4 : 0), 3};
        memcpy(hdr, g, 10); hlen = 10;
    }
    int rc = ZB200_OK;
    if (hlen > *out_len) rc = ZB200_ERR_OUTPUT; else { memcpy(out, hdr, hlen); pos = hlen; }
    uint32_t crc = 0, adler = 1;
    for (size_t k = 0; k < np; ++k) {
        const size_t len = k + 1 < np ? piece : n - k * piece;
        if (cudaEventSynchronize(ev_out[k]) != cudaSuccess) { rc = ZB200_ERR_CUDA; break; }
        const uint64_t total = ctx->h_pipe[2 * k];
        const uint32_t *sums = (const uint32_t *)(ctx->h_pipe + 2 * k + 1);
        crc = zb200_crc32_combine(crc, sums[0], len);
        adler = zb200_adler32_combine(adler, sums[1], (int64_t)len);
        if (rc == ZB200_OK && pos + total > *out_len) rc = ZB200_ERR_OUTPUT;
        if (rc == ZB200_OK && cudaMemcpyAsync(out + pos, ctx->d_io_out + ooff[k], (size_t)total, cudaMemcpyDeviceToHost, ctx->back_stream) != cudaSuccess) rc = ZB200_ERR_CUDA;
        pos += (size_t)total;
    }
    cudaStreamSynchronize(ctx->back_stream);
    cudaStreamSynchronize(s0);
    if (fly) for (auto a : ctx->aux_stream) cudaStreamSynchronize(a);
    if (rc == ZB200_OK && bits_used && finish) {                   // deflateUsed: left by the last piece in its slot's cell
        const size_t so = fly && ((np - 1) & 1) ? 32 : 0;
        if (cudaMemcpy(ctx->h_small, ctx->d_small + so + 19, 8, cudaMemcpyDeviceToHost) == cudaSuccess) *bits_used = ((const uint32_t *)ctx->h_small)[0];
        else rc = ZB200_ERR_CUDA;
    }
    cleanup();
    if (rc == ZB200_OK && finish) {
        uint8_t tr[8];
        size_t tlen = 0;
        if (frame == ZB200_FRAME_ZLIB) { for (int i = 0; i < 4; ++i) tr[i] = (uint8_t)(adler >> (24 - 8 * i)); tlen = 4; }
        else if (frame == ZB200_FRAME_GZIP) { for (int i = 0; i < 4; ++i) { tr[i] = (uint8_t)(crc >> (8 * i)); tr[4 + i] = (uint8_t)((uint32_t)n >> (8 * i)); } tlen = 8; }
        if (pos + tlen > *out_len) rc = ZB200_ERR_OUTPUT; else memcpy(out + pos, tr, tlen);
        pos += tlen;
    }
    if (in_crc) *in_crc = crc;
    if (in_adler) *in_adler = adler;
    if (rc == ZB200_ERR_OUTPUT) { *out_len = pos; set_error("deflate: %zu bytes do not fit the output buffer", pos); return rc; }
    if (rc != ZB200_OK) { set_error("deflate: pipelined transfer failed"); return rc; }
    *out_len = pos;
    return ZB200_OK;
}

// One host-buffer deflate call with every option (DeflateOpts).  A preset dictionary (opts.skip = dict_len > 0): `in` starts
// with the dictionary, the first chunk is compressed behind it as a single-chunk call and the rest follows as usual.
// *bits_used (optional): bits in use in the last byte written (deflateUsed).
static int deflate_host_impl(zb200_ctx *ctx, const uint8_t *in, size_t n, size_t chunk_size, const DeflateOpts &opts, int frame,
                             int finish, uint8_t *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc, uint32_t *bits_used) {
    const int level = opts.level, strategy = opts.strategy;
    if (opts.skip && !opts.carry) {
        const size_t dl = opts.skip;
        if (dl > n || dl > (size_t)kWSize || frame != ZB200_FRAME_RAW || chunk_size < 1) return ZB200_ERR_PARAM;
        const size_t body = n - dl, first = body < chunk_size ? body : chunk_size;
        DeflateOpts o1 = opts;
        size_t cap1 = *out_len;
        uint32_t a1 = 1, c1 = 0;
        const size_t S1 = dl + first ? dl + first : 1;
        const size_t bound1 = zb200_deflate_bound(dl + first, S1, ZB200_FRAME_RAW);
        int r = ensure_io(ctx, dl + first + 16, bound1 + 16);
        if (r) return r;
        cudaStream_t s = ctx->stream;
        if ((r = h2d_auto(ctx, ctx->d_io_in, in, dl + first, s))) return r;
        r = deflate_launch_opts(ctx, ctx->d_io_in, dl + first, S1, o1, ZB200_FRAME_RAW, (finish && first == body) ? 1 : 0, ctx->d_io_out, bound1,
                                nullptr, ctx->d_small + 17, (uint32_t *)(ctx->d_small + 18), s);
        if (r) return r;
        ZB_CUDA(cudaMemcpyAsync(ctx->h_small, ctx->d_small + 17, 24, cudaMemcpyDeviceToHost, s));
        ZB_CUDA(cudaStreamSynchronize(s));
        const uint64_t total = ctx->h_small[0];
        c1 = ((const uint32_t *)(ctx->h_small + 1))[0]; a1 = ((const uint32_t *)(ctx->h_small + 1))[1];
        if (bits_used) *bits_used = ((const uint32_t *)(ctx->h_small + 2))[0];
        if (total > cap1) { *out_len = (size_t)total; set_error("deflate: %llu bytes do not fit the output buffer", (unsigned long long)total); return ZB200_ERR_OUTPUT; }
        if ((r = d2h_auto(ctx, out, ctx->d_io_out, (size_t)total, s))) return r;
        ZB_CUDA(cudaStreamSynchronize(s));
        size_t pos = (size_t)total;
        if (first < body) {
            DeflateOpts o2 = opts;
            o2.skip = 0; o2.first_bit = 0;
            size_t cap2 = *out_len - pos;
            uint32_t a2 = 1, c2 = 0;
            r = deflate_host_impl(ctx, in + dl + first, body - first, chunk_size, o2, ZB200_FRAME_RAW, finish, out + pos, &cap2, &a2, &c2, bits_used);
            if (r) { if (r == ZB200_ERR_OUTPUT) *out_len = pos + cap2; return r; }
            c1 = zb200_crc32_combine(c1, c2, body - first);
            a1 = zb200_adler32_combine(a1, a2, (int64_t)(body - first));
            pos += cap2;
        }
        if (in_crc) *in_crc = c1;
        if (in_adler) *in_adler = a1;
        *out_len = pos;
        return ZB200_OK;
    }
    if (n >= ((size_t)256 << 20) && chunk_size >= 1 && chunk_size <= ((size_t)16 << 20) && level >= 0 && level <= 9 &&
        strategy >= 0 && strategy <= 4 && frame >= 0 && frame <= 3 && !opts.skip && is_pinned(in) && is_pinned(out)) {
        if (bits_used) *bits_used = 8;                             // (read back below only on the single-shot path)
        return deflate_host_pipelined(ctx, in, n, chunk_size, opts, frame, finish, out, out_len, in_adler, in_crc, bits_used);
    }
    const size_t bound = zb200_deflate_bound(n, chunk_size ? chunk_size : 1, frame);
    int r = ensure_io(ctx, n + 16, bound + 16);
    if (r) return r;
    cudaStream_t s = ctx->stream;
    if ((r = h2d_auto(ctx, ctx->d_io_in, in, n, s))) return r;
    uint64_t *d_total = ctx->d_small + 17;
    uint32_t *d_sums = (uint32_t *)(ctx->d_small + 18);
    r = deflate_launch_opts(ctx, ctx->d_io_in, n, chunk_size, opts, frame, finish, ctx->d_io_out, bound, nullptr, d_total, d_sums, s);
    if (r) return r;
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, ctx->d_small + 17, 24, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint64_t total = ctx->h_small[0];
    const uint32_t *hs = (const uint32_t *)(ctx->h_small + 1);
    if (in_crc) *in_crc = hs[0];
    if (in_adler) *in_adler = hs[1];
    if (bits_used) *bits_used = ((const uint32_t *)(ctx->h_small + 2))[0];
    if (total > *out_len) { *out_len = (size_t)total; set_error("deflate: %llu bytes do not fit the output buffer", (unsigned long long)total); return ZB200_ERR_OUTPUT; }
    if ((r = d2h_auto(ctx, out, ctx->d_io_out, (size_t)total, s))) return r;
    ZB_CUDA(cudaStreamSynchronize(s));
    *out_len = (size_t)total;
    return ZB200_OK;
}

int zb200_deflate_host(zb200_ctx *ctx, const void *in, size_t n, size_t chunk_size, int level, int strategy,
                       int frame, int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc) {
    if (!ctx || (!in && n) || !out || !out_len) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    DeflateOpts o;
    o.level = level; o.strategy = strategy; o.carry = (frame & ZB200_CHUNK_CARRY) != 0; o.exact_fast = (frame & ZB200_EXACT_FAST) != 0;
    return deflate_host_impl(ctx, (const uint8_t *)(in ? in : (const void *)""), n, chunk_size, o, frame & 0xff, finish, (uint8_t *)out, out_len, in_adler, in_crc, nullptr);
}

int zb200_deflate_host_opts(zb200_ctx *ctx, const void *in, size_t n, size_t chunk_size, const zb200_deflate_opts *opts,
                            int frame, int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc,
                            uint32_t *bits_used) {
    if (!ctx || (!in && n) || !out || !out_len || !opts) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    DeflateOpts o;
    o.level = opts->level; o.strategy = opts->strategy; o.window_bits = opts->window_bits ? opts->window_bits : 15;
    o.mem_level = opts->mem_level ? opts->mem_level : 8; o.skip = opts->dict_len; o.first_bit = opts->first_bit;
    o.carry = (frame & ZB200_CHUNK_CARRY) != 0; o.exact_fast = (frame & ZB200_EXACT_FAST) != 0;
    return deflate_host_impl(ctx, (const uint8_t *)(in ? in : (const void *)""), n, chunk_size, o, frame & 0xff, finish, (uint8_t *)out, out_len, in_adler, in_crc, bits_used);
}

int zb200_deflate_host_dict(zb200_ctx *ctx, const void *in, size_t n, size_t dict_len, int level, int strategy,
                            int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc) {
    if (!ctx || !in || !out || !out_len || dict_len > n || dict_len > (size_t)kWSize || n > 0x40000000ull) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    DeflateOpts o;
    o.level = level; o.strategy = strategy; o.skip = dict_len;
    return deflate_host_impl(ctx, (const uint8_t *)in, n, n - dict_len ? n - dict_len : 1, o, ZB200_FRAME_RAW, finish, (uint8_t *)out, out_len, in_adler, in_crc, nullptr);
}

}  // extern "C"
