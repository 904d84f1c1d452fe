// zb_zlib_api.cu — the zlib.h-compatible host layer (include/zb200_zlib.h) over
// the engine.  Mirrors the reference's interface for the hot path: same names,
// argument meaning and return codes (zlib.h, deflate.c:371-430,954-1263,
// inflate.c:141-250,590-1264, compress.c, uncompr.c), so the reference's own
// callers (examples/zpipe.c, src/wasm_module.c) link against it unchanged.
// Host code only: buffering, framing, return-code mapping.  All compression,
// decompression and checksumming is done by the kernels.
#include "zb_internal.h"
#include "../../include/zb200_zlib.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <deque>
#include <memory>
#include <utility>
#include <new>

using namespace zb;

namespace {

// The engine contexts behind the zlib API (device: $ZB200_DEVICE, default 0).  A context owns one set of streams, pinned
// stages and device buffers, and its calls are serialised (zb_internal.h CtxUse) — so host threads that call the zlib names
// at the same time are spread over a small POOL of contexts: a thread is given a slot the first time it calls (round robin
// over $ZB200_CONTEXTS slots, default 4, at most 16) and keeps it; a slot's context is created when its first thread
// arrives (a single-threaded process only ever has slot 0).  Every zlib.h call finishes its device work before it returns,
// so a z_stream may move between threads — and so between contexts — from call to call, as with the reference (zlib.h:
// "thread safe" = one thread per stream at a time, FAQ:151-160).
constexpr int kMaxContexts = 16;
std::mutex g_mu;
zb200_ctx *g_pool[kMaxContexts] = {nullptr};
int g_ctx_err = 0;
std::atomic<unsigned> g_threads{0};

int pool_slots() {
    static const int n = [] { const char *e = getenv("ZB200_CONTEXTS"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : v > kMaxContexts ? kMaxContexts : v; }();
    return n;
}

zb200_ctx *api_ctx() {
    thread_local int slot = -1;
    if (slot < 0) slot = (int)(g_threads.fetch_add(1) % (unsigned)pool_slots());
    std::lock_guard<std::mutex> g(g_mu);
    if (g_ctx_err) return nullptr;                              // no usable device: said once, true for every thread
    auto make = [](int *err) -> zb200_ctx * {
        const char *e = getenv("ZB200_DEVICE");
        zb200_ctx *c = nullptr;
        *err = zb200_create(e ? atoi(e) : 0, &c);
        return *err == ZB200_OK ? c : nullptr;
    };
    int err = 0;
    if (!g_pool[slot]) g_pool[slot] = make(&err);
    if (g_pool[slot]) return g_pool[slot];
    if (!g_pool[0]) g_pool[0] = make(&err);                     // (a later slot that cannot be had shares the first one)
    if (!g_pool[0]) g_ctx_err = err;
    return g_pool[0];
}

// $ZB200_DEVICES = "all" or a list "0,1,3": large one-shot jobs behind the zlib API (compress2, deflate() flushes,
// crc32_z / adler32_z of >= 32 MiB) are sharded over those GPUs (zb200_multi_*, zb_multi.cu).  Unset: one GPU.
zb200_multi *api_multi() {
    static zb200_multi *m = [] () -> zb200_multi * {
        const char *e = getenv("ZB200_DEVICES");
        if (!e || !*e) return nullptr;
        std::vector<int> dev;
        if (strcmp(e, "all") != 0) {
            for (const char *p = e; *p;) {
                char *end = nullptr;
                const long v = strtol(p, &end, 10);
                if (end == p) break;
                dev.push_back((int)v);
                p = *end == ',' ? end + 1 : end;
            }
            if (dev.empty()) return nullptr;
        }
        zb200_multi *mm = nullptr;
        if (zb200_multi_create(dev.empty() ? nullptr : dev.data(), (int)dev.size(), &mm) != ZB200_OK) return nullptr;
        if (zb200_multi_count(mm) < 2) { zb200_multi_destroy(mm); return nullptr; }
        return mm;
    }();
    return m;
}
constexpr size_t kMultiMin = (size_t)32 << 20;

}  // namespace
namespace zb { zb200_ctx *zlib_api_ctx() { return api_ctx(); } }
namespace {

size_t api_chunk() {
    static size_t c = [] {
        const char *e = getenv("ZB200_CHUNK");
        long v = e ? atol(e) : 0;
        return (size_t)(v >= 1024 ? v : 262144);
    }();
    return c;
}

// How one call's input is cut.  Up to $ZB200_SINGLE_RUN_MAX bytes it is NOT cut: one run of blocks, no flush marker
// inside — for levels 4-9 byte for byte what the reference's one-shot compress2() / deflate(Z_FINISH) emits
// (compress.c:22-59).  The default is the engine's largest chunk, 1 GiB, for the lazy levels (4-9, any strategy but
// Z_RLE / Z_HUFFMAN_ONLY): their ordered phases are shared by many CTAs (zb_deflate.cu dfl_parse_multi_kernel, chain
// ranges), a single run costs the device what the same bytes in chunks cost.  The greedy levels (whose streams are not
// the reference's byte for byte in any case) walk a single run in one CTA: 1 MiB (config C1's size).  Longer inputs
// are cut into $ZB200_CHUNK-byte Z_FULL_FLUSH runs that are compressed concurrently (the reference's bytes for that
// chunking).
size_t run_chunk(size_t n, int level, int strategy) {
    static long long knob = [] { const char *e = getenv("ZB200_SINGLE_RUN_MAX"); return e ? atoll(e) : -1ll; }();
    const bool lazy = level >= 4 && strategy != Z_RLE && strategy != Z_HUFFMAN_ONLY;
    long long v = knob >= 0 ? knob : lazy ? (1ll << 30) : (1ll << 20);
    if (v > (1ll << 30)) v = 1ll << 30;
    if (!lazy && knob < 0 && v > (1ll << 20)) v = 1ll << 20;
    return (n && n <= (size_t)v && n > api_chunk()) ? n : api_chunk();
}

// $ZB200_CHUNK_CARRY=1: where a call's input IS cut into chunks (longer than a single run, sharded over GPUs, every call at
// levels 1-3 beyond 1 MiB), every chunk is compressed behind the 32 KiB before it (zb200.h ZB200_CHUNK_CARRY: sync points
// instead of full-flush points; 0.7-1.4 % smaller, 2-3 % more device time).  Off by default: the chunks of the default
// form decode independently, which this library's own inflate exploits.
int api_frame(int frame) {
    static const int carry = [] { const char *e = getenv("ZB200_CHUNK_CARRY"); return e && atoi(e) > 0 ? ZB200_CHUNK_CARRY : 0; }();
    // $ZB200_EXACT_FAST=1: levels 1-3 emit the reference's own bytes (zb200.h ZB200_EXACT_FAST: one thread per chunk, slow)
    static const int exact = [] { const char *e = getenv("ZB200_EXACT_FAST"); return e && atoi(e) > 0 ? ZB200_EXACT_FAST : 0; }();
    return frame | carry | exact;
}

[[noreturn]] void die_no_device(const char *fn) {
    fprintf(stderr, "zlib-b200: %s: no usable CUDA device (%s); this library has no CPU path\n", fn, zb200_last_error());
    abort();
}

const char *const kErrMsg[10] = {"need dictionary", "stream end", "", "file error", "stream error",
                                 "data error", "insufficient memory", "buffer error", "incompatible version", ""};

constexpr uint32_t kDeflateMagic = 0x5a42444cu, kInflateMagic = 0x5a42494eu;

struct DeflateStream {
    uint32_t magic;
    int level, strategy, wrap;           // wrap: 0 raw, 1 zlib, 2 gzip
    std::vector<uint8_t> in;             // input accepted but not yet compressed
    std::vector<uint8_t> pending;        // compressed bytes not yet handed to the caller
    size_t pending_pos;
    bool header_done, finished, trailer_done;
    uint32_t crc, adler;                 // running checksums of all compressed input
    uint64_t total_in_hashed;
    int last_flush;
    std::vector<uint8_t> dict;           // history ahead of the next byte: a preset dictionary not yet consumed, or the
                                         // last <= 32768 bytes before a Z_SYNC_FLUSH / Z_PARTIAL_FLUSH / Z_BLOCK point
    bool keep_history;                   // the flush in progress does not reset the history (deflate.c:1211-1218)
    uint32_t dictid;                     // Adler-32 of the whole dictionary (zlib header DICTID)
    bool have_dictid;
    gz_headerp gzhead;                   // deflateSetHeader: the caller's header fields (read when the header is written)
    int tune[4]; bool tuned;             // deflateTune: good_length, max_lazy, nice_length, max_chain
    uint64_t prime_hold; unsigned prime_bits;   // deflatePrime: bits waiting to go out ahead of the next block (deflate.c:731)
    int window_bits, mem_level;          // deflateInit2_'s windowBits (8..15) and memLevel (1..9)
    int bi_used;                         // deflateUsed: trees.c:187,470, deflate.c:1758
    alloc_func zalloc; free_func zfree; voidpf opaque;   // who allocated this state
};

// deflate.c:393-406 / inflate.c:187-199: the stream state comes from the caller's zalloc (NULL: malloc), and the pointers
// the library settled on are written back into the z_stream.  Working buffers are device / pinned memory (and host
// vectors) that the caller's allocator cannot provide; they stay with the library.
voidpf default_alloc(voidpf, uInt items, uInt size) { return malloc((size_t)items * size); }
void default_free(voidpf, voidpf p) { free(p); }
template <class T> T *state_new(z_streamp strm) {
    if (strm->zalloc == (alloc_func)0) { strm->zalloc = default_alloc; strm->opaque = (voidpf)0; }
    if (strm->zfree == (free_func)0) strm->zfree = default_free;
    void *mem = strm->zalloc(strm->opaque, 1, (uInt)sizeof(T));
    if (!mem) return nullptr;
    T *s = new (mem) T();
    s->zalloc = strm->zalloc; s->zfree = strm->zfree; s->opaque = strm->opaque;
    return s;
}
template <class T> void state_delete(T *s) {
    free_func f = s->zfree; voidpf o = s->opaque;
    s->~T();
    f(o, s);
}
template <class T> T *state_clone(z_streamp dest, const T &src) {
    void *mem = src.zalloc(src.opaque, 1, (uInt)sizeof(T));
    (void)dest;
    return mem ? new (mem) T(src) : nullptr;
}

// resize() without the zero fill: output buffers are sized for the worst case and then written by DMA
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U, class... A> void construct(U *p, A &&...a) {
        if (sizeof...(A) == 0) ::new ((void *)p) U; else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};

struct InflateStream {
    uint32_t magic;
    int wrap;                            // ZB200_WRAP_*
    int kind;                            // resolved wrapper kind once the header is parsed
    std::vector<uint8_t> in;             // all compressed input seen so far
    uint8_t *d_in, *d_out;               // device mirrors (grow-only)
    size_t d_in_cap, d_out_cap, d_in_have;
    std::vector<uint8_t, NoInitAlloc<uint8_t>> out;   // decoded bytes not yet delivered
    size_t out_pos;                      // next byte of `out` to deliver
    uint64_t out_base;                   // stream offset of out[0]
    uint64_t decoded;                    // valid decoded bytes on the device
    uint64_t resume_bit, resume_out;
    bool started, done;
    int error;                           // sticky ZB200_INF_* data error
    uint64_t in_used;
    uint32_t check;
    gz_headerp gzhead;                   // inflateGetHeader: where the gzip header fields go
    bool verify;                         // inflateValidate: compare the trailer's check value
    // A stream continued after a run-parallel step (parallel_step): the input before in_erased and the output before
    // out_before have left the buffers; from there on the engine sees a RAW stream behind a 32 KiB history, and the
    // wrapper's trailer is checked here (check_before = check value of the output before the base, in stream_kind).
    bool rebased, seq_ready, trailer_pending;
    int stream_kind;                     // 0 raw, 1 zlib, 2 gzip
    uint64_t in_erased, out_before;
    uint32_t check_before;
    std::vector<uint8_t> hist;           // the last <= 32768 bytes of output produced so far
    size_t dict_len;                     // preset dictionary: d_out[0 .. dict_len), the output follows it
    uint32_t dictid;                     // DICTID of the zlib header that asked for one
    bool retry;                          // inflateSetDictionary was called: decode again with no new input
    int hdr_kind;                        // wrapper the first two input bytes select (0: raw / not seen yet), for inflateSync
    // inflate(Z_BLOCK): deflate-block boundaries decoded but not yet reported — the caller's stream position in bits, the
    // output position, BFINAL of the block that ends there
    struct Bound { uint64_t bit_abs, out_abs; bool last; };
    std::deque<Bound> bounds;
    bool want_bounds;
    int64_t last_bound_bit;
    unsigned sync_have;                  // inflateSync: pattern bytes matched so far (inflate.c:1352 syncsearch), 0..4
    bool syncing;                        // ... a search is under way (state->mode == SYNC)
    uint32_t prime_hold; unsigned prime_bits;   // inflatePrime: bits ahead of the first input byte (inflate.c:223)
    bool prime_byte;                     // the primed bits have been put ahead of the input ...
    size_t prime_pref;                   // ... as this many synthetic bytes (not part of total_in)
    alloc_func zalloc; free_func zfree; voidpf opaque;   // who allocated this state
};

int dev_grow(uint8_t **p, size_t *cap, size_t need, size_t keep, cudaStream_t s) {
    if (need <= *cap) return ZB200_OK;
    size_t want = need + (need >> 1) + 65536;
    uint8_t *n = nullptr;
    if (cudaMalloc((void **)&n, want) != cudaSuccess) { cudaGetLastError(); return ZB200_ERR_NOMEM; }
    if (*p && keep) ZB_CUDA(cudaMemcpyAsync(n, *p, keep, cudaMemcpyDeviceToDevice, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    if (*p) cudaFree(*p);
    *p = n; *cap = want;
    return ZB200_OK;
}

int map_engine_error(int r) { return r == ZB200_ERR_NOMEM ? Z_MEM_ERROR : r == ZB200_ERR_OUTPUT ? Z_BUF_ERROR : Z_STREAM_ERROR; }

// Compress everything buffered in `st.in` (as `finish` ? the end of the stream : a
// run of Z_FULL_FLUSH-terminated chunks) and append the bytes to st.pending.
int compress_buffered(DeflateStream &st, bool finish) {
    zb200_ctx *ctx = api_ctx();
    if (!ctx) return Z_STREAM_ERROR;
    const size_t n = st.in.size();
    if (n == 0 && !finish) return Z_OK;
    const size_t chunk = run_chunk(n, st.level, st.strategy);
    struct TuneScope {                                         // the engine reads the override on this thread
        explicit TuneScope(const int *t) { deflate_tune_set(t); }
        ~TuneScope() { deflate_tune_set(nullptr); }
    } tune_scope(st.tuned ? st.tune : nullptr);
    // deflatePrime (deflate.c:731-757): whole bytes of primed bits go out as they are, a rest of k bits occupies the low
    // bits of the next byte and the engine starts its first block at bit k of it
    while (st.prime_bits >= 8) { st.pending.push_back((uint8_t)st.prime_hold); st.prime_hold >>= 8; st.prime_bits -= 8; }
    const unsigned first_bit = st.prime_bits;
    const uint8_t first_val = (uint8_t)(st.prime_hold & ((1u << first_bit) - 1u));
    st.prime_bits = 0; st.prime_hold = 0;
    size_t cap = zb200_deflate_bound(n, chunk, ZB200_FRAME_RAW) + 16;
    if (!st.dict.empty()) cap += zb200_deflate_bound(st.dict.size() + chunk, st.dict.size() + chunk, ZB200_FRAME_RAW);
    const size_t at = st.pending.size();
    st.pending.resize(at + cap);
    uint32_t adler = 1, crc = 0, used = 8;
    int r;
    zb200_deflate_opts o;
    o.level = st.level; o.strategy = st.strategy; o.window_bits = st.window_bits; o.mem_level = st.mem_level;
    o.dict_len = 0; o.first_bit = first_bit;
    const bool plain = st.window_bits == 15 && st.mem_level == 8 && !first_bit;
    if (!st.dict.empty() && n) {
        // deflate.c:550-632: the dictionary is window content ahead of the first byte: the first chunk goes to the engine
        // behind it (history only), the rest follows as usual
        const size_t dl = st.dict.size();
        std::vector<uint8_t> joined(dl + n);
        memcpy(joined.data(), st.dict.data(), dl);
        memcpy(joined.data() + dl, st.in.data(), n);
        o.dict_len = (uint32_t)dl;
        r = zb200_deflate_host_opts(ctx, joined.data(), dl + n, chunk, &o, api_frame(ZB200_FRAME_RAW), finish ? 1 : 0, st.pending.data() + at, &cap, &adler, &crc, &used);
    } else if (n >= kMultiMin && !st.tuned && plain && api_multi()) {
        r = zb200_multi_deflate_host(api_multi(), st.in.data(), n, chunk, st.level, st.strategy, api_frame(ZB200_FRAME_RAW), finish ? 1 : 0,
                                     st.pending.data() + at, &cap, &adler, &crc);
    } else {
        r = zb200_deflate_host_opts(ctx, n ? st.in.data() : (const uint8_t *)"", n, chunk, &o, api_frame(ZB200_FRAME_RAW), finish ? 1 : 0,
                                    st.pending.data() + at, &cap, &adler, &crc, &used);
    }
    if (r == ZB200_OK && first_bit && cap) st.pending[at] |= first_val;
    if (r == ZB200_OK && finish) st.bi_used = (int)used;           // deflateUsed (deflate.c:723)
    if (r != ZB200_OK) { st.pending.resize(at); return map_engine_error(r); }
    st.pending.resize(at + cap);
    st.crc = zb200_crc32_combine(st.crc, crc, n);
    st.adler = zb200_adler32_combine(st.adler, adler, (int64_t)n);
    st.total_in_hashed += n;
    if (st.keep_history) {                                     // what a sync flush leaves in the window: the last 32 KiB
        std::vector<uint8_t> h;
        if (n >= 32768) h.assign(st.in.end() - 32768, st.in.end());
        else {
            const size_t from_old = st.dict.size() + n > 32768 ? 32768 - n : st.dict.size();
            h.assign(st.dict.end() - (long)from_old, st.dict.end());
            h.insert(h.end(), st.in.begin(), st.in.end());
        }
        st.dict.swap(h);
    } else if (n) st.dict.clear();
    st.in.clear();
    return Z_OK;
}

void put_header(DeflateStream &st) {
    if (st.wrap == 1) {                                        // deflate.c:1004-1037
        const unsigned lf = (st.strategy >= Z_HUFFMAN_ONLY || st.level < 2) ? 0 : st.level < 6 ? 1 : st.level == 6 ? 2 : 3;
        unsigned hdr = ((8u + ((unsigned)(st.window_bits - 8) << 4)) << 8) | (lf << 6);   // deflate.c:1006: Z_DEFLATED + ((w_bits - 8) << 4)
        if (st.have_dictid) hdr |= 0x20;                       // PRESET_DICT, deflate.c:1026
        hdr += 31 - hdr % 31;
        st.pending.push_back((uint8_t)(hdr >> 8)); st.pending.push_back((uint8_t)hdr);
        if (st.have_dictid)                                    // deflate.c:1031-1034: DICTID, most significant byte first
            for (int i = 3; i >= 0; --i) st.pending.push_back((uint8_t)(st.dictid >> (8 * i)));
    } else if (st.wrap == 2 && st.gzhead == Z_NULL) {          // deflate.c:1042-1054
        const uint8_t g[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0,
                               (uint8_t)(st.level == 9 ? 2 : (st.strategy >= Z_HUFFMAN_ONLY || st.level < 2) ? 4 : 0), 3};
        st.pending.insert(st.pending.end(), g, g + 10);
    } else if (st.wrap == 2) {                                 // deflate.c:1056-1170: the caller's fields
        const gz_header &h = *st.gzhead;
        const size_t at = st.pending.size();
        const uint8_t g[10] = {0x1f, 0x8b, 8,
                               (uint8_t)((h.text ? 1 : 0) + (h.hcrc ? 2 : 0) + (h.extra == Z_NULL ? 0 : 4) + (h.name == Z_NULL ? 0 : 8) + (h.comment == Z_NULL ? 0 : 16)),
                               (uint8_t)h.time, (uint8_t)(h.time >> 8), (uint8_t)(h.time >> 16), (uint8_t)(h.time >> 24),
                               (uint8_t)(st.level == 9 ? 2 : (st.strategy >= Z_HUFFMAN_ONLY || st.level < 2) ? 4 : 0), (uint8_t)(h.os & 0xff)};
        st.pending.insert(st.pending.end(), g, g + 10);
        if (h.extra != Z_NULL) {
            st.pending.push_back((uint8_t)h.extra_len); st.pending.push_back((uint8_t)(h.extra_len >> 8));
            st.pending.insert(st.pending.end(), h.extra, h.extra + (h.extra_len & 0xffff));
        }
        if (h.name != Z_NULL) st.pending.insert(st.pending.end(), h.name, h.name + strlen((const char *)h.name) + 1);
        if (h.comment != Z_NULL) st.pending.insert(st.pending.end(), h.comment, h.comment + strlen((const char *)h.comment) + 1);
        if (h.hcrc) {
            const z_crc_t *t = get_crc_table();
            uint32_t c = 0xffffffffu;
            for (size_t i = at; i < st.pending.size(); ++i) c = t[(c ^ st.pending[i]) & 0xff] ^ (c >> 8);
            c = ~c;
            st.pending.push_back((uint8_t)c); st.pending.push_back((uint8_t)(c >> 8));
        }
    }
    st.header_done = true;
}

void put_trailer(DeflateStream &st) {
    if (st.wrap == 1) {                                        // deflate.c:1254-1255
        for (int i = 3; i >= 0; --i) st.pending.push_back((uint8_t)(st.adler >> (8 * i)));
    } else if (st.wrap == 2) {                                 // deflate.c:1241-1250
        for (int i = 0; i < 4; ++i) st.pending.push_back((uint8_t)(st.crc >> (8 * i)));
        for (int i = 0; i < 4; ++i) st.pending.push_back((uint8_t)(st.total_in_hashed >> (8 * i)));
    }
    st.trailer_done = true;
}

DeflateStream *dstate(z_streamp strm) {
    if (!strm || !strm->state) return nullptr;
    DeflateStream *s = reinterpret_cast<DeflateStream *>(strm->state);
    return s->magic == kDeflateMagic ? s : nullptr;
}
InflateStream *istate(z_streamp strm) {
    if (!strm || !strm->state) return nullptr;
    InflateStream *s = reinterpret_cast<InflateStream *>(strm->state);
    return s->magic == kInflateMagic ? s : nullptr;
}

void inflate_reset_state(InflateStream &s) {
    s.kind = 0; s.in.clear(); s.d_in_have = 0; s.out.clear(); s.out_pos = 0; s.out_base = 0; s.decoded = 0;
    s.resume_bit = s.resume_out = 0; s.started = s.done = false; s.error = 0; s.in_used = 0; s.check = 0;
    s.dict_len = 0; s.dictid = 0; s.retry = false; s.gzhead = Z_NULL; s.verify = true;
    s.rebased = s.seq_ready = s.trailer_pending = false; s.stream_kind = 0; s.in_erased = s.out_before = 0; s.check_before = 0; s.hist.clear();
    s.bounds.clear(); s.want_bounds = false; s.last_bound_bit = -1; s.hdr_kind = 0; s.sync_have = 0; s.syncing = false; s.prime_hold = 0; s.prime_bits = 0; s.prime_byte = false; s.prime_pref = 0;
}

void hist_push(InflateStream &s, const uint8_t *p, size_t n) {   // keep the last 32 KiB of output
    if (n >= 32768) { s.hist.assign(p + (n - 32768), p + n); return; }
    if (s.hist.size() + n > 32768) s.hist.erase(s.hist.begin(), s.hist.begin() + (long)(s.hist.size() + n - 32768));
    s.hist.insert(s.hist.end(), p, p + n);
}

// The deflate data of a rebased stream ended at s.in[end] with `tail_len` bytes of output since the base whose check
// value (in stream_kind) is tail_check: verify the wrapper's trailer (inflate.c:1183-1219).  Returns the status;
// ZB200_INF_TRUNCATED while the trailer has not arrived in full.
int finish_rebased(InflateStream &s, uint32_t tail_check, uint64_t tail_len, size_t end) {
    const uint32_t total = s.stream_kind == 1 ? zb200_adler32_combine(s.check_before, tail_check, (int64_t)tail_len)
                                              : zb200_crc32_combine(s.check_before, tail_check, tail_len);
    const uint64_t out_total = s.out_before + tail_len;
    const size_t need = s.stream_kind == 2 ? 8 : s.stream_kind == 1 ? 4 : 0;
    if (s.in.size() < end + need) return ZB200_INF_TRUNCATED;
    const uint8_t *t = s.in.data() + end;
    int st = ZB200_INF_OK;
    if (s.stream_kind == 2) {
        const uint32_t c = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        const uint32_t z = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        if (c != total) st = ZB200_INF_DATA_CHECK;
        else if (z != (uint32_t)out_total) st = ZB200_INF_LENGTH_CHECK;
    } else if (s.stream_kind == 1) {
        const uint32_t c = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | (uint32_t)t[3];
        if (c != total) st = ZB200_INF_DATA_CHECK;
    }
    if (!s.verify) st = ZB200_INF_OK;                           // inflateValidate(strm, 0): the trailer is read, not compared
    if (st == ZB200_INF_OK) { s.done = true; s.in_used = s.in_erased + end + need; s.check = total; }
    return st;
}

// inflate.c:671-808 as far as inflateGetHeader needs it: the gzip header fields out of the input seen so far.
// Returns false while the header is incomplete.
bool fill_gz_header(const std::vector<uint8_t> &in, gz_header &h) {
    const size_t n = in.size();
    if (n < 10) return false;
    const unsigned flg = in[3];
    size_t p = 10;
    const uint8_t *ex = nullptr, *nm = nullptr, *cm = nullptr;
    size_t exl = 0, nml = 0, cml = 0;
    if (flg & 4) {
        if (p + 2 > n) return false;
        exl = in[p] | (in[p + 1] << 8); p += 2;
        if (p + exl > n) return false;
        ex = in.data() + p; p += exl;
    }
    if (flg & 8) { nm = in.data() + p; while (p < n && in[p]) ++p; if (p >= n) return false; ++p; nml = (size_t)(in.data() + p - nm); }
    if (flg & 16) { cm = in.data() + p; while (p < n && in[p]) ++p; if (p >= n) return false; ++p; cml = (size_t)(in.data() + p - cm); }
    if ((flg & 2) && p + 2 > n) return false;
    h.text = (int)(flg & 1);
    h.time = (uLong)in[4] | ((uLong)in[5] << 8) | ((uLong)in[6] << 16) | ((uLong)in[7] << 24);
    h.xflags = in[8]; h.os = in[9];
    if (ex) { h.extra_len = (uInt)exl; if (h.extra != Z_NULL) memcpy(h.extra, ex, exl < h.extra_max ? exl : h.extra_max); }
    else h.extra = Z_NULL;
    if (nm) { if (h.name != Z_NULL) memcpy(h.name, nm, nml < h.name_max ? nml : h.name_max); }
    else h.name = Z_NULL;
    if (cm) { if (h.comment != Z_NULL) memcpy(h.comment, cm, cml < h.comm_max ? cml : h.comm_max); }
    else h.comment = Z_NULL;
    h.hcrc = (int)((flg >> 1) & 1);
    h.done = 1;
    return true;
}

// Bounded memory for a stream without flush points (the reference needs a 32 KiB window, inflate.c:368-412): the
// one-member path RE-BASES at a deflate-block boundary once enough has piled up behind it — the input before the
// boundary and the output delivered so far leave the buffers, the check value of that output is folded into
// check_before, and the engine continues at the boundary as a raw stream behind the last 32 KiB of output (the
// dictionary mechanism), exactly as parallel_step does at flush points.  rbit: bit offset of the boundary in s.in;
// rout: output bytes since the current base that lie before it (all delivered, the tail of them in s.hist).
constexpr uint64_t kRebaseIn = (uint64_t)1 << 20, kRebaseOut = (uint64_t)4 << 20;    // re-base after 1 MiB in / 4 MiB out
constexpr size_t kDevOutCap = (size_t)1 << 30;                                        // ... and rather than growing the device output past 1 GiB
constexpr size_t kTakeMax = (size_t)256 << 20, kOwedMax = (size_t)256 << 20;           // input taken per inflate() call / output owed before more is taken
int rebase_here(InflateStream &s, zb200_ctx *ctx, cudaStream_t st, uint64_t rbit, uint64_t rout) {
    const int kind = s.rebased ? s.stream_kind : s.kind;
    uint32_t part = kind == 1 ? 1u : 0u;
    if (rout) {
        uint8_t *base = (uint8_t *)ctx->d_scratch;
        uint32_t *d_sum = (uint32_t *)(base + 512);
        if (checksum_launch(ctx, s.d_out + s.dict_len, nullptr, nullptr, rout, 1, kind == 1 ? ZB200_ADLER32 : ZB200_CRC32, 0, 1, d_sum, d_sum + 1,
                            (CkAccum *)(base + 768), st) != ZB200_OK) return -1;
        if (cudaMemcpyAsync(ctx->h_small, d_sum, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        part = ((const uint32_t *)ctx->h_small)[kind == 1 ? 1 : 0];
    }
    if (!s.rebased) { s.stream_kind = kind; s.check_before = part; s.out_before = rout; s.rebased = true; }
    else {
        s.check_before = kind == 1 ? zb200_adler32_combine(s.check_before, part, (int64_t)rout) : zb200_crc32_combine(s.check_before, part, rout);
        s.out_before += rout;
    }
    const size_t drop = (size_t)(rbit >> 3);
    s.in_erased += drop;
    s.in.erase(s.in.begin(), s.in.begin() + (long)drop);
    s.d_in_have = 0;
    s.resume_bit = rbit & 7u; s.resume_out = 0; s.decoded = 0; s.seq_ready = false;
    return 0;
}

// One decode attempt over everything received so far, resuming at the last
// block boundary.  Updates the stream state; returns a ZB200_INF_* status.
int inflate_attempt(InflateStream &s) {
    zb200_ctx *ctx = api_ctx();
    if (!ctx) return -1;
    static const int dbg = [] { const char *e = getenv("ZB200_API_DEBUG"); return e ? atoi(e) : 0; }();
    if (dbg) fprintf(stderr, "[inflate_attempt] in %zu decoded %llu resume_bit %llu rebased %d seq_ready %d\n", s.in.size(),
                     (unsigned long long)s.decoded, (unsigned long long)s.resume_bit, (int)s.rebased, (int)s.seq_ready);
    CtxUse use(ctx, ctx->stream);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    cudaStream_t st = ctx->stream;
    constexpr uint32_t kBlogCap = 2048;                          // block boundaries logged per attempt (a fuller log ends the attempt there)
    const size_t blog_at = (1024 + inflate_work_bytes(1) + 255) & ~(size_t)255;
    if (ensure_scratch(ctx, blog_at + 16 + 16 * (size_t)kBlogCap) != ZB200_OK) return -1;
    for (;;) {
    const size_t n = s.in.size();
    if (dev_grow(&s.d_in, &s.d_in_cap, n + 16, s.d_in_have, st)) return -1;
    if (n > s.d_in_have) {
        if (cudaMemcpyAsync(s.d_in + s.d_in_have, s.in.data() + s.d_in_have, n - s.d_in_have, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        s.d_in_have = n;
    }
    if (s.rebased && !s.seq_ready) {                            // the one-member path takes over behind the history
        s.dict_len = s.hist.size();
        if (dev_grow(&s.d_out, &s.d_out_cap, n * 4 + (1u << 20) + s.dict_len, 0, st)) return -1;
        if (s.dict_len && cudaMemcpyAsync(s.d_out, s.hist.data(), s.dict_len, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        s.seq_ready = true;
    }
    if (s.d_out_cap == 0 && dev_grow(&s.d_out, &s.d_out_cap, n * 4 + (1u << 20) + s.dict_len, 0, st)) return -1;
    bool again = false;                                          // re-based mid-call: go round once more from the new base
    for (;;) {
        zb200_member m;
        m.in_off = 0; m.in_len = n; m.out_off = s.dict_len; m.out_cap = s.d_out_cap - s.dict_len;
        m.resume_bit = s.resume_bit; m.resume_out = s.resume_out; m.dict_len = s.dict_len;
        uint8_t *base = (uint8_t *)ctx->d_scratch;
        zb200_member *d_m = (zb200_member *)base;
        zb200_member_result *d_r = (zb200_member_result *)(base + 256);
        zb200_member_result *h_r = (zb200_member_result *)ctx->h_small;
        memcpy(ctx->h_small + 40, &m, sizeof m);                 // pinned bounce for the descriptor
        if (cudaMemcpyAsync(d_m, ctx->h_small + 40, sizeof m, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        const int wrap = s.rebased ? ZB200_WRAP_RAW : s.resume_bit ? s.kind : s.wrap;
        uint64_t *d_blog = s.want_bounds ? (uint64_t *)(base + blog_at) : nullptr;
        if (inflate_launch(ctx, s.d_in, s.d_out, d_m, 1, wrap, (s.verify && !s.rebased) ? 1 : 0, d_r, base + 1024, st, d_blog, kBlogCap) != ZB200_OK) return -1;
        if (cudaMemcpyAsync(h_r, d_r, sizeof *h_r, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        const zb200_member_result r = *h_r;
        if (d_blog) {                                            // the block boundaries this attempt passed, in the caller's coordinates
            uint64_t cnt = 0;
            if (cudaMemcpy(&cnt, d_blog, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
            if (cnt > kBlogCap) cnt = kBlogCap;
            std::vector<uint64_t> e(2 * (size_t)cnt);
            if (cnt && cudaMemcpy(e.data(), d_blog + 2, 16 * (size_t)cnt, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
            const uint64_t valid_out = r.status == ZB200_INF_OK ? r.out_len : (r.resume_bit ? r.resume_out : 0);
            for (uint64_t i = 0; i < cnt; ++i) {
                const uint64_t bit = e[2 * i], pos = e[2 * i + 1] & ~(1ull << 63);
                const bool last = (e[2 * i + 1] >> 63) != 0;
                const int64_t bit_abs = (int64_t)bit + 8 * ((int64_t)s.in_erased - (int64_t)s.prime_pref);
                const uint64_t out_rel = pos - s.dict_len;
                if (out_rel > valid_out) break;                  // (behind what this attempt hands out)
                if (bit_abs <= 0 || bit_abs <= s.last_bound_bit) continue;  // (inside primed bits / logged by an earlier attempt; a raw
                                                                            //  stream's very start is no stop: inflate.c:627-630 enters at TYPEDO)
                s.last_bound_bit = bit_abs;
                s.bounds.push_back({(uint64_t)bit_abs, s.out_before + out_rel, last});
            }
        }
        if (r.status == ZB200_INF_OUTPUT_FULL) {                 // grow the device output and go again from the checkpoint ...
            if (!s.rebased) s.kind = (int)r.wrap_kind;
            s.resume_bit = r.resume_bit; s.resume_out = r.resume_out;
            if (r.resume_bit == 0) { s.resume_bit = 0; s.resume_out = 0; }
            if (s.d_out_cap >= kDevOutCap && r.resume_bit && r.resume_out > s.decoded) {
                // ... or, once the buffer is large, hand out what is complete and re-base at the checkpoint instead
                const size_t add = (size_t)(r.resume_out - s.decoded), at = s.out.size();
                s.out.resize(at + add);
                if (cudaMemcpyAsync(s.out.data() + at, s.d_out + s.dict_len + s.decoded, add, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
                if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
                hist_push(s, s.out.data() + at, add);
                if (rebase_here(s, ctx, st, r.resume_bit, r.resume_out)) return -1;
                again = true;
                break;
            }
            if (dev_grow(&s.d_out, &s.d_out_cap, s.d_out_cap * 2 + (1u << 20), (size_t)r.out_len + s.dict_len, st)) return -1;
            continue;
        }
        // fetch the newly valid bytes.  An attempt that ran out of input delivers up to the last block boundary only:
        // the open block is decoded again from its start by the next attempt, and what this one produced of it is
        // not handed out (tools/repro_slices.py: with 7-byte slices ten such bytes differed from the final decode).
        const uint64_t valid = r.status == ZB200_INF_TRUNCATED ? (r.resume_bit ? r.resume_out : 0) : r.out_len;
        if (valid > s.decoded) {
            const size_t add = (size_t)(valid - s.decoded);
            const size_t at = s.out.size();
            s.out.resize(at + add);
            if (cudaMemcpyAsync(s.out.data() + at, s.d_out + s.dict_len + s.decoded, add, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
            if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
            hist_push(s, s.out.data() + at, add);
            s.decoded = valid;
        }
        // bounded memory: enough has piled up behind the last block boundary -> continue from there as a re-based stream
        if (r.status == ZB200_INF_TRUNCATED && r.resume_bit && ((r.resume_bit >> 3) >= kRebaseIn || r.resume_out >= kRebaseOut) &&
            (s.rebased || r.wrap_kind == 0 || r.resume_bit > 16)) {
            if (!s.rebased) s.kind = (int)r.wrap_kind;
            if (rebase_here(s, ctx, st, r.resume_bit, r.resume_out)) return -1;
            return ZB200_INF_TRUNCATED;
        }
        if (s.rebased) {
            // (a data error keeps the last block boundary too: inflateSync starts its search there)
            if (r.status != ZB200_INF_OK && r.status != ZB200_INF_OUTPUT_FULL && r.resume_bit) { s.resume_bit = r.resume_bit; s.resume_out = r.resume_out; }
            if (r.status != ZB200_INF_OK) return r.status;
            uint32_t tail = r.check;                             // CRC-32 of the bytes since the base (raw members get a CRC)
            if (s.stream_kind == 1) {                            // a zlib wrapper wants their Adler-32
                uint32_t *d_sum = (uint32_t *)(base + 512);
                if (checksum_launch(ctx, s.d_out + s.dict_len, nullptr, nullptr, r.out_len, 1, ZB200_ADLER32, 0, 1, d_sum, d_sum + 1,
                                    (CkAccum *)(base + 768), st) != ZB200_OK) return -1;
                if (cudaMemcpyAsync(ctx->h_small, d_sum, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
                if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
                tail = ((const uint32_t *)ctx->h_small)[1];
            }
            return finish_rebased(s, tail, r.out_len, (size_t)r.in_used);
        }
        s.kind = (int)r.wrap_kind;
        if (r.status == ZB200_INF_NEED_DICT) s.dictid = r.check;
        if (r.status == ZB200_INF_TRUNCATED) {
            if (r.resume_bit) { s.resume_bit = r.resume_bit; s.resume_out = r.resume_out; }
        } else if (r.status != ZB200_INF_OK) {                    // a data error: the last block boundary is where inflateSync starts its search
            if (r.resume_bit) { s.resume_bit = r.resume_bit; s.resume_out = r.resume_out; }
        }
        if (r.status == ZB200_INF_OK) {
            s.done = true; s.in_used = s.in_erased + r.in_used; s.check = r.check;
        }
        return r.status;
    }
    if (!again) return -1;                                       // (not reached: the inner loop leaves by return or with `again`)
    }
}

// A call that brings (or has piled up) at least 256 KiB of undecoded input at a point where a run may start:
// the runs between flush points are decoded in one batch (inflate_stream_parallel).  Whatever prefix of runs is
// verified is delivered, the stream is re-based behind it — the input and output before the base leave the
// buffers — and the rest waits for the next call; the last run of a complete stream ends it.  Returns 1 when the
// step made this call's progress (status in *st_out), 0 when the one-member path should take the call.
// `direct` (optional): the caller's own output buffer, `direct_cap` bytes of room.  When nothing is queued in s.out and the
// step's output fits, it is decoded straight into it (*direct_len bytes) instead of through the queue — a queue of the
// stream's size costs its page faults and one more pass over the bytes (128 MiB in one call: 56 -> 31 ms).
int parallel_step(InflateStream &s, int *st_out, uint8_t *direct = nullptr, size_t direct_cap = 0, size_t *direct_len = nullptr) {
    // (a re-based stream waits at bit resume_bit < 8 of its first buffered byte — a block boundary — with nothing decoded yet)
    const bool at_base = s.rebased ? s.resume_bit < 8 : s.resume_bit == 0;
    if (!s.verify || s.decoded || !at_base || s.seq_ready || (s.dict_len && !s.rebased) || s.in.size() < 262144) return 0;
    zb200_ctx *ctx = api_ctx();
    if (!ctx) return 0;
    const size_t at = s.out.size();
    size_t out_len = 0, in_used = 0;
    int status = 0, applicable = 0;
    uint32_t check = 0, end_bit = 0;
    const StreamContinuation cont = {s.hist.data(), s.hist.size(), s.rebased ? (uint32_t)s.resume_bit : 0u, s.stream_kind};
    // the step's output goes into the caller's buffer when nothing is queued and it fits, else into the queue, grown to the
    // size the decoders then know (StreamOutAlt: no second decode either way)
    bool to_caller = direct && direct_len && at == 0 && direct_cap >= 65536;
    struct Grow { InflateStream *s; size_t at; } gr = {&s, at};
    StreamOutAlt alt = {[](void *self, size_t need) -> uint8_t * {
                            Grow *g = (Grow *)self;
                            try { g->s->out.resize(g->at + need); } catch (...) { return nullptr; }
                            return g->s->out.data() + g->at;
                        }, &gr, false};
    {
        CtxUse use(ctx, ctx->stream);
        if (cudaSetDevice(ctx->device) != cudaSuccess) return 0;
        if (!to_caller) s.out.resize(at + 65536);
        const int r = inflate_stream_parallel(ctx, s.in.data(), s.in.size(), s.rebased ? ZB200_WRAP_RAW : s.wrap,
                                              to_caller ? direct : s.out.data() + at, to_caller ? direct_cap : 65536,
                                              &out_len, &status, &in_used, &check, &applicable, s.rebased ? &cont : nullptr, 2, &end_bit, &alt);
        if (r != ZB200_OK) applicable = 0;
    }
    static const int dbg = [] { const char *e = getenv("ZB200_API_DEBUG"); return e ? atoi(e) : 0; }();
    if (dbg) fprintf(stderr, "[parallel_step] in %zu at %zu direct_cap %zu -> applicable %d status %d out_len %zu in_used %zu end_bit %u alt %d\n",
                     s.in.size(), at, direct_cap, applicable, status, out_len, in_used, end_bit, (int)alt.used);
    if (!applicable || status == ZB200_INF_OUTPUT_FULL || (status != ZB200_INF_OK && out_len == 0)) { s.out.resize(at); return 0; }
    if (to_caller && !alt.used) { *direct_len = out_len; hist_push(s, direct, out_len); }
    else { s.out.resize(at + out_len); hist_push(s, s.out.data() + at, out_len); }
    if (!s.rebased) {
        const int kind = s.wrap == ZB200_WRAP_RAW ? 0 : ((s.wrap & ZB200_WRAP_GZIP) && s.in[0] == 0x1f && s.in[1] == 0x8b) ? 2 : 1;
        if (status == ZB200_INF_OK) {                           // the whole stream, trailer verified
            s.kind = kind; s.done = true; s.in_used = in_used; s.check = check;
            *st_out = ZB200_INF_OK;
            return 1;
        }
        if (status == ZB200_INF_DATA_CHECK || status == ZB200_INF_LENGTH_CHECK) { s.kind = kind; *st_out = status; return 1; }   // whole stream, bad trailer
        s.stream_kind = kind; s.kind = kind; s.rebased = true;
        s.check_before = check; s.out_before = out_len;
        if (applicable == 2) s.trailer_pending = true;          // all of the deflate data is out; the trailer has not arrived in full
    } else if (status == ZB200_INF_OK) {                        // the deflate data ends in this step: the trailer is checked here
        const int fs = finish_rebased(s, check, out_len, in_used);
        if (fs == ZB200_INF_TRUNCATED) {                        // ... once it has arrived
            s.check_before = s.stream_kind == 1 ? zb200_adler32_combine(s.check_before, check, (int64_t)out_len) : zb200_crc32_combine(s.check_before, check, out_len);
            s.out_before += out_len; s.trailer_pending = true;
            s.in_erased += in_used; s.in.erase(s.in.begin(), s.in.begin() + (long)in_used); s.d_in_have = 0;
        }
        *st_out = fs;
        return 1;
    } else {
        s.check_before = s.stream_kind == 1 ? zb200_adler32_combine(s.check_before, check, (int64_t)out_len) : zb200_crc32_combine(s.check_before, check, out_len);
        s.out_before += out_len;
    }
    s.in_erased += in_used;
    s.in.erase(s.in.begin(), s.in.begin() + (long)in_used);
    s.d_in_have = 0;
    s.resume_bit = end_bit; s.resume_out = 0;                   // (a prefix of chunks ends at a block boundary inside a byte)
    *st_out = ZB200_INF_TRUNCATED;                              // more to come; an error met behind the prefix shows up when its run is reached
    return 1;
}

}  // namespace

extern "C" {

const char *zlibVersion(void) { return ZLIB_VERSION; }
const char *zError(int err) { return (err >= -6 && err <= 2) ? kErrMsg[2 - err] : ""; }   // zutil.c:131

// ---------------------------------------------------------------------------
int deflateInit2_(z_streamp strm, int level, int method, int windowBits, int memLevel, int strategy,
                  const char *version, int stream_size) {
    if (version == Z_NULL || version[0] != ZLIB_VERSION[0] || stream_size != (int)sizeof(z_stream))
        return Z_VERSION_ERROR;                                 // deflate.c:386-389
    if (strm == Z_NULL) return Z_STREAM_ERROR;
    strm->msg = Z_NULL;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    int wrap = 1;
    if (windowBits < 0) { wrap = 0; if (windowBits < -15) return Z_STREAM_ERROR; windowBits = -windowBits; }
    else if (windowBits > 15) { wrap = 2; windowBits -= 16; }
    if (memLevel < 1 || memLevel > 9 || method != Z_DEFLATED || windowBits < 8 || windowBits > 15 || level < 0 ||
        level > 9 || strategy < 0 || strategy > Z_FIXED || (windowBits == 8 && wrap != 1))
        return Z_STREAM_ERROR;                                  // deflate.c:426-430
    if (!api_ctx()) { strm->msg = "zlib-b200: no usable CUDA device (no CPU path)"; return Z_STREAM_ERROR; }
    if (windowBits == 8) windowBits = 9;                        // deflate.c:431: until 256-byte window bug fixed
    DeflateStream *s = state_new<DeflateStream>(strm);
    if (!s) return Z_MEM_ERROR;
    s->magic = kDeflateMagic; s->level = level; s->strategy = strategy; s->wrap = wrap;
    s->window_bits = windowBits; s->mem_level = memLevel; s->gzhead = Z_NULL;
    strm->state = reinterpret_cast<struct internal_state *>(s);
    return deflateReset(strm);
}

int deflateInit_(z_streamp strm, int level, const char *version, int stream_size) {
    return deflateInit2_(strm, level, Z_DEFLATED, 15, 8, Z_DEFAULT_STRATEGY, version, stream_size);
}

int deflateReset(z_streamp strm) {
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    s->in.clear(); s->pending.clear(); s->pending_pos = 0;
    s->header_done = s->finished = s->trailer_done = false;
    s->crc = 0; s->adler = 1; s->total_in_hashed = 0; s->last_flush = -2;
    s->dict.clear(); s->dictid = 0; s->have_dictid = false; s->keep_history = false;   // (gzhead stays: deflate.c:644-673 does not touch it)
    s->bi_used = 0; s->prime_hold = 0; s->prime_bits = 0;
    s->tuned = false;                                           // lm_init reloads the level's table values (deflate.c:1307-1326)
    strm->total_in = strm->total_out = 0; strm->msg = Z_NULL; strm->data_type = Z_UNKNOWN;
    strm->adler = s->wrap == 2 ? 0 : 1;                         // deflate.c:656-660
    return Z_OK;
}

int deflateParams(z_streamp strm, int level, int strategy) {
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    if (level < 0 || level > 9 || strategy < 0 || strategy > Z_FIXED) return Z_STREAM_ERROR;
    if ((level != s->level || strategy != s->strategy) && !s->in.empty()) {
        if (!s->header_done) put_header(*s);                    // deflate.c:779-790: flush what was taken with the old setting
        s->keep_history = true;                                 // (a Z_BLOCK flush there: the window stays)
        const int r = compress_buffered(*s, false);
        if (r != Z_OK) return r;
    }
    if (level != s->level) s->tuned = false;                    // deflate.c:792-800: a new level reloads its table values
    s->level = level; s->strategy = strategy;
    return Z_OK;
}

// deflate.c:550-632.  Accepted where the reference accepts it at the start of a stream: before the first
// deflate() call (zlib / raw wrappers; gzip streams take no dictionary).  Mid-stream use on a raw stream
// (deflate.c:566, lookahead == 0 after a full flush) is not offered: Z_STREAM_ERROR.
int deflateSetDictionary(z_streamp strm, const Bytef *dictionary, uInt dictLength) {
    DeflateStream *s = dstate(strm);
    if (!s || dictionary == Z_NULL) return Z_STREAM_ERROR;
    if (s->wrap == 2 || s->header_done || !s->in.empty() || s->total_in_hashed || strm->total_in || s->finished) return Z_STREAM_ERROR;
    if (s->wrap == 1) {                                         // deflate.c:570-571
        zb200_ctx *ctx = api_ctx();
        if (!ctx) return Z_STREAM_ERROR;
        uint32_t a = (uint32_t)strm->adler;
        if (dictLength && zb200_checksum_host(ctx, dictionary, dictLength, ZB200_ADLER32, 0, (uint32_t)strm->adler, nullptr, &a) != ZB200_OK)
            return Z_STREAM_ERROR;
        strm->adler = a;
        s->dictid = a; s->have_dictid = true;
    }
    const uInt keep = dictLength > 32768u ? 32768u : dictLength;   // deflate.c:575-583: the tail of a long dictionary
    s->dict.assign(dictionary + (dictLength - keep), dictionary + dictLength);
    return Z_OK;
}

int deflateTune(z_streamp strm, int good_length, int max_lazy, int nice_length, int max_chain) {   // deflate.c:805-816
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    s->tune[0] = good_length; s->tune[1] = max_lazy; s->tune[2] = nice_length; s->tune[3] = max_chain;
    s->tuned = true;
    return Z_OK;
}

int deflateGetDictionary(z_streamp strm, Bytef *dictionary, uInt *dictLength) {   // deflate.c:638: what the next byte would be compressed against
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (dictionary != Z_NULL && !s->dict.empty()) memcpy(dictionary, s->dict.data(), s->dict.size());
    if (dictLength != Z_NULL) *dictLength = (uInt)s->dict.size();
    return Z_OK;
}

int deflateSetHeader(z_streamp strm, gz_headerp head) {      // deflate.c:692-697
    DeflateStream *s = dstate(strm);
    if (!s || s->wrap != 2) return Z_STREAM_ERROR;
    s->gzhead = head;
    return Z_OK;
}

int deflatePending(z_streamp strm, unsigned *pending, int *bits) {   // deflate.c:703-713; output here is always whole bytes
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (pending != Z_NULL) *pending = (unsigned)(s->pending.size() - s->pending_pos) + s->prime_bits / 8;
    if (bits != Z_NULL) *bits = (int)(s->prime_bits & 7u);      // bi_valid: primed bits short of a byte
    return Z_OK;
}

// deflate.c:731-757: up to 16 bits per call ahead of the next deflate block.  They wait here (the reference moves them
// into its bit buffer) until the next block run is compressed; with more than 48 waiting the call is refused like a full
// pending buffer.
int deflatePrime(z_streamp strm, int bits, int value) {
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (bits < 0 || bits > 16 || s->prime_bits + (unsigned)bits > 48) return Z_BUF_ERROR;
    s->prime_hold |= (uint64_t)((unsigned)value & ((1u << bits) - 1u)) << s->prime_bits;
    s->prime_bits += (unsigned)bits;
    return Z_OK;
}

int deflateUsed(z_streamp strm, int *bits) {                 // deflate.c:723-729: bits of the last byte handed out that are in use
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (bits != Z_NULL) *bits = s->bi_used;
    return Z_OK;
}

int deflateResetKeep(z_streamp strm) { return deflateReset(strm); }   // deflate.c:644: no match-finder state to keep here

int deflateCopy(z_streamp dest, z_streamp source) {          // deflate.c:1297-1345
    DeflateStream *s = dstate(source);
    if (!s || dest == Z_NULL) return Z_STREAM_ERROR;
    DeflateStream *d = state_clone(dest, *s);
    if (!d) return Z_MEM_ERROR;
    *dest = *source;
    dest->state = reinterpret_cast<struct internal_state *>(d);
    return Z_OK;
}

uLong zlibCompileFlags(void) {                               // zutil.c:32-113: type sizes; no debug, no asm, gz* present
    auto code = [](size_t n) -> uLong { return n == 2 ? 0 : n == 4 ? 1 : n == 8 ? 2 : 3; };
    return code(sizeof(uInt)) | (code(sizeof(uLong)) << 2) | (code(sizeof(voidpf)) << 4) | (code(sizeof(z_off_t)) << 6);
}

// deflate.c:842-905 / compress.c:72.  The reference's tight bound (n + n/4096 + n/16384 + n/2^25 + 13) counts the 5-byte
// stored-block headers of ITS blocking; this engine's streams are cut into chunks, and a chunk of incompressible bytes
// costs its own stored blocks — one per sym_limit literals (lit_bufsize - 1, deflate.c:455,512) plus the short one that
// ends the chunk — and its 5-byte flush marker.  Non-default windowBits / memLevel get the reference's conservative
// fixed-code bound (deflate.c:852-855) on top of the same per-chunk terms.
static uLong bound_for(uLong n, int window_bits, int mem_level) {
    const uLong chunk = (uLong)api_chunk();
    const uLong sym_limit = (1ul << (mem_level + 6)) - 1;
    const uLong nch = n / chunk + 1;
    const uLong per_chunk = 5 * ((chunk + sym_limit - 1) / sym_limit + 1) + 6;
    uLong body = n + nch * per_chunk;
    if (window_bits != 15 || mem_level != 8) body += (n >> 3) + (n >> 8) + (n >> 9) + 4;
    const uLong ref = n + (n >> 12) + (n >> 14) + (n >> 25) + 13;
    return (body > ref ? body : ref) + 18;
}
uLong deflateBound(z_streamp strm, uLong n) {
    DeflateStream *s = dstate(strm);
    uLong b = bound_for(n, s ? s->window_bits : 15, s ? s->mem_level : 8);
    if (s && s->wrap == 2 && s->gzhead != Z_NULL) {             // deflate.c:869-888: the caller's gzip header fields
        const gz_header &h = *s->gzhead;
        if (h.extra != Z_NULL) b += 2 + h.extra_len;
        if (h.name != Z_NULL) b += strlen((const char *)h.name) + 1;
        if (h.comment != Z_NULL) b += strlen((const char *)h.comment) + 1;
        if (h.hcrc) b += 2;
    }
    return b;
}

int deflate(z_streamp strm, int flush) {
    DeflateStream *s = dstate(strm);
    if (!s || flush > Z_BLOCK || flush < 0) return Z_STREAM_ERROR;
    if (strm->next_out == Z_NULL || (strm->avail_in != 0 && strm->next_in == Z_NULL) || (s->finished && flush != Z_FINISH)) {
        strm->msg = kErrMsg[2 - Z_STREAM_ERROR];
        return Z_STREAM_ERROR;                                  // deflate.c:962-966
    }
    if (strm->avail_out == 0) { strm->msg = kErrMsg[2 - Z_BUF_ERROR]; return Z_BUF_ERROR; }
    const uInt in0 = strm->avail_in, out0 = strm->avail_out;
    if (strm->avail_in) {
        if (s->finished) { strm->msg = kErrMsg[2 - Z_BUF_ERROR]; return Z_BUF_ERROR; }   // deflate.c:1184-1187
        s->in.insert(s->in.end(), strm->next_in, strm->next_in + strm->avail_in);
        strm->next_in += strm->avail_in; strm->total_in += strm->avail_in; strm->avail_in = 0;
    }
    if (!s->finished) {
        const bool want_flush = flush != Z_NO_FLUSH;
        // bound host buffering: hand what has piled up to the GPU — as one run of blocks (levels 4-9) or whole chunks,
        // ending on a full-flush marker — once it reaches $ZB200_STREAM_HOLD_MIB (default 256: a stream up to that size
        // comes out as the reference's one run however it is fed)
        static const size_t hold = [] { const char *e = getenv("ZB200_STREAM_HOLD_MIB"); const long v = e ? atol(e) : 0; return (size_t)(v >= 1 ? v : 256) << 20; }();
        const bool spill = !want_flush && s->in.size() > hold;   // (a stream of exactly `hold` bytes is still one run)
        if (want_flush || spill) {
            if (!s->header_done) put_header(*s);
            // Z_SYNC_FLUSH / Z_PARTIAL_FLUSH / Z_BLOCK keep the window (deflate.c:1211-1218 clears the hash only for
            // Z_FULL_FLUSH): the first chunk after such a point is compressed behind the last 32 KiB before it
            // (a spill is this library's own cut, not the caller's: with $ZB200_CHUNK_CARRY=1 it keeps the window as well — the reference
            //  would not have cut there at all)
            s->keep_history = spill ? (api_frame(0) & ZB200_CHUNK_CARRY) != 0 : (flush == Z_SYNC_FLUSH || flush == Z_PARTIAL_FLUSH || flush == Z_BLOCK);
            int r;
            if (spill) {
                const size_t keep = s->in.size() % api_chunk();
                std::vector<uint8_t> tail(s->in.end() - (long)keep, s->in.end());
                s->in.resize(s->in.size() - keep);
                r = compress_buffered(*s, false);
                s->in = tail;
            } else {
                r = compress_buffered(*s, flush == Z_FINISH);
            }
            if (r != Z_OK) { strm->msg = kErrMsg[2 - r]; return r; }
            if (flush == Z_FINISH) { s->finished = true; put_trailer(*s); }
            if (s->wrap) strm->adler = s->wrap == 2 ? s->crc : s->adler;   // (a raw stream leaves strm->adler to the caller: deflate.c:218-239;
                                                                           //  examples/gzappend.c keeps its own CRC there)
        }
    }
    s->last_flush = flush;
    const size_t avail = s->pending.size() - s->pending_pos;
    const size_t k = avail < strm->avail_out ? avail : strm->avail_out;
    if (k) {
        memcpy(strm->next_out, s->pending.data() + s->pending_pos, k);
        strm->next_out += k; strm->avail_out -= (uInt)k; strm->total_out += k; s->pending_pos += k;
        if (s->pending_pos == s->pending.size()) { s->pending.clear(); s->pending_pos = 0; }
    }
    if (s->finished && s->pending.empty()) return Z_STREAM_END;
    if (in0 == 0 && out0 == strm->avail_out && flush != Z_FINISH) { strm->msg = kErrMsg[2 - Z_BUF_ERROR]; return Z_BUF_ERROR; }
    return Z_OK;
}

int deflateEnd(z_streamp strm) {
    DeflateStream *s = dstate(strm);
    if (!s) return Z_STREAM_ERROR;
    const bool busy = !s->finished && (s->header_done || !s->in.empty());
    s->magic = 0;
    state_delete(s);
    strm->state = Z_NULL;
    return busy ? Z_DATA_ERROR : Z_OK;                          // deflate.c:1284
}

// ---------------------------------------------------------------------------
int inflateInit2_(z_streamp strm, int windowBits, const char *version, int stream_size) {
    if (version == Z_NULL || version[0] != ZLIB_VERSION[0] || stream_size != (int)sizeof(z_stream))
        return Z_VERSION_ERROR;                                 // inflate.c:183-185
    if (strm == Z_NULL) return Z_STREAM_ERROR;
    strm->msg = Z_NULL;
    if (!api_ctx()) { strm->msg = "zlib-b200: no usable CUDA device (no CPU path)"; return Z_STREAM_ERROR; }
    InflateStream *s = state_new<InflateStream>(strm);
    if (!s) return Z_MEM_ERROR;
    s->magic = kInflateMagic; s->d_in = s->d_out = nullptr; s->d_in_cap = s->d_out_cap = 0;
    strm->state = reinterpret_cast<struct internal_state *>(s);
    const int r = inflateReset2(strm, windowBits);
    if (r != Z_OK) { state_delete(s); strm->state = Z_NULL; }
    return r;
}

int inflateInit_(z_streamp strm, const char *version, int stream_size) { return inflateInit2_(strm, 15, version, stream_size); }

int inflateReset2(z_streamp strm, int windowBits) {
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    int wrap;                                                   // inflate.c:152-172
    if (windowBits < 0) { if (windowBits < -15) return Z_STREAM_ERROR; wrap = ZB200_WRAP_RAW; windowBits = -windowBits; }
    else if (windowBits >= 48) return Z_STREAM_ERROR;
    else if (windowBits >= 32) { wrap = ZB200_WRAP_AUTO; windowBits -= 32; }
    else if (windowBits >= 16) { wrap = ZB200_WRAP_GZIP; windowBits -= 16; }
    else wrap = ZB200_WRAP_ZLIB;
    if (windowBits && (windowBits < 8 || windowBits > 15)) return Z_STREAM_ERROR;
    s->wrap = wrap;
    return inflateReset(strm);
}

int inflateReset(z_streamp strm) {
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    inflate_reset_state(*s);
    strm->total_in = strm->total_out = 0; strm->msg = Z_NULL; strm->data_type = 0;
    strm->adler = s->wrap == ZB200_WRAP_RAW ? 0 : 1;
    return Z_OK;
}

int inflate(z_streamp strm, int flush) {
    InflateStream *s = istate(strm);
    if (!s || strm->next_out == Z_NULL || (strm->next_in == Z_NULL && strm->avail_in != 0)) return Z_STREAM_ERROR;   // inflate.c:611-613
    const Bytef *const in_ptr0 = strm->next_in;
    const uInt in0 = strm->avail_in, out0 = strm->avail_out;
    const uLong total_in0 = strm->total_in;
    // Z_BLOCK / Z_TREES (zlib.h:540-560, inflate.c:824-826): stop at every deflate-block boundary.  The engine then logs the
    // boundaries it passes; a call hands out output up to the next one only, and reports as consumed only the input up
    // to it — what lies behind stays in this stream's buffer AND with the caller, who passes it again (the bytes are
    // recognised by their position and not taken twice).
    const bool block_mode = flush == Z_BLOCK || flush == Z_TREES;
    if (block_mode) s->want_bounds = true;
    int status = s->error ? s->error : (s->done ? ZB200_INF_OK : ZB200_INF_TRUNCATED);
    const uint64_t have_abs = s->in_erased + s->in.size() - s->prime_pref;       // caller's stream position up to which input is held
    const uint64_t again = have_abs > total_in0 ? have_abs - total_in0 : 0;      // bytes of this call that were given back earlier
    const uInt skip = again < in0 ? (uInt)again : in0;
    const uInt fresh = in0 - skip;
    // a call without new input while the cut run of a re-based stream waits: its complete blocks are owed to the caller
    const bool flush_tail = fresh == 0 && s->rebased && !s->seq_ready && !s->trailer_pending && !s->in.empty() && s->out.empty() && s->bounds.empty();
    const bool owed_full = s->out.size() - s->out_pos > kOwedMax;   // the caller drains first (bounds the host-side queue)
    const uInt take = fresh > kTakeMax ? (uInt)kTakeMax : fresh;
    uint64_t fed_abs = total_in0 + skip;                            // ... after this call's new bytes
    if (!s->done && !s->error && !owed_full && (take || s->retry || flush_tail)) {
        s->retry = false;
        if (s->prime_bits && s->in.empty() && !s->in_erased) {
            // inflatePrime: the k = prime_bits % 8 oldest bits sit in the top of a synthetic first byte that the engine
            // enters at bit 8 - k (raw streams only); the whole bytes behind them go ahead of the input as they are
            const unsigned k = s->prime_bits & 7u;
            if (k) {
                s->in.push_back((uint8_t)((s->prime_hold & ((1u << k) - 1u)) << (8 - k)));
                s->prime_hold >>= k; s->prime_bits -= k;
                s->resume_bit = 8 - k; s->resume_out = 0; s->kind = 0;
            }
            for (; s->prime_bits; s->prime_bits -= 8, s->prime_hold >>= 8) s->in.push_back((uint8_t)s->prime_hold);
            s->prime_pref = s->in.size();
            s->prime_byte = true;
        }
        s->in.insert(s->in.end(), in_ptr0 + skip, in_ptr0 + skip + take);
        fed_abs += take;
        if (!s->hdr_kind && !s->rebased && !s->in_erased && s->wrap != ZB200_WRAP_RAW && s->in.size() >= 2)
            s->hdr_kind = ((s->wrap & ZB200_WRAP_GZIP) && s->in[0] == 0x1f && s->in[1] == 0x8b) ? 2 : 1;   // inflate.c:622-669
        if (s->gzhead != Z_NULL && s->gzhead->done == 0 && !s->rebased) {   // inflateGetHeader (inflate.c:1331-1345), while the header bytes are here
            if (s->in.size() >= 2 && !(s->in[0] == 0x1f && s->in[1] == 0x8b)) s->gzhead->done = -1;
            else if (s->in.size() >= 2 && (s->wrap & ZB200_WRAP_GZIP)) fill_gz_header(s->in, *s->gzhead);
        }
        status = -2;
        size_t direct_len = 0;                                  // bytes a batch step decoded straight into the caller's buffer
        if (s->trailer_pending) status = finish_rebased(*s, s->stream_kind == 1 ? 1u : 0u, 0, 0);   // (combining with an empty tail leaves check_before)
        else if (flush_tail || s->want_bounds || parallel_step(*s, &status, strm->next_out, strm->avail_out, &direct_len) == 0) status = -2;   // (the run-parallel step logs no block boundaries)
        if (direct_len) { strm->next_out += direct_len; strm->avail_out -= (uInt)direct_len; strm->total_out += (uLong)direct_len; }
        if (status == -2) status = inflate_attempt(*s);
        if (status < 0) { strm->msg = "zlib-b200: device error"; return Z_STREAM_ERROR; }
        if (status != ZB200_INF_OK && status != ZB200_INF_TRUNCATED) s->error = status;
        if (status == ZB200_INF_OK) strm->adler = s->check;
    }
    // ---- how far this call goes: output up to `stop_out`, input up to `cons_abs` ----
    uint64_t cons_abs = s->done ? s->in_used - s->prime_pref : fed_abs;          // (bytes behind the stream's end are given back)
    uint64_t stop_out = ~0ull;
    const bool have_bound = block_mode && !s->bounds.empty();
    if (have_bound) {
        const InflateStream::Bound &nb = s->bounds.front();
        stop_out = nb.out_abs;
        const uint64_t upto = (nb.bit_abs + 7) >> 3;
        if (upto < cons_abs) cons_abs = upto;
    }
    if (cons_abs < total_in0) cons_abs = total_in0;                               // (never backwards: a boundary behind what an earlier call reported)
    if (cons_abs > total_in0 + in0) cons_abs = total_in0 + in0;
    // deliver decoded bytes
    const size_t avail = s->out.size() - s->out_pos;
    size_t k = avail < strm->avail_out ? avail : strm->avail_out;
    if (stop_out != ~0ull && (uint64_t)strm->total_out + k > stop_out) k = (size_t)(stop_out - strm->total_out);
    if (k) {
        memcpy(strm->next_out, s->out.data() + s->out_pos, k);
        strm->next_out += k; strm->avail_out -= (uInt)k; strm->total_out += k; s->out_pos += k;
        if (s->out_pos == s->out.size()) { s->out.clear(); s->out_pos = 0; }
    }
    bool at_bound = false;
    if (have_bound && (uint64_t)strm->total_out == stop_out) {                   // the block's output is out in full: report its end
        const InflateStream::Bound nb = s->bounds.front();
        s->bounds.pop_front();
        at_bound = true;
        strm->data_type = (int)((8 - (nb.bit_abs & 7)) & 7) + 128 + (nb.last ? 64 : 0);
    } else {
        strm->data_type = s->done ? 64 : 0;
        while (!block_mode && !s->bounds.empty() && s->bounds.front().out_abs < (uint64_t)strm->total_out) s->bounds.pop_front();
    }
    {
        const uint64_t used = cons_abs - total_in0;
        strm->next_in = in_ptr0 + used; strm->avail_in = in0 - (uInt)used; strm->total_in = (uLong)cons_abs;
    }
    const bool drained = s->out.empty();
    if (at_bound) return (in0 == strm->avail_in && out0 == strm->avail_out) ? Z_BUF_ERROR : Z_OK;
    if (s->error && drained) {
        if (s->error == ZB200_INF_NEED_DICT) { strm->adler = s->dictid; return Z_NEED_DICT; }   // inflate.c:667-669
        strm->msg = zb200_inflate_msg(s->error);
        return Z_DATA_ERROR;
    }
    if (s->done && drained && s->bounds.empty()) return Z_STREAM_END;
    if (s->done && drained && !block_mode) { s->bounds.clear(); return Z_STREAM_END; }
    if ((in0 == strm->avail_in && out0 == strm->avail_out) || (flush == Z_FINISH && !(s->done && drained)))
        return Z_BUF_ERROR;                                     // inflate.c:1259-1261
    return Z_OK;
}

// inflate.c:1278-1312.  zlib streams: only when inflate() has just returned Z_NEED_DICT, and the dictionary's
// Adler-32 must be the header's DICTID (else Z_DATA_ERROR); raw streams: before any output has been produced.
int inflateSetDictionary(z_streamp strm, const Bytef *dictionary, uInt dictLength) {
    InflateStream *s = istate(strm);
    if (!s || dictionary == Z_NULL) return Z_STREAM_ERROR;
    zb200_ctx *ctx = api_ctx();
    if (!ctx) return Z_STREAM_ERROR;
    if (s->wrap != ZB200_WRAP_RAW) {
        if (s->error != ZB200_INF_NEED_DICT) return Z_STREAM_ERROR;
        uint32_t a = 1;
        if (dictLength && zb200_checksum_host(ctx, dictionary, dictLength, ZB200_ADLER32, 0, 1, nullptr, &a) != ZB200_OK) return Z_STREAM_ERROR;
        if (a != s->dictid) return Z_DATA_ERROR;
    } else if (s->decoded || s->done || s->error || s->dict_len) return Z_STREAM_ERROR;
    const size_t keep = dictLength > 32768u ? 32768u : dictLength;
    {
        CtxUse use(ctx, ctx->stream);
        if (cudaSetDevice(ctx->device) != cudaSuccess) return Z_STREAM_ERROR;
        cudaStream_t st = ctx->stream;
        if (dev_grow(&s->d_out, &s->d_out_cap, keep + (1u << 20), 0, st)) return Z_MEM_ERROR;
        if (keep && cudaMemcpyAsync(s->d_out, dictionary + (dictLength - keep), keep, cudaMemcpyHostToDevice, st) != cudaSuccess) return Z_STREAM_ERROR;
        if (cudaStreamSynchronize(st) != cudaSuccess) return Z_STREAM_ERROR;
    }
    s->dict_len = keep;
    hist_push(*s, dictionary + (dictLength - keep), keep);      // (it is window content: inflateGetDictionary returns it)
    if (s->error == ZB200_INF_NEED_DICT) { s->error = 0; s->retry = true; }
    if (s->wrap != ZB200_WRAP_RAW) strm->adler = 1;               // inflate.c:671: adler32(0L, Z_NULL, 0)
    return Z_OK;
}

int inflateGetDictionary(z_streamp strm, Bytef *dictionary, uInt *dictLength) {   // inflate.c:1258-1276: the sliding window = the last 32 KiB produced
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (!s->hist.empty() && dictionary != Z_NULL) memcpy(dictionary, s->hist.data(), s->hist.size());
    if (dictLength != Z_NULL) *dictLength = (uInt)s->hist.size();
    return Z_OK;
}

int inflateGetHeader(z_streamp strm, gz_headerp head) {      // inflate.c:1331-1345
    InflateStream *s = istate(strm);
    if (!s || !(s->wrap & ZB200_WRAP_GZIP)) return Z_STREAM_ERROR;
    s->gzhead = head;
    if (head != Z_NULL) head->done = 0;
    return Z_OK;
}

int inflateResetKeep(z_streamp strm) { return inflateReset(strm); }   // inflate.c:102

int inflateValidate(z_streamp strm, int check) {             // inflate.c:1495-1508
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    s->verify = check != 0;
    return Z_OK;
}

int inflateUndermine(z_streamp strm, int subvert) {          // inflate.c:1478-1493 without INFLATE_ALLOW_INVALID_DISTANCE_TOOFAR_ARRR
    (void)subvert;
    return istate(strm) ? Z_DATA_ERROR : Z_STREAM_ERROR;
}

long inflateMark(z_streamp strm) {                           // inflate.c:1510-1521: this inflate() stops between blocks only (back = -1, nothing pending)
    (void)strm;
    return -(1L << 16);
}

// inflate.c:1352-1372 syncsearch: the pattern 00 00 FF FF, `have` bytes of it matched before buf[0].
static size_t sync_search(unsigned *have, const uint8_t *buf, size_t len) {
    unsigned got = *have;
    size_t next = 0;
    while (next < len && got < 4) {
        if (buf[next] == (got < 2 ? 0 : 0xff)) got++;
        else if (buf[next]) got = 0;
        else got = 4 - got;
        next++;
    }
    *have = got;
    return next;
}

// inflate.c:1375-1421.  The reference starts its search at the point its decoder stopped; this inflate() takes all the
// input of a call, so the search starts in the input it still holds — from the last block boundary decoded — and goes
// on in next_in.  What follows the pattern continues as a stream of deflate blocks with no history and no check value
// (state->wrap &= ~4), the wrapper's trailer still being read at the end; a stream that never saw its header goes on raw.
// total_in already counts the bytes earlier calls took.
int inflateSync(z_streamp strm) {
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (!s->syncing) {                                          // first call: what the stream holds but has not decoded
        const size_t from = s->done ? s->in.size() : (size_t)((s->resume_bit + 7) >> 3);
        if (from < s->in.size()) s->in.erase(s->in.begin(), s->in.begin() + (long)from); else s->in.clear();
        s->sync_have = 0; s->syncing = true;
    } else s->in.clear();
    if (strm->avail_in == 0 && s->in.empty()) return Z_BUF_ERROR;
    std::vector<uint8_t> rest;
    size_t k = sync_search(&s->sync_have, s->in.data(), s->in.size());
    if (s->sync_have == 4) rest.assign(s->in.begin() + (long)k, s->in.end());
    else {
        k = sync_search(&s->sync_have, strm->next_in, strm->avail_in);
        strm->next_in += k; strm->avail_in -= (uInt)k; strm->total_in += k;
    }
    s->in.clear();
    if (s->sync_have != 4) return Z_DATA_ERROR;
    const int kind = s->rebased ? s->stream_kind : s->hdr_kind;   // flags == -1: no header yet, treat as raw
    const int wrap = s->wrap;
    const gz_headerp head = s->gzhead;
    inflate_reset_state(*s);
    s->wrap = wrap; s->gzhead = head; s->started = true; s->kind = kind; s->hdr_kind = kind;
    s->rebased = true; s->stream_kind = kind; s->verify = false;
    s->in.swap(rest);
    return Z_OK;
}

// inflate.c:1431-1437: "inflate is waiting for the length bytes of an empty stored block" — true when everything taken
// so far decodes up to a block boundary that is followed by nothing but a stored-block header (BFINAL 0, type 00)
// and its padding to the byte boundary.
int inflateSyncPoint(z_streamp strm) {
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (s->done || s->error || s->in.empty() || !s->out.empty()) return 0;
    if (s->resume_bit == 0 && !s->rebased && s->wrap != ZB200_WRAP_RAW) return 0;
    const uint64_t end = (uint64_t)s->in.size() * 8, rb = s->resume_bit;
    if (end < rb + 3 || end > rb + 10) return 0;
    for (uint64_t b = rb; b < end; ++b) if ((s->in[(size_t)(b >> 3)] >> (b & 7)) & 1) return 0;
    return 1;
}

// inflate.c:223-240: bits ahead of the first input byte.  Offered where zran-style callers use it — on a stream that has
// taken no input yet; a part of a byte only on a raw stream (the engine can enter a member at a bit offset, a wrapper
// header is parsed from a byte boundary).
int inflatePrime(z_streamp strm, int bits, int value) {
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (bits == 0) return Z_OK;
    if (bits < 0) { s->prime_hold = 0; s->prime_bits = 0; return Z_OK; }
    if (bits > 16 || s->prime_bits + (unsigned)bits > 32) return Z_STREAM_ERROR;
    if (!s->in.empty() || s->in_erased || s->done || s->decoded || s->prime_byte) return Z_STREAM_ERROR;
    if (((s->prime_bits + (unsigned)bits) & 7u) && s->wrap != ZB200_WRAP_RAW) return Z_STREAM_ERROR;
    s->prime_hold += ((uint32_t)value & ((1u << bits) - 1u)) << s->prime_bits;
    s->prime_bits += (unsigned)bits;
    return Z_OK;
}

// inflate.c:1521-1526 counts the decode-table entries of the open block; the tables live in shared memory of whichever
// warp decodes the block and are gone when the kernel returns.
unsigned long inflateCodesUsed(z_streamp strm) { return istate(strm) ? 0ul : (unsigned long)-1; }

int inflateCopy(z_streamp dest, z_streamp source) {          // inflate.c:1433-1476
    InflateStream *s = istate(source);
    if (!s || dest == Z_NULL) return Z_STREAM_ERROR;
    zb200_ctx *ctx = api_ctx();
    if (!ctx) return Z_STREAM_ERROR;
    InflateStream *d = state_clone(dest, *s);
    if (!d) return Z_MEM_ERROR;
    d->d_in = d->d_out = nullptr;
    {
        CtxUse use(ctx, ctx->stream);
        bool ok = cudaSetDevice(ctx->device) == cudaSuccess;
        if (ok && s->d_in) ok = cudaMalloc((void **)&d->d_in, s->d_in_cap) == cudaSuccess &&
                                cudaMemcpyAsync(d->d_in, s->d_in, s->d_in_have, cudaMemcpyDeviceToDevice, ctx->stream) == cudaSuccess;
        if (ok && s->d_out) ok = cudaMalloc((void **)&d->d_out, s->d_out_cap) == cudaSuccess &&
                                 cudaMemcpyAsync(d->d_out, s->d_out, (size_t)s->decoded + s->dict_len, cudaMemcpyDeviceToDevice, ctx->stream) == cudaSuccess;
        if (ok) ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            if (d->d_in) cudaFree(d->d_in);
            if (d->d_out) cudaFree(d->d_out);
            state_delete(d);
            return Z_MEM_ERROR;
        }
    }
    *dest = *source;
    dest->state = reinterpret_cast<struct internal_state *>(d);
    return Z_OK;
}

// ---- inflateBack (infback.c): a raw stream pulled through in() and pushed through out(), window by window.
// The reference decodes straight into the caller's window; here the caller's window is the staging buffer
// between this library's inflate() and out().
struct BackState { uint32_t magic; unsigned char *window; unsigned wsize; };
constexpr uint32_t kBackMagic = 0x5a42424bu;

int inflateBackInit_(z_streamp strm, int windowBits, unsigned char *window, const char *version, int stream_size) {
    if (version == Z_NULL || version[0] != ZLIB_VERSION[0] || stream_size != (int)sizeof(z_stream)) return Z_VERSION_ERROR;
    if (strm == Z_NULL || window == Z_NULL || windowBits < 8 || windowBits > 15) return Z_STREAM_ERROR;   // infback.c:33-35
    strm->msg = Z_NULL;
    if (!api_ctx()) { strm->msg = "zlib-b200: no usable CUDA device (no CPU path)"; return Z_STREAM_ERROR; }
    BackState *b = new (std::nothrow) BackState();
    if (!b) return Z_MEM_ERROR;
    b->magic = kBackMagic; b->window = window; b->wsize = 1u << windowBits;
    strm->state = reinterpret_cast<struct internal_state *>(b);
    return Z_OK;
}

int inflateBack(z_streamp strm, in_func in, void *in_desc, out_func out, void *out_desc) {
    BackState *b = strm ? reinterpret_cast<BackState *>(strm->state) : nullptr;
    if (!b || b->magic != kBackMagic || !in || !out) return Z_STREAM_ERROR;
    strm->msg = Z_NULL;
    z_stream inner;
    memset(&inner, 0, sizeof inner);
    if (inflateInit2_(&inner, -15, ZLIB_VERSION, (int)sizeof(z_stream)) != Z_OK) return Z_MEM_ERROR;
    const unsigned char *next = strm->next_in;
    unsigned have = next != Z_NULL ? strm->avail_in : 0;
    int ret;
    for (;;) {
        if (have == 0) {
            have = in(in_desc, &next);
            if (have == 0) { next = Z_NULL; ret = Z_BUF_ERROR; break; }          // infback.c:PULL: input exhausted
        }
        inner.next_in = const_cast<Bytef *>(next); inner.avail_in = have;
        int r;
        bool out_failed = false;
        do {
            inner.next_out = b->window; inner.avail_out = b->wsize;
            r = inflate(&inner, Z_NO_FLUSH);
            const unsigned got = b->wsize - inner.avail_out;
            if (got && out(out_desc, b->window, got)) { out_failed = true; break; }   // infback.c:ROOM: output error
        } while (inner.avail_out == 0 && r != Z_STREAM_END && r != Z_DATA_ERROR);
        next += have - inner.avail_in; have = inner.avail_in;
        if (out_failed) { ret = Z_BUF_ERROR; break; }
        if (r == Z_STREAM_END) { ret = Z_STREAM_END; break; }
        if (r == Z_DATA_ERROR || r == Z_NEED_DICT) { strm->msg = inner.msg; ret = Z_DATA_ERROR; break; }
        if (r == Z_STREAM_ERROR || r == Z_MEM_ERROR) { strm->msg = inner.msg; ret = r; break; }
    }
    strm->next_in = const_cast<Bytef *>(next); strm->avail_in = have;
    inflateEnd(&inner);
    return ret;
}

int inflateBackEnd(z_streamp strm) {
    BackState *b = strm ? reinterpret_cast<BackState *>(strm->state) : nullptr;
    if (!b || b->magic != kBackMagic) return Z_STREAM_ERROR;
    b->magic = 0;
    delete b;
    strm->state = Z_NULL;
    return Z_OK;
}

int inflateEnd(z_streamp strm) {
    InflateStream *s = istate(strm);
    if (!s) return Z_STREAM_ERROR;
    if (s->d_in) cudaFree(s->d_in);
    if (s->d_out) cudaFree(s->d_out);
    s->magic = 0;
    state_delete(s);
    strm->state = Z_NULL;
    return Z_OK;
}

// ---------------------------------------------------------------------------
uLong compressBound(uLong n) { return bound_for(n, 15, 8) - 18 + 6; }

int compress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen, int level) {
    zb200_ctx *ctx = api_ctx();                                 // compress.c:22-59 in one GPU call
    if (!ctx) return Z_STREAM_ERROR;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    if (level < 0 || level > 9 || !dest || !destLen || (!source && sourceLen)) return Z_STREAM_ERROR;
    size_t cap = *destLen;
    const int r = (sourceLen >= kMultiMin && api_multi())
                      ? zb200_multi_deflate_host(api_multi(), source, sourceLen, api_chunk(), level, 0, api_frame(ZB200_FRAME_ZLIB), 1, dest, &cap, nullptr, nullptr)
                      : zb200_deflate_host(ctx, source ? source : (const Bytef *)"", sourceLen, run_chunk(sourceLen, level == Z_DEFAULT_COMPRESSION ? 6 : level, 0), level,
                                           0, api_frame(ZB200_FRAME_ZLIB), 1, dest, &cap, nullptr, nullptr);
    if (r == ZB200_ERR_OUTPUT) return Z_BUF_ERROR;
    if (r != ZB200_OK) return map_engine_error(r);
    *destLen = cap;
    return Z_OK;
}

int compress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen) {
    return compress2(dest, destLen, source, sourceLen, Z_DEFAULT_COMPRESSION);
}

int uncompress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong *sourceLen) {
    zb200_ctx *ctx = api_ctx();                                 // uncompr.c:27-80
    if (!ctx) return Z_STREAM_ERROR;
    if (!destLen || !sourceLen || (!source && *sourceLen)) return Z_STREAM_ERROR;
    Bytef one[1];
    const bool probe = (*destLen == 0);                         // uncompr.c:37-42: detect "would need output"
    zb200_member m;
    m.in_off = 0; m.in_len = *sourceLen; m.out_off = 0; m.out_cap = probe ? 1 : *destLen; m.resume_bit = m.resume_out = 0; m.dict_len = 0;
    zb200_member_result res;
    // streams with flush points (all of this library's deflate output) are decoded run by run in parallel
    const int r = (!probe && *sourceLen >= 65536)
                      ? zb200_inflate_stream_host(ctx, source, *sourceLen, ZB200_WRAP_ZLIB, dest, *destLen, &res)
                      : zb200_inflate_host(ctx, source, probe ? one : dest, &m, 1, ZB200_WRAP_ZLIB, 1, &res);
    if (r != ZB200_OK) return map_engine_error(r);
    if (res.status == ZB200_INF_OUTPUT_FULL && res.out_len > *destLen) {   // the parallel path copies nothing then; uncompr.c fills dest as far as it goes
        const int r2 = zb200_inflate_host(ctx, source, dest, &m, 1, ZB200_WRAP_ZLIB, 1, &res);
        if (r2 != ZB200_OK) return map_engine_error(r2);
    }
    *sourceLen = res.status == ZB200_INF_OK ? (uLong)res.in_used : *sourceLen;
    *destLen = probe ? 0 : (uLong)res.out_len;
    switch (res.status) {
    case ZB200_INF_OK: return probe && res.out_len ? Z_BUF_ERROR : Z_OK;
    case ZB200_INF_OUTPUT_FULL: return Z_BUF_ERROR;
    case ZB200_INF_TRUNCATED: return Z_DATA_ERROR;              // uncompr.c:76-79
    default: return Z_DATA_ERROR;
    }
}

int uncompress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen) {
    return uncompress2(dest, destLen, source, &sourceLen);
}

// ---------------------------------------------------------------------------
uLong crc32_z(uLong crc, const Bytef *buf, z_size_t len) {
    if (buf == Z_NULL) return 0;                                // crc32.c:697
    zb200_ctx *ctx = api_ctx();
    if (!ctx) die_no_device("crc32");
    uint32_t c = (uint32_t)crc;
    if (len >= kMultiMin && api_multi()) {
        if (zb200_multi_checksum_host(api_multi(), buf, len, ZB200_CRC32, (uint32_t)crc, 1, &c, nullptr) != ZB200_OK) die_no_device("crc32");
        return c;
    }
    if (zb200_checksum_host(ctx, buf, len, ZB200_CRC32, (uint32_t)crc, 1, &c, nullptr) != ZB200_OK) die_no_device("crc32");
    return c;
}
uLong crc32(uLong crc, const Bytef *buf, uInt len) { return crc32_z(crc, buf, len); }
uLong adler32_z(uLong adler, const Bytef *buf, z_size_t len) {
    if (buf == Z_NULL) return 1;                                // adler32.c:81
    zb200_ctx *ctx = api_ctx();
    if (!ctx) die_no_device("adler32");
    uint32_t a = (uint32_t)adler;
    if (len >= kMultiMin && api_multi()) {
        if (zb200_multi_checksum_host(api_multi(), buf, len, ZB200_ADLER32, 0, (uint32_t)adler, nullptr, &a) != ZB200_OK) die_no_device("adler32");
        return a;
    }
    if (zb200_checksum_host(ctx, buf, len, ZB200_ADLER32, 0, (uint32_t)adler, nullptr, &a) != ZB200_OK) die_no_device("adler32");
    return a;
}
uLong adler32(uLong adler, const Bytef *buf, uInt len) { return adler32_z(adler, buf, len); }
uLong crc32_combine(uLong c1, uLong c2, z_off_t len2) { return zb200_crc32_combine((uint32_t)c1, (uint32_t)c2, (uint64_t)len2); }
uLong crc32_combine_gen(z_off_t len2) { return zb200_crc32_combine_gen((uint64_t)len2); }
uLong crc32_combine_op(uLong c1, uLong c2, uLong op) { return zb200_crc32_combine_op((uint32_t)c1, (uint32_t)c2, (uint32_t)op); }
uLong adler32_combine(uLong a1, uLong a2, z_off_t len2) { return zb200_adler32_combine((uint32_t)a1, (uint32_t)a2, (int64_t)len2); }
// crc32.c:1021,1034, adler32.c:162: what zlib.h renames the combine functions to under _FILE_OFFSET_BITS=64 (z_off64_t == long here)
uLong crc32_combine64(uLong c1, uLong c2, long len2) { return crc32_combine(c1, c2, len2); }
uLong crc32_combine_gen64(long len2) { return crc32_combine_gen(len2); }
uLong adler32_combine64(uLong a1, uLong a2, long len2) { return adler32_combine(a1, a2, len2); }

// crc32.c:549 get_crc_table: the 256-entry byte-wise table of the reflected polynomial
// 0xedb88320 (crc32.c:149), generated rather than transcribed (crc32.h:5-58 holds the same values).
const z_crc_t *get_crc_table(void) {
    static z_crc_t table[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (unsigned n = 0; n < 256; ++n) {
            unsigned c = n;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[n] = c;
        }
    });
    return table;
}

}  // extern "C"
