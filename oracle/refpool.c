/* refpool.c — TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The CPU baseline BASELINE.md §3 asks for: the UNMODIFIED reference
 * (oracle/_ref/libzref.so, zlib 1.3.1.1-motley compiled in place, symbols
 * prefixed z_) driven by a pthread pool —
 *   deflate : one z_stream per thread, deflateReset per chunk, deflate(Z_FULL_FLUSH)
 *             per chunk and Z_FINISH on the stream's last chunk (deflate.c:954-1263);
 *             after a full flush the reference's state equals a reset one
 *             (deflate.c:1219-1231), so the chunks laid end to end are byte for byte
 *             the stream ONE z_stream emits for the same chunking
 *   inflate : one z_stream per thread, inflateReset2 per member (inflate.c:590)
 *   checksum: crc32_z + adler32_z per range, merged with crc32_combine /
 *             adler32_combine (crc32.c:694,1021; adler32.c:61,133)
 * threads = what the caller passes (bench.py: the host's core count, printed),
 * clock_gettime(CLOCK_MONOTONIC), best of `reps`.
 *
 * Built by oracle/Makefile into oracle/_ref/librefpool.so against the reference's own
 * zlib.h; only bench.py's cpu_baseline / --impl reference legs and tests/ load it.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "zlib.h"

typedef struct {
    double best_s;          /* best wall time of `reps` passes                         */
    double first_s;         /* the first (cold) pass                                   */
    uint64_t out_bytes;     /* bytes produced by one pass                              */
    int32_t threads;
    int32_t err;            /* 0, or the first zlib error code seen                    */
} rp_result;

typedef struct { uint64_t in_off, in_len, out_off, out_cap; } rp_member;

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* ---- a tiny work-sharing pool: `n_items` items handed out by an atomic counter ---- */
typedef struct job job;
typedef void (*item_fn)(job *, int tid, size_t item, void **tls);
struct job {
    item_fn fn;
    size_t n_items;
    size_t next;            /* atomic */
    int err;                /* atomic, first error */
    /* deflate */
    const uint8_t *in; size_t n, chunk; int level, strategy, wbits, mem_level;
    uint8_t *slots; size_t slot; uint64_t *sizes;
    /* inflate */
    const rp_member *members; uint8_t *out; uint64_t *out_lens; int allow_open;
    const rp_member *dmembers;   /* deflate of members */
    /* checksum */
    uint32_t *crcs, *adlers; size_t piece;
};
typedef struct { job *j; int tid; } worker_arg;

static void *worker(void *p) {
    worker_arg *a = (worker_arg *)p;
    job *j = a->j;
    void *tls = NULL;
    for (;;) {
        size_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n_items) break;
        j->fn(j, a->tid, i, &tls);
    }
    if (tls) {                                   /* a z_stream: either kind ends the same way here */
        z_stream *s = (z_stream *)tls;
        if (j->members) inflateEnd(s); else deflateEnd(s);
        free(s);
    }
    return NULL;
}

static int run_pool(job *j, int threads) {
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    worker_arg *args = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)threads);
    j->next = 0;
    int started = 0;
    for (int t = 0; t < threads; ++t) {
        args[t].j = j; args[t].tid = t;
        if (pthread_create(&th[t], NULL, worker, &args[t]) != 0) break;
        ++started;
    }
    if (started == 0) { worker_arg a = {j, 0}; worker(&a); }
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th); free(args);
    return j->err;
}

static void set_err(job *j, int e) { int z = 0; __atomic_compare_exchange_n(&j->err, &z, e, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED); }

/* ---- deflate -------------------------------------------------------------------- */
static void deflate_item(job *j, int tid, size_t c, void **tls) {
    (void)tid;
    z_stream *s = (z_stream *)*tls;
    if (!s) {
        s = (z_stream *)calloc(1, sizeof(z_stream));
        int r = deflateInit2(s, j->level, Z_DEFLATED, j->wbits, j->mem_level, j->strategy);
        if (r != Z_OK) { set_err(j, r); free(s); return; }
        *tls = s;
    } else {
        deflateReset(s);
    }
    const size_t off = c * j->chunk;
    const size_t len = j->n - off < j->chunk ? j->n - off : j->chunk;
    const int last = off + len >= j->n;
    s->next_in = (Bytef *)(j->in + off); s->avail_in = (uInt)len;
    s->next_out = j->slots + c * j->slot; s->avail_out = (uInt)j->slot;
    int r = deflate(s, last ? Z_FINISH : Z_FULL_FLUSH);
    if ((last && r != Z_STREAM_END) || (!last && (r != Z_OK || s->avail_out == 0))) { set_err(j, r ? r : Z_BUF_ERROR); return; }
    j->sizes[c] = j->slot - s->avail_out;
}

/* Compress in[0..n) as ONE raw deflate stream cut into `chunk`-byte Z_FULL_FLUSH runs (windowBits -15,
 * memLevel 8), chunk-parallel.  out (optional, out_cap bytes) receives the stream, chunk_end[i]
 * (optional) the exclusive end offset of chunk i inside it.  The timed region is the compression
 * into per-chunk slots; laying the slots end to end afterwards is one memcpy per chunk, untimed. */
int rp_deflate(const uint8_t *in, size_t n, size_t chunk, int level, int strategy, int threads, int reps,
               uint8_t *out, size_t out_cap, uint64_t *chunk_end, rp_result *res) {
    if (!in || !n || !chunk || !res) return Z_STREAM_ERROR;
    job j;
    memset(&j, 0, sizeof j);
    j.fn = deflate_item; j.in = in; j.n = n; j.chunk = chunk; j.level = level; j.strategy = strategy;
    j.wbits = -15; j.mem_level = 8;
    j.n_items = (n + chunk - 1) / chunk;
    j.slot = chunk + (chunk >> 3) + 1024;                /* above deflateBound for any level/strategy */
    j.slots = (uint8_t *)malloc(j.n_items * j.slot);
    j.sizes = (uint64_t *)calloc(j.n_items, sizeof(uint64_t));
    if (!j.slots || !j.sizes) { free(j.slots); free(j.sizes); return Z_MEM_ERROR; }
    res->best_s = 1e30; res->first_s = 0; res->threads = threads; res->err = 0;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        const double t0 = now_s();
        run_pool(&j, threads);
        const double dt = now_s() - t0;
        if (r == 0) res->first_s = dt;
        if (dt < res->best_s) res->best_s = dt;
        if (j.err) break;
    }
    uint64_t pos = 0;
    for (size_t c = 0; c < j.n_items && !j.err; ++c) {
        if (out) {
            if (pos + j.sizes[c] > out_cap) { j.err = Z_BUF_ERROR; break; }
            memcpy(out + pos, j.slots + c * j.slot, j.sizes[c]);
        }
        pos += j.sizes[c];
        if (chunk_end) chunk_end[c] = pos;
    }
    res->out_bytes = pos; res->err = j.err;
    free(j.slots); free(j.sizes);
    return j.err;
}

/* Independent members of any sizes (config C3's gzip members: deflateInit2(level, 15+16, 8, strategy) + Z_FINISH each),
 * member-parallel: member i = in[in_off, in_off+in_len) -> out[out_off ..), at most out_cap bytes; out_lens[i] = its size. */
static void deflate_member_item(job *j, int tid, size_t m, void **tls) {
    (void)tid;
    z_stream *s = (z_stream *)*tls;
    if (!s) {
        s = (z_stream *)calloc(1, sizeof(z_stream));
        int r = deflateInit2(s, j->level, Z_DEFLATED, j->wbits, j->mem_level, j->strategy);
        if (r != Z_OK) { set_err(j, r); free(s); return; }
        *tls = s;
    } else {
        deflateReset(s);
    }
    const rp_member *mb = &j->dmembers[m];
    s->next_in = (Bytef *)(j->in + mb->in_off); s->avail_in = (uInt)mb->in_len;
    s->next_out = j->out + mb->out_off; s->avail_out = (uInt)mb->out_cap;
    int r = deflate(s, Z_FINISH);
    if (r != Z_STREAM_END) { set_err(j, r ? r : Z_BUF_ERROR); return; }
    j->out_lens[m] = mb->out_cap - s->avail_out;
}
int rp_deflate_members(const uint8_t *in, const rp_member *members, size_t n_members, int level, int strategy, int wbits,
                       int threads, int reps, uint8_t *out, uint64_t *out_lens, rp_result *res) {
    if (!in || !members || !n_members || !out || !out_lens || !res) return Z_STREAM_ERROR;
    job j;
    memset(&j, 0, sizeof j);
    j.fn = deflate_member_item; j.in = in; j.dmembers = members; j.n_items = n_members; j.level = level; j.strategy = strategy;
    j.wbits = wbits; j.mem_level = 8; j.out = out; j.out_lens = out_lens;
    res->best_s = 1e30; res->first_s = 0; res->threads = threads; res->err = 0;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        const double t0 = now_s();
        run_pool(&j, threads);
        const double dt = now_s() - t0;
        if (r == 0) res->first_s = dt;
        if (dt < res->best_s) res->best_s = dt;
        if (j.err) break;
    }
    uint64_t tot = 0;
    for (size_t m = 0; m < n_members; ++m) tot += out_lens[m];
    res->out_bytes = tot; res->err = j.err;
    return j.err;
}

/* ---- inflate -------------------------------------------------------------------- */
static void inflate_item(job *j, int tid, size_t m, void **tls) {
    (void)tid;
    z_stream *s = (z_stream *)*tls;
    if (!s) {
        s = (z_stream *)calloc(1, sizeof(z_stream));
        int r = inflateInit2(s, j->wbits);
        if (r != Z_OK) { set_err(j, r); free(s); return; }
        *tls = s;
    } else {
        inflateReset2(s, j->wbits);
    }
    const rp_member *mb = &j->members[m];
    uint64_t in_left = mb->in_len, out_left = mb->out_cap;
    const uint8_t *ip = j->in + mb->in_off;
    uint8_t *op = j->out + mb->out_off;
    int r = Z_OK;
    s->avail_in = 0; s->avail_out = 0;
    while (r == Z_OK) {                                   /* uInt windows over 64-bit lengths (uncompr.c:27-85) */
        if (s->avail_in == 0) { uInt k = in_left > 0x40000000u ? 0x40000000u : (uInt)in_left; s->next_in = (Bytef *)ip; s->avail_in = k; ip += k; in_left -= k; }
        if (s->avail_out == 0) { uInt k = out_left > 0x40000000u ? 0x40000000u : (uInt)out_left; s->next_out = op; s->avail_out = k; op += k; out_left -= k; }
        r = inflate(s, Z_NO_FLUSH);
        if (r == Z_OK && s->avail_in == 0 && in_left == 0) r = Z_BUF_ERROR;
        if (r == Z_OK && s->avail_out == 0 && out_left == 0) r = Z_BUF_ERROR;
    }
    /* allow_open: a run of blocks that ends at a flush point instead of a final block (one Z_FULL_FLUSH chunk of a
     * longer stream) is complete when all its input is consumed and all its expected output produced */
    if (r != Z_STREAM_END && !(j->allow_open && r == Z_BUF_ERROR && s->avail_in == 0 && in_left == 0 && s->total_out == mb->out_cap)) {
        set_err(j, r); return;
    }
    j->out_lens[m] = s->total_out;
}

/* Decode n_members independent members (wbits as for inflateInit2: 31 gzip, 15 zlib, -15 raw), member-parallel. */
static int inflate_pool(const uint8_t *in, const rp_member *members, size_t n_members, int wbits, int threads, int reps,
                        uint8_t *out, uint64_t *out_lens, rp_result *res, int allow_open) {
    if (!in || !members || !n_members || !out || !res) return Z_STREAM_ERROR;
    job j;
    memset(&j, 0, sizeof j);
    j.allow_open = allow_open;
    j.fn = inflate_item; j.in = in; j.members = members; j.n_items = n_members; j.wbits = wbits; j.out = out;
    j.out_lens = out_lens ? out_lens : (uint64_t *)calloc(n_members, sizeof(uint64_t));
    res->best_s = 1e30; res->first_s = 0; res->threads = threads; res->err = 0;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        const double t0 = now_s();
        run_pool(&j, threads);
        const double dt = now_s() - t0;
        if (r == 0) res->first_s = dt;
        if (dt < res->best_s) res->best_s = dt;
        if (j.err) break;
    }
    uint64_t tot = 0;
    for (size_t m = 0; m < n_members; ++m) tot += j.out_lens[m];
    res->out_bytes = tot; res->err = j.err;
    if (!out_lens) free(j.out_lens);
    return j.err;
}

int rp_inflate(const uint8_t *in, const rp_member *members, size_t n_members, int wbits, int threads, int reps,
               uint8_t *out, uint64_t *out_lens, rp_result *res) {
    return inflate_pool(in, members, n_members, wbits, threads, reps, out, out_lens, res, 0);
}
/* ... members that are Z_FULL_FLUSH-terminated runs of a longer raw stream (no final block): out_cap must be exact. */
int rp_inflate_open(const uint8_t *in, const rp_member *members, size_t n_members, int wbits, int threads, int reps,
                    uint8_t *out, uint64_t *out_lens, rp_result *res) {
    return inflate_pool(in, members, n_members, wbits, threads, reps, out, out_lens, res, 1);
}

/* ---- checksums ------------------------------------------------------------------- */
static void checksum_item(job *j, int tid, size_t i, void **tls) {
    (void)tid; (void)tls;
    const size_t off = i * j->piece;
    const size_t len = j->n - off < j->piece ? j->n - off : j->piece;
    if (j->crcs) j->crcs[i] = (uint32_t)crc32_z(0, j->in + off, len);
    if (j->adlers) j->adlers[i] = (uint32_t)adler32_z(1, j->in + off, len);
}

/* which: bit 0 CRC-32, bit 1 Adler-32.  One contiguous range per thread, partials folded left to right. */
int rp_checksum(const uint8_t *in, size_t n, int which, int threads, int reps, uint32_t *crc, uint32_t *adler, rp_result *res) {
    if (!in || !res) return Z_STREAM_ERROR;
    if (threads < 1) threads = 1;
    job j;
    memset(&j, 0, sizeof j);
    j.fn = checksum_item; j.in = in; j.n = n;
    j.piece = ((n + (size_t)threads - 1) / (size_t)threads + 63) & ~(size_t)63;
    if (j.piece == 0) j.piece = 64;
    j.n_items = n ? (n + j.piece - 1) / j.piece : 0;
    j.crcs = (which & 1) ? (uint32_t *)calloc(j.n_items + 1, 4) : NULL;
    j.adlers = (which & 2) ? (uint32_t *)calloc(j.n_items + 1, 4) : NULL;
    res->best_s = 1e30; res->first_s = 0; res->threads = threads; res->err = 0;
    uint32_t c = 0, a = 1;
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        const double t0 = now_s();
        run_pool(&j, threads);
        c = 0; a = 1;
        for (size_t i = 0; i < j.n_items; ++i) {
            const size_t off = i * j.piece;
            const size_t len = n - off < j.piece ? n - off : j.piece;
            if (j.crcs) c = (uint32_t)crc32_combine(c, j.crcs[i], (z_off_t)len);
            if (j.adlers) a = (uint32_t)adler32_combine(a, j.adlers[i], (z_off_t)len);
        }
        const double dt = now_s() - t0;
        if (r == 0) res->first_s = dt;
        if (dt < res->best_s) res->best_s = dt;
    }
    if (crc) *crc = c;
    if (adler) *adler = a;
    res->out_bytes = n;
    free(j.crcs); free(j.adlers);
    return 0;
}

/* The reference's own one-shot compress2 over a 64-bit length (compress.c:22-59 loops over uInt windows the
 * same way): ONE z_stream, no flush points — what history carry-over is compared against. */
int rp_compress2_whole(const uint8_t *in, size_t n, int level, int strategy, int wbits, int mem_level,
                       uint8_t *out, size_t out_cap, size_t *out_len, double *seconds) {
    z_stream s;
    memset(&s, 0, sizeof s);
    int r = deflateInit2(&s, level, Z_DEFLATED, wbits, mem_level, strategy);
    if (r != Z_OK) return r;
    const double t0 = now_s();
    size_t in_left = n, out_left = out_cap;
    const uint8_t *ip = in;
    uint8_t *op = out;
    s.avail_in = 0; s.avail_out = 0;
    do {
        if (s.avail_in == 0) { uInt k = in_left > 0x40000000u ? 0x40000000u : (uInt)in_left; s.next_in = (Bytef *)ip; s.avail_in = k; ip += k; in_left -= k; }
        if (s.avail_out == 0) { uInt k = out_left > 0x40000000u ? 0x40000000u : (uInt)out_left; s.next_out = op; s.avail_out = k; op += k; out_left -= k; }
        r = deflate(&s, in_left ? Z_NO_FLUSH : Z_FINISH);
    } while (r == Z_OK);
    if (seconds) *seconds = now_s() - t0;
    if (out_len) *out_len = (size_t)(op - out) - s.avail_out;
    deflateEnd(&s);
    return r == Z_STREAM_END ? Z_OK : (r == Z_OK || r == Z_BUF_ERROR ? Z_BUF_ERROR : r);
}
