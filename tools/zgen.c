/* zgen.c — deterministic synthetic inputs for tests and bench.
 *
 * Not part of the product library and not part of the oracle: this only makes
 * bytes.  Every generator is BLOCK-PARALLEL: the buffer is cut into 64 KiB
 * blocks and block b is a pure function of (seed, b), so the same bytes come
 * out for any thread count, any prefix length, and any split across GPUs/ranks
 * (a rank generates only its own byte range by passing `first_block`).
 *
 *   ZG_TEXT    order-0 word text  : 4096-word vocabulary, skewed pick, ' ' / '\n'
 *   ZG_MARKOV  order-1 word text  : with p=1/2 the next word comes from a fixed
 *                                   16-entry successor table of the previous word
 *   ZG_RANDOM  xorshift64* bytes  (incompressible)
 *   ZG_MIXED   seeded segments of 1..16 blocks: 40% markov text, 20% random,
 *              15% zero / long runs, 15% u32 counters + slowly varying floats,
 *              10% PNG-filter-like small signed deltas
 *   ZG_BYTES   50% word text + 50% random, alternating blocks (config C2)
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define ZG_BLOCK 65536u
enum { ZG_TEXT = 0, ZG_MARKOV = 1, ZG_RANDOM = 2, ZG_MIXED = 3, ZG_BYTES = 4 };

static inline uint64_t zg_next(uint64_t *s) {          /* xorshift64* */
    uint64_t x = *s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1Dull;
}
static inline uint64_t zg_mix(uint64_t z) {            /* splitmix64 finalizer */
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct {
    uint64_t seed;
    uint8_t word[4096][12]; uint8_t wlen[4096];
    uint16_t succ[4096][16];
} zg_vocab;

static zg_vocab *zg_voc;                               /* one cached vocabulary */

static void zg_vocab_build(zg_vocab *v, uint64_t seed) {
    uint64_t s = zg_mix(seed ^ 0x766f636162ull) | 1;
    v->seed = seed;
    for (int i = 0; i < 4096; i++) {
        uint64_t r = zg_next(&s);
        int len = 2 + (int)(r % 9);                    /* 2..10 */
        v->wlen[i] = (uint8_t)len;
        for (int k = 0; k < len; k++) { r = zg_next(&s); v->word[i][k] = (uint8_t)('a' + (r >> 33) % 26); }
    }
    for (int i = 0; i < 4096; i++)
        for (int k = 0; k < 16; k++) v->succ[i][k] = (uint16_t)((zg_next(&s) >> 20) & 0xfff);
}

static inline unsigned zg_pick(uint64_t r) {           /* skewed vocabulary index */
    unsigned k = (unsigned)(r & 0xfff);
    if (r & 0x1000) k &= 0xff;
    if (r & 0x2000) k &= 0x3f;
    return k;
}

static void zg_block_text(const zg_vocab *v, uint8_t *dst, size_t n, uint64_t s, int markov) {
    size_t o = 0; unsigned prev = zg_pick(zg_next(&s));
    while (o < n) {
        uint64_t r = zg_next(&s);
        unsigned w = (markov && (r & 0x4000)) ? v->succ[prev][(r >> 16) & 15] : zg_pick(r);
        prev = w;
        unsigned len = v->wlen[w];
        for (unsigned k = 0; k < len && o < n; k++) dst[o++] = v->word[w][k];
        if (o < n) dst[o++] = ((r >> 40) % 17 == 0) ? '\n' : ' ';
    }
}
static void zg_block_random(uint8_t *dst, size_t n, uint64_t s) {
    size_t o = 0;
    for (; o + 8 <= n; o += 8) { uint64_t r = zg_next(&s); memcpy(dst + o, &r, 8); }
    if (o < n) { uint64_t r = zg_next(&s); memcpy(dst + o, &r, n - o); }
}
static void zg_block_runs(uint8_t *dst, size_t n, uint64_t s) {
    size_t o = 0;
    while (o < n) {
        uint64_t r = zg_next(&s);
        size_t run = 1 + (size_t)(r % 4000);
        uint8_t b = (r & 0x30000) ? 0 : (uint8_t)(r >> 24);
        if (run > n - o) run = n - o;
        memset(dst + o, b, run); o += run;
    }
}
static void zg_block_numeric(uint8_t *dst, size_t n, uint64_t s, uint64_t blk) {
    uint32_t ctr = (uint32_t)(blk * (ZG_BLOCK / 8));
    float f = (float)(zg_next(&s) % 1000);
    size_t o = 0;
    while (o < n) {
        uint8_t rec[8];
        uint32_t c = ctr++;
        f += (float)((int)(zg_next(&s) % 7) - 3) * 0.125f;
        memcpy(rec, &c, 4); memcpy(rec + 4, &f, 4);
        size_t k = n - o < 8 ? n - o : 8;
        memcpy(dst + o, rec, k); o += k;
    }
}
static void zg_block_deltas(uint8_t *dst, size_t n, uint64_t s) {
    for (size_t o = 0; o < n; o++) {
        uint64_t r = zg_next(&s);
        int d = (int)(r % 5) - 2;
        if ((r >> 8) % 23 == 0) d = (int)((r >> 16) % 33) - 16;
        dst[o] = (uint8_t)d;
    }
}

static int zg_mixed_kind(uint64_t seed, uint64_t blk) {
    /* segments: walk back to the segment start; segment lengths are 1..16
     * blocks and are found by hashing 16-block groups so any block can locate
     * its segment without a sequential pass */
    uint64_t grp = blk >> 4, h = zg_mix(seed ^ (grp * 0x9E3779B97F4A7C15ull));
    unsigned cut = 1 + (unsigned)(h & 15);             /* group = [0,cut) + [cut,16) */
    uint64_t segid = grp * 2 + ((blk & 15) >= cut);
    unsigned p = (unsigned)(zg_mix(seed ^ segid ^ 0x6d69786564ull) % 100);
    return p < 40 ? 0 : p < 60 ? 1 : p < 75 ? 2 : p < 90 ? 3 : 4;
}

/* Fill dst[0..n) with the bytes of blocks first_block, first_block+1, ...
 * (n need not be a multiple of the block size: the last block is cut). */
void zgen_fill(uint8_t *dst, size_t n, int kind, uint64_t seed, uint64_t first_block) {
    if (kind != ZG_RANDOM && (!zg_voc || zg_voc->seed != seed)) {
        if (!zg_voc) zg_voc = (zg_vocab *)malloc(sizeof *zg_voc);
        zg_vocab_build(zg_voc, seed);
    }
    const zg_vocab *v = zg_voc;
    long nblk = (long)((n + ZG_BLOCK - 1) / ZG_BLOCK);
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < nblk; i++) {
        uint64_t blk = first_block + (uint64_t)i;
        size_t off = (size_t)i * ZG_BLOCK, len = n - off < ZG_BLOCK ? n - off : ZG_BLOCK;
        uint64_t s = zg_mix(seed ^ zg_mix(blk)) | 1;
        uint8_t *p = dst + off;
        switch (kind) {
        case ZG_TEXT:   zg_block_text(v, p, len, s, 0); break;
        case ZG_MARKOV: zg_block_text(v, p, len, s, 1); break;
        case ZG_RANDOM: zg_block_random(p, len, s); break;
        case ZG_BYTES:  if (blk & 1) zg_block_random(p, len, s); else zg_block_text(v, p, len, s, 0); break;
        default:
            switch (zg_mixed_kind(seed, blk)) {
            case 0: zg_block_text(v, p, len, s, 1); break;
            case 1: zg_block_random(p, len, s); break;
            case 2: zg_block_runs(p, len, s); break;
            case 3: zg_block_numeric(p, len, s, blk); break;
            default: zg_block_deltas(p, len, s); break;
            }
        }
    }
}

/* Member / chunk size schedule for config C3: log-uniform in [lo, hi]. */
uint64_t zgen_member_size(uint64_t seed, uint64_t index, uint64_t lo, uint64_t hi) {
    uint64_t r = zg_mix(seed ^ zg_mix(index ^ 0x6d656d62ull));
    double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
    double ratio = (double)hi / (double)lo, x = 1.0;
    /* lo * ratio^u without libm: 32 square-root-free steps of exponentiation by bits */
    double base = ratio;                               /* ratio^(2^-k) via Newton sqrt */
    for (int k = 0; k < 30; k++) {
        double sq = base > 1 ? base / 2 + 0.5 : 1.0;   /* sqrt(base) by Newton */
        for (int it = 0; it < 30; it++) sq = 0.5 * (sq + base / sq);
        base = sq;
        u *= 2.0;
        if (u >= 1.0) { x *= base; u -= 1.0; }
    }
    uint64_t v = (uint64_t)((double)lo * x);
    return v < lo ? lo : v > hi ? hi : v;
}
