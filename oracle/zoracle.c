/* zoracle.c — CPU ORACLE for the zlib hot path.  TEST INFRASTRUCTURE ONLY.
 * See zoracle.h for the role of this file and how it is pinned to the
 * reference.  Plain C, scalar, single thread.  Written from the reference's
 * algorithm descriptions; citations are file:line under /root/reference.
 */
#include "zoracle.h"
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* CRC-32                                                                   */
/* ======================================================================== */
#define ZO_POLY 0xedb88320u            /* crc32.c:149  reflected polynomial  */

static uint32_t zo_crc_tab[256];
static uint32_t zo_x2n_tab[32];
static int zo_crc_ready;

/* crc32.c:155-170: a(x)*b(x) mod p(x); bit 31 is the x^0 coefficient. */
uint32_t zo_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 0x80000000u, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ ZO_POLY : b >> 1;
    }
    return p;
}

static void zo_crc_init(void) {
    /* crc32.c:303-323 make_crc_table: byte table and x^(2^n) table */
    for (unsigned i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ ZO_POLY : c >> 1;
        zo_crc_tab[i] = c;
    }
    uint32_t p = 0x40000000u;          /* x^1 */
    zo_x2n_tab[0] = p;
    for (int n = 1; n < 32; n++) zo_x2n_tab[n] = p = zo_multmodp(p, p);
    zo_crc_ready = 1;
}

/* crc32.c:176-187: x^(n * 2^k) mod p(x) */
uint32_t zo_x2nmodp(uint64_t n, unsigned k) {
    if (!zo_crc_ready) zo_crc_init();
    uint32_t p = 0x80000000u;          /* x^0 == 1 */
    while (n) {
        if (n & 1) p = zo_multmodp(zo_x2n_tab[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

/* crc32.c:694-1010 computes this same function with a braided word-at-a-time
 * schedule; the value is defined by the byte recurrence (crc32.c:716-719). */
uint32_t zo_crc32(uint32_t crc, const uint8_t *buf, size_t len) {
    if (!zo_crc_ready) zo_crc_init();
    if (buf == NULL) return 0;         /* crc32.c:697 */
    uint32_t c = ~crc;
    while (len--) c = zo_crc_tab[(c ^ *buf++) & 0xff] ^ (c >> 8);
    return ~c;
}

uint32_t zo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    return zo_multmodp(zo_x2nmodp(len2, 3), crc1) ^ crc2;     /* crc32.c:1021-1026 */
}
uint32_t zo_crc32_combine_gen(uint64_t len2) { return zo_x2nmodp(len2, 3); } /* :1034 */
uint32_t zo_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op) {
    return zo_multmodp(op, crc1) ^ crc2;                                     /* :1047 */
}

/* ======================================================================== */
/* Adler-32                                                                 */
/* ======================================================================== */
#define ZO_BASE 65521u                 /* adler32.c:10 */
#define ZO_NMAX 5552                   /* adler32.c:11 */

uint32_t zo_adler32(uint32_t adler, const uint8_t *buf, size_t len) {
    if (buf == NULL) return 1;         /* adler32.c:81-82 */
    uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;
    while (len) {                      /* adler32.c:94-121: defer the modulo NMAX bytes */
        size_t k = len < ZO_NMAX ? len : ZO_NMAX;
        len -= k;
        while (k--) { a += *buf++; b += a; }
        a %= ZO_BASE; b %= ZO_BASE;
    }
    return a | (b << 16);
}

uint32_t zo_adler32_combine(uint32_t a1, uint32_t a2, int64_t len2) {
    if (len2 < 0) return 0xffffffffu;  /* adler32.c:139-140 */
    uint32_t rem = (uint32_t)(len2 % ZO_BASE);
    uint32_t s1 = a1 & 0xffff;
    uint32_t s2 = (rem * s1) % ZO_BASE;
    s1 += (a2 & 0xffff) + ZO_BASE - 1;
    s2 += ((a1 >> 16) & 0xffff) + ((a2 >> 16) & 0xffff) + ZO_BASE - rem;
    if (s1 >= ZO_BASE) s1 -= ZO_BASE;
    if (s1 >= ZO_BASE) s1 -= ZO_BASE;
    if (s2 >= (ZO_BASE << 1)) s2 -= (ZO_BASE << 1);
    if (s2 >= ZO_BASE) s2 -= ZO_BASE;
    return s1 | (s2 << 16);
}

/* ======================================================================== */
/* Shared deflate-format constants (trees.c:62-72, inftrees.c:55-68)        */
/* ======================================================================== */
static const uint8_t zo_len_extra[29] =
    {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint8_t zo_dist_extra[30] =
    {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const uint8_t zo_cl_order[19] =
    {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
static uint16_t zo_len_base[29];       /* match length (3..258) base per code */
static uint16_t zo_dist_base[30];      /* distance (1..32768) base per code   */
static uint8_t  zo_len_code[256];      /* (len-3) -> length code 0..28        */
static int zo_fmt_ready;

static void zo_fmt_init(void) {
    /* trees.c:344-372 tr_static_init: bases from the extra-bit tables */
    unsigned l = 0;
    for (int c = 0; c < 28; c++) {
        zo_len_base[c] = (uint16_t)(l + 3);
        for (unsigned k = 0; k < (1u << zo_len_extra[c]); k++) zo_len_code[l++] = (uint8_t)c;
    }
    zo_len_base[28] = 258;
    zo_len_code[255] = 28;             /* trees.c:356-360: 258 gets its own code */
    unsigned d = 0;
    for (int c = 0; c < 30; c++) {
        zo_dist_base[c] = (uint16_t)(d + 1);
        d += 1u << zo_dist_extra[c];
    }
    zo_fmt_ready = 1;
}

static int zo_dist_code(unsigned dist) {   /* dist 1..32768 -> code; deflate.h:317 d_code */
    int c = 29;
    while (zo_dist_base[c] > dist) c--;
    return c;
}

/* ======================================================================== */
/* Inflate                                                                  */
/* ======================================================================== */
static const char *const zo_msgs[ZO_E_COUNT] = {
    "", "incorrect header check", "unknown compression method", "invalid window size",
    "unknown header flags set", "header crc mismatch", "invalid block type",
    "invalid stored block lengths", "too many length or distance symbols",
    "invalid code lengths set", "invalid bit length repeat",
    "invalid code -- missing end-of-block", "invalid literal/lengths set",
    "invalid distances set", "invalid literal/length code", "invalid distance code",
    "invalid distance too far back", "incorrect data check", "incorrect length check",
    "truncated input", "output buffer full", "need dictionary"
};
const char *zo_inflate_msg(int e) { return (e >= 0 && e < ZO_E_COUNT) ? zo_msgs[e] : "?"; }

typedef struct {
    const uint8_t *src; size_t n, pos;
    uint64_t acc; int cnt;             /* LSB-first bit accumulator (inflate.c:470-505) */
} zo_bits;

/* Ensure `need` (<=32) bits; returns 0 if the input ran dry. */
static int zo_need(zo_bits *b, int need) {
    while (b->cnt < need) {
        if (b->pos >= b->n) return 0;
        b->acc |= (uint64_t)b->src[b->pos++] << b->cnt;
        b->cnt += 8;
    }
    return 1;
}
static uint32_t zo_take(zo_bits *b, int k) {
    uint32_t v = (uint32_t)(b->acc & ((1ull << k) - 1));
    b->acc >>= k; b->cnt -= k;
    return v;
}

/* A canonical Huffman code held as count-per-length + symbols sorted by
 * (length, symbol) — the same canonical assignment inftrees.c:136-147 sorts
 * into `work[]`, decoded bit-serially instead of through lookup tables. */
typedef struct { uint16_t count[16]; uint16_t sym[320]; int complete; int nonempty; } zo_code;

/* inftrees.c:100-133: returns <0 over-subscribed, >0 incomplete, 0 complete */
static int zo_code_build(zo_code *h, const uint16_t *lens, int n) {
    uint16_t offs[16];
    memset(h->count, 0, sizeof h->count);
    for (int i = 0; i < n; i++) h->count[lens[i]]++;
    h->nonempty = (h->count[0] != n);
    int left = 1;
    for (int len = 1; len <= 15; len++) {
        left <<= 1;
        left -= h->count[len];
        if (left < 0) return -1;
    }
    offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + h->count[len];
    for (int i = 0; i < n; i++) if (lens[i]) h->sym[offs[lens[i]]++] = (uint16_t)i;
    h->complete = (left == 0);
    return left;
}

/* Decode one symbol. -1: ran out of input; -2: code not in the (incomplete) set. */
static int zo_code_decode(zo_bits *b, const zo_code *h) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; len++) {
        if (!zo_need(b, 1)) return -1;
        code |= (int)zo_take(b, 1);
        int count = h->count[len];
        if (code - count < first) return h->sym[index + (code - first)];
        /* inftrees.c:111-119,289-295: an empty or single-code incomplete set
         * decodes every other 1-bit pattern to the "invalid code" marker */
        if (!h->complete) return -2;
        index += count; first += count;
        first <<= 1; code <<= 1;
    }
    return -2;
}

static zo_code zo_fix_lit, zo_fix_dist;
static int zo_fix_ready;
static void zo_fixed_init(void) {      /* inflate.c:252-290 fixedtables */
    uint16_t lens[288];
    int i = 0;
    for (; i < 144; i++) lens[i] = 8;
    for (; i < 256; i++) lens[i] = 9;
    for (; i < 280; i++) lens[i] = 7;
    for (; i < 288; i++) lens[i] = 8;
    zo_code_build(&zo_fix_lit, lens, 288);
    for (i = 0; i < 32; i++) lens[i] = 5;
    zo_code_build(&zo_fix_dist, lens, 32);
    zo_fix_ready = 1;
}

/* inflate.c:1026-1182 (LEN..LIT) / inffast.c:100-287: decode symbols of one block */
static int zo_block_codes(zo_bits *b, uint8_t *dst, size_t cap, size_t *outp,
                          const zo_code *lc, const zo_code *dc) {
    size_t out = *outp;
    for (;;) {
        int sym = zo_code_decode(b, lc);
        if (sym == -1) return ZO_E_TRUNCATED;
        if (sym == -2 || sym > 285) return ZO_E_LITLEN_CODE;    /* 286/287: inftrees.c:57-60 op 64 */
        if (sym < 256) {
            if (out >= cap) return ZO_E_OUTPUT_FULL;
            dst[out++] = (uint8_t)sym;
            continue;
        }
        if (sym == 256) break;
        sym -= 257;
        int eb = zo_len_extra[sym];
        if (!zo_need(b, eb)) return ZO_E_TRUNCATED;
        unsigned len = zo_len_base[sym] + zo_take(b, eb);
        int ds = zo_code_decode(b, dc);
        if (ds == -1) return ZO_E_TRUNCATED;
        if (ds == -2 || ds > 29) return ZO_E_DIST_CODE;         /* 30/31: inftrees.c:65-68 */
        eb = zo_dist_extra[ds];
        if (!zo_need(b, eb)) return ZO_E_TRUNCATED;
        size_t dist = zo_dist_base[ds] + zo_take(b, eb);
        if (dist > out) return ZO_E_DIST_FAR;                   /* inffast.c:152-161 */
        if (out + len > cap) return ZO_E_OUTPUT_FULL;
        for (unsigned k = 0; k < len; k++, out++) dst[out] = dst[out - dist];  /* byte-serial: inffast.c:249-260 */
    }
    *outp = out;
    return ZO_OK;
}

/* inflate.c:898-1022: read a dynamic block header and build both codes */
static int zo_dynamic(zo_bits *b, zo_code *lc, zo_code *dc) {
    uint16_t lens[320];
    if (!zo_need(b, 14)) return ZO_E_TRUNCATED;
    int nlen = (int)zo_take(b, 5) + 257, ndist = (int)zo_take(b, 5) + 1, ncode = (int)zo_take(b, 4) + 4;
    if (nlen > 286 || ndist > 30) return ZO_E_TOO_MANY_SYMS;
    int i;
    for (i = 0; i < ncode; i++) {
        if (!zo_need(b, 3)) return ZO_E_TRUNCATED;
        lens[zo_cl_order[i]] = (uint16_t)zo_take(b, 3);
    }
    for (; i < 19; i++) lens[zo_cl_order[i]] = 0;
    zo_code cl;
    /* inftrees.c:131: an incomplete code-length code is always an error;
     * inftrees.c:111-119: an empty one is accepted here and fails at first use */
    int r = zo_code_build(&cl, lens, 19);
    if (r < 0 || (r > 0 && cl.nonempty)) return ZO_E_CODE_LENGTHS;
    i = 0;
    while (i < nlen + ndist) {
        int sym;
        if (!cl.nonempty) {
            /* empty code-length code: every lookup hits the {op 64, bits 1, val 0}
             * marker (inftrees.c:112-118), which CODELENS (inflate.c:940-947)
             * reads as "length 0, drop 1 bit" */
            if (!zo_need(b, 1)) return ZO_E_TRUNCATED;
            zo_take(b, 1);
            sym = 0;
        } else {
            sym = zo_code_decode(b, &cl);
        }
        if (sym == -1) return ZO_E_TRUNCATED;
        if (sym == -2) return ZO_E_CODE_LENGTHS;   /* unreachable: cl is complete here */
        if (sym < 16) { lens[i++] = (uint16_t)sym; continue; }
        int rep, val = 0;
        if (sym == 16) {
            if (!zo_need(b, 2)) return ZO_E_TRUNCATED;
            if (i == 0) return ZO_E_BIT_REPEAT;
            val = lens[i - 1];
            rep = 3 + (int)zo_take(b, 2);
        } else if (sym == 17) {
            if (!zo_need(b, 3)) return ZO_E_TRUNCATED;
            rep = 3 + (int)zo_take(b, 3);
        } else {
            if (!zo_need(b, 7)) return ZO_E_TRUNCATED;
            rep = 11 + (int)zo_take(b, 7);
        }
        if (i + rep > nlen + ndist) return ZO_E_BIT_REPEAT;
        while (rep--) lens[i++] = (uint16_t)val;
    }
    if (lens[256] == 0) return ZO_E_NO_EOB;
    /* inftrees.c:131-132: lit/len and distance codes may be incomplete only
     * when they consist of a single 1-bit code (max == 1) */
    r = zo_code_build(lc, lens, nlen);
    if (r < 0 || (r > 0 && !(lc->count[1] == 1 && nlen - lc->count[0] == 1))) return ZO_E_LITLEN_SET;
    r = zo_code_build(dc, lens + nlen, ndist);
    if (r < 0 || (r > 0 && dc->nonempty && !(dc->count[1] == 1 && ndist - dc->count[0] == 1)))
        return ZO_E_DIST_SET;
    return ZO_OK;
}

static int zo_inflate_raw(zo_bits *b, uint8_t *dst, size_t cap, size_t *outp) {
    if (!zo_fmt_ready) zo_fmt_init();
    if (!zo_fix_ready) zo_fixed_init();
    int last;
    do {
        if (!zo_need(b, 3)) return ZO_E_TRUNCATED;
        last = (int)zo_take(b, 1);
        int type = (int)zo_take(b, 2);                         /* inflate.c:827-862 */
        int e;
        if (type == 0) {                                       /* inflate.c:863-897 */
            zo_take(b, b->cnt & 7);
            if (!zo_need(b, 32)) return ZO_E_TRUNCATED;
            uint32_t v = zo_take(b, 32);
            unsigned len = v & 0xffff;
            if (len != ((v >> 16) ^ 0xffff)) return ZO_E_STORED_LEN;
            /* bit accumulator is empty and byte-aligned here */
            if (len > b->n - b->pos) {
                size_t can = b->n - b->pos;
                if (can > cap - *outp) return ZO_E_OUTPUT_FULL;
                return ZO_E_TRUNCATED;
            }
            if (len > cap - *outp) return ZO_E_OUTPUT_FULL;
            memcpy(dst + *outp, b->src + b->pos, len);
            b->pos += len; *outp += len;
            e = ZO_OK;
        } else if (type == 1) {
            e = zo_block_codes(b, dst, cap, outp, &zo_fix_lit, &zo_fix_dist);
        } else if (type == 2) {
            zo_code lc, dc;
            e = zo_dynamic(b, &lc, &dc);
            if (e == ZO_OK) e = zo_block_codes(b, dst, cap, outp, &lc, &dc);
        } else {
            return ZO_E_BLOCK_TYPE;
        }
        if (e) return e;
    } while (!last);
    return ZO_OK;
}

int zo_inflate(const uint8_t *src, size_t srclen, uint8_t *dst, size_t dstcap,
               int wrap, size_t *consumed, size_t *produced) {
    zo_bits b = { src, srclen, 0, 0, 0 };
    size_t out = 0;
    int e = ZO_OK, gz = 0;
    if (consumed) *consumed = 0;
    if (produced) *produced = 0;
    if (wrap) {                                                /* inflate.c:622-669 HEAD */
        if (srclen < 2) return ZO_E_TRUNCATED;
        unsigned h0 = src[0], h1 = src[1];
        if ((wrap & 2) && h0 == 0x1f && h1 == 0x8b) {          /* inflate.c:629 gzip magic */
            gz = 1;
            if (srclen < 10) return ZO_E_TRUNCATED;
            if (src[2] != 8) return ZO_E_METHOD;               /* inflate.c:674 */
            unsigned flg = src[3];
            if (flg & 0xe0) return ZO_E_GZ_FLAGS;              /* inflate.c:679 */
            size_t p = 10;
            if (flg & 4) {                                     /* FEXTRA inflate.c:712-747 */
                if (p + 2 > srclen) return ZO_E_TRUNCATED;
                size_t xl = src[p] | (src[p + 1] << 8);
                p += 2 + xl;
            }
            if (flg & 8)  { do { if (p >= srclen) return ZO_E_TRUNCATED; } while (src[p++]); }   /* FNAME */
            if (flg & 16) { do { if (p >= srclen) return ZO_E_TRUNCATED; } while (src[p++]); }   /* FCOMMENT */
            if (flg & 2) {                                     /* FHCRC inflate.c:791-800 */
                if (p + 2 > srclen) return ZO_E_TRUNCATED;
                uint32_t hc = zo_crc32(0, src, p) & 0xffff;
                if (hc != (uint32_t)(src[p] | (src[p + 1] << 8))) return ZO_E_GZ_HCRC;
                p += 2;
            }
            if (p > srclen) return ZO_E_TRUNCATED;
            b.pos = p;
        } else {
            if (!(wrap & 1)) return ZO_E_HEADER_CHECK;         /* inflate.c:640-647 */
            if (((h0 << 8) + h1) % 31) return ZO_E_HEADER_CHECK;
            if ((h0 & 0xf) != 8) return ZO_E_METHOD;
            if ((h0 >> 4) + 8 > 15) return ZO_E_WINDOW;
            if (h1 & 0x20) return ZO_E_NEED_DICT;
            b.pos = 2;
        }
    }
    e = zo_inflate_raw(&b, dst, dstcap, &out);
    if (produced) *produced = out;
    if (e) { if (consumed) *consumed = b.pos; return e; }
    /* give back whole unused bytes (inflate.c:1185 NEEDBITS after BYTEBITS) */
    b.pos -= (size_t)(b.cnt >> 3); b.cnt = 0; b.acc = 0;
    if (wrap) {                                                /* inflate.c:1183-1219 CHECK/LENGTH */
        size_t need = gz ? 8 : 4;
        if (srclen - b.pos < need) { if (consumed) *consumed = b.pos; return ZO_E_TRUNCATED; }
        const uint8_t *t = src + b.pos;
        if (gz) {
            uint32_t c = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
            uint32_t l = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
            if (c != zo_crc32(0, dst, out)) e = ZO_E_DATA_CHECK;
            else if (l != (uint32_t)out) e = ZO_E_LENGTH_CHECK;
            b.pos += e == ZO_E_DATA_CHECK ? 4 : 8;
        } else {
            uint32_t c = ((uint32_t)t[0] << 24) | (t[1] << 16) | (t[2] << 8) | t[3];
            if (c != zo_adler32(1, dst, out)) e = ZO_E_DATA_CHECK;
            b.pos += 4;
        }
    }
    if (consumed) *consumed = b.pos;
    return e;
}

/* ======================================================================== */
/* Deflate                                                                  */
/* ======================================================================== */
#define ZO_WSIZE     32768u            /* windowBits 15 */
#define ZO_WMASK     (ZO_WSIZE - 1)
#define ZO_HBITS     15                /* memLevel 8: deflate.c:444 */
#define ZO_HMASK     ((1u << ZO_HBITS) - 1)
#define ZO_HSHIFT    5                 /* deflate.c:447 (15+3-1)/3 */
#define ZO_MINM      3
#define ZO_MAXM      258
#define ZO_MINLOOK   (ZO_MAXM + ZO_MINM + 1)        /* deflate.h:293 */
#define ZO_MAXDIST   (ZO_WSIZE - ZO_MINLOOK)        /* deflate.h:298 */
#define ZO_TOOFAR    4096              /* deflate.c:88-90 */
#define ZO_SYMLIM    16383             /* deflate.c:455,512: lit_bufsize - 1 */

typedef struct { uint16_t good, lazy, nice, chain; int slow; } zo_cfg;
static const zo_cfg zo_cfgs[10] = {    /* deflate.c:112-124 */
    {0,0,0,0,0}, {4,4,8,4,0}, {4,5,16,8,0}, {4,6,32,32,0}, {4,4,16,16,1},
    {8,16,32,32,1}, {8,16,128,128,1}, {8,32,128,256,1}, {32,128,258,1024,1},
    {32,258,258,4096,1}
};

typedef struct { uint16_t fc; uint16_t dl; } zo_ct;   /* fc: Freq|Code, dl: Dad|Len  (deflate.h:75-90) */

typedef struct {
    /* bit writer (trees.c:274-286 send_bits; LSB first) */
    uint8_t *out; size_t cap, pos; int ovf;
    uint64_t acc; int cnt;
    /* symbol buffer (deflate.h:354-372) */
    uint16_t sdist[ZO_SYMLIM + 1]; uint8_t slc[ZO_SYMLIM + 1]; unsigned nsym;
    /* trees */
    zo_ct lt[2 * 286 + 1], dt[2 * 30 + 1], bt[2 * 19 + 1];
    int heap[2 * 286 + 1]; int heap_len, heap_max; uint8_t depth[2 * 286 + 1];
    uint16_t bl_count[16];
    uint32_t opt_len, static_len;
    int l_max, d_max;
    int level, strategy;
} zo_def;

static zo_ct zo_sl[288], zo_sd[30];    /* static trees (trees.h:3,64) */
static int zo_static_ready;

static unsigned zo_rev(unsigned code, int len) {     /* trees.c:154 bi_reverse */
    unsigned r = 0;
    do { r |= code & 1; code >>= 1; r <<= 1; } while (--len > 0);
    return r >> 1;
}

/* trees.c:203-232 gen_codes: canonical codes from bl_count, bit-reversed */
static void zo_gen_codes(zo_ct *t, int max_code, const uint16_t *blc) {
    uint16_t next[16]; unsigned code = 0;
    for (int b = 1; b <= 15; b++) { code = (code + blc[b - 1]) << 1; next[b] = (uint16_t)code; }
    for (int n = 0; n <= max_code; n++) {
        int len = t[n].dl;
        if (len) t[n].fc = (uint16_t)zo_rev(next[len]++, len);
    }
}

static void zo_static_init(void) {     /* trees.c:374-390 */
    uint16_t blc[16] = {0};
    int n = 0;
    while (n <= 143) zo_sl[n++].dl = 8, blc[8]++;
    while (n <= 255) zo_sl[n++].dl = 9, blc[9]++;
    while (n <= 279) zo_sl[n++].dl = 7, blc[7]++;
    while (n <= 287) zo_sl[n++].dl = 8, blc[8]++;
    zo_gen_codes(zo_sl, 287, blc);
    for (n = 0; n < 30; n++) { zo_sd[n].dl = 5; zo_sd[n].fc = (uint16_t)zo_rev((unsigned)n, 5); }
    zo_static_ready = 1;
}

static void zo_put(zo_def *s, unsigned v, int len) {
    s->acc |= (uint64_t)v << s->cnt; s->cnt += len;
    while (s->cnt >= 8) {
        if (s->pos < s->cap) s->out[s->pos] = (uint8_t)s->acc; else s->ovf = 1;
        s->pos++; s->acc >>= 8; s->cnt -= 8;
    }
}
static void zo_align(zo_def *s) { if (s->cnt) zo_put(s, 0, 8 - s->cnt); }   /* trees.c:181 bi_windup */

static void zo_init_block(zo_def *s) {                /* trees.c:480 init_block */
    for (int n = 0; n < 286; n++) s->lt[n].fc = 0;
    for (int n = 0; n < 30; n++) s->dt[n].fc = 0;
    for (int n = 0; n < 19; n++) s->bt[n].fc = 0;
    s->lt[256].fc = 1;
    s->opt_len = s->static_len = 0; s->nsym = 0;
}

/* trees.c:499-501 */
#define ZO_SMALLER(t, n, m) ((t)[n].fc < (t)[m].fc || ((t)[n].fc == (t)[m].fc && s->depth[n] <= s->depth[m]))

static void zo_sift(zo_def *s, zo_ct *t, int k) {     /* trees.c:509 pqdownheap */
    int v = s->heap[k];
    for (int j = k << 1; j <= s->heap_len; j <<= 1) {
        if (j < s->heap_len && ZO_SMALLER(t, s->heap[j + 1], s->heap[j])) j++;
        if (ZO_SMALLER(t, v, s->heap[j])) break;
        s->heap[k] = s->heap[j]; k = j;
    }
    s->heap[k] = v;
}

/* trees.c:540-613 gen_bitlen */
static void zo_gen_bitlen(zo_def *s, zo_ct *t, int max_code, const zo_ct *st,
                          const uint8_t *extra, int base, int max_length) {
    int h, overflow = 0;
    for (int b = 0; b <= 15; b++) s->bl_count[b] = 0;
    t[s->heap[s->heap_max]].dl = 0;
    for (h = s->heap_max + 1; h < 2 * 286 + 1; h++) {
        int n = s->heap[h];
        int bits = t[t[n].dl].dl + 1;
        if (bits > max_length) bits = max_length, overflow++;
        t[n].dl = (uint16_t)bits;
        if (n > max_code) continue;
        s->bl_count[bits]++;
        int xb = (n >= base) ? extra[n - base] : 0;
        uint32_t f = t[n].fc;
        s->opt_len += f * (unsigned)(bits + xb);
        if (st) s->static_len += f * (unsigned)(st[n].dl + xb);
    }
    if (!overflow) return;
    do {
        int bits = max_length - 1;
        while (s->bl_count[bits] == 0) bits--;
        s->bl_count[bits]--; s->bl_count[bits + 1] += 2; s->bl_count[max_length]--;
        overflow -= 2;
    } while (overflow > 0);
    for (int bits = max_length; bits != 0; bits--) {
        int n = s->bl_count[bits];
        while (n != 0) {
            int m = s->heap[--h];
            if (m > max_code) continue;
            if (t[m].dl != (unsigned)bits) {
                s->opt_len += ((uint32_t)bits - t[m].dl) * t[m].fc;
                t[m].dl = (uint16_t)bits;
            }
            n--;
        }
    }
}

/* trees.c:627-706 build_tree; returns max_code */
static int zo_build_tree(zo_def *s, zo_ct *t, int elems, const zo_ct *st,
                         const uint8_t *extra, int base, int max_length) {
    int max_code = -1, node;
    s->heap_len = 0; s->heap_max = 2 * 286 + 1;
    for (int n = 0; n < elems; n++) {
        if (t[n].fc) { s->heap[++s->heap_len] = max_code = n; s->depth[n] = 0; }
        else t[n].dl = 0;
    }
    while (s->heap_len < 2) {
        node = s->heap[++s->heap_len] = (max_code < 2 ? ++max_code : 0);
        t[node].fc = 1; s->depth[node] = 0;
        s->opt_len--; if (st) s->static_len -= st[node].dl;
    }
    for (int n = s->heap_len / 2; n >= 1; n--) zo_sift(s, t, n);
    node = elems;
    do {
        int n = s->heap[1];
        s->heap[1] = s->heap[s->heap_len--];
        zo_sift(s, t, 1);
        int m = s->heap[1];
        s->heap[--s->heap_max] = n;
        s->heap[--s->heap_max] = m;
        t[node].fc = (uint16_t)(t[n].fc + t[m].fc);
        s->depth[node] = (uint8_t)((s->depth[n] >= s->depth[m] ? s->depth[n] : s->depth[m]) + 1);
        t[n].dl = t[m].dl = (uint16_t)node;
        s->heap[1] = node++;
        zo_sift(s, t, 1);
    } while (s->heap_len >= 2);
    s->heap[--s->heap_max] = s->heap[1];
    zo_gen_bitlen(s, t, max_code, st, extra, base, max_length);
    zo_gen_codes(t, max_code, s->bl_count);
    return max_code;
}

/* trees.c:712-745 scan_tree / :753-795 send_tree share this run-length walk.
 * emit==0 tallies bl_tree frequencies, emit==1 writes the codes. */
static void zo_walk_tree(zo_def *s, zo_ct *t, int max_code, int emit) {
    int prevlen = -1, nextlen = t[0].dl, count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) max_count = 138, min_count = 3;
    if (!emit) t[max_code + 1].dl = 0xffff;
    for (int n = 0; n <= max_code; n++) {
        int curlen = nextlen; nextlen = t[n + 1].dl;
        if (++count < max_count && curlen == nextlen) continue;
        if (count < min_count) {
            if (emit) { do zo_put(s, s->bt[curlen].fc, s->bt[curlen].dl); while (--count); }
            else s->bt[curlen].fc += (uint16_t)count;
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (emit) { zo_put(s, s->bt[curlen].fc, s->bt[curlen].dl); count--; }
                else s->bt[curlen].fc++;
            }
            if (emit) { zo_put(s, s->bt[16].fc, s->bt[16].dl); zo_put(s, (unsigned)count - 3, 2); }
            else s->bt[16].fc++;
        } else if (count <= 10) {
            if (emit) { zo_put(s, s->bt[17].fc, s->bt[17].dl); zo_put(s, (unsigned)count - 3, 3); }
            else s->bt[17].fc++;
        } else {
            if (emit) { zo_put(s, s->bt[18].fc, s->bt[18].dl); zo_put(s, (unsigned)count - 11, 7); }
            else s->bt[18].fc++;
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) max_count = 138, min_count = 3;
        else if (curlen == nextlen) max_count = 6, min_count = 3;
        else max_count = 7, min_count = 4;
    }
}

/* trees.c:900-951 compress_block */
static void zo_emit_symbols(zo_def *s, const zo_ct *lt, const zo_ct *dt) {
    for (unsigned i = 0; i < s->nsym; i++) {
        unsigned dist = s->sdist[i], lc = s->slc[i];
        if (dist == 0) { zo_put(s, lt[lc].fc, lt[lc].dl); continue; }
        unsigned c = zo_len_code[lc];
        zo_put(s, lt[c + 257].fc, lt[c + 257].dl);
        if (zo_len_extra[c]) zo_put(s, lc + 3 - zo_len_base[c], zo_len_extra[c]);
        c = (unsigned)zo_dist_code(dist);
        zo_put(s, dt[c].fc, dt[c].dl);
        if (zo_dist_extra[c]) zo_put(s, dist - zo_dist_base[c], zo_dist_extra[c]);
    }
    zo_put(s, lt[256].fc, lt[256].dl);
}

/* trees.c:860-875 _tr_stored_block */
static void zo_stored(zo_def *s, const uint8_t *buf, unsigned len, int last) {
    zo_put(s, (unsigned)last, 3);
    zo_align(s);
    zo_put(s, len & 0xffff, 16);
    zo_put(s, ~len & 0xffff, 16);
    for (unsigned i = 0; i < len; i++) zo_put(s, buf[i], 8);
}

/* trees.c:997-1089 _tr_flush_block.  buf==NULL reproduces block_start<0
 * (deflate.c:1597-1600: the window slid past the start of the block). */
static void zo_flush_block(zo_def *s, const uint8_t *buf, uint32_t stored_len, int last) {
    static const uint8_t bl_extra[19] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,3,7};
    s->l_max = zo_build_tree(s, s->lt, 286, zo_sl, zo_len_extra, 257, 15);
    s->d_max = zo_build_tree(s, s->dt, 30, zo_sd, zo_dist_extra, 0, 15);
    zo_walk_tree(s, s->lt, s->l_max, 0);               /* trees.c:800-829 build_bl_tree */
    zo_walk_tree(s, s->dt, s->d_max, 0);
    zo_build_tree(s, s->bt, 19, NULL, bl_extra, 0, 7);
    int mb;
    for (mb = 18; mb >= 3; mb--) if (s->bt[zo_cl_order[mb]].dl) break;
    s->opt_len += 3 * ((uint32_t)mb + 1) + 5 + 5 + 4;
    uint32_t opt_lenb = (s->opt_len + 3 + 7) >> 3, static_lenb = (s->static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb || s->strategy == 4) opt_lenb = static_lenb;
    if (stored_len + 4 <= opt_lenb && buf != NULL) {
        zo_stored(s, buf, stored_len, last);
    } else if (static_lenb == opt_lenb) {
        zo_put(s, 2u + (unsigned)last, 3);
        zo_emit_symbols(s, zo_sl, zo_sd);
    } else {
        zo_put(s, 4u + (unsigned)last, 3);
        zo_put(s, (unsigned)s->l_max + 1 - 257, 5);    /* trees.c:833-855 send_all_trees */
        zo_put(s, (unsigned)s->d_max + 1 - 1, 5);
        zo_put(s, (unsigned)mb + 1 - 4, 4);
        for (int r = 0; r <= mb; r++) zo_put(s, s->bt[zo_cl_order[r]].dl, 3);
        zo_walk_tree(s, s->lt, s->l_max, 1);
        zo_walk_tree(s, s->dt, s->d_max, 1);
        zo_emit_symbols(s, s->lt, s->dt);
    }
    zo_init_block(s);
    if (last) zo_align(s);
}

static int zo_tally_lit(zo_def *s, unsigned c) {
    s->sdist[s->nsym] = 0; s->slc[s->nsym++] = (uint8_t)c; s->lt[c].fc++;
    return s->nsym == ZO_SYMLIM;
}
static int zo_tally_match(zo_def *s, unsigned dist, unsigned len) {
    s->sdist[s->nsym] = (uint16_t)dist; s->slc[s->nsym++] = (uint8_t)(len - 3);
    s->lt[zo_len_code[len - 3] + 257].fc++; s->dt[zo_dist_code(dist)].fc++;
    return s->nsym == ZO_SYMLIM;
}

/* Match-finder state in the reference's own window coordinates: positions are
 * indices into a 64 KB window whose origin `base` advances by 32 KB per slide
 * (deflate.c:251-368 fill_window, :187-209 slide_hash).  The whole chunk is
 * always available, so the window is refilled to 64 KB (or end of input) each
 * time lookahead drops under MIN_LOOKAHEAD. */
typedef struct {
    const uint8_t *in; size_t n;
    size_t base;                       /* absolute offset of window index 0 */
    unsigned strstart, lookahead;      /* window coordinates */
    long block_start;
    uint16_t head[1u << ZO_HBITS], prev[ZO_WSIZE];
    unsigned ins_h;
    unsigned match_start;
} zo_win;

#define ZO_W(w, i) ((w)->in[(w)->base + (i)])

static void zo_fill(zo_win *w) {
    if (w->strstart >= ZO_WSIZE + ZO_MAXDIST) {        /* deflate.c:277-287 */
        w->base += ZO_WSIZE; w->strstart -= ZO_WSIZE; w->match_start -= ZO_WSIZE;
        w->block_start -= (long)ZO_WSIZE;
        for (unsigned i = 0; i < (1u << ZO_HBITS); i++)
            w->head[i] = (uint16_t)(w->head[i] >= ZO_WSIZE ? w->head[i] - ZO_WSIZE : 0);
        for (unsigned i = 0; i < ZO_WSIZE; i++)
            w->prev[i] = (uint16_t)(w->prev[i] >= ZO_WSIZE ? w->prev[i] - ZO_WSIZE : 0);
    }
    size_t end = w->base + 2 * ZO_WSIZE;               /* deflate.c:257,303: fill to window_size */
    if (end > w->n) end = w->n;
    w->lookahead = (unsigned)(end - (w->base + w->strstart));
}

static unsigned zo_insert(zo_win *w, unsigned str) {   /* deflate.c:160-163 INSERT_STRING */
    w->ins_h = ((w->ins_h << ZO_HSHIFT) ^ ZO_W(w, str + 2)) & ZO_HMASK;
    unsigned h = w->prev[str & ZO_WMASK] = w->head[w->ins_h];
    w->head[w->ins_h] = (uint16_t)str;
    return h;
}

/* deflate.c:1356-1497 longest_match.  Bytes past the end of input are treated
 * as never matching, which the reference's lookahead clamps make equivalent
 * (deflate.c:1398-1407,1494-1496). */
static unsigned zo_longest(zo_win *w, unsigned cur, unsigned prev_length,
                           const zo_cfg *c) {
    unsigned chain = c->chain, best = prev_length, nice = c->nice;
    unsigned limit = w->strstart > ZO_MAXDIST ? w->strstart - ZO_MAXDIST : 0;
    unsigned maxlen = w->lookahead < ZO_MAXM ? w->lookahead : ZO_MAXM;
    if (prev_length >= c->good) chain >>= 2;
    if (nice > w->lookahead) nice = w->lookahead;
    const uint8_t *scan = &ZO_W(w, w->strstart);
    do {
        const uint8_t *m = &ZO_W(w, cur);
        unsigned len = 0;
        while (len < maxlen && m[len] == scan[len]) len++;
        if (len > best) {
            w->match_start = cur; best = len;
            if (len >= nice) break;
        }
    } while ((cur = w->prev[cur & ZO_WMASK]) > limit && --chain != 0);
    return best <= w->lookahead ? best : w->lookahead;
}

static void zo_flush(zo_def *s, zo_win *w, int last) { /* deflate.c:1597-1606 FLUSH_BLOCK_ONLY */
    zo_flush_block(s, w->block_start >= 0 ? &ZO_W(w, w->block_start) : NULL,
                   (uint32_t)((long)w->strstart - w->block_start), last);
    w->block_start = (long)w->strstart;
}

size_t zo_deflate_chunk(const uint8_t *in, size_t n, int level, int strategy,
                        int last, uint8_t *out, size_t outcap) {
    if (!zo_fmt_ready) zo_fmt_init();
    if (!zo_static_ready) zo_static_init();
    if (level < 1 || level > 9 || strategy < 0 || strategy > 4) return (size_t)-1;
    zo_def *s = (zo_def *)calloc(1, sizeof *s);
    zo_win *w = (zo_win *)calloc(1, sizeof *w);
    if (!s || !w) { free(s); free(w); return (size_t)-1; }
    const zo_cfg *c = &zo_cfgs[level];
    s->out = out; s->cap = outcap; s->level = level; s->strategy = strategy;
    zo_init_block(s);
    w->in = in; w->n = n;
    zo_fill(w);
    if (w->lookahead >= 2) w->ins_h = ((unsigned)ZO_W(w, 0) << ZO_HSHIFT ^ ZO_W(w, 1)) & ZO_HMASK; /* deflate.c:306-311 */
    int bflush;
    if (strategy == 2) {                               /* deflate.c:2122-2152 deflate_huff */
        for (;;) {
            if (w->lookahead == 0) { zo_fill(w); if (w->lookahead == 0) break; }
            bflush = zo_tally_lit(s, ZO_W(w, w->strstart));
            w->lookahead--; w->strstart++;
            if (bflush) zo_flush(s, w, 0);
        }
    } else if (strategy == 3) {                        /* deflate.c:2051-2115 deflate_rle */
        for (;;) {
            if (w->lookahead <= ZO_MAXM) { zo_fill(w); if (w->lookahead == 0) break; }
            unsigned ml = 0;
            if (w->lookahead >= ZO_MINM && w->base + w->strstart > 0) {
                const uint8_t *p = &ZO_W(w, w->strstart);
                unsigned pv = p[-1];
                unsigned maxlen = w->lookahead < ZO_MAXM ? w->lookahead : ZO_MAXM;
                while (ml < maxlen && p[ml] == pv) ml++;
            }
            if (ml >= ZO_MINM) {
                bflush = zo_tally_match(s, 1, ml);
                w->lookahead -= ml; w->strstart += ml;
            } else {
                bflush = zo_tally_lit(s, ZO_W(w, w->strstart));
                w->lookahead--; w->strstart++;
            }
            if (bflush) zo_flush(s, w, 0);
        }
    } else if (!c->slow) {                             /* deflate.c:1824-1915 deflate_fast */
        for (;;) {
            if (w->lookahead < ZO_MINLOOK) { zo_fill(w); if (w->lookahead == 0) break; }
            unsigned hash_head = 0, ml = 0;
            if (w->lookahead >= ZO_MINM) hash_head = zo_insert(w, w->strstart);
            if (hash_head != 0 && w->strstart - hash_head <= ZO_MAXDIST)
                ml = zo_longest(w, hash_head, ZO_MINM - 1, c);
            if (ml >= ZO_MINM) {
                bflush = zo_tally_match(s, w->strstart - w->match_start, ml);
                w->lookahead -= ml;
                if (ml <= c->lazy && w->lookahead >= ZO_MINM) {   /* lazy == max_insert_length */
                    ml--;
                    do { w->strstart++; zo_insert(w, w->strstart); } while (--ml != 0);
                    w->strstart++;
                } else {
                    w->strstart += ml;
                    w->ins_h = ZO_W(w, w->strstart);
                    w->ins_h = ((w->ins_h << ZO_HSHIFT) ^ (w->lookahead >= 2 ? ZO_W(w, w->strstart + 1) : 0)) & ZO_HMASK;
                }
            } else {
                bflush = zo_tally_lit(s, ZO_W(w, w->strstart));
                w->lookahead--; w->strstart++;
            }
            if (bflush) zo_flush(s, w, 0);
        }
    } else {                                           /* deflate.c:1923-2043 deflate_slow */
        unsigned match_length = ZO_MINM - 1, prev_length, prev_match;
        int match_available = 0;
        for (;;) {
            if (w->lookahead < ZO_MINLOOK) { zo_fill(w); if (w->lookahead == 0) break; }
            unsigned hash_head = 0;
            if (w->lookahead >= ZO_MINM) hash_head = zo_insert(w, w->strstart);
            prev_length = match_length; prev_match = w->match_start;
            match_length = ZO_MINM - 1;
            if (hash_head != 0 && prev_length < c->lazy && w->strstart - hash_head <= ZO_MAXDIST) {
                match_length = zo_longest(w, hash_head, prev_length, c);
                if (match_length <= 5 && (strategy == 1 ||
                    (match_length == ZO_MINM && w->strstart - w->match_start > ZO_TOOFAR)))
                    match_length = ZO_MINM - 1;
            }
            if (prev_length >= ZO_MINM && match_length <= prev_length) {
                unsigned max_insert = w->strstart + w->lookahead - ZO_MINM;
                bflush = zo_tally_match(s, w->strstart - 1 - prev_match, prev_length);
                w->lookahead -= prev_length - 1;
                prev_length -= 2;
                do { if (++w->strstart <= max_insert) zo_insert(w, w->strstart); } while (--prev_length != 0);
                match_available = 0; match_length = ZO_MINM - 1;
                w->strstart++;
                if (bflush) zo_flush(s, w, 0);
            } else if (match_available) {
                bflush = zo_tally_lit(s, ZO_W(w, w->strstart - 1));
                if (bflush) zo_flush(s, w, 0);
                w->strstart++; w->lookahead--;
            } else {
                match_available = 1; w->strstart++; w->lookahead--;
            }
        }
        if (match_available) zo_tally_lit(s, ZO_W(w, w->strstart - 1));
    }
    if (last) zo_flush(s, w, 1);                       /* deflate.c:1908-1910 */
    else {
        if (s->nsym) zo_flush(s, w, 0);                /* deflate.c:1912-1913 */
        zo_stored(s, NULL, 0, 0);                      /* deflate.c:1214-1215: 00 00 FF FF marker */
    }
    size_t r = s->ovf ? (size_t)-1 : s->pos;
    free(s); free(w);
    return r;
}

size_t zo_compress_bound(size_t n) { return n + (n >> 12) + (n >> 14) + (n >> 25) + 13; }

size_t zo_deflate_stream(const uint8_t *in, size_t n, int level, int strategy,
                         int wrap, size_t chunk, uint8_t *out, size_t outcap) {
    size_t pos = 0;
    if (chunk == 0) chunk = n ? n : 1;
    if (wrap == 1) {                                   /* deflate.c:1004-1037 */
        if (outcap < 2) return (size_t)-1;
        unsigned lf = (strategy >= 2 || level < 2) ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3;
        unsigned hdr = (0x78u << 8) | (lf << 6);
        hdr += 31 - hdr % 31;
        out[pos++] = (uint8_t)(hdr >> 8); out[pos++] = (uint8_t)hdr;
    } else if (wrap == 2) {                            /* deflate.c:1042-1054 */
        if (outcap < 10) return (size_t)-1;
        static const uint8_t g[8] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0};
        memcpy(out, g, 8); pos = 8;
        out[pos++] = (uint8_t)(level == 9 ? 2 : (strategy >= 2 || level < 2) ? 4 : 0);
        out[pos++] = 3;                                /* OS_CODE Unix: zutil.h:184 */
    }
    size_t off = 0;
    do {
        size_t k = n - off < chunk ? n - off : chunk;
        int last = (off + k == n);
        size_t r = zo_deflate_chunk(in + off, k, level, strategy, last, out + pos, outcap - pos);
        if (r == (size_t)-1) return r;
        pos += r; off += k;
    } while (off < n);
    if (wrap == 1) {                                   /* deflate.c:1254-1255 */
        if (outcap - pos < 4) return (size_t)-1;
        uint32_t a = zo_adler32(1, in, n);
        out[pos++] = (uint8_t)(a >> 24); out[pos++] = (uint8_t)(a >> 16);
        out[pos++] = (uint8_t)(a >> 8);  out[pos++] = (uint8_t)a;
    } else if (wrap == 2) {                            /* deflate.c:1241-1250 */
        if (outcap - pos < 8) return (size_t)-1;
        uint32_t c = zo_crc32(0, in, n), l = (uint32_t)n;
        for (int i = 0; i < 4; i++) out[pos++] = (uint8_t)(c >> (8 * i));
        for (int i = 0; i < 4; i++) out[pos++] = (uint8_t)(l >> (8 * i));
    }
    return pos;
}
