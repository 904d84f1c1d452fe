/* zoracle.h — CPU ORACLE for the zlib hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithms (zlib 1.3.1.1-motley as
 * shipped by discere-os/zlib.wasm) for the path BASELINE.json names.  Nothing
 * under oracle/ is linked into, loaded by, or called from the product library
 * (zlib_wasm_b200/libzb200.so).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it.
 *
 * Parity pin: every function here is checked (tests/test_oracle.py) against the
 * UNMODIFIED reference compiled from /root/reference into oracle/_ref/libzref.so
 * (byte-identical deflate streams at levels 1..9 x strategies, identical
 * checksums/combines, identical inflate output and error class) and against
 * the only known-answer vectors the reference tree holds for this path:
 * contrib/puff/zeros.raw and the malformed-stream vectors of
 * contrib/puff/Makefile:19-38 (committed under tests/golden/).
 *
 * Citations are file:line relative to /root/reference.
 */
#ifndef ZORACLE_H
#define ZORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- checksums ---------------------------------------------------------- */
uint32_t zo_crc32(uint32_t crc, const uint8_t *buf, size_t len);          /* crc32.c:694 crc32_z   */
uint32_t zo_adler32(uint32_t adler, const uint8_t *buf, size_t len);      /* adler32.c:61 adler32_z */
uint32_t zo_multmodp(uint32_t a, uint32_t b);                             /* crc32.c:155 */
uint32_t zo_x2nmodp(uint64_t n, unsigned k);                              /* crc32.c:176 */
uint32_t zo_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);   /* crc32.c:1021 */
uint32_t zo_crc32_combine_gen(uint64_t len2);                             /* crc32.c:1034 */
uint32_t zo_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op);  /* crc32.c:1047 */
uint32_t zo_adler32_combine(uint32_t a1, uint32_t a2, int64_t len2);      /* adler32.c:133 */

/* ---- inflate ------------------------------------------------------------ */
/* error classes: index into zo_inflate_msg(); 0 = ok.  Strings are the
 * reference's strm->msg literals (inflate.c:645-1212, inffast.c:158-283). */
enum {
    ZO_OK = 0,
    ZO_E_HEADER_CHECK,      /* "incorrect header check"               inflate.c:645 */
    ZO_E_METHOD,            /* "unknown compression method"           inflate.c:650 */
    ZO_E_WINDOW,            /* "invalid window size"                  inflate.c:660 */
    ZO_E_GZ_FLAGS,          /* "unknown header flags set"             inflate.c:683 */
    ZO_E_GZ_HCRC,           /* "header crc mismatch"                  inflate.c:796 */
    ZO_E_BLOCK_TYPE,        /* "invalid block type"                   inflate.c:857 */
    ZO_E_STORED_LEN,        /* "invalid stored block lengths"         inflate.c:867 */
    ZO_E_TOO_MANY_SYMS,     /* "too many length or distance symbols"  inflate.c:907 */
    ZO_E_CODE_LENGTHS,      /* "invalid code lengths set"             inflate.c:930 */
    ZO_E_BIT_REPEAT,        /* "invalid bit length repeat"            inflate.c:953,976 */
    ZO_E_NO_EOB,            /* "invalid code -- missing end-of-block" inflate.c:992 */
    ZO_E_LITLEN_SET,        /* "invalid literal/lengths set"          inflate.c:1005 */
    ZO_E_DIST_SET,          /* "invalid distances set"                inflate.c:1014 */
    ZO_E_LITLEN_CODE,       /* "invalid literal/length code"          inflate.c:1077, inffast.c:283 */
    ZO_E_DIST_CODE,         /* "invalid distance code"                inflate.c:1119, inffast.c:268 */
    ZO_E_DIST_FAR,          /* "invalid distance too far back"        inflate.c:1139, inffast.c:158 */
    ZO_E_DATA_CHECK,        /* "incorrect data check"                 inflate.c:1197 */
    ZO_E_LENGTH_CHECK,      /* "incorrect length check"               inflate.c:1212 */
    ZO_E_TRUNCATED,         /* input exhausted before end of stream  -> Z_BUF_ERROR */
    ZO_E_OUTPUT_FULL,       /* output space exhausted                -> Z_BUF_ERROR */
    ZO_E_NEED_DICT,         /* FDICT set                              inflate.c:664 */
    ZO_E_COUNT
};
const char *zo_inflate_msg(int err);

/* wrap: 0 raw deflate, 1 zlib (RFC1950), 2 gzip (RFC1952), 3 auto zlib/gzip. */
int zo_inflate(const uint8_t *src, size_t srclen, uint8_t *dst, size_t dstcap,
               int wrap, size_t *consumed, size_t *produced);

/* ---- deflate ------------------------------------------------------------ */
/* One Z_FULL_FLUSH-bounded run: the block sequence the reference emits for
 * deflate(strm, last ? Z_FINISH : Z_FULL_FLUSH) over `n` fresh bytes with an
 * empty history (deflate.c:954, 1190-1233), without zlib/gzip framing.
 * memLevel 8, windowBits 15.  level 1..9; strategy 0..4 (zlib.h:196-200).
 * Returns bytes written or (size_t)-1 if outcap is too small. */
size_t zo_deflate_chunk(const uint8_t *in, size_t n, int level, int strategy,
                        int last, uint8_t *out, size_t outcap);

/* Whole stream = header + chunks of `chunk` bytes (Z_FULL_FLUSH between,
 * Z_FINISH on the last) + trailer.  wrap: 0 raw, 1 zlib, 2 gzip. */
size_t zo_deflate_stream(const uint8_t *in, size_t n, int level, int strategy,
                         int wrap, size_t chunk, uint8_t *out, size_t outcap);

size_t zo_compress_bound(size_t n);                                        /* compress.c:72 */

#ifdef __cplusplus
}
#endif
#endif
