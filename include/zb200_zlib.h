/* zb200_zlib.h — the zlib.h-compatible surface of libzb200.so.
 *
 * libzb200.so exports, with the reference's names, signatures and z_stream
 * layout (zlib.h:90-110 of /root/reference, zlib 1.3.1.1-motley), the part of
 * the zlib API that sits on the GPU hot path, plus the C exports of the
 * reference's src/wasm_module.c.  A program built against the reference's own
 * zlib.h links against libzb200.so unchanged; this header only exists so that
 * the drop-in surface is declared somewhere in this repo, and it backs off if
 * the real zlib.h has already been included.
 *
 * Behavioural notes (details in DESIGN.md / INTEGRATION.md):
 *  - every call runs on the GPU; there is no CPU path.  Init functions return
 *    Z_STREAM_ERROR with strm->msg set when no device is usable; crc32/adler32
 *    (which have no error channel) print a message and abort().
 *  - deflate() buffers input and compresses at flush points in independent
 *    chunks (ZB200_CHUNK bytes, default 262144), i.e. the stream the reference
 *    emits for deflate(Z_FULL_FLUSH) every chunk.  Levels 4..9 are byte-identical
 *    to that reference stream; any zlib-compatible inflate decodes all levels.
 *  - inflate() accepts input in arbitrary slices and resumes at deflate block
 *    boundaries on the device.
 */
#ifndef ZB200_ZLIB_H
#define ZB200_ZLIB_H
#include <stddef.h>
#include <stdarg.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#ifndef ZLIB_H   /* the reference's zlib.h was not included: declare the ABI types ourselves */

#define ZLIB_VERSION "1.3.1.1-motley"                    /* zlib.h:44 */

typedef unsigned char Bytef;                              /* zconf.h:396-436 */
typedef unsigned int uInt;
typedef unsigned long uLong;
typedef uLong uLongf;
typedef void *voidpf;
typedef size_t z_size_t;
typedef unsigned int z_crc_t;            /* zconf.h:396-436 */
typedef long z_off_t;

typedef voidpf (*alloc_func)(voidpf opaque, uInt items, uInt size);
typedef void (*free_func)(voidpf opaque, voidpf address);
struct internal_state;

typedef struct z_stream_s {                               /* zlib.h:90-110 */
    const Bytef *next_in;  uInt avail_in;  uLong total_in;
    Bytef *next_out;       uInt avail_out; uLong total_out;
    const char *msg;
    struct internal_state *state;
    alloc_func zalloc; free_func zfree; voidpf opaque;
    int data_type;
    uLong adler;
    uLong reserved;
} z_stream;
typedef z_stream *z_streamp;
typedef void *voidp;
typedef const void *voidpc;
struct gzFile_s { unsigned have; unsigned char *next; long pos; };   /* zlib.h:1819-1823 */
typedef struct gz_header_s {                              /* zlib.h:116-131 */
    int text; uLong time; int xflags; int os;
    Bytef *extra; uInt extra_len; uInt extra_max;
    Bytef *name; uInt name_max;
    Bytef *comment; uInt comm_max;
    int hcrc; int done;
} gz_header;
typedef gz_header *gz_headerp;
typedef struct gzFile_s *gzFile;

#define Z_NO_FLUSH 0                                      /* zlib.h:172-189 */
#define Z_PARTIAL_FLUSH 1
#define Z_SYNC_FLUSH 2
#define Z_FULL_FLUSH 3
#define Z_FINISH 4
#define Z_BLOCK 5
#define Z_TREES 6
#define Z_OK 0
#define Z_STREAM_END 1
#define Z_NEED_DICT 2
#define Z_ERRNO (-1)
#define Z_STREAM_ERROR (-2)
#define Z_DATA_ERROR (-3)
#define Z_MEM_ERROR (-4)
#define Z_BUF_ERROR (-5)
#define Z_VERSION_ERROR (-6)
#define Z_NO_COMPRESSION 0                                /* zlib.h:191-200 */
#define Z_BEST_SPEED 1
#define Z_BEST_COMPRESSION 9
#define Z_DEFAULT_COMPRESSION (-1)
#define Z_FILTERED 1
#define Z_HUFFMAN_ONLY 2
#define Z_RLE 3
#define Z_FIXED 4
#define Z_DEFAULT_STRATEGY 0
#define Z_BINARY 0
#define Z_TEXT 1
#define Z_UNKNOWN 2
#define Z_DEFLATED 8
#define Z_NULL 0

#define deflateInit(strm, level) deflateInit_((strm), (level), ZLIB_VERSION, (int)sizeof(z_stream))      /* zlib.h:1832-1841 */
#define inflateInit(strm) inflateInit_((strm), ZLIB_VERSION, (int)sizeof(z_stream))
#define deflateInit2(strm, level, method, windowBits, memLevel, strategy) \
    deflateInit2_((strm), (level), (method), (windowBits), (memLevel), (strategy), ZLIB_VERSION, (int)sizeof(z_stream))
typedef unsigned (*in_func)(void *, const unsigned char **);                            /* zlib.h:1116-1118 */
typedef int (*out_func)(void *, unsigned char *, unsigned);
#define inflateBackInit(strm, windowBits, window) inflateBackInit_((strm), (windowBits), (window), ZLIB_VERSION, (int)sizeof(z_stream))
#define inflateInit2(strm, windowBits) inflateInit2_((strm), (windowBits), ZLIB_VERSION, (int)sizeof(z_stream))

#endif /* ZLIB_H */

/* ---- basic + advanced stream API (zlib.h:220-616,1832-1841) ---- */
const char *zlibVersion(void);
int deflateInit_(z_streamp strm, int level, const char *version, int stream_size);
int deflateInit2_(z_streamp strm, int level, int method, int windowBits, int memLevel, int strategy,
                  const char *version, int stream_size);
int deflate(z_streamp strm, int flush);
int deflateEnd(z_streamp strm);
int deflateReset(z_streamp strm);
int deflateParams(z_streamp strm, int level, int strategy);
int deflateSetDictionary(z_streamp strm, const Bytef *dictionary, uInt dictLength);   /* zlib.h:612, deflate.c:550 */
uLong deflateBound(z_streamp strm, uLong sourceLen);
int inflateInit_(z_streamp strm, const char *version, int stream_size);
int inflateInit2_(z_streamp strm, int windowBits, const char *version, int stream_size);
int inflate(z_streamp strm, int flush);
int inflateEnd(z_streamp strm);
int inflateReset(z_streamp strm);
int inflateReset2(z_streamp strm, int windowBits);
int inflateSetDictionary(z_streamp strm, const Bytef *dictionary, uInt dictLength);   /* zlib.h:887, inflate.c:1278 */
int deflateGetDictionary(z_streamp strm, Bytef *dictionary, uInt *dictLength);        /* zlib.h:655, deflate.c:638 */
int inflateGetDictionary(z_streamp strm, Bytef *dictionary, uInt *dictLength);        /* zlib.h:910, inflate.c:1258 */
int deflateSetHeader(z_streamp strm, gz_headerp head);                                 /* zlib.h:802, deflate.c:692 */
int inflateGetHeader(z_streamp strm, gz_headerp head);                                 /* zlib.h:1040, inflate.c:1331 */
int deflateTune(z_streamp strm, int good_length, int max_lazy, int nice_length, int max_chain);   /* zlib.h:744, deflate.c:805 */
int deflatePending(z_streamp strm, unsigned *pending, int *bits);                      /* zlib.h:779, deflate.c:703 */
int deflateCopy(z_streamp dest, z_streamp source);                                     /* zlib.h:677, deflate.c:1297 */
int inflateCopy(z_streamp dest, z_streamp source);                                     /* zlib.h:929, inflate.c:1433 */
int deflateResetKeep(z_streamp strm);                                                  /* deflate.c:644 */
int inflateResetKeep(z_streamp strm);                                                  /* inflate.c:102 */
int inflateValidate(z_streamp strm, int check);                                        /* zlib.h:1807, inflate.c:1495 */
int inflateUndermine(z_streamp strm, int subvert);                                     /* inflate.c:1478 */
long inflateMark(z_streamp strm);
int inflateSync(z_streamp strm);                                                       /* zlib.h:914, inflate.c:1375 */
int inflateSyncPoint(z_streamp strm);                                                  /* inflate.c:1431 */
int inflatePrime(z_streamp strm, int bits, int value);                                 /* zlib.h:996, inflate.c:223 */
unsigned long inflateCodesUsed(z_streamp strm);                                        /* inflate.c:1521 */
int deflatePrime(z_streamp strm, int bits, int value);                                 /* zlib.h:785, deflate.c:731 */
int deflateUsed(z_streamp strm, int *bits);                                            /* deflate.c:723 */                                                      /* zlib.h:1012, inflate.c:1510 */
int inflateBackInit_(z_streamp strm, int windowBits, unsigned char *window, const char *version, int stream_size);   /* infback.c:25 */
int inflateBack(z_streamp strm, in_func in, void *in_desc, out_func out, void *out_desc);                           /* infback.c:242 */
int inflateBackEnd(z_streamp strm);                                                                                  /* infback.c:622 */
uLong zlibCompileFlags(void);                                                          /* zlib.h:1227, zutil.c:32 */

/* ---- gz* file layer (zlib.h:1300-1823; gzlib.c, gzread.c, gzwrite.c, gzclose.c) ---- */
gzFile gzopen(const char *path, const char *mode);
gzFile gzdopen(int fd, const char *mode);
int gzbuffer(gzFile file, unsigned size);
int gzsetparams(gzFile file, int level, int strategy);
int gzread(gzFile file, voidp buf, unsigned len);
z_size_t gzfread(voidp buf, z_size_t size, z_size_t nitems, gzFile file);
int gzwrite(gzFile file, voidpc buf, unsigned len);
z_size_t gzfwrite(voidpc buf, z_size_t size, z_size_t nitems, gzFile file);
int gzprintf(gzFile file, const char *format, ...);
int gzvprintf(gzFile file, const char *format, va_list va);  /* zlib.h:1496, gzwrite.c */
gzFile gzopen64(const char *path, const char *mode);         /* zlib.h:1893-1912, gzlib.c:268,342: the LFS names */
long gzseek64(gzFile file, long offset, int whence);
long gztell64(gzFile file);
long gzoffset64(gzFile file);
int gzputs(gzFile file, const char *s);
char *gzgets(gzFile file, char *buf, int len);
int gzputc(gzFile file, int c);
int gzgetc(gzFile file);
int gzgetc_(gzFile file);                                /* zlib.h:1825: what the gzgetc() macro falls back to */
int gzungetc(int c, gzFile file);
int gzflush(gzFile file, int flush);
z_off_t gzseek(gzFile file, z_off_t offset, int whence);
int gzrewind(gzFile file);
z_off_t gztell(gzFile file);
z_off_t gzoffset(gzFile file);
int gzeof(gzFile file);
int gzdirect(gzFile file);
int gzclose(gzFile file);
int gzclose_r(gzFile file);
int gzclose_w(gzFile file);
const char *gzerror(gzFile file, int *errnum);
void gzclearerr(gzFile file);

/* ---- utility (compress.c:22-75, uncompr.c:27-85) ---- */
int compress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen);
int compress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen, int level);
uLong compressBound(uLong sourceLen);
int uncompress(Bytef *dest, uLongf *destLen, const Bytef *source, uLong sourceLen);
int uncompress2(Bytef *dest, uLongf *destLen, const Bytef *source, uLong *sourceLen);

/* ---- checksums (crc32.c:694-1049, adler32.c:61-164) ---- */
uLong crc32(uLong crc, const Bytef *buf, uInt len);
uLong crc32_z(uLong crc, const Bytef *buf, z_size_t len);
uLong crc32_combine(uLong crc1, uLong crc2, z_off_t len2);
uLong crc32_combine_gen(z_off_t len2);
uLong crc32_combine_op(uLong crc1, uLong crc2, uLong op);
uLong adler32(uLong adler, const Bytef *buf, uInt len);
uLong adler32_z(uLong adler, const Bytef *buf, z_size_t len);
uLong adler32_combine(uLong adler1, uLong adler2, z_off_t len2);
uLong crc32_combine64(uLong crc1, uLong crc2, long len2);                /* crc32.c:1021; zlib.h:1893-1912 LFS names */
uLong crc32_combine_gen64(long len2);                                    /* crc32.c:1034 */
uLong adler32_combine64(uLong adler1, uLong adler2, long len2);          /* adler32.c:162 */
const char *zError(int err);
const z_crc_t *get_crc_table(void);                                      /* zlib.h:1935, crc32.c:549 */

/* ---- the reference's WASM C exports (src/wasm_module.c:34-311) ---- */
typedef struct zlib_stream_s zlib_stream_t;                                  /* wasm_module.c:146-150 */
int zlib_compress_buffer(const unsigned char *src, unsigned long src_len, unsigned char *dest,
                         unsigned long *dest_len, int level);                /* :35 */
int zlib_decompress_buffer(const unsigned char *src, unsigned long src_len, unsigned char *dest,
                           unsigned long *dest_len);                         /* :53 */
unsigned long zlib_crc32(unsigned long crc, const unsigned char *buf, unsigned int len);     /* :66 */
unsigned long zlib_adler32(unsigned long adler, const unsigned char *buf, unsigned int len); /* :74 */
unsigned long zlib_compress_bound(unsigned long source_len);                 /* :82 */
const char *zlib_get_version(void);                                          /* :90 */
int zlib_has_simd(void);                                                     /* :96 */
double zlib_benchmark_crc32(const char *data, int size, int iterations);     /* :101 ops/sec */
double zlib_benchmark_compression(const char *data, int size, int iterations, int level);   /* :116 */
void zlib_get_performance_info(int *has_simd, int *crc32_threshold, int *compression_threshold);  /* :137 */
zlib_stream_t *zlib_deflate_init(int level, int window_bits, int mem_level, int strategy);  /* :153 */
int zlib_deflate_process(zlib_stream_t *ctx, const unsigned char *in, unsigned int in_len,
                         unsigned char *out, unsigned int out_len, int flush);               /* :179 */
void zlib_deflate_end(zlib_stream_t *ctx);                                   /* :196 */
zlib_stream_t *zlib_inflate_init(int window_bits);                           /* :209 */
int zlib_inflate_process(zlib_stream_t *ctx, const unsigned char *in, unsigned int in_len,
                         unsigned char *out, unsigned int out_len);          /* :232 */
void zlib_inflate_end(zlib_stream_t *ctx);                                   /* :249 */
unsigned int zlib_stream_avail_in(zlib_stream_t *ctx);                       /* :262 */
unsigned int zlib_stream_avail_out(zlib_stream_t *ctx);                      /* :269 */
unsigned long zlib_stream_total_in(zlib_stream_t *ctx);                      /* :276 */
unsigned long zlib_stream_total_out(zlib_stream_t *ctx);                     /* :283 */
/* src/zlib_simd_optimized.c:354-405, src/zlib_simd_compression.c:279-284: raw-deflate one-shots */
int zlib_compress_simd(const unsigned char *in, size_t n, unsigned char *out, size_t *out_len, int level);
int zlib_compress_simd_full(const unsigned char *in, size_t n, unsigned char *out, size_t *out_len, int level);
unsigned int zlib_crc32_simd_optimized(unsigned int crc, const unsigned char *data, size_t len);
unsigned int zlib_crc32_simd_enhanced(unsigned int crc, const unsigned char *data, size_t len);
int zlib_simd_capabilities(void);
int zlib_simd_capabilities_enhanced(void);                                   /* zlib_simd_optimized.c:408 */
unsigned int zlib_adler32_simd(unsigned int adler, const unsigned char *data, size_t len);
/* src/wasm_module_side.c:61-81, src/zlib_simd_compression.c:348,387, src/zlib_simd_optimized.c:420 */
int zlib_compress_simd_buffer(const unsigned char *src, unsigned long src_len, unsigned char *dest,
                              unsigned long *dest_len, int level);
unsigned long zlib_crc32_simd(unsigned long crc, const unsigned char *buf, unsigned int len);
double zlib_benchmark_simd_compression(const unsigned char *data, size_t len, int iterations);   /* MiB/s */
void zlib_simd_analysis(const unsigned char *input, size_t input_len, double *compression_ratio, double *simd_speedup,
                        double *memory_efficiency);
void zlib_simd_performance_analysis(const unsigned char *input, size_t input_len, double *compression_speedup,
                                    double *crc32_speedup, double *adler32_speedup);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ZB200_ZLIB_H */
