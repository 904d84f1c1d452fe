// zb_wasm_exports.cu — the C exports of the reference's WASM shim
// (src/wasm_module.c:34-311, src/wasm_module_side.c:17-81 and the delegating
// entry points of src/zlib_simd_optimized.c:354-405), re-hosted on the B200
// engine with the same names, argument checks and return conventions.
// emscripten_get_now() is replaced by clock_gettime; EMSCRIPTEN_KEEPALIVE by
// default visibility.  Everything below is thin glue over zb_zlib_api.cu.
#include "zb_internal.h"
#include "../../include/zb200_zlib.h"
#include <stdlib.h>
#include <time.h>

struct zlib_stream_s {             // wasm_module.c:146-150
    z_stream stream;
    int initialized;
};

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

extern "C" {

int zlib_compress_buffer(const unsigned char *src, unsigned long src_len, unsigned char *dest,
                         unsigned long *dest_len, int level) {
    if (!src || !dest || !dest_len || src_len == 0) return Z_STREAM_ERROR;      // wasm_module.c:37-39
    if (level < 0 || level > 9) level = Z_DEFAULT_COMPRESSION;                  // :41-43
    return compress2(dest, dest_len, src, src_len, level);
}

int zlib_decompress_buffer(const unsigned char *src, unsigned long src_len, unsigned char *dest,
                           unsigned long *dest_len) {
    if (!src || !dest || !dest_len || src_len == 0) return Z_STREAM_ERROR;      // :55-57
    return uncompress(dest, dest_len, src, src_len);
}

unsigned long zlib_crc32(unsigned long crc, const unsigned char *buf, unsigned int len) { return crc32(crc, buf, len); }
unsigned long zlib_adler32(unsigned long adler, const unsigned char *buf, unsigned int len) { return adler32(adler, buf, len); }
unsigned long zlib_compress_bound(unsigned long source_len) { return compressBound(source_len); }
const char *zlib_get_version(void) { return zlibVersion(); }
int zlib_has_simd(void) { return 0; }                                           // :96-98 (no WASM SIMD here either)

double zlib_benchmark_crc32(const char *data, int size, int iterations) {       // :101-113, ops/sec
    if (!data || size <= 0 || iterations <= 0) return -1.0;
    const double t0 = now_ms();
    for (int i = 0; i < iterations; i++) {
        volatile uLong r = crc32(0L, (const Bytef *)data, (uInt)size);
        (void)r;
    }
    return iterations / ((now_ms() - t0) / 1000.0);
}

double zlib_benchmark_compression(const char *data, int size, int iterations, int level) {   // :116-134
    if (!data || size <= 0 || iterations <= 0) return -1.0;
    uLongf dest_len = compressBound((uLong)size);
    char *dest = (char *)malloc(dest_len);
    if (!dest) return -1.0;
    const double t0 = now_ms();
    for (int i = 0; i < iterations; i++) {
        uLongf cur = dest_len;
        int r = compress2((Bytef *)dest, &cur, (const Bytef *)data, (uLong)size, level);
        (void)r;
    }
    const double dt = (now_ms() - t0) / 1000.0;
    free(dest);
    return iterations / dt;
}

void zlib_get_performance_info(int *has_simd, int *crc32_threshold, int *compression_threshold) {   // :137-141
    if (has_simd) *has_simd = 0;
    if (crc32_threshold) *crc32_threshold = 1024;
    if (compression_threshold) *compression_threshold = 4096;
}

zlib_stream_t *zlib_deflate_init(int level, int window_bits, int mem_level, int strategy) {   // :153-176
    zlib_stream_t *ctx = (zlib_stream_t *)calloc(1, sizeof(zlib_stream_t));
    if (!ctx) return NULL;
    if (level < 0 || level > 9) level = Z_DEFAULT_COMPRESSION;
    if (window_bits < 8 || window_bits > 15) window_bits = 15;
    if (mem_level < 1 || mem_level > 9) mem_level = 8;
    if (deflateInit2(&ctx->stream, level, Z_DEFLATED, window_bits, mem_level, strategy) != Z_OK) {
        free(ctx);
        return NULL;
    }
    ctx->initialized = 1;
    return ctx;
}

int zlib_deflate_process(zlib_stream_t *ctx, const unsigned char *in, unsigned int in_len,
                         unsigned char *out, unsigned int out_len, int flush) {               // :179-193
    if (!ctx || !ctx->initialized) return Z_STREAM_ERROR;
    ctx->stream.next_in = (const Bytef *)in;
    ctx->stream.avail_in = in_len;
    ctx->stream.next_out = out;
    ctx->stream.avail_out = out_len;
    return deflate(&ctx->stream, flush);
}

void zlib_deflate_end(zlib_stream_t *ctx) {                                                  // :196-203
    if (ctx) {
        if (ctx->initialized) deflateEnd(&ctx->stream);
        free(ctx);
    }
}

zlib_stream_t *zlib_inflate_init(int window_bits) {                                           // :209-229
    zlib_stream_t *ctx = (zlib_stream_t *)calloc(1, sizeof(zlib_stream_t));
    if (!ctx) return NULL;
    if (window_bits < 8 || window_bits > 15) window_bits = 15;
    if (inflateInit2(&ctx->stream, window_bits) != Z_OK) {
        free(ctx);
        return NULL;
    }
    ctx->initialized = 1;
    return ctx;
}

int zlib_inflate_process(zlib_stream_t *ctx, const unsigned char *in, unsigned int in_len,
                         unsigned char *out, unsigned int out_len) {                          // :232-246
    if (!ctx || !ctx->initialized) return Z_STREAM_ERROR;
    ctx->stream.next_in = (const Bytef *)in;
    ctx->stream.avail_in = in_len;
    ctx->stream.next_out = out;
    ctx->stream.avail_out = out_len;
    return inflate(&ctx->stream, Z_NO_FLUSH);
}

void zlib_inflate_end(zlib_stream_t *ctx) {                                                  // :249-256
    if (ctx) {
        if (ctx->initialized) inflateEnd(&ctx->stream);
        free(ctx);
    }
}

unsigned int zlib_stream_avail_in(zlib_stream_t *ctx) { return ctx ? ctx->stream.avail_in : 0; }      // :262-288
unsigned int zlib_stream_avail_out(zlib_stream_t *ctx) { return ctx ? ctx->stream.avail_out : 0; }
unsigned long zlib_stream_total_in(zlib_stream_t *ctx) { return ctx ? ctx->stream.total_in : 0; }
unsigned long zlib_stream_total_out(zlib_stream_t *ctx) { return ctx ? ctx->stream.total_out : 0; }

// src/zlib_simd_optimized.c:354-383: raw deflate (windowBits -15), one Z_FINISH call, Z_OK on success.
int zlib_compress_simd_full(const unsigned char *in, size_t n, unsigned char *out, size_t *out_len, int level) {
    if (!in || !out || !out_len) return Z_STREAM_ERROR;
    z_stream s;
    memset(&s, 0, sizeof s);
    if (level < 0 || level > 9) level = Z_DEFAULT_COMPRESSION;
    int r = deflateInit2(&s, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    if (r != Z_OK) return r;
    s.next_in = in; s.avail_in = (uInt)n; s.next_out = out; s.avail_out = (uInt)*out_len;
    r = deflate(&s, Z_FINISH);
    *out_len = s.total_out;
    deflateEnd(&s);
    return r == Z_STREAM_END ? Z_OK : (r == Z_OK ? Z_BUF_ERROR : r);
}
int zlib_compress_simd(const unsigned char *in, size_t n, unsigned char *out, size_t *out_len, int level) {
    return zlib_compress_simd_full(in, n, out, out_len, level);                 // zlib_simd_compression.c:280-284
}
unsigned int zlib_crc32_simd_enhanced(unsigned int crc, const unsigned char *data, size_t len) {   // zlib_simd_optimized.c:387-405
    return (unsigned int)crc32_z(crc, data, len);
}
unsigned int zlib_crc32_simd_optimized(unsigned int crc, const unsigned char *data, size_t len) {
    return (unsigned int)crc32_z(crc, data, len);
}
int zlib_simd_capabilities(void) { return 1; }
int zlib_simd_capabilities_enhanced(void) { return 0; }                          // zlib_simd_optimized.c:408-415 (no __wasm_simd128__ here)
unsigned int zlib_adler32_simd(unsigned int adler, const unsigned char *data, size_t len) {   // zlib_simd_optimized.c: delegates to adler32
    return (unsigned int)adler32_z(adler, data, len);
}

// src/wasm_module_side.c:61-70: the reference sends buffers of >= 8192 bytes down the "SIMD" path — RAW deflate —
// only when compiled with __wasm_simd128__; this host build, like the reference's fallback build, always takes compress2.
int zlib_compress_simd_buffer(const unsigned char *src, unsigned long src_len, unsigned char *dest,
                              unsigned long *dest_len, int level) {
    return compress2(dest, dest_len, src, src_len, level);
}
unsigned long zlib_crc32_simd(unsigned long crc, const unsigned char *buf, unsigned int len) {   // wasm_module_side.c:74-81, wasm_module.c:296
    return crc32(crc, buf, len);
}

// src/zlib_simd_compression.c:348-377: MB/s (MiB) of `iterations` raw-deflate one-shots at the default level.
double zlib_benchmark_simd_compression(const unsigned char *data, size_t len, int iterations) {
    if (!data || len == 0 || iterations <= 0) return -1.0;
    const size_t cap = compressBound((uLong)len) + 64;     // (the reference's len + len/10 + 64 is below this engine's bound)
    unsigned char *out = (unsigned char *)malloc(cap);
    if (!out) return -1.0;
    const double t0 = now_ms();
    for (int i = 0; i < iterations; i++) {
        size_t olen = cap;
        if (zlib_compress_simd(data, len, out, &olen, Z_DEFAULT_COMPRESSION) != Z_OK) { free(out); return -1.0; }
    }
    const double dt = (now_ms() - t0) / 1000.0;
    free(out);
    return ((double)len * iterations / dt) / (1024.0 * 1024.0);
}

// src/zlib_simd_compression.c:387-418: ratio of compress2, time of compress2 / time of zlib_compress_simd.
void zlib_simd_analysis(const unsigned char *input, size_t input_len, double *compression_ratio, double *simd_speedup,
                        double *memory_efficiency) {
    if (!input || input_len == 0) return;
    uLongf a_len = compressBound((uLong)input_len);
    size_t b_len = a_len;
    unsigned char *a = (unsigned char *)malloc(a_len), *b = (unsigned char *)malloc(b_len);
    if (!a || !b) { free(a); free(b); return; }
    double t = now_ms();
    compress2(a, &a_len, input, (uLong)input_len, Z_DEFAULT_COMPRESSION);
    const double t_scalar = now_ms() - t;
    t = now_ms();
    zlib_compress_simd(input, input_len, b, &b_len, Z_DEFAULT_COMPRESSION);
    const double t_simd = now_ms() - t;
    if (compression_ratio) *compression_ratio = (double)input_len / (double)a_len;
    if (simd_speedup) *simd_speedup = t_simd > 0 ? t_scalar / t_simd : 1.0;
    if (memory_efficiency) *memory_efficiency = 1.0;
    free(a); free(b);
}

// src/zlib_simd_optimized.c:420-470: the three "speed-ups" (plain call time / *_simd call time; both run on the GPU here).
void zlib_simd_performance_analysis(const unsigned char *input, size_t input_len, double *compression_speedup,
                                    double *crc32_speedup, double *adler32_speedup) {
    if (!input || input_len == 0) return;
    double t = now_ms();
    volatile uLong c1 = crc32_z(0, input, input_len);
    const double t_crc = now_ms() - t;
    t = now_ms();
    volatile unsigned c2 = zlib_crc32_simd_enhanced(0, input, input_len);
    const double t_crc_simd = now_ms() - t;
    t = now_ms();
    volatile uLong a1 = adler32_z(1, input, input_len);
    const double t_ad = now_ms() - t;
    t = now_ms();
    volatile unsigned a2 = zlib_adler32_simd(1, input, input_len);
    const double t_ad_simd = now_ms() - t;
    (void)c1; (void)c2; (void)a1; (void)a2;
    if (crc32_speedup) *crc32_speedup = t_crc_simd > 0 ? t_crc / t_crc_simd : 1.0;
    if (adler32_speedup) *adler32_speedup = t_ad_simd > 0 ? t_ad / t_ad_simd : 1.0;
    uLongf cap = compressBound((uLong)input_len);
    unsigned char *out = (unsigned char *)malloc(cap);
    if (!out) { if (compression_speedup) *compression_speedup = 1.0; return; }
    uLongf l1 = cap;
    t = now_ms();
    compress2(out, &l1, input, (uLong)input_len, Z_DEFAULT_COMPRESSION);
    const double t_c = now_ms() - t;
    size_t l2 = cap;
    t = now_ms();
    zlib_compress_simd_full(input, input_len, out, &l2, Z_DEFAULT_COMPRESSION);
    const double t_cs = now_ms() - t;
    if (compression_speedup) *compression_speedup = t_cs > 0 ? t_c / t_cs : 1.0;
    free(out);
}

}  // extern "C"
