/* zb200.h — C ABI of the B200 engine behind the zlib drop-in.
 *
 * Plain pointers and sizes only; no CUDA or torch types in any signature
 * (`stream` is a cudaStream_t passed as void*, NULL = the context's own
 * stream).  Every entry point names the reference interface it stands in for
 * (file:line under /root/reference, zlib 1.3.1.1-motley as shipped by
 * discere-os/zlib.wasm).  The zlib.h-compatible surface (deflateInit2_,
 * deflate, inflate, compress2, uncompress, crc32, adler32 ... and the
 * src/wasm_module.c `zlib_*` exports) lives in the same shared object and is
 * declared in include/zb200_zlib.h; it is a thin host layer over these calls.
 *
 * Conventions
 *   - return value: ZB200_OK (0) or a negative ZB200_ERR_* code; never a CPU
 *     fallback — without a usable CUDA device every call fails with
 *     ZB200_ERR_NO_DEVICE.
 *   - "_dev" entry points take DEVICE pointers and are asynchronous on
 *     `stream` unless they return a value through a host pointer, in which
 *     case they synchronise that stream before returning.
 *     A context owns ONE set of device working buffers (scratch, accumulators).  Calls on one context are
 *     serialised: on the host by the context's mutex while they enqueue, on the device by an event — a call
 *     on another stream than the previous one first waits (cudaStreamWaitEvent) for that one to finish.
 *     Callers that want two jobs to overlap on the device use two contexts.
 *   - "_host" entry points take HOST pointers, stage through the context's
 *     pinned buffers with cudaMemcpyAsync, and are synchronous.
 *   - unit of parallel work: a CHUNK (deflate: one Z_FULL_FLUSH-bounded block
 *     run, deflate.c:1211-1226) or a MEMBER (inflate: one self-contained raw /
 *     zlib / gzip stream, zlib.h:888-893).
 */
#ifndef ZB200_H
#define ZB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

#define ZB200_OK               0
#define ZB200_ERR_NO_DEVICE   (-101)  /* no CUDA device / driver: there is no CPU path */
#define ZB200_ERR_CUDA        (-102)  /* a CUDA runtime call failed: see zb200_last_error() */
#define ZB200_ERR_PARAM       (-103)
#define ZB200_ERR_NOMEM       (-104)
#define ZB200_ERR_OUTPUT      (-105)  /* output capacity too small */

typedef struct zb200_ctx zb200_ctx;

/* ---- context ------------------------------------------------------------ */
int         zb200_device_count(void);
int         zb200_create(int device, zb200_ctx **out);   /* one context per GPU; owns stream, pinned staging, scratch */
void        zb200_destroy(zb200_ctx *ctx);
int         zb200_ctx_device(const zb200_ctx *ctx);
int         zb200_sync(zb200_ctx *ctx, void *stream);
const char *zb200_last_error(void);                      /* thread-local text of the last failure */
const char *zb200_version(void);
/* pinned host memory: *_host entry points DMA straight out of / into buffers
 * obtained here (or registered with cudaHostRegister) instead of staging. */
void       *zb200_host_alloc(size_t bytes);
void        zb200_host_free(void *p);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
uint64_t    zb200_launch_count(void);

/* per-kernel timing for bench.py's roofline: while enabled, every kernel the deflate / inflate / checksum pipelines
 * launch is bracketed by CUDA events on the launching stream.  zb200_profile_read waits for them, adds the spans
 * up per kernel name (ms over all launches since the last read) and clears the list.  Off by default. */
typedef struct { char name[40]; double ms; uint64_t launches; } zb200_kernel_time;
int         zb200_profile_enable(zb200_ctx *ctx, int on);
int         zb200_profile_read(zb200_ctx *ctx, zb200_kernel_time *out, size_t cap, size_t *n);

/* ---- checksums: crc32.c:694 crc32_z, adler32.c:61 adler32_z --------------
 * which: bit 0 = CRC-32, bit 1 = Adler-32 (3 = one fused pass over the data). */
#define ZB200_CRC32   1
#define ZB200_ADLER32 2

/* d_out2[0] = crc32(init_crc, data), d_out2[1] = adler32(init_adler, data); device memory, async. */
int zb200_checksum_dev(zb200_ctx *ctx, const void *d_data, size_t len, int which,
                       uint32_t init_crc, uint32_t init_adler, uint32_t *d_out2, void *stream);
/* same, result returned to the host (synchronises). */
int zb200_checksum_dev_sync(zb200_ctx *ctx, const void *d_data, size_t len, int which,
                            uint32_t init_crc, uint32_t init_adler, uint32_t *crc, uint32_t *adler, void *stream);
/* per-segment checksums (one per gzip member / chunk): segment i = d_base[off[i] .. off[i]+len[i]). */
int zb200_checksum_segments_dev(zb200_ctx *ctx, const void *d_base, const uint64_t *d_off,
                                const uint64_t *d_len, size_t nseg, int which,
                                uint32_t *d_crc, uint32_t *d_adler, void *stream);
/* host buffer, H2D through pinned staging inside the call. */
int zb200_checksum_host(zb200_ctx *ctx, const void *data, size_t len, int which,
                        uint32_t init_crc, uint32_t init_adler, uint32_t *crc, uint32_t *adler);

/* crc32.c:1021-1049 / adler32.c:133: pure host arithmetic (no device needed). */
uint32_t zb200_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
uint32_t zb200_crc32_combine_gen(uint64_t len2);
uint32_t zb200_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op);
uint32_t zb200_adler32_combine(uint32_t adler1, uint32_t adler2, int64_t len2);

/* ---- deflate: deflate.c:954 deflate() over Z_FULL_FLUSH-bounded chunks ---- */
#define ZB200_FRAME_RAW          0   /* chunks joined by 00 00 FF FF markers, last chunk BFINAL (windowBits -15) */
#define ZB200_FRAME_ZLIB         1   /* + 2-byte header, Adler-32 trailer (deflate.c:1004-1037,1254) */
#define ZB200_FRAME_GZIP         2   /* + 10-byte header, CRC-32/ISIZE trailer (deflate.c:1042-1054,1241) */
#define ZB200_FRAME_GZIP_MEMBERS 3   /* every chunk its own gzip member (config C3 layout) */
/* OR-ed into `frame` (frames 0-2, levels 1-9): history CARRIED from chunk to chunk.  Every chunk is still its own unit of
 * parallel work and still ends on the byte-aligned 00 00 FF FF marker, but it is compressed behind the w_size bytes before
 * it, as if by deflateSetDictionary(previous 32 KiB) + deflate(chunk, Z_SYNC_FLUSH) (deflate.c:550-632,1211-1218: what
 * pigz does per block): the chunking costs 5 bytes per chunk instead of 0.7-1.3 % of the size.  Levels 4-9, Z_RLE and
 * Z_HUFFMAN_ONLY emit exactly those bytes.  The result is ONE run of blocks with sync points: a decoder needs the window
 * across chunks (any inflate has it; zb200_inflate_stream_host decodes it at its block headers).  With a dictionary
 * (zb200_deflate_opts.dict_len) the first chunk is compressed behind that. */
#define ZB200_CHUNK_CARRY        0x100
/* OR-ed into `frame` (levels 1-3, chunks of at least 1 << (memLevel + 7) bytes, no dictionary, no carried history): the greedy
 * levels with the reference's OWN hash chains — deflate_fast inserts only where its loop stands and inside short matches
 * (deflate.c:1873-1897), so its chains depend on its parse and a chunk is one serial walk: one thread per chunk, the bytes of
 * the reference's deflate() for the same chunking, at a small fraction of the default path's speed (which inserts every
 * position and finds more: 0.95-1.00 x the reference's size, not its bytes).  Ignored where it does not apply. */
#define ZB200_EXACT_FAST         0x200

/* Worst-case output bytes for n input bytes cut into chunk_size chunks (compress.c:72, deflate.c:842). */
size_t zb200_deflate_bound(size_t n, size_t chunk_size, int frame);
/* Bytes of device scratch one sub-batch of a call of this shape works in; the engine keeps two sub-batches in flight
 * (and the host pipeline two pieces), so it may hold up to four times this. */
size_t zb200_deflate_scratch_bytes(size_t n, size_t chunk_size);

/* Compress d_in[0..n).  level 1..9 (deflate.c:112-124), strategy 0..4 (zlib.h:196-200),
 * finish: 1 = last chunk carries BFINAL + trailer (Z_FINISH), 0 = every chunk ends
 * with the sync marker (Z_FULL_FLUSH) and no trailer is written.
 * d_out must be 4-byte aligned.  d_chunk_end (optional, nchunks entries) receives the
 * exclusive end offset of every chunk inside d_out.  d_total receives the stream
 * length (device uint64).  Asynchronous. */
int zb200_deflate_dev(zb200_ctx *ctx, const void *d_in, size_t n, size_t chunk_size,
                      int level, int strategy, int frame, int finish,
                      void *d_out, size_t out_cap, uint64_t *d_chunk_end, uint64_t *d_total,
                      void *stream);
/* Host buffers: H2D, compress, D2H inside.  *out_len in: capacity, out: bytes written. */
int zb200_deflate_host(zb200_ctx *ctx, const void *in, size_t n, size_t chunk_size,
                       int level, int strategy, int frame, int finish,
                       void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc);

/* Every option of deflateInit2_ / deflateSetDictionary / deflatePrime for one call (deflate.c:371-512,550-632,731-757).
 *   window_bits 9..15 (0 = 15): w_size = 1 << window_bits, MAX_DIST = w_size - 262, CINFO of the zlib header
 *   mem_level   1..9  (0 = 8):  hash_bits = mem_level + 7, lit_bufsize = 1 << (mem_level + 6) => blocks of lit_bufsize - 1 symbols
 *   dict_len    the first dict_len (<= 32768) bytes of `in` are history only (raw frame); the FIRST chunk is compressed
 *               behind them, later chunks start afresh as always
 *   first_bit   0..7: the stream starts at this bit of out[0] (deflatePrime put that many bits there; the caller ORs
 *               them in afterwards); raw frame only
 * Levels 4-9, Z_RLE and Z_HUFFMAN_ONLY emit the reference's bytes for the same settings and chunking.
 * bits_used (optional): bits in use in the last byte written (deflate.c:723 deflateUsed).  Host pointers; synchronous. */
typedef struct {
    int32_t level, strategy, window_bits, mem_level;
    uint32_t dict_len, first_bit;
} zb200_deflate_opts;
int zb200_deflate_host_opts(zb200_ctx *ctx, const void *in, size_t n, size_t chunk_size, const zb200_deflate_opts *opts,
                            int frame, int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc,
                            uint32_t *bits_used);

/* ---- inflate: inflate.c:590 inflate(), one warp per member ---------------- */
#define ZB200_WRAP_RAW   0
#define ZB200_WRAP_ZLIB  1
#define ZB200_WRAP_GZIP  2
#define ZB200_WRAP_AUTO  3   /* zlib or gzip by magic (inflateInit2 windowBits+32) */

/* Per-member status: 0 = stream end reached and checks passed; otherwise the
 * class of the reference's strm->msg (inflate.c:645-1212, inffast.c:158-283). */
enum {
    ZB200_INF_OK = 0,
    ZB200_INF_HEADER_CHECK, ZB200_INF_METHOD, ZB200_INF_WINDOW, ZB200_INF_GZ_FLAGS, ZB200_INF_GZ_HCRC,
    ZB200_INF_BLOCK_TYPE, ZB200_INF_STORED_LEN, ZB200_INF_TOO_MANY_SYMS, ZB200_INF_CODE_LENGTHS,
    ZB200_INF_BIT_REPEAT, ZB200_INF_NO_EOB, ZB200_INF_LITLEN_SET, ZB200_INF_DIST_SET,
    ZB200_INF_LITLEN_CODE, ZB200_INF_DIST_CODE, ZB200_INF_DIST_FAR, ZB200_INF_DATA_CHECK,
    ZB200_INF_LENGTH_CHECK, ZB200_INF_TRUNCATED, ZB200_INF_OUTPUT_FULL, ZB200_INF_NEED_DICT,
    ZB200_INF_COUNT
};
const char *zb200_inflate_msg(int status);   /* the reference's message literal for a status */

typedef struct {
    uint64_t in_off, in_len;     /* compressed bytes of this member inside d_in  */
    uint64_t out_off, out_cap;   /* where its output goes inside d_out           */
    uint64_t resume_bit;         /* 0 = fresh member (parse the wrapper header); else continue at this */
    uint64_t resume_out;         /*     block boundary with resume_out bytes of output already in place */
    uint64_t dict_len;           /* preset dictionary (inflateSetDictionary, inflate.c:1278-1312): the dict_len */
                                 /*     (<= 32768) bytes just BEFORE out_off in d_out; distances may reach into */
                                 /*     them; they are not part of the output, its length or its check value.   */
                                 /*     A zlib header with FDICT and dict_len == 0 gives ZB200_INF_NEED_DICT    */
                                 /*     with the DICTID in `check` (inflate.c:660-669)                          */
} zb200_member;

typedef struct {
    int32_t  status;             /* ZB200_INF_*                                   */
    uint32_t wrap_kind;          /* resolved wrapper: 0 raw, 1 zlib, 2 gzip       */
    uint32_t check;              /* CRC-32 (gzip/raw) or Adler-32 (zlib) computed over the output */
    uint32_t isize;              /* gzip ISIZE field as stored in the trailer     */
    uint64_t out_len;            /* valid bytes produced                          */
    uint64_t in_used;            /* bytes consumed incl. header/trailer (status 0) */
    uint64_t resume_bit;         /* last deflate-block boundary reached: bit offset in the member ... */
    uint64_t resume_out;         /* ... and the output bytes complete there (streaming resume point) */
} zb200_member_result;

/* Decode n_members independent members.  verify!=0 also checks the trailer
 * (CRC-32 + ISIZE for gzip, Adler-32 for zlib) against the produced bytes.
 * For a member that resumes (resume_bit != 0) `wrap` must be the resolved kind.
 * Asynchronous; d_members / d_results are device arrays. */
int zb200_inflate_dev(zb200_ctx *ctx, const void *d_in, void *d_out,
                      const zb200_member *d_members, size_t n_members, int wrap, int verify,
                      zb200_member_result *d_results, void *stream);
/* Host buffers, member table on the host; H2D/D2H inside; synchronous. */
int zb200_inflate_host(zb200_ctx *ctx, const void *in, void *out,
                       const zb200_member *members, size_t n_members, int wrap, int verify,
                       zb200_member_result *results);

/* One raw-deflate chunk compressed with a preset dictionary (deflateSetDictionary, deflate.c:550-632): the
 * first dict_len (<= 32768) bytes of `in` are history only — hashed and searched like any window content,
 * never emitted — and in[dict_len .. n) is compressed as one Z_FULL_FLUSH- (finish = 0) or Z_FINISH-terminated
 * run of blocks whose matches may reach into the dictionary.  Levels 4-9 and Z_RLE / Z_HUFFMAN_ONLY emit the
 * reference's bytes.  in_adler / in_crc cover in[dict_len .. n).  Host pointers; synchronous. */
int zb200_deflate_host_dict(zb200_ctx *ctx, const void *in, size_t n, size_t dict_len, int level, int strategy,
                            int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc);

/* A whole gzip FILE — any number of members back to back (RFC 1952 2.2; what gzread.c:76-234 walks one
 * member at a time) — decoded without an index: member starts are discovered on the device (header
 * candidates, each sized by the ISIZE in front of the next one), all members are inflated in one batch and the
 * chain is verified from the front; candidates that turn out to lie inside a member are dropped and the batch
 * redone.  Bytes after the last valid member that are no gzip member are ignored, as gzread does.
 *   out / out_cap   the concatenated output; ZB200_ERR_OUTPUT with *out_len = the size needed if it does not fit
 *   inf_status      ZB200_INF_OK, or the status of the first member that failed (*out_len = valid bytes before it)
 *   members         optional: the discovered table (in_off, in_len, out_off, out_cap = output length), up to
 *                   max_members entries; *n_members = members found
 * Host pointers; synchronous. */
int zb200_gunzip_host(zb200_ctx *ctx, const void *in, size_t n, void *out, size_t out_cap, size_t *out_len,
                      int *inf_status, zb200_member *members, size_t max_members, size_t *n_members);

/* ONE raw / zlib / gzip stream decoded in parallel at its flush points: every byte-aligned 00 00 FF FF (the empty
 * stored block deflate.c:1211-1226 emits for Z_SYNC_FLUSH / Z_FULL_FLUSH — all of this library's deflate output,
 * pigz -i files, zlib's own full flushes) is a candidate boundary; the runs between candidates are inflated as
 * one batch and the chain is verified from the front (a run that stops mid-block had a false successor, a run that
 * reaches behind its start follows a sync flush: both are merged and redone).  A stream with no (or few) flush points
 * — what the reference's compress2 / gzip write: inflate.c decodes it serially — is decoded chunk by chunk in parallel at
 * its dynamic and stored block headers instead (csrc/zb_inflate_blocks.cuh: every bit position is tested for a header that
 * validates itself, chunks are counted, chained from the front, decoded at their places, their matches resolved by pointer
 * jumping).  A stream with nothing to split on, or damaged, takes the one-member path of zb200_inflate_host — same
 * statuses either way.
 * result: status, wrap_kind, check (computed), out_len (when it exceeds out_cap: ZB200_INF_OUTPUT_FULL), in_used.
 * Host pointers; synchronous. */
int zb200_inflate_stream_host(zb200_ctx *ctx, const void *in, size_t n, int wrap, void *out, size_t out_cap,
                              zb200_member_result *result);

/* ---- one process, every GPU of the box ----
 * A zb200_multi holds one context per device (devices == NULL: all visible ones).  Each call below gives every
 * GPU a contiguous range of the units — checksum bytes, deflate chunks, inflate members — on its own host thread
 * and does the exchange step on the host: partial checksums folded with the combine functions, compressed pieces
 * laid end to end (whole chunks per piece: the bytes are those of the single-GPU call), header / trailer written
 * last.  Host pointers (pinned memory is DMA-ed directly); synchronous; same arguments as the single-GPU calls. */
typedef struct zb200_multi zb200_multi;
int zb200_multi_create(const int *devices, int n_devices, zb200_multi **out);
void zb200_multi_destroy(zb200_multi *m);
int zb200_multi_count(const zb200_multi *m);
int zb200_multi_checksum_host(zb200_multi *m, const void *data, size_t n, int which, uint32_t init_crc, uint32_t init_adler,
                              uint32_t *crc, uint32_t *adler);
int zb200_multi_deflate_host(zb200_multi *m, const void *in, size_t n, size_t chunk_size, int level, int strategy, int frame,
                             int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc);
int zb200_multi_inflate_host(zb200_multi *m, const void *in, void *out, const zb200_member *members, size_t n_members,
                             int wrap, int verify, zb200_member_result *results);

/* Self-test of the warp-parallel decode-table construction (csrc/zb_inflate_tables.cuh) against the
 * serial one that follows inftrees.c:32-299: for each of n_cases sets of code lengths (lens: 320 bytes
 * per case = nlen literal/length lengths followed by ndist distance lengths; counts: nlen, ndist per
 * case) verdict[i] = 0 when both give the same status and, where accepted, identical tables.
 * Host pointers; synchronous.  Used by tests/test_gpu_inflate.py. */
int zb200_selftest_tables(zb200_ctx *ctx, const uint8_t *lens, const uint32_t *counts, size_t n_cases,
                          uint32_t *verdict);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* ZB200_H */
