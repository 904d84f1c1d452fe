// zb_internal.h — engine-private declarations shared by the CUDA translation
// units and the host API layers.  Nothing here is part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>
#include <atomic>
#include <vector>

#include "../../include/zb200.h"
#include "zb_gf2.h"

namespace zb {

// ---- error plumbing --------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define ZB_CUDA(expr)                                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            ::zb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                            cudaGetErrorString(_e));                               \
            return ZB200_ERR_CUDA;                                                 \
        }                                                                          \
    } while (0)

#define ZB_LAUNCHED() (::zb::g_launches.fetch_add(1, std::memory_order_relaxed))
#define ZB_CHECK_LAUNCH() ZB_CUDA(cudaGetLastError())

// ---- device-resident constant tables --------------------------------------
// CRC "Horner" tables: row j, entry b = (b << 8j) * x^(8*stride) mod p, for the
// two strides the kernels use (see zb_checksum.cu), plus x^(2^k) and x^32.
struct CrcTables {
    uint32_t big[4][256];   // stride 16 KiB  (1024 threads x 16 B)
    uint32_t seg[4][256];   // stride  4 KiB  ( 256 threads x 16 B)
    uint32_t x2n[32];
    uint32_t x32;
    uint32_t pad[31];
};

struct CkAccum {            // per-segment accumulators, zeroed before each launch
    unsigned long long a;   // sum of per-part Adler s1 contributions (each < 65521)
    unsigned long long b;   // sum of per-part Adler s2 contributions
    uint32_t crc;           // xor of weighted pure-CRC partials
    uint32_t pad;
};

}  // namespace zb

// ---- the context -----------------------------------------------------------
struct zb200_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;          // the context's own stream
    cudaStream_t copy_stream = nullptr;     // H2D/D2H staging stream
    cudaStream_t back_stream = nullptr;     // D2H of finished pieces while later ones are still computed
    static constexpr int kAux = 3;
    cudaStream_t aux_stream[kAux] = {nullptr, nullptr, nullptr};   // deflate: sub-batch j of a call runs on stream j % K (K in flight)
    uint64_t *d_pipe = nullptr, *h_pipe = nullptr;   // per-piece (total, crc|adler<<32) of the pipelined host entry points: 2 x 80 u64
    std::mutex mu;                          // serialises use of the scratch/staging below

    zb::CrcTables *d_crc_tables = nullptr;
    void *d_deflate_tables = nullptr;       // zb::DeflateDeviceTables (zb_deflate.cu)
    void *d_inflate_tables = nullptr;       // zb::InflateDeviceTables (zb_inflate.cu)

    // grow-only device scratch (checksum accumulators, deflate/inflate workspaces)
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    // small device result area + its pinned mirror
    uint64_t *d_small = nullptr;            // 64 x u64
    uint64_t *h_small = nullptr;
    // pinned staging ring for *_host entry points
    static constexpr int kStages = 2;
    uint8_t *h_stage[kStages] = {nullptr, nullptr};
    cudaEvent_t stage_ev[kStages] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // device I/O buffers for *_host entry points (grow-only)
    uint8_t *d_io_in = nullptr;  size_t io_in_bytes = 0;
    uint8_t *d_io_out = nullptr; size_t io_out_bytes = 0;
    // The context's device state (scratch, accumulators, I/O buffers) is shared by every call.  ctx->mu serialises the
    // HOST side (enqueuing); the device side is ordered by this event: a call that runs on another stream than the
    // previous one first waits for it (zb::CtxUse).
    cudaEvent_t busy_ev = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy_valid = false;
    // per-kernel timing (zb200_profile_*): while on, every kernel launch of the pipelines is preceded by an event
    // on its stream; a kernel's time is the span to the next mark
    bool prof_on = false;
    struct ProfMark { const char *name; cudaEvent_t ev; cudaStream_t s; };
    std::vector<ProfMark> prof_marks;
    std::vector<cudaEvent_t> prof_pool;
};

namespace zb {

int ensure_scratch(zb200_ctx *ctx, size_t bytes);
int ensure_io(zb200_ctx *ctx, size_t in_bytes, size_t out_bytes);
int h2d_staged(zb200_ctx *ctx, void *d_dst, const void *h_src, size_t n, cudaStream_t s);
int d2h_staged(zb200_ctx *ctx, void *h_dst, const void *d_src, size_t n, cudaStream_t s);
// pinned host memory is DMA-ed directly (async); pageable memory goes through the stages
bool is_pinned(const void *p);
// memcpy between pageable caller memory and the pinned stages, split over a few host threads from 4 MiB up (zb_engine.cu)
void host_copy(void *dst, const void *src, size_t n);
int h2d_auto(zb200_ctx *ctx, void *d_dst, const void *h_src, size_t n, cudaStream_t s);
int d2h_auto(zb200_ctx *ctx, void *h_dst, const void *d_src, size_t n, cudaStream_t s);
// zb_engine.cu: mark the start of the kernel `name` on stream s (name == nullptr: end of a pipeline); no-op unless profiling
void prof_mark_slow(zb200_ctx *ctx, cudaStream_t s, const char *name);
inline void prof_mark(zb200_ctx *ctx, cudaStream_t s, const char *name) { if (ctx->prof_on) prof_mark_slow(ctx, s, name); }
inline cudaStream_t pick_stream(zb200_ctx *ctx, void *stream) {
    return stream ? (cudaStream_t)stream : ctx->stream;
}
// One use of a context: holds ctx->mu for the scope and orders the work enqueued on stream `s` after whatever the
// previous use left running on a different stream (cudaStreamWaitEvent: no host synchronisation).
struct CtxUse {
    zb200_ctx *c; cudaStream_t s; std::unique_lock<std::mutex> lk;
    CtxUse(zb200_ctx *ctx, cudaStream_t stream) : c(ctx), s(stream), lk(ctx->mu) {
        if (c->busy_valid && c->busy_stream != s && c->busy_ev) cudaStreamWaitEvent(s, c->busy_ev, 0);
    }
    ~CtxUse() {
        if (c->busy_ev && cudaEventRecord(c->busy_ev, s) == cudaSuccess) { c->busy_stream = s; c->busy_valid = true; }
        else cudaGetLastError();
    }
    CtxUse(const CtxUse &) = delete;
    CtxUse &operator=(const CtxUse &) = delete;
};

// zb_checksum.cu
int checksum_init(zb200_ctx *ctx);
int checksum_launch(zb200_ctx *ctx, const uint8_t *d_base, const uint64_t *d_off, const uint64_t *d_len,
                    uint64_t single_len, size_t nseg, int which, uint32_t init_crc, uint32_t init_adler,
                    uint32_t *d_crc, uint32_t *d_adler, CkAccum *d_acc, cudaStream_t s,
                    const uint32_t *d_init2 = nullptr);

// zb_inflate.cu: one stream decoded in parallel at its flush points (host pointers); *applicable = 0 -> take the one-member path,
// 1 -> *status etc. are set, 2 -> ... and the deflate data was seen to end (a TRUNCATED status then means: the trailer is missing)
// cont != nullptr: src continues a raw stream at a block boundary (bit `bit0` of src[0]) whose last hist_len (<= 32768)
// output bytes are `hist`; *check then covers the bytes produced by this call only (check_kind 1: Adler-32, else CRC-32).
struct StreamContinuation { const uint8_t *hist; size_t hist_len; uint32_t bit0; int check_kind; };
// alt (optional): where the output goes when it does not fit `out` — grow(self, need) returns a buffer of at least `need` bytes
// (or nullptr: ZB200_INF_OUTPUT_FULL as without it); `used` tells the caller which of the two holds the bytes.
struct StreamOutAlt { uint8_t *(*grow)(void *self, size_t need); void *self; bool used; };
int inflate_stream_parallel(zb200_ctx *ctx, const uint8_t *src, size_t n, int wrap, uint8_t *out, size_t out_cap,
                            size_t *out_len, int *status, size_t *in_used, uint32_t *check, int *applicable,
                            const StreamContinuation *cont = nullptr, int blocks_mode = 0, uint32_t *end_bit = nullptr,
                            StreamOutAlt *alt = nullptr);
// blocks_mode: what to do with a stream that has no (or too few) flush points — 0: leave it to the caller's one-member path,
// 1: decode it in parallel at its dynamic block headers (zb_inflate_blocks.cuh) when the whole stream is there, 2: also when
// only a prefix can be delivered (ZB200_INF_TRUNCATED; the prefix ends at bit *end_bit of src[*in_used], a block boundary).

// zb_deflate.cu
int deflate_init(zb200_ctx *ctx);
// deflateTune: {good_length, max_lazy, nice_length, max_chain} for the deflate calls this thread makes next (nullptr: the level's own)
const int *deflate_tune_override();
void deflate_tune_set(const int *four);
// everything deflateInit2_ / deflateSetDictionary / deflatePrime can ask of one engine call
struct DeflateOpts {
    int level = 6, strategy = 0;
    int window_bits = 15, mem_level = 8;     // deflate.c:440-455: w_size, hash_bits, lit_bufsize
    size_t skip = 0;                         // preset dictionary at the head of a single raw chunk
    unsigned first_bit = 0;                  // deflatePrime: the stream starts at this bit (0..7) of out[0]
    int slot = 0;                            // flight slot (0 / 1) of a caller that keeps two calls in flight on two streams of one
    size_t slot_bytes = 0;                   //   context, and the scratch bytes slot 0 may use (slot 1's part starts there)
    bool exact_fast = false;                 // ZB200_EXACT_FAST: levels 1-3 with the reference's own parse-dependent chains (one thread per chunk)
    bool carry = false;                      // ZB200_CHUNK_CARRY: every chunk is compressed behind the w_size bytes before it; `skip` is
                                             // then the history in front of the FIRST chunk (a dictionary, or the previous piece's tail)
};
// The 8 bytes at ctx->d_small + 19 receive {bits in use in the stream's last byte (deflateUsed), 0}.
int deflate_launch_opts(zb200_ctx *ctx, const uint8_t *d_in, size_t n, size_t S, const DeflateOpts &o, int frame,
                        int finish, uint8_t *d_out, size_t out_cap, uint64_t *d_chunk_end, uint64_t *d_total,
                        uint32_t *d_sums_out, cudaStream_t s);
inline int deflate_launch(zb200_ctx *ctx, const uint8_t *d_in, size_t n, size_t S, int level, int strategy, int frame,
                          int finish, uint8_t *d_out, size_t out_cap, uint64_t *d_chunk_end, uint64_t *d_total,
                          uint32_t *d_sums_out, cudaStream_t s, size_t skip = 0) {
    DeflateOpts o;
    o.level = level; o.strategy = strategy; o.skip = skip;
    return deflate_launch_opts(ctx, d_in, n, S, o, frame, finish, d_out, out_cap, d_chunk_end, d_total, d_sums_out, s);
}
// chunks the chain kernel keeps resident at once (one wave), and a piece size for the host pipelines: `want` bytes rounded to
// whole chunks and, from one wave up, down to whole waves (the ordered kernels cost a full wave for any part of one)
size_t deflate_wave_chunks(zb200_ctx *ctx, int mem_level);
size_t deflate_piece_bytes(zb200_ctx *ctx, size_t want, size_t S, int mem_level, int level);
// zb_zlib_api.cu: the process-wide context behind the zlib.h surface (nullptr without a usable device)
zb200_ctx *zlib_api_ctx();
// zb_inflate.cu
int inflate_init(zb200_ctx *ctx);
size_t inflate_work_bytes(size_t n_members);
// d_blog (optional, single-member calls): blog_cap entries of (bit offset, output position | BFINAL << 63) for every block
// boundary passed, count in d_blog[0], entries from d_blog[2] (zb_inflate.cuh InflateState::blog)
int inflate_launch(zb200_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, const zb200_member *d_members,
                   size_t n, int wrap, int verify, zb200_member_result *d_results, void *d_work, cudaStream_t s,
                   uint64_t *d_blog = nullptr, uint32_t blog_cap = 0);

}  // namespace zb
