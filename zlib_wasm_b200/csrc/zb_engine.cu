// zb_engine.cu — context management, pinned staging, and the checksum part of
// the C ABI (include/zb200.h).  The deflate / inflate entry points live next to
// their kernels (zb_deflate.cu, zb_inflate.cu).
#include "zb_internal.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <condition_variable>
#include <thread>

namespace zb {

// Pageable caller memory passes through the pinned stages by memcpy, and ONE host thread copies at 10-12 GB/s: that, not
// the link (55 GB/s) nor the kernels, bounded every zlib.h call on a large malloc'd buffer (crc32 of 512 MiB: 38 ms of which
// 36 were this memcpy).  Copies of 4 MiB and more are therefore split over the calling thread and three helpers (started
// on first use, parked on a condition variable in between; one split copy at a time — a second caller copies alone).
namespace {
struct CopyPool {
    static constexpr int kHelpers = 3;
    std::mutex mu, call_mu;
    std::condition_variable cv_work, cv_done;
    struct Job { uint8_t *d; const uint8_t *s; size_t n; } jobs[kHelpers];
    uint64_t gen = 0;
    int pending = 0;
    bool started = false;
    void helper(int i) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_work.wait(lk, [&] { return gen != seen; });
            seen = gen;
            const Job j = jobs[i];
            lk.unlock();
            if (j.n) memcpy(j.d, j.s, j.n);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    void copy(void *dst, const void *src, size_t n) {
        static const bool off = [] { const char *e = getenv("ZB200_COPY_THREADS"); return e && atoi(e) <= 1; }();
        if (n < ((size_t)4 << 20) || off || !call_mu.try_lock()) { memcpy(dst, src, n); return; }
        uint8_t *d = (uint8_t *)dst;
        const uint8_t *sp = (const uint8_t *)src;
        const size_t part = ((n / (kHelpers + 1)) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!started) { for (int i = 0; i < kHelpers; ++i) std::thread([this, i] { helper(i); }).detach(); started = true; }
            for (int i = 0; i < kHelpers; ++i) {
                const size_t a = part * (size_t)(i + 1), b = a + part < n ? a + part : n;
                jobs[i] = Job{d + a, sp + a, a < n ? (i + 1 == kHelpers ? n - a : b - a) : 0};
            }
            pending = kHelpers;
            ++gen;
        }
        cv_work.notify_all();
        memcpy(d, sp, part < n ? part : n);
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return pending == 0; });
        }
        call_mu.unlock();
    }
};
CopyPool &copy_pool() { static CopyPool *p = new CopyPool; return *p; }   // (never destroyed: its helpers outlive main)
}  // namespace
void host_copy(void *dst, const void *src, size_t n) { copy_pool().copy(dst, src, n); }

static thread_local char t_err[512];
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

int ensure_scratch(zb200_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return ZB200_OK;
    if (ctx->d_scratch) {
        ZB_CUDA(cudaDeviceSynchronize());   // nobody may still be using the old block
        ZB_CUDA(cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    }
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc(&ctx->d_scratch, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&ctx->d_scratch, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("device scratch of %zu bytes: %s", bytes, cudaGetErrorString(e));
            return ZB200_ERR_NOMEM;
        }
    }
    ctx->scratch_bytes = want;
    return ZB200_OK;
}

static int grow(uint8_t **p, size_t *have, size_t need) {
    if (need <= *have) return ZB200_OK;
    if (*p) { ZB_CUDA(cudaDeviceSynchronize()); ZB_CUDA(cudaFree(*p)); *p = nullptr; *have = 0; }
    size_t want = need + (need >> 3) + 4096;
    cudaError_t e = cudaMalloc((void **)p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device buffer of %zu bytes: %s", want, cudaGetErrorString(e));
        return ZB200_ERR_NOMEM;
    }
    *have = want;
    return ZB200_OK;
}

int ensure_io(zb200_ctx *ctx, size_t in_bytes, size_t out_bytes) {
    int r = grow(&ctx->d_io_in, &ctx->io_in_bytes, in_bytes);
    if (r) return r;
    return grow(&ctx->d_io_out, &ctx->io_out_bytes, out_bytes);
}

// Pageable host memory -> device through the two pinned stages: memcpy into
// stage k on the CPU overlaps the DMA of stage k^1.
int h2d_staged(zb200_ctx *ctx, void *d_dst, const void *h_src, size_t n, cudaStream_t s) {
    const uint8_t *src = (const uint8_t *)h_src;
    uint8_t *dst = (uint8_t *)d_dst;
    int k = 0;
    for (size_t off = 0; off < n;) {
        size_t m = n - off < ctx->stage_bytes ? n - off : ctx->stage_bytes;
        ZB_CUDA(cudaEventSynchronize(ctx->stage_ev[k]));
        host_copy(ctx->h_stage[k], src + off, m);
        ZB_CUDA(cudaMemcpyAsync(dst + off, ctx->h_stage[k], m, cudaMemcpyHostToDevice, s));
        ZB_CUDA(cudaEventRecord(ctx->stage_ev[k], s));
        off += m; k ^= 1;
    }
    return ZB200_OK;
}

int d2h_staged(zb200_ctx *ctx, void *h_dst, const void *d_src, size_t n, cudaStream_t s) {
    uint8_t *dst = (uint8_t *)h_dst;
    const uint8_t *src = (const uint8_t *)d_src;
    size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
    int k = 0;
    for (size_t off = 0; off < n;) {
        size_t m = n - off < ctx->stage_bytes ? n - off : ctx->stage_bytes;
        if (pend_n[k]) {
            ZB_CUDA(cudaEventSynchronize(ctx->stage_ev[k]));
            host_copy(dst + pend_off[k], ctx->h_stage[k], pend_n[k]);
            pend_n[k] = 0;
        } else {
            ZB_CUDA(cudaEventSynchronize(ctx->stage_ev[k]));
        }
        ZB_CUDA(cudaMemcpyAsync(ctx->h_stage[k], src + off, m, cudaMemcpyDeviceToHost, s));
        ZB_CUDA(cudaEventRecord(ctx->stage_ev[k], s));
        pend_off[k] = off; pend_n[k] = m;
        off += m; k ^= 1;
    }
    for (int i = 0; i < 2; ++i, k ^= 1)
        if (pend_n[k]) {
            ZB_CUDA(cudaEventSynchronize(ctx->stage_ev[k]));
            host_copy(dst + pend_off[k], ctx->h_stage[k], pend_n[k]);
            pend_n[k] = 0;
        }
    return ZB200_OK;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

int h2d_auto(zb200_ctx *ctx, void *d_dst, const void *h_src, size_t n, cudaStream_t s) {
    if (n == 0) return ZB200_OK;
    if (is_pinned(h_src)) { ZB_CUDA(cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, s)); return ZB200_OK; }
    return h2d_staged(ctx, d_dst, h_src, n, s);
}

int d2h_auto(zb200_ctx *ctx, void *h_dst, const void *d_src, size_t n, cudaStream_t s) {
    if (n == 0) return ZB200_OK;
    if (is_pinned(h_dst)) { ZB_CUDA(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, s)); return ZB200_OK; }
    return d2h_staged(ctx, h_dst, d_src, n, s);
}

void prof_mark_slow(zb200_ctx *ctx, cudaStream_t s, const char *name) {
    cudaEvent_t ev = nullptr;
    if (!ctx->prof_pool.empty()) { ev = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else if (cudaEventCreate(&ev) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaEventRecord(ev, s) != cudaSuccess) { cudaGetLastError(); ctx->prof_pool.push_back(ev); return; }
    ctx->prof_marks.push_back({name, ev, s});
}

}  // namespace zb

using namespace zb;

extern "C" {

int zb200_profile_enable(zb200_ctx *ctx, int on) {
    if (!ctx) return ZB200_ERR_PARAM;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->prof_on = on != 0;
    return ZB200_OK;
}

int zb200_profile_read(zb200_ctx *ctx, zb200_kernel_time *out, size_t cap, size_t *n_out) {
    if (!ctx || !n_out || (!out && cap)) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->mu);
    size_t n = 0;
    auto &m = ctx->prof_marks;
    for (size_t i = 0; i < m.size(); ++i) ZB_CUDA(cudaEventSynchronize(m[i].ev));
    for (size_t i = 0; i + 1 < m.size(); ++i) {
        if (!m[i].name || m[i + 1].s != m[i].s) continue;          // a pipeline ended here
        float ms = 0;
        if (cudaEventElapsedTime(&ms, m[i].ev, m[i + 1].ev) != cudaSuccess) { cudaGetLastError(); continue; }
        size_t k = 0;
        for (; k < n; ++k) if (!strncmp(out[k].name, m[i].name, sizeof out[k].name - 1)) break;
        if (k == n) {
            if (n == cap) continue;
            memset(&out[n], 0, sizeof out[n]);
            strncpy(out[n].name, m[i].name, sizeof out[n].name - 1);
            ++n;
        }
        out[k].ms += ms; out[k].launches += 1;
    }
    for (auto &x : m) ctx->prof_pool.push_back(x.ev);
    m.clear();
    *n_out = n;
    return ZB200_OK;
}

const char *zb200_version(void) { return "zlib-b200 0.1 (sm_100a) / zlib 1.3.1.1-motley API"; }
const char *zb200_last_error(void) { return t_err; }
uint64_t zb200_launch_count(void) { return g_launches.load(); }

int zb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int zb200_create(int device, zb200_ctx **out) {
    if (!out) return ZB200_ERR_PARAM;
    *out = nullptr;
    int n = zb200_device_count();
    if (n <= 0) { set_error("no CUDA device available; this library has no CPU path"); return ZB200_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return ZB200_ERR_PARAM; }
    ZB_CUDA(cudaSetDevice(device));
    zb200_ctx *ctx = new (std::nothrow) zb200_ctx();
    if (!ctx) return ZB200_ERR_NOMEM;
    ctx->device = device;
    int r = ZB200_OK;
    auto fail = [&](int code) { zb200_destroy(ctx); return code; };
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return fail(ZB200_ERR_NO_DEVICE);
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->back_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    for (auto &a : ctx->aux_stream) if (cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    if (cudaEventCreateWithFlags(&ctx->busy_ev, cudaEventDisableTiming) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    if (cudaMalloc((void **)&ctx->d_pipe, 160 * sizeof(uint64_t)) != cudaSuccess) return fail(ZB200_ERR_NOMEM);
    if (cudaMallocHost((void **)&ctx->h_pipe, 160 * sizeof(uint64_t)) != cudaSuccess) return fail(ZB200_ERR_NOMEM);
    ctx->stage_bytes = 32u << 20;
    for (int k = 0; k < zb200_ctx::kStages; ++k) {
        if (cudaMallocHost((void **)&ctx->h_stage[k], ctx->stage_bytes) != cudaSuccess) return fail(ZB200_ERR_NOMEM);
        if (cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming) != cudaSuccess) return fail(ZB200_ERR_CUDA);
    }
    if (cudaMalloc((void **)&ctx->d_small, 64 * sizeof(uint64_t)) != cudaSuccess) return fail(ZB200_ERR_NOMEM);
    if (cudaMallocHost((void **)&ctx->h_small, 64 * sizeof(uint64_t)) != cudaSuccess) return fail(ZB200_ERR_NOMEM);
    ctx->h_small[32] = 0; ctx->h_small[33] = 2; ctx->h_small[34] = 10;   // constants DMA-ed by deflate_launch
    if ((r = checksum_init(ctx)) != ZB200_OK) return fail(r);
    if ((r = deflate_init(ctx)) != ZB200_OK) return fail(r);
    if ((r = inflate_init(ctx)) != ZB200_OK) return fail(r);
    if ((r = ensure_scratch(ctx, 1u << 20)) != ZB200_OK) return fail(r);
    *out = ctx;
    return ZB200_OK;
}

void zb200_destroy(zb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int k = 0; k < zb200_ctx::kStages; ++k) {
        if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]);
        if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
    }
    if (ctx->d_small) cudaFree(ctx->d_small);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->d_crc_tables) cudaFree(ctx->d_crc_tables);
    if (ctx->d_deflate_tables) cudaFree(ctx->d_deflate_tables);
    if (ctx->d_inflate_tables) cudaFree(ctx->d_inflate_tables);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_io_in) cudaFree(ctx->d_io_in);
    if (ctx->d_io_out) cudaFree(ctx->d_io_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->back_stream) cudaStreamDestroy(ctx->back_stream);
    for (auto a : ctx->aux_stream) if (a) cudaStreamDestroy(a);
    if (ctx->busy_ev) cudaEventDestroy(ctx->busy_ev);
    if (ctx->d_pipe) cudaFree(ctx->d_pipe);
    if (ctx->h_pipe) cudaFreeHost(ctx->h_pipe);
    for (auto &x : ctx->prof_marks) cudaEventDestroy(x.ev);
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    cudaGetLastError();
    delete ctx;
}

int zb200_ctx_device(const zb200_ctx *ctx) { return ctx ? ctx->device : -1; }

int zb200_sync(zb200_ctx *ctx, void *stream) {
    if (!ctx) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    ZB_CUDA(cudaStreamSynchronize(pick_stream(ctx, stream)));
    return ZB200_OK;
}

// ---- checksums --------------------------------------------------------------
int zb200_checksum_dev(zb200_ctx *ctx, const void *d_data, size_t len, int which,
                       uint32_t init_crc, uint32_t init_adler, uint32_t *d_out2, void *stream) {
    if (!ctx || !d_out2 || (!d_data && len)) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    CtxUse use(ctx, s);
    // accumulators: one CkAccum at the front of the small area (slots 8..11)
    CkAccum *acc = (CkAccum *)(ctx->d_small + 8);
    return checksum_launch(ctx, (const uint8_t *)d_data, nullptr, nullptr, len, 1, which, init_crc, init_adler,
                           d_out2, d_out2 + 1, acc, s);
}

int zb200_checksum_dev_sync(zb200_ctx *ctx, const void *d_data, size_t len, int which,
                            uint32_t init_crc, uint32_t init_adler, uint32_t *crc, uint32_t *adler, void *stream) {
    if (!ctx || (!d_data && len)) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = pick_stream(ctx, stream);
    CtxUse use(ctx, s);
    uint32_t *d_out2 = (uint32_t *)ctx->d_small;
    int r = checksum_launch(ctx, (const uint8_t *)d_data, nullptr, nullptr, len, 1, which, init_crc, init_adler,
                            d_out2, d_out2 + 1, (CkAccum *)(ctx->d_small + 8), s);
    if (r) return r;
    ZB_CUDA(cudaMemcpyAsync(ctx->h_small, d_out2, 8, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    const uint32_t *h = (const uint32_t *)ctx->h_small;
    if (crc && (which & ZB200_CRC32)) *crc = h[0];
    if (adler && (which & ZB200_ADLER32)) *adler = h[1];
    return ZB200_OK;
}

int zb200_checksum_segments_dev(zb200_ctx *ctx, const void *d_base, const uint64_t *d_off,
                                const uint64_t *d_len, size_t nseg, int which,
                                uint32_t *d_crc, uint32_t *d_adler, void *stream) {
    if (!ctx || !d_off || !d_len) return ZB200_ERR_PARAM;
    if (nseg == 0) return ZB200_OK;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, pick_stream(ctx, stream));
    int r = ensure_scratch(ctx, nseg * sizeof(CkAccum));
    if (r) return r;
    return checksum_launch(ctx, (const uint8_t *)d_base, d_off, d_len, 0, nseg, which, 0, 1,
                           d_crc, d_adler, (CkAccum *)ctx->d_scratch, pick_stream(ctx, stream));
}

int zb200_checksum_host(zb200_ctx *ctx, const void *data, size_t len, int which,
                        uint32_t init_crc, uint32_t init_adler, uint32_t *crc, uint32_t *adler) {
    if (!ctx || (!data && len)) return ZB200_ERR_PARAM;
    ZB_CUDA(cudaSetDevice(ctx->device));
    CtxUse use(ctx, ctx->stream);
    // Stream the buffer through the device in pieces, chaining the running
    // values on the device (crc32/adler32 are running checksums, zlib.h:1711-1768):
    // host memory of any size needs O(piece) device memory and one final 8-byte
    // read-back.  Pinned callers (zb200_host_alloc / cudaHostRegister) are DMA-ed
    // directly; pageable memory goes through the two pinned stages.
    const size_t piece = ctx->stage_bytes;
    int r = ensure_io(ctx, piece, 0);
    if (r) return r;
    cudaStream_t s = ctx->stream;
    cudaPointerAttributes attr;
    bool pinned = false;
    if (len && cudaPointerGetAttributes(&attr, data) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    uint32_t *d_run = (uint32_t *)ctx->d_small;            // running (crc, adler) on the device
    uint32_t *h_run = (uint32_t *)ctx->h_small;
    h_run[0] = init_crc; h_run[1] = init_adler;
    ZB_CUDA(cudaMemcpyAsync(d_run, h_run, 8, cudaMemcpyHostToDevice, s));
    const uint8_t *src = (const uint8_t *)data;
    CkAccum *acc = (CkAccum *)(ctx->d_small + 8);
    int k = 0;
    for (size_t off = 0; off < len; ) {
        size_t m = len - off < piece ? len - off : piece;
        if (pinned) {
            ZB_CUDA(cudaMemcpyAsync(ctx->d_io_in, src + off, m, cudaMemcpyHostToDevice, s));
        } else {
            ZB_CUDA(cudaEventSynchronize(ctx->stage_ev[k]));
            host_copy(ctx->h_stage[k], src + off, m);
            ZB_CUDA(cudaMemcpyAsync(ctx->d_io_in, ctx->h_stage[k], m, cudaMemcpyHostToDevice, s));
            ZB_CUDA(cudaEventRecord(ctx->stage_ev[k], s));
        }
        r = checksum_launch(ctx, ctx->d_io_in, nullptr, nullptr, m, 1, which, 0, 1, d_run, d_run + 1, acc, s, d_run);
        if (r) return r;
        off += m; k ^= 1;
    }
    ZB_CUDA(cudaMemcpyAsync(h_run, d_run, 8, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    if (crc && (which & ZB200_CRC32)) *crc = h_run[0];
    if (adler && (which & ZB200_ADLER32)) *adler = h_run[1];
    return ZB200_OK;
}

void *zb200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void zb200_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- combine (pure host arithmetic) -------------------------------------------
static const uint32_t *host_x2n() {
    static X2nTable t;
    static std::once_flag once;
    std::call_once(once, [] { gf2_fill_x2n(t); });
    return t.v;
}

uint32_t zb200_crc32_combine_gen(uint64_t len2) { return gf2_xpow(host_x2n(), len2, 3); }        // crc32.c:1034
uint32_t zb200_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op) { return gf2_mul(op, crc1) ^ crc2; }  // :1047
uint32_t zb200_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {                      // :1021
    return gf2_mul(gf2_xpow(host_x2n(), len2, 3), crc1) ^ crc2;
}

uint32_t zb200_adler32_combine(uint32_t adler1, uint32_t adler2, int64_t len2) {                 // adler32.c:133-155
    if (len2 < 0) return 0xffffffffu;
    const uint64_t P = kAdlerBase;
    const uint64_t rem = (uint64_t)len2 % P;
    const uint64_t a1 = adler1 & 0xffff, b1 = (adler1 >> 16) & 0xffff;
    const uint64_t a2 = adler2 & 0xffff, b2 = (adler2 >> 16) & 0xffff;
    // s1 = a1 + a2 - 1 ; s2 = b1 + b2 + len2*(a1 - 1)   (mod P)
    const uint64_t s1 = (a1 + a2 + P - 1) % P;
    const uint64_t s2 = (b1 + b2 + rem * a1 + P - rem) % P;
    return (uint32_t)(s1 | (s2 << 16));
}

}  // extern "C"
