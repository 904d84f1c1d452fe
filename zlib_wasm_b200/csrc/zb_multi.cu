// zb_multi.cu — one process, every GPU of the box (the path shards with no data-path collective).
//
// bench.py drives one process per GPU under torchrun; a caller that links the library wants the same
// sharding behind one call.  A zb200_multi holds one engine context per device; each entry point gives
// every context a contiguous range of the units (checksum bytes, deflate chunks, inflate members), runs
// the ranges on one host thread per GPU, and does the exchange step on the host: partial checksums are
// folded left to right with crc32_combine / adler32_combine (crc32.c:1021, adler32.c:133), compressed
// pieces are laid end to end (every piece but the last ends on a full-flush marker, so the concatenation
// is the very stream one GPU would emit), the stream header and trailer (deflate.c:1004-1054,1239-1256)
// are written here.  Host code only.
#include "zb_internal.h"
#include <condition_variable>
#include <string.h>
#include <thread>
#include <vector>

using namespace zb;

struct zb200_multi {
    std::vector<zb200_ctx *> ctx;
};

namespace {

struct Rendezvous {                         // every thread arrives once; all leave together
    std::mutex m; std::condition_variable cv; size_t expected, arrived = 0;
    explicit Rendezvous(size_t n) : expected(n) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> l(m);
        if (++arrived == expected) cv.notify_all();
        else cv.wait(l, [&] { return arrived >= expected; });
    }
};

template <class F> void on_all(size_t n, F f) {
    std::vector<std::thread> th;
    for (size_t k = 1; k < n; ++k) th.emplace_back(f, k);
    f(0);
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

int zb200_multi_create(const int *devices, int n_devices, zb200_multi **out) {
    if (!out || n_devices < 0) return ZB200_ERR_PARAM;
    *out = nullptr;
    const int have = zb200_device_count();
    if (have <= 0) { set_error("no CUDA device available; this library has no CPU path"); return ZB200_ERR_NO_DEVICE; }
    std::vector<int> dev;
    if (devices && n_devices) dev.assign(devices, devices + n_devices);
    else for (int d = 0; d < have; ++d) dev.push_back(d);
    zb200_multi *m = new (std::nothrow) zb200_multi();
    if (!m) return ZB200_ERR_NOMEM;
    for (int d : dev) {
        zb200_ctx *c = nullptr;
        const int r = zb200_create(d, &c);
        if (r != ZB200_OK) { zb200_multi_destroy(m); return r; }
        m->ctx.push_back(c);
    }
    *out = m;
    return ZB200_OK;
}

void zb200_multi_destroy(zb200_multi *m) {
    if (!m) return;
    for (zb200_ctx *c : m->ctx) zb200_destroy(c);
    delete m;
}

int zb200_multi_count(const zb200_multi *m) { return m ? (int)m->ctx.size() : 0; }

int zb200_multi_checksum_host(zb200_multi *m, const void *data, size_t n, int which, uint32_t init_crc, uint32_t init_adler,
                              uint32_t *crc, uint32_t *adler) {
    if (!m || m->ctx.empty() || (!data && n)) return ZB200_ERR_PARAM;
    const size_t g = m->ctx.size();
    const size_t per = ((n + g - 1) / g + 4095) & ~(size_t)4095;
    std::vector<uint32_t> pc(g, 0), pa(g, 1);
    std::vector<size_t> len(g, 0);
    std::vector<int> rc(g, ZB200_OK);
    on_all(g, [&](size_t k) {
        const size_t off = k * per < n ? k * per : n;
        len[k] = n - off < per ? n - off : per;
        if (!len[k] && k) return;
        rc[k] = zb200_checksum_host(m->ctx[k], (const uint8_t *)data + off, len[k], which, k ? 0u : init_crc, k ? 1u : init_adler, &pc[k], &pa[k]);
    });
    uint32_t c = pc[0], a = pa[0];
    for (size_t k = 0; k < g; ++k) if (rc[k] != ZB200_OK) return rc[k];
    for (size_t k = 1; k < g; ++k) {
        if (!len[k]) continue;
        c = zb200_crc32_combine(c, pc[k], len[k]);
        a = zb200_adler32_combine(a, pa[k], (int64_t)len[k]);
    }
    if (crc) *crc = c;
    if (adler) *adler = a;
    return ZB200_OK;
}

int zb200_multi_deflate_host(zb200_multi *m, const void *in, size_t n, size_t chunk_size, int level, int strategy, int frame,
                             int finish, void *out, size_t *out_len, uint32_t *in_adler, uint32_t *in_crc) {
    // ZB200_CHUNK_CARRY: a GPU's first chunk is compressed behind the 32 KiB before its piece (copied along), like every other chunk
    const size_t W = ((frame & ZB200_CHUNK_CARRY) && (frame & 0xff) != ZB200_FRAME_GZIP_MEMBERS && level >= 1) ? 32768 : 0;
    const bool exact_fast = (frame & ZB200_EXACT_FAST) != 0;
    frame &= 0xff;
    if (!m || m->ctx.empty() || (!in && n) || !out || !out_len || chunk_size == 0 || frame < 0 || frame > 3) return ZB200_ERR_PARAM;
    const size_t S = chunk_size, g0 = m->ctx.size();
    const size_t nch = (n + S - 1) / S;
    const size_t g = nch < g0 ? (nch ? nch : 1) : g0;                  // no more pieces than chunks
    const size_t per = (nch + g - 1) / g * S;                         // bytes per piece: whole chunks
    const int pframe = frame == ZB200_FRAME_GZIP_MEMBERS ? ZB200_FRAME_GZIP_MEMBERS : ZB200_FRAME_RAW;
    uint8_t hdr[10];
    size_t hlen = 0;
    if (frame == ZB200_FRAME_ZLIB) {                                  // deflate.c:1004-1037
        const unsigned lf = (strategy >= 2 || level < 2) ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3;
        unsigned h = (0x78u << 8) | (lf << 6);
        h += 31 - h % 31;
        hdr[0] = (uint8_t)(h >> 8); hdr[1] = (uint8_t)h; hlen = 2;
    } else if (frame == ZB200_FRAME_GZIP) {                           // deflate.c:1042-1054
        const uint8_t gz[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(level == 9 ? 2 : (strategy >= 2 || level < 2) ? 4 : 0), 3};
        memcpy(hdr, gz, 10); hlen = 10;
    }
    std::vector<size_t> off(g), len(g), total(g, 0), final_off(g, 0);
    std::vector<uint32_t> pc(g, 0), pa(g, 1);
    std::vector<int> rc(g, ZB200_OK);
    Rendezvous sized(g);
    size_t need = 0;
    const size_t cap = *out_len;
    constexpr size_t kSub = 16;                                       // sub-pieces per GPU at most (ctx->d_pipe holds 16 records)
    on_all(g, [&](size_t k) {
        zb200_ctx *ctx = m->ctx[k];
        off[k] = k * per < n ? k * per : n;
        len[k] = n - off[k] < per ? n - off[k] : per;
        const bool last = k + 1 == g;
        // A GPU's piece goes through in sub-pieces of >= 128 MiB (whole chunks): sub-piece j+1 travels to the device
        // while sub-piece j is compressed.  The copies back wait for the rendezvous: a piece's place in the caller's
        // buffer depends on the sizes of all pieces before it.
        size_t sub = len[k] / 8 > ((size_t)128 << 20) ? len[k] / 8 : ((size_t)128 << 20);
        sub = deflate_piece_bytes(ctx, sub, S, 8, level);
        const size_t ns = len[k] ? (len[k] + sub - 1) / sub : 1;
        size_t sub_off[kSub + 1] = {0}, sub_total[kSub] = {0};
        cudaStream_t s = nullptr;
        {
            int r = cudaSetDevice(ctx->device) == cudaSuccess ? ZB200_OK : ZB200_ERR_CUDA;
            CtxUse lk(ctx, ctx->stream);
            size_t bound_total = 0;
            for (size_t j = 0; j < ns; ++j) {
                const size_t l = j + 1 < ns ? sub : len[k] - j * sub;
                sub_off[j] = bound_total;
                bound_total += zb200_deflate_bound(l, S, pframe);
            }
            const size_t H = off[k] < W ? off[k] : W;                          // history in front of this GPU's piece
            if (!r) r = ensure_io(ctx, H + len[k] + 16, bound_total + 16);
            if (!r) r = ensure_scratch(ctx, W ? zb200_deflate_scratch_bytes(((ns > 1 ? sub : len[k]) / S + 1) * (S + W), S + W)
                                              : zb200_deflate_scratch_bytes(ns > 1 ? sub : len[k], S));   // no reallocation (= implicit sync) mid-pipeline
            s = ctx->stream;
            const bool pinned = is_pinned(in);
            cudaEvent_t ev[kSub];
            size_t nev = 0;
            for (size_t j = 0; j < ns && !r; ++j) {
                const size_t l = j + 1 < ns ? sub : len[k] - j * sub;
                const size_t h = j ? 0 : H;                           // (the first sub-piece brings the history along)
                const uint8_t *src = (const uint8_t *)in + off[k] + j * sub - h;
                uint8_t *dst = ctx->d_io_in + H + j * sub - h;
                if (pinned && ns > 1) {
                    if (cudaEventCreateWithFlags(&ev[nev], cudaEventDisableTiming) != cudaSuccess) { r = ZB200_ERR_CUDA; break; }
                    ++nev;
                    if (cudaMemcpyAsync(dst, src, l + h, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess ||
                        cudaEventRecord(ev[j], ctx->copy_stream) != cudaSuccess) r = ZB200_ERR_CUDA;
                } else r = h2d_auto(ctx, dst, src, l + h, s);
            }
            for (size_t j = 0; j < ns && !r; ++j) {
                const size_t l = j + 1 < ns ? sub : len[k] - j * sub;
                if (pinned && ns > 1) cudaStreamWaitEvent(s, ev[j], 0);
                DeflateOpts o;
                o.level = level; o.strategy = strategy; o.carry = W != 0; o.exact_fast = exact_fast; o.skip = j ? (H + j * sub < W ? H + j * sub : W) : H;
                r = deflate_launch_opts(ctx, ctx->d_io_in + H + j * sub - o.skip, l + o.skip, S, o, pframe, (finish && last && j + 1 == ns) ? 1 : 0,
                                        ctx->d_io_out + sub_off[j], zb200_deflate_bound(l, S, pframe), nullptr, ctx->d_pipe + 2 * j,
                                        (uint32_t *)(ctx->d_pipe + 2 * j + 1), s);
                if (!r && cudaMemcpyAsync(ctx->h_pipe + 2 * j, ctx->d_pipe + 2 * j, 16, cudaMemcpyDeviceToHost, s) != cudaSuccess) r = ZB200_ERR_CUDA;
            }
            if (cudaStreamSynchronize(s) != cudaSuccess && !r) r = ZB200_ERR_CUDA;
            if (cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess && !r) r = ZB200_ERR_CUDA;
            for (size_t j = 0; j < nev; ++j) cudaEventDestroy(ev[j]);
            if (!r) {
                uint32_t c = 0, a = 1;
                for (size_t j = 0; j < ns; ++j) {
                    const size_t l = j + 1 < ns ? sub : len[k] - j * sub;
                    sub_total[j] = (size_t)ctx->h_pipe[2 * j];
                    const uint32_t *hs = (const uint32_t *)(ctx->h_pipe + 2 * j + 1);
                    c = j ? zb200_crc32_combine(c, hs[0], l) : hs[0];
                    a = j ? zb200_adler32_combine(a, hs[1], (int64_t)l) : hs[1];
                    total[k] += sub_total[j];
                }
                pc[k] = c; pa[k] = a;
            }
            rc[k] = r;
        }
        sized.arrive_and_wait();                                      // every piece knows its size: the exchange step
        bool ok = true;
        size_t at = hlen;
        for (size_t j = 0; j < g; ++j) { if (rc[j] != ZB200_OK) ok = false; if (j < k) at += total[j]; }
        final_off[k] = at;
        if (k + 1 == g) need = at + total[k] + ((finish && frame == ZB200_FRAME_ZLIB) ? 4 : (finish && frame == ZB200_FRAME_GZIP) ? 8 : 0);
        if (!ok || at + total[k] > cap) return;
        if (cudaSetDevice(ctx->device) != cudaSuccess) { rc[k] = ZB200_ERR_CUDA; return; }
        CtxUse lk(ctx, ctx->stream);
        int r = ZB200_OK;
        for (size_t j = 0; j < ns && !r; ++j) {
            if (sub_total[j]) r = d2h_auto(ctx, (uint8_t *)out + at, ctx->d_io_out + sub_off[j], sub_total[j], s);
            at += sub_total[j];
        }
        if (!r && cudaStreamSynchronize(s) != cudaSuccess) r = ZB200_ERR_CUDA;
        rc[k] = r;
    });
    for (size_t k = 0; k < g; ++k) if (rc[k] != ZB200_OK) return rc[k];
    uint32_t c = pc[0], a = pa[0];
    for (size_t k = 1; k < g; ++k) {
        c = zb200_crc32_combine(c, pc[k], len[k]);
        a = zb200_adler32_combine(a, pa[k], (int64_t)len[k]);
    }
    if (in_crc) *in_crc = c;
    if (in_adler) *in_adler = a;
    if (need > cap) { *out_len = need; set_error("deflate: %zu bytes do not fit the output buffer", need); return ZB200_ERR_OUTPUT; }
    memcpy(out, hdr, hlen);
    size_t pos = final_off[g - 1] + total[g - 1];
    if (finish && frame == ZB200_FRAME_ZLIB) for (int i = 0; i < 4; ++i) ((uint8_t *)out)[pos++] = (uint8_t)(a >> (24 - 8 * i));   // deflate.c:1254-1255
    if (finish && frame == ZB200_FRAME_GZIP) {                                                                                    // deflate.c:1241-1250
        for (int i = 0; i < 4; ++i) ((uint8_t *)out)[pos++] = (uint8_t)(c >> (8 * i));
        for (int i = 0; i < 4; ++i) ((uint8_t *)out)[pos++] = (uint8_t)((uint32_t)n >> (8 * i));
    }
    *out_len = pos;
    return ZB200_OK;
}

int zb200_multi_inflate_host(zb200_multi *m, const void *in, void *out, const zb200_member *members, size_t n_members,
                             int wrap, int verify, zb200_member_result *results) {
    if (!m || m->ctx.empty() || !members || !results) return ZB200_ERR_PARAM;
    if (n_members == 0) return ZB200_OK;
    const size_t g0 = m->ctx.size(), g = n_members < g0 ? n_members : g0;
    // contiguous member ranges of about equal output capacity
    uint64_t total = 0;
    for (size_t i = 0; i < n_members; ++i) total += members[i].out_cap + 1;
    std::vector<size_t> lo(g + 1, n_members);
    lo[0] = 0;
    {
        uint64_t run = 0;
        size_t k = 1;
        for (size_t i = 0; i < n_members && k < g; ++i) {
            run += members[i].out_cap + 1;
            if (run >= total * k / g) lo[k++] = i + 1;
        }
    }
    std::vector<int> rc(g, ZB200_OK);
    on_all(g, [&](size_t k) {
        const size_t a = lo[k], b = lo[k + 1];
        if (a >= b) return;
        // the range's own window of the two buffers (the single-GPU entry moves [0, max offset) otherwise)
        uint64_t in_lo = ~0ull, out_lo = ~0ull;
        for (size_t i = a; i < b; ++i) {
            if (members[i].in_off < in_lo) in_lo = members[i].in_off;
            const uint64_t o = members[i].out_off - members[i].dict_len;
            if (o < out_lo) out_lo = o;
        }
        std::vector<zb200_member> tab(members + a, members + b);
        for (auto &t : tab) { t.in_off -= in_lo; t.out_off -= out_lo; }
        rc[k] = zb200_inflate_host(m->ctx[k], (const uint8_t *)in + in_lo, (uint8_t *)out + out_lo, tab.data(), tab.size(), wrap, verify, results + a);
    });
    for (size_t k = 0; k < g; ++k) if (rc[k] != ZB200_OK) return rc[k];
    return ZB200_OK;
}

}  // extern "C"
