// Host replay of the inflate state machine (TEST ONLY): the same
// zb_inflate.cuh the kernel runs on lane 0, with the warp-wide copy events
// executed by plain loops.  Never linked into libzb200.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../../zlib_wasm_b200/csrc/zb_inflate.cuh"

using namespace zb;

static FormatTables g_fmt;
static uint32_t g_flit[512], g_fdist[64];
static int g_ready;

static void setup() {
    format_fill(g_fmt);
    uint8_t lens[288]; uint16_t work[320];
    int i = 0;
    for (; i < 144; ++i) lens[i] = 8;
    for (; i < 256; ++i) lens[i] = 9;
    for (; i < 280; ++i) lens[i] = 7;
    for (; i < 288; ++i) lens[i] = 8;
    build_decode_table(TBL_LITLEN, lens, 288, g_flit, 512, kLitRoot, work, g_fmt);
    for (i = 0; i < 32; ++i) lens[i] = 5;
    build_decode_table(TBL_DIST, lens, 32, g_fdist, 64, kDistRoot, work, g_fmt);
    g_ready = 1;
}

extern "C" int emul_inflate(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, int wrap,
                            uint64_t resume_bit, uint64_t resume_out,
                            uint64_t *in_used, uint64_t *out_len, uint32_t *check, uint32_t *isize,
                            uint64_t *ck_bit, uint64_t *ck_out, int *kind) {
    if (!g_ready) setup();
    static InflateScratch scr;
    InflateState st;
    st.init(src, n, dst, cap, &scr, g_flit, g_fdist, &g_fmt);
    int hs = ZB200_INF_OK;
    if (resume_bit) st.resume(resume_bit, resume_out, wrap);
    else hs = st.parse_header(wrap);
    if (hs == ZB200_INF_OK) {
        static QueuedMatch q[kQueue];
        auto copy = [&](const QueuedMatch &m) {
            const uint32_t dist = qm_dist(m.packed), len = qm_len(m.packed);
            for (uint32_t i = 0; i < len; ++i) dst[m.dst + i] = dst[m.dst - dist + (dist >= len ? i : i % dist)];
        };
        for (;;) {
            InflateEvent ev = st.run_batch(q);
            if (ev.kind == EV_DONE) break;
            if (ev.kind == EV_BATCH) {
                // the device runs independent matches concurrently: replay them in REVERSE
                // order (any order must give the same bytes), then the dependent ones in order
                for (int i = (int)ev.len - 1; i >= 0; --i) if (!qm_dep(q[i].packed)) copy(q[i]);
                for (uint32_t i = 0; i < ev.len; ++i) if (qm_dep(q[i].packed)) copy(q[i]);
            } else {
                memcpy(dst + ev.dst, src + ev.src, ev.len);
            }
        }
    } else st.status = hs;
    *in_used = st.in_used; *out_len = st.pos; *check = st.stored_check; *isize = st.stored_isize;
    *ck_bit = st.ck_bit; *ck_out = st.ck_out; *kind = st.wrap_kind;
    return st.status;
}

extern "C" const char *emul_msg(int s);
