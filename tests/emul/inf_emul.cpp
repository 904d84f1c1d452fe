// Host replay of the inflate state machine (TEST ONLY): the same
// zb_inflate.cuh the kernel runs on lane 0, with the warp-wide copy events
// executed by plain loops.  Never linked into libzb200.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "../../zlib_wasm_b200/csrc/zb_inflate.cuh"
#include "../../zlib_wasm_b200/csrc/zb_inflate_round.cuh"
#include <vector>

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


// ---- the warp-parallel rounds (zb_inflate_round.cuh) replayed with loops over the 32 lanes.
// The per-lane phases are the product's; only the cross-lane glue (shuffles, ballots,
// scans) of zb_inflate.cu huff_rounds_warp is restated here.
struct RoundStats { uint64_t rounds, fix_passes, fix_lane_runs, serial_returns, matches, dep_matches, copy_passes; };
static RoundStats g_stats;
extern "C" void emul_round_stats(uint64_t *o) { memcpy(o, &g_stats, sizeof g_stats); }

template <int NL>
static int host_rounds(const uint8_t *src, uint64_t in_len, uint8_t *dst, uint64_t out_cap, uint64_t &bitpos, uint64_t &pos,
                       const uint32_t *lt, const uint32_t *dt, int force_lg) {
    static RoundSharedT<NL> rs;
    static std::vector<QueuedMatch> gq(kRoundQueueCap * (NL / 32));
    const uint32_t bias = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    const uint64_t total_bits = bias + in_len * 8, nwords = (total_bits + 31) >> 5;
    int lg_cap = kRoundLgMax;
    for (;;) {
        const uint64_t B = bias + bitpos;
        if (B >= total_bits) { g_stats.serial_returns++; return 1; }
        int lg = round_pick_lg(total_bits - B, NL);
        if (lg < 0) { g_stats.serial_returns++; return 1; }
        if (force_lg >= kRoundLgMin && force_lg < lg) lg = force_lg;
        if (lg > lg_cap) lg = lg_cap;                   // an overfull round shortens the subsequences for the rest of the block
        const uint32_t S = 32u << lg, W = 1u << lg, stride = stage_row_stride(lg);
        const uint64_t W0 = B >> 5;
        uint32_t *stage = rs.stage;
        for (uint32_t k = 0; k < NL * stride; ++k) {
            const uint32_t row = k / stride, col = k - row * stride;
            const uint64_t w = W0 + row * W + col;
            uint32_t v = 0;
            if (w < nwords) {                           // byte-wise: the host buffer has no alignment slack to read whole words from
                for (int b = 0; b < 4; ++b) {
                    const uint64_t byte = w * 4 + b;    // relative to the aligned base
                    const uint64_t lo = bias / 8, hi = lo + in_len;
                    if (byte >= lo && byte < hi) v |= (uint32_t)src[byte - lo] << (8 * b);
                }
            }
            stage[k] = v;
        }
        RoundLane r[NL];
        for (uint32_t l = 0; l < NL; ++l) round_speculate(r[l], l, lg, l ? 0u : (uint32_t)(B & 31u), stage, rs, lt, dt);
        g_stats.rounds++;
        for (;;) {
            uint32_t nvalid = NL;
            for (uint32_t l = 0; l < NL; ++l) if (r[l].stop != STOP_NONE) { nvalid = l + 1; break; }
            uint32_t t[NL]; bool need[NL]; bool any = false;
            for (uint32_t l = 0; l < NL; ++l) {
                t[l] = l ? r[l - 1].end - S : 0;
                need[l] = l > 0 && l < nvalid && t[l] != r[l].start;
                any |= need[l];
            }
            if (!any) break;
            g_stats.fix_passes++;
            for (uint32_t l = 0; l < NL; ++l) if (need[l]) { round_fix(r[l], l, lg, t[l], stage, rs, lt, dt); g_stats.fix_lane_runs++; }
        }
        uint32_t last = NL - 1;
        for (uint32_t l = 0; l < NL; ++l) if (r[l].stop != STOP_NONE) { last = l; break; }
        const uint64_t end_abs = W0 * 32 + (uint64_t)last * S + r[last].end;
        if (r[last].stop == STOP_BAD) { g_stats.serial_returns++; return 1; }
        if (end_abs > total_bits) { g_stats.serial_returns++; return 1; }
        uint32_t off_o[NL + 1], off_m[NL + 1];
        off_o[0] = off_m[0] = 0;
        for (uint32_t l = 0; l < NL; ++l) {
            off_o[l + 1] = off_o[l] + (l <= last ? r[l].out : 0);
            off_m[l + 1] = off_m[l] + (l <= last ? r[l].m : 0);
        }
        if (off_o[NL] > out_cap - pos) { g_stats.serial_returns++; return 1; }
        if (off_m[NL] > kRoundQueueCap * (NL / 32)) { lg_cap = lg >= kRoundLgMin + 2 ? lg - 2 : kRoundLgMin; continue; }   // as on the device
        bool err = false;
        for (int l = (int)last; l >= 0; --l)             // any lane order must do: run them backwards
            err |= round_emit(r[l], (uint32_t)l, lg, stage, lt, dt, dst, (uint32_t)pos + off_o[l], gq.data(), off_m[l]) != 0;
        if (err) { g_stats.serial_returns++; return 1; }
        // P4 as on the device: waves of NL matches; those whose source ends before the wave's first destination
        // at once, the others in passes — whoever waits for no unfinished wave-mate copies (any order within a pass
        // must do: run them backwards).  A lane's own copy = 16-byte load-then-store steps where the device uses them.
        const uint32_t count = off_m[NL];
        auto copy = [&](const QueuedMatch &m) {
            const uint32_t dist = qm_dist(m.packed), len = qm_len(m.packed);
            if (dist >= len || dist >= 16) {
                for (uint32_t i = 0; i < len; i += 16) {
                    uint8_t t[16];
                    const uint32_t n = len - i < 16 ? len - i : 16;
                    memcpy(t, dst + m.dst - dist + i, n);
                    memcpy(dst + m.dst + i, t, n);
                }
            } else
                for (uint32_t i = 0; i < len; ++i) dst[m.dst + i] = dst[m.dst - dist + i % dist];
        };
        for (uint32_t base = 0; base < count; base += NL) {
            const uint32_t n = count - base < NL ? count - base : NL;
            const uint32_t first = gq[base].dst;
            bool undone[NL];
            std::vector<uint32_t> waits[NL];
            uint32_t left = 0;
            for (uint32_t k = 0; k < n; ++k) {
                const QueuedMatch &m = gq[base + k];
                const uint32_t dist = qm_dist(m.packed), len = qm_len(m.packed);
                undone[k] = m.dst - dist + (len < dist ? len : dist) > first;
                g_stats.matches++; g_stats.dep_matches += undone[k];
                left += undone[k];
            }
            for (int k = (int)n - 1; k >= 0; --k) if (!undone[k]) copy(gq[base + k]);
            for (uint32_t k = 0; k < n; ++k) if (undone[k]) {
                const QueuedMatch &m = gq[base + k];
                const uint32_t dist = qm_dist(m.packed), len = qm_len(m.packed);
                const uint32_t srcb = m.dst - dist, srce = srcb + (len < dist ? len : dist);
                for (uint32_t j = 0; j < k; ++j)
                    if (undone[j] && gq[base + j].dst < srce && gq[base + j].dst + qm_len(gq[base + j].packed) > srcb) waits[k].push_back(j);
            }
            while (left) {
                bool ready[NL];
                for (uint32_t k = 0; k < n; ++k) {
                    ready[k] = undone[k];
                    for (uint32_t j : waits[k]) if (undone[j]) ready[k] = false;
                }
                for (int k = (int)n - 1; k >= 0; --k) if (ready[k]) { copy(gq[base + k]); undone[k] = false; --left; }
                g_stats.copy_passes++;
            }
        }
        pos += off_o[NL];
        bitpos = end_abs - bias;
        if (r[last].stop == STOP_EOB) return 0;
    }
}

extern "C" int emul_inflate_rounds(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, int wrap, int force_lg_and_lanes,
                                   uint64_t *in_used, uint64_t *out_len, uint32_t *check, uint32_t *isize,
                                   uint64_t *ck_bit, uint64_t *ck_out, int *kind) {
    if (!g_ready) setup();
    static InflateScratch scr;
    InflateState st;
    st.init(src, n, dst, cap < 0xfffffff0ull ? cap : 0xfffffff0ull, &scr, g_flit, g_fdist, &g_fmt);
    st.huff_external = 1;
    const int lanes = force_lg_and_lanes >= 1000 ? 128 : 32, force_lg = force_lg_and_lanes >= 1000 ? force_lg_and_lanes - 1000 - 1 : force_lg_and_lanes;
    int hs = st.parse_header(wrap);
    if (hs == ZB200_INF_OK) {
        static QueuedMatch q[kQueue];
        auto copy = [&](const QueuedMatch &m) {
            const uint32_t dist = qm_dist(m.packed), len = qm_len(m.packed);
            for (uint32_t i = 0; i < len; ++i) dst[m.dst + i] = dst[m.dst - dist + (dist >= len ? i : i % dist)];
        };
        for (;;) {
            InflateEvent ev = st.run_batch(q);
            if (ev.kind == EV_DONE) break;
            if (ev.kind == EV_HUFF) {
                uint64_t bp = ev.src, op = ev.dst;
                const uint32_t *lt = ev.len ? g_flit : scr.lit, *dt = ev.len ? g_fdist : scr.dist;
                int r = lanes == 128 ? host_rounds<128>(src, n, dst, cap < 0xfffffff0ull ? cap : 0xfffffff0ull, bp, op, lt, dt, force_lg) : 1;
                if (r == 1) r = host_rounds<32>(src, n, dst, cap < 0xfffffff0ull ? cap : 0xfffffff0ull, bp, op, lt, dt, force_lg);   // the team hands tails to one warp
                if (r < 0) return -100;
                st.seek(bp, op, r);
            } else if (ev.kind == EV_BATCH) {
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
