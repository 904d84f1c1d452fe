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

// ---- one member decoded chunk by chunk (zb_inflate_blocks.cuh) replayed on the host ------------------------------
// The product's header test over every bit position, the product's state machine in its two chunk modes (count /
// list; the serial path stands in for the rounds, which are replayed above), then the chain, the source pointers, the
// pointer jumping and the gather restated with plain loops as zb_inflate.cu inflate_stream_blocks drives them.
#include "../../zlib_wasm_b200/csrc/zb_inflate_blocks.cuh"
#include <algorithm>

static std::vector<uint32_t> words_of(const uint8_t *src, uint64_t n) {
    std::vector<uint32_t> w((n + 3) / 4 + 4, 0u);
    memcpy(w.data(), src, n);
    return w;
}

extern "C" uint64_t emul_blk_candidates(const uint8_t *src, uint64_t n, uint64_t bit_lo, uint64_t *list, uint64_t cap, uint64_t *quick_pass) {
    if (!g_ready) setup();
    const std::vector<uint32_t> w = words_of(src, n);
    const uint64_t nwords = (n + 3) / 4, bit_hi = n * 8;
    uint64_t cnt = 0, qp = 0;
    uint8_t tab[128];
    for (uint64_t b = bit_lo; b + 20 <= bit_hi; ++b) {
        const uint64_t x = blk_bits64(w.data(), nwords, b);
        if (!blk_quick((uint32_t)x)) continue;
        if (!blk_cl_complete(blk_bits64(w.data(), nwords, b + 17), (((uint32_t)x >> 13) & 15u) + 4u)) continue;
        ++qp;
        uint64_t end;
        if (!blk_header_valid(w.data(), nwords, bit_hi, b, tab, g_fmt.cl_order, &end)) continue;
        if (cnt < cap) list[cnt] = b;
        ++cnt;
    }
    for (uint64_t P = 4; P + 4 <= n; ++P) {              // stored blocks, as blk_scan_kernel looks for them
        const uint32_t len = src[P] | (src[P + 1] << 8), nlen = src[P + 2] | (src[P + 3] << 8);
        if ((len ^ nlen) != 0xffffu || P * 8 < bit_lo + 3) continue;
        const uint32_t before = (uint32_t)src[P - 2] | ((uint32_t)src[P - 1] << 8);
        const uint32_t ns = blk_stored_starts(before, P * 8, bit_lo);
        for (uint32_t k = 0; k < ns; ++k, ++cnt) if (cnt < cap) list[cnt] = P * 8 - 3 - k;
    }
    if (cnt <= cap) std::sort(list, list + cnt);
    if (quick_pass) *quick_pass = qp;
    return cnt;
}

struct ChunkOut { int status; uint64_t out_len, resume_bit, resume_out, nm, in_used; uint32_t check, isize; };
// mode 1: count, mode 2: list (literals stored into dst, matches appended to ml)
static ChunkOut host_chunk(const uint8_t *src, uint64_t n, int kind, uint64_t start_bit, uint64_t out_pos, const uint64_t *cand, uint32_t nc,
                           int mode, uint8_t *dst, uint64_t cap, std::vector<QueuedMatch> *ml) {
    static InflateScratch scr;
    InflateState st;
    st.init(src, n, dst, cap, &scr, g_flit, g_fdist, &g_fmt);
    st.cand = cand; st.cand_n = nc; st.count_only = mode == 1;
    int hs = ZB200_INF_OK;
    if (start_bit) st.resume(start_bit, out_pos, kind);
    else { hs = st.parse_header(kind); st.start_bit = st.bitpos(); }
    ChunkOut o;
    memset(&o, 0, sizeof o);
    if (hs == ZB200_INF_OK) {
        static QueuedMatch q[kQueue];
        for (;;) {
            InflateEvent ev = st.run_batch(q);
            if (ev.kind == EV_DONE) break;
            if (ev.kind == EV_BATCH) {
                if (mode == 2) for (uint32_t i = 0; i < ev.len; ++i) ml->push_back(q[i]);
                o.nm += ev.len;
            } else if (mode == 2) memcpy(dst + ev.dst, src + ev.src, ev.len);
        }
    } else st.status = hs;
    o.status = st.status; o.out_len = st.pos; o.resume_bit = st.ck_bit; o.resume_out = st.ck_out; o.in_used = st.in_used;
    o.check = st.stored_check; o.isize = st.stored_isize;
    return o;
}

// Returns the number of chunks on the chain (0: not applicable), -1 on an inconsistency between the two decodes.
// stats: [0] candidates, [1] chunks on the chain, [2] jump passes, [3] matches
extern "C" int emul_inflate_blocks(const uint8_t *src, uint64_t n, int wrap, uint8_t *dst, uint64_t cap,
                                   uint64_t *out_len, int *status, uint64_t *in_used, uint32_t *check, uint32_t *isize, uint64_t *stats) {
    if (!g_ready) setup();
    InflateState hs;
    hs.init(src, n, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    if (hs.parse_header(wrap) != ZB200_INF_OK) return 0;
    const int kind = hs.wrap_kind;
    const uint64_t bit0 = hs.next * 8;
    std::vector<uint64_t> cand(n / 8 + 64);
    const uint64_t nc64 = emul_blk_candidates(src, n, bit0, cand.data(), cand.size(), nullptr);
    if (nc64 > cand.size()) return 0;
    cand.resize(nc64);
    const uint32_t nc = (uint32_t)nc64;
    stats[0] = nc;
    if (nc < 2) return 0;
    std::vector<ChunkOut> res(nc + 1);
    for (uint32_t k = 0; k <= nc; ++k)
        res[k] = host_chunk(src, n, kind, k ? cand[k - 1] : 0, 0, cand.data(), nc, 1, nullptr, 0xfffffff0ull, nullptr);
    struct Link { uint32_t idx; uint64_t out_off; };
    std::vector<Link> chain;
    uint64_t total = 0;
    int fin = ZB200_INF_TRUNCATED;
    for (uint32_t cur = 0;;) {
        const ChunkOut &q = res[cur];
        const uint64_t begin = cur ? cand[cur - 1] : bit0;
        if (q.status == ZB200_INF_OK) { chain.push_back({cur, total}); total += q.out_len; fin = ZB200_INF_OK; break; }
        if (q.status != ZB200_INF_TRUNCATED || q.resume_out != q.out_len || q.resume_bit <= begin) break;
        const auto it = std::lower_bound(cand.begin(), cand.end(), q.resume_bit);
        if (it == cand.end() || *it != q.resume_bit) break;
        chain.push_back({cur, total}); total += q.out_len;
        cur = (uint32_t)(it - cand.begin()) + 1;
    }
    stats[1] = chain.size();
    if (chain.size() < 2 || fin != ZB200_INF_OK) return 0;
    *out_len = total; *status = fin;
    if (total > cap) { *status = ZB200_INF_OUTPUT_FULL; return (int)chain.size(); }
    std::vector<QueuedMatch> ml;
    for (size_t k = 0; k < chain.size(); ++k) {
        const uint32_t i = chain[k].idx;
        const ChunkOut b = host_chunk(src, n, kind, i ? cand[i - 1] : 0, chain[k].out_off, cand.data(), nc, 2, dst, total, &ml);
        const ChunkOut &a = res[i];
        if (a.status != b.status || b.out_len != chain[k].out_off + a.out_len || a.resume_bit != b.resume_bit) return -1;
        if (k + 1 == chain.size()) { *in_used = b.in_used; *check = b.check; *isize = b.isize; }
    }
    stats[3] = ml.size();
    std::vector<uint32_t> sp(total, kSrcLiteral);
    for (const QueuedMatch &m : ml) {
        const uint32_t len = qm_len(m.packed), dist = qm_dist(m.packed);
        for (uint32_t i = 0; i < len; ++i) sp[m.dst + i] = m.dst - dist + (i < dist ? i : i % dist);
    }
    uint64_t passes = 0;
    for (bool any = true; any; ++passes) {
        any = false;
        for (uint64_t p = total; p-- > 0;) {                 // (backwards: the least favourable order for in-place updates)
            const uint32_t s = sp[p];
            if (s == kSrcLiteral) continue;
            const uint32_t t = sp[s];
            if (t != kSrcLiteral) { sp[p] = t; any = true; }
        }
    }
    stats[2] = passes;
    for (uint64_t p = 0; p < total; ++p) if (sp[p] != kSrcLiteral) dst[p] = dst[sp[p]];
    return (int)chain.size();
}
