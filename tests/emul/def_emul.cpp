// Host replay of the deflate pipeline phases (TEST ONLY): the same
// zb_deflate.cuh cores the kernels run, driven by plain loops.  Never linked
// into libzb200.so.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "../../zlib_wasm_b200/csrc/zb_deflate.cuh"

using namespace zb;

static FormatTables g_fmt;
static StaticTrees g_st;
static int g_ready;

struct BitSink {
    uint8_t *out; size_t cap; uint64_t bit;
    void put(uint64_t v, uint32_t nb) {
        for (uint32_t i = 0; i < nb; ++i, ++bit)
            if ((v >> i) & 1) { if ((bit >> 3) < cap) out[bit >> 3] |= (uint8_t)(1u << (bit & 7)); }
    }
};

// ONE long chunk worked on by many CTAs (zb_deflate.cu: chain kernel in ranges, dfl_parse_multi_kernel): emul_set_multi(G,
// range) makes the replay below hand the lazy parse from CTA to CTA over G x kSegLanes segments (cold settle, provisional
// link, settle against the predecessor's provisional end, compare with its true end, count link) and build the chain
// links range by range behind w_size re-inserted positions; the matches and the parse must still be the serial ones.
static uint32_t g_multi_G = 0, g_multi_range = 0;
extern "C" void emul_set_multi(uint32_t G, uint32_t range) { g_multi_G = G; g_multi_range = range; }

// deflate_fast with the reference's own chains (zb_deflate.cuh fast_exact_chunk: the opt-in exact form of levels 1-3)
static int g_exact_fast = 0;
extern "C" void emul_set_exact_fast(int on) { g_exact_fast = on; }

// One chunk -> raw deflate bytes (blocks + marker when !final).  Returns size or -1.
// `skip`: the first skip bytes are a preset dictionary (history only).
// window_bits / mem_level: deflateInit2_'s (deflate.c:440-455).
extern "C" long emul_deflate_chunk_opts(const uint8_t *data, uint32_t n, uint32_t skip, int level, int strategy, int window_bits,
                                        int mem_level, int final_chunk, uint8_t *out, size_t cap, uint32_t *stats /* nsyms, nblocks */) {
    if (!g_ready) { format_fill(g_fmt); static_trees_fill(g_st); g_ready = 1; }
    const DeflateParams prm = deflate_params(level, strategy, window_bits, mem_level);
    std::vector<uint16_t> prev(n + 1, 0);
    std::vector<uint32_t> mf(n + 1, 0), mq(n + 1, 0);
    const bool exact_fast = g_exact_fast && prm.mode == MODE_FAST && skip == 0;
    if (exact_fast) {
    } else if (prm.mode == MODE_FAST || prm.mode == MODE_SLOW) {
        std::vector<int32_t> head(1u << prm.hash_bits, -1);
        for (uint32_t p = 0; p + kMinMatch <= n; ++p) {
            const uint32_t h = hash3(data + p, prm);
            const int32_t q = head[h];
            prev[p] = (q > 0 && p - (uint32_t)q <= 65535u) ? (uint16_t)(p - (uint32_t)q) : 0;
            head[h] = (int32_t)p;
        }
        if (g_multi_range) {                                   // links as the ranged chain kernel draws them: the matches must not change
            std::vector<uint16_t> prev2(n + 1, 0);
            for (uint32_t lo_r = 0; lo_r < n; lo_r += g_multi_range) {
                const uint32_t hi_r = n - lo_r < g_multi_range ? n : lo_r + g_multi_range, warm = lo_r > prm.w_size ? lo_r - prm.w_size : 0;
                std::fill(head.begin(), head.end(), -1);
                for (uint32_t p = warm; p < hi_r && p + kMinMatch <= n; ++p) {
                    const uint32_t h = hash3(data + p, prm);
                    const int32_t q = head[h];
                    if (p >= lo_r) prev2[p] = (q > 0 && p - (uint32_t)q <= 65535u) ? (uint16_t)(p - (uint32_t)q) : 0;
                    head[h] = (int32_t)p;
                }
            }
            for (uint32_t p = 0; p < n; ++p) {
                const MatchPair a = match_at(data, n, prev.data(), p, prm), b = match_at(data, n, prev2.data(), p, prm);
                if (a.full != b.full || a.quarter != b.quarter) return -11;
            }
            prev.swap(prev2);
        }
        for (uint32_t p = 0; p < n; ++p) {
            MatchPair r = match_at(data, n, prev.data(), p, prm); mf[p] = r.full; mq[p] = r.quarter;
            if (prm.mode == MODE_FAST && prm.level <= 2 && p + kUniformTail <= n) {   // the kernel's branch-free walk must agree
                const PlainWin pw{data, prev.data()};
                const uint32_t u = prm.level == 1 ? match_uniform<4, 8>(pw, p, prm.max_dist) : match_uniform<8, 16>(pw, p, prm.max_dist);
                if (u != r.full) return -10;
            }
        }
    } else if (prm.mode == MODE_RLE) {
        for (uint32_t p = 0; p < n; ++p) mf[p] = rle_at(data, n, p);
    }
    std::vector<uint32_t> syms(n + 2);
    std::vector<BlockInfo> blocks(max_blocks_for(n, prm.sym_limit));
    struct { uint32_t nsyms, nblocks; } sink;
    // the kernel's segmented parse (zb_deflate.cuh seg_*), the lanes replayed by loops;
    // the serial whole-chunk parse must give the same symbols and blocks
    if (exact_fast) {
        std::vector<uint32_t> head(1u << prm.hash_bits, 0);
        fast_exact_chunk(data, n, prm, final_chunk != 0, head.data(), prev.data(), syms.data(), blocks.data(), sink.nsyms, sink.nblocks);
    } else if (prm.mode != MODE_SLOW) {
        // greedy rules: exit tables per tile (zb_deflate.cu dfl_parse_greedy_kernel), the lanes replayed by loops
        const bool use_m = prm.mode != MODE_HUFF;
        std::vector<uint32_t> mfv(kGtSlots); std::vector<uint16_t> lc(kGtSlots);
        uint32_t entry = skip, nsyms = 0;
        struct GA { const uint32_t *mfv; const uint8_t *data; uint32_t t0;
                    uint32_t mf(uint32_t p) const { return mfv[gt_slot(p - t0)]; } uint32_t byte(uint32_t p) const { return data[p]; } };
        for (uint32_t t0 = skip / kGtTile * kGtTile; t0 < n; t0 += kGtTile) {
            for (uint32_t r = 0; r < kGtTile; ++r) mfv[gt_slot(r)] = (use_m && t0 + r < n) ? mf[t0 + r] : 0u;
            for (int l = 31; l >= 0; --l) gt_fill((uint32_t)l, n - t0, mfv.data(), lc.data());
            uint32_t exit_rel = 0, total = 0;
            for (int l = 31; l >= 0; --l) {
                uint32_t my_entry, my_first;
                exit_rel = gt_hop((uint32_t)l, entry - t0, n - t0, lc.data(), my_entry, my_first, total);
                if (my_entry == 0xffffffffu) continue;
                GA acc{mfv.data(), data, t0};
                const uint32_t seg_end = ((uint32_t)l + 1) * kGtSeg, lim = n - t0;
                uint32_t r = my_entry, g = nsyms + my_first;
                while (r < seg_end && r < lim) {
                    const uint32_t m = mfv[gt_slot(r)];
                    syms[g] = greedy_symbol(t0 + r, use_m, acc, g, blocks.data(), n, prm);
                    ++g; r += m ? (m >> 16) : 1u;
                }
            }
            entry = t0 + exit_rel; nsyms += total;
        }
        sink.nsyms = nsyms; sink.nblocks = seg_finish(blocks.data(), nsyms, false, n, prm, final_chunk != 0, skip);
    } else if (g_multi_G >= 2) {
        const uint32_t G = g_multi_G, m = n - skip;
        struct WinAcc { const uint8_t *data; const uint32_t *mfull, *mquarter; uint32_t *out; uint32_t at;
                        uint32_t mf(uint32_t p) const { return mfull[p]; } uint32_t mq(uint32_t p) const { return mquarter[p]; }
                        uint32_t byte(uint32_t p) const { return data[p]; } void put(uint32_t sym) { if (out) out[at++] = sym; }
                        uint32_t windows(uint32_t) const { return 1; } uint32_t open(uint32_t, uint32_t) { return 0xffffffffu; } bool any(bool b) const { return b; } };
        WinAcc cacc{data, mf.data(), mq.data(), nullptr, 0};
        std::vector<SegGeom> geo(G);
        std::vector<std::vector<SegRec>> rec(G, std::vector<SegRec>((kSegRecs - 1) * kSegLanes));
        std::vector<std::vector<SegLane>> lanes(G, std::vector<SegLane>(kSegLanes));
        std::vector<SegState> prov(G), fin(G), pred(G);
        std::vector<uint32_t> total(G, 0), before(G, 0);
        auto settle = [&](uint32_t k) -> int {
            SegGeom &g = geo[k];
            for (int pass = 0;; ++pass) {
                if (pass > (int)kSegLanes + 8) return -20;
                SegState t[kSegLanes]; bool need[kSegLanes], any = false;
                for (uint32_t l = 0; l < g.nact; ++l) {
                    t[l] = l ? lanes[k][l - 1].end : pred[k];
                    need[l] = t[l].p != lanes[k][l].start.p || t[l].w0 != lanes[k][l].start.w0;
                    any |= need[l];
                }
                if (!any) return 0;
                for (uint32_t l = 0; l < g.nact; ++l) if (need[l]) seg_fix(lanes[k][l], l, g, n, prm, cacc, rec[k].data(), t[l]);
            }
        };
        for (uint32_t k = 0; k < G; ++k) {                        // every CTA on its own: speculate, settle from a cold start
            SegGeom &g = geo[k];
            const uint64_t all = (uint64_t)G * kSegLanes * kSegRecs;
            g.blk = (uint32_t)((m + all - 1) / all); if (g.blk < 32) g.blk = 32;
            g.seg = g.blk * kSegRecs;
            const uint32_t nact_total = m ? (m + g.seg - 1) / g.seg : 1;
            const uint64_t first = (uint64_t)k * kSegLanes;
            g.nact = nact_total > first ? (nact_total - first < kSegLanes ? (uint32_t)(nact_total - first) : kSegLanes) : 0u;
            g.lo = g.nact ? skip + (uint32_t)first * g.seg : n;
            for (uint32_t l = 0; l < g.nact; ++l) seg_speculate(lanes[k][l], l, g, n, prm, cacc, rec[k].data());
            pred[k] = seg_cold(g.lo);
            if (settle(k)) return -20;
            prov[k] = g.nact ? lanes[k][g.nact - 1].end : pred[k];
        }
        uint32_t refixed = 0;
        for (uint32_t k = 0; k < G; ++k) {                        // the chain of links, in ticket order
            if (k) {
                pred[k] = prov[k - 1];
                if (settle(k)) return -20;
                if (fin[k - 1].p != pred[k].p || fin[k - 1].w0 != pred[k].w0) { pred[k] = fin[k - 1]; ++refixed; if (settle(k)) return -20; }
            }
            fin[k] = geo[k].nact ? lanes[k][geo[k].nact - 1].end : pred[k];
            for (uint32_t l = 0; l < geo[k].nact; ++l) total[k] += lanes[k][l].count;
            before[k] = k ? before[k - 1] + total[k - 1] : 0;
        }
        for (int k = (int)G - 1; k >= 0; --k) {                   // any CTA order must do for the emit
            uint32_t first = before[k];
            for (uint32_t l = 0; l < geo[k].nact; ++l) {
                WinAcc eacc{data, mf.data(), mq.data(), syms.data(), first};
                seg_emit(lanes[k][l], l, geo[k], n, prm, eacc, blocks.data(), first);
                first += lanes[k][l].count;
                if (eacc.at != first) return -21;
            }
        }
        const uint32_t all_syms = before[G - 1] + total[G - 1];
        const bool pending = (fin[G - 1].w0 >> 25) & 1u;
        if (pending) syms[all_syms] = data[n - 1];
        sink.nsyms = all_syms + (pending ? 1 : 0);
        sink.nblocks = seg_finish(blocks.data(), all_syms, pending, n, prm, final_chunk != 0, skip);
        if (stats) stats[2] = refixed;
    } else {
        const SegGeom g = seg_geometry(n, skip);
        std::vector<SegRec> rec((kSegRecs - 1) * kSegLanes);
        SegLane r[kSegLanes];
        // operands paced in windows of 16 positions at an arbitrary alignment, as the kernel stages them
        struct WinAcc { const uint8_t *data; const uint32_t *mfull, *mquarter; uint32_t *out; uint32_t at, skew, lo, hi;
                        uint32_t mf(uint32_t p) const { if (p < lo || p >= hi) abort(); return mfull[p]; }
                        uint32_t mq(uint32_t p) const { if (p < lo || p >= hi) abort(); return mquarter[p]; }
                        uint32_t byte(uint32_t p) const { if (p + 1 < lo || p >= hi) abort(); return data[p]; }
                        void put(uint32_t sym) { if (out) out[at++] = sym; }
                        uint32_t windows(uint32_t seg) const { return seg / 16 + 2; }
                        uint32_t open(uint32_t w, uint32_t s0) { const uint32_t a = (s0 + skew) / 16 * 16 + w * 16; lo = a > skew ? a - skew : 0; hi = a + 16 - skew; return hi; }
                        bool any(bool b) const { return b; } };
        WinAcc cacc{data, mf.data(), mq.data(), nullptr, 0, (uint32_t)(n % 13), 0, 0};
        for (uint32_t l = 0; l < g.nact; ++l) seg_speculate(r[l], l, g, n, prm, cacc, rec.data());
        for (int pass = 0;; ++pass) {
            if (pass > (int)kSegLanes + 8) return -20;
            SegState t[kSegLanes]; bool need[kSegLanes], any = false;
            for (uint32_t l = 0; l < g.nact; ++l) {
                if (l) t[l] = r[l - 1].end;
                need[l] = l > 0 && (t[l].p != r[l].start.p || t[l].w0 != r[l].start.w0);
                any |= need[l];
            }
            if (!any) break;
            for (uint32_t l = 0; l < g.nact; ++l) if (need[l]) seg_fix(r[l], l, g, n, prm, cacc, rec.data(), t[l]);
        }
        uint32_t first[kSegLanes + 1]; first[0] = 0;
        for (uint32_t l = 0; l < g.nact; ++l) first[l + 1] = first[l] + r[l].count;
        const uint32_t total = first[g.nact];
        for (int l = (int)g.nact - 1; l >= 0; --l) {               // any lane order must do
            WinAcc eacc{data, mf.data(), mq.data(), syms.data(), first[l], (uint32_t)(n % 13), 0, 0};
            seg_emit(r[l], (uint32_t)l, g, n, prm, eacc, blocks.data(), first[l]);
            if (eacc.at != first[l + 1]) return -21;
        }
        const bool pending = prm.mode == MODE_SLOW && ((r[g.nact - 1].end.w0 >> 25) & 1u);
        if (pending) syms[total] = data[n - 1];
        sink.nsyms = total + (pending ? 1 : 0);
        sink.nblocks = seg_finish(blocks.data(), total, pending, n, prm, final_chunk != 0, skip);
    }
    if (!exact_fast) {   // cross-check against the serial whole-chunk parse
        std::vector<uint32_t> syms2(n + 2);
        std::vector<BlockInfo> blocks2(max_blocks_for(n, prm.sym_limit));
        uint32_t ns2 = 0, nb2 = 0;
        parse_chunk(data, n, mf.data(), mq.data(), prm, final_chunk != 0, syms2.data(), blocks2.data(), ns2, nb2, skip);
        if (ns2 != sink.nsyms || nb2 != sink.nblocks) return -22;
        if (memcmp(syms2.data(), syms.data(), 4ull * ns2)) return -23;
        for (uint32_t k = 0; k < nb2; ++k) {
            const BlockInfo &x = blocks[k], &y = blocks2[k];
            if (x.sym_start != y.sym_start || x.sym_count != y.sym_count || x.byte_start != y.byte_start ||
                x.byte_len != y.byte_len || x.flags != y.flags) return -24;
        }
    }
    if (sink.nblocks > blocks.size()) return -2;
    if (stats) { stats[0] = sink.nsyms; stats[1] = sink.nblocks; }
    memset(out, 0, cap);
    BitSink bs{out, cap, 0};
    static TreeWork w;
    static BlockCode code;
    for (uint32_t b = 0; b < sink.nblocks; ++b) {
        const BlockInfo &bi = blocks[b];
        for (int i = 0; i < 286; ++i) w.lt[i].fc = 0;
        for (int i = 0; i < 30; ++i) w.dt[i].fc = 0;
        w.lt[256].fc = 1;
        for (uint32_t i = 0; i < bi.sym_count; ++i) {
            const uint32_t s = syms[bi.sym_start + i], dist = s >> 16, lc = s & 0xffff;
            if (!dist) w.lt[lc].fc++;
            else { w.lt[257 + g_fmt.len_code[lc]].fc++; w.dt[dist_to_code(g_fmt, dist)].fc++; }
        }
        block_build(w, bi, strategy | (level == 0 ? 0x100 : 0), g_st, g_fmt, code);   // as deflate_launch passes it
        const uint64_t start = bs.bit;
        const uint32_t last = bi.flags & BLK_LAST;
        bs.put(last | (code.type << 1), 3);
        if (code.type == 0) {
            bs.bit = (bs.bit + 7) & ~7ull;
            bs.put(bi.byte_len & 0xffff, 16); bs.put(~bi.byte_len & 0xffff, 16);
            for (uint32_t i = 0; i < bi.byte_len; ++i) bs.put(data[bi.byte_start + i], 8);
        } else {
            for (uint32_t i = 0; i < code.hdr_bits; ++i) bs.put((code.hdr[i >> 5] >> (i & 31)) & 1, 1);
            for (uint32_t i = 0; i < bi.sym_count; ++i) {
                uint32_t nb; const uint64_t v = symbol_bits(syms[bi.sym_start + i], code.lcode, code.llen, code.dcode, code.dlen, g_fmt, nb);
                bs.put(v, nb);
            }
            bs.put(code.lcode[256], code.llen[256]);
            if (bs.bit - start != code.body_bits) return -3;      // trees.c:1075 bits_sent == compressed_len
        }
        if (last) bs.bit = (bs.bit + 7) & ~7ull;
        if (bs.bit != block_end_bit(bi, code, start)) return -4;
    }
    if (!final_chunk) { bs.put(0, 3); bs.bit = (bs.bit + 7) & ~7ull; bs.put(0xffff0000u, 32); }
    const size_t bytes = (size_t)((bs.bit + 7) >> 3);
    return bytes > cap ? -1 : (long)bytes;
}

extern "C" long emul_deflate_chunk_dict(const uint8_t *data, uint32_t n, uint32_t skip, int level, int strategy, int final_chunk,
                                        uint8_t *out, size_t cap, uint32_t *stats) {
    return emul_deflate_chunk_opts(data, n, skip, level, strategy, 15, 8, final_chunk, out, cap, stats);
}
extern "C" long emul_deflate_chunk(const uint8_t *data, uint32_t n, int level, int strategy, int final_chunk,
                                   uint8_t *out, size_t cap, uint32_t *stats /* nsyms, nblocks */) {
    return emul_deflate_chunk_dict(data, n, 0, level, strategy, final_chunk, out, cap, stats);
}

// The kernel's construction (keyed heap, tree_build_fast) against the transliterated one (tree_build), on one set of
// frequencies: 0 when type, sizes, every code and length, and the dynamic header agree.
extern "C" int emul_tree_compare(const uint16_t *lfreq /* 286 */, const uint16_t *dfreq /* 30 */, uint32_t byte_len, int strategy) {
    if (!g_ready) { format_fill(g_fmt); static_trees_fill(g_st); g_ready = 1; }
    static TreeWork wa, wb;
    static BlockCode a, b;
    memset(&wa, 0, sizeof wa); memset(&wb, 0, sizeof wb); memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    for (int i = 0; i < 286; ++i) wa.lt[i].fc = wb.lt[i].fc = lfreq[i];
    for (int i = 0; i < 30; ++i) wa.dt[i].fc = wb.dt[i].fc = dfreq[i];
    wa.lt[256].fc = wb.lt[256].fc = 1;
    BlockInfo blk;
    memset(&blk, 0, sizeof blk);
    blk.byte_len = byte_len; blk.flags = BLK_STORED_OK;
    block_build(wa, blk, strategy, g_st, g_fmt, a);
    block_build_fast(wb, blk, strategy, g_st, g_fmt, b);
    if (a.type != b.type) return 1;
    if (a.hdr_bits != b.hdr_bits || a.body_bits != b.body_bits) return 2;
    if (a.type == 0) return 0;
    if (memcmp(a.lcode, b.lcode, sizeof a.lcode) || memcmp(a.llen, b.llen, sizeof a.llen)) return 3;
    if (memcmp(a.dcode, b.dcode, sizeof a.dcode) || memcmp(a.dlen, b.dlen, sizeof a.dlen)) return 4;
    if (memcmp(a.hdr, b.hdr, sizeof a.hdr)) return 5;
    return 0;
}
