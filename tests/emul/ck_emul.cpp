// Host replay of the checksum kernels' thread decomposition (TEST ONLY).
// Compiles zb_checksum.cuh with g++ and walks every (part, thread) pair the way
// ck_big_kernel / ck_seg_kernel do, so the GF(2) weighting and the Adler position
// algebra are validated on the CPU before any GPU time is spent.  Never linked
// into libzb200.so.
#include <stdint.h>
#include <stddef.h>
#include "../../zlib_wasm_b200/csrc/zb_checksum.cuh"

using namespace zb;

extern "C" int emul_checksum(const uint8_t *data, uint64_t len, uint32_t T, uint32_t parts, int which,
                             uint32_t init_crc, uint32_t init_adler, uint32_t *crc, uint32_t *adler) {
    X2nTable x2n;
    gf2_fill_x2n(x2n);
    static uint32_t tab[4][256];
    ck_fill_horner(tab, x2n.v, T);
    const uint32_t x32 = x2n.v[5];
    auto tf = [&](int j, uint32_t v) -> uint32_t { return tab[j][(v >> (8 * j)) & 0xff]; };
    uint32_t acc_crc = 0;
    uint64_t acc_a = 0, acc_b = 0;
    for (uint32_t p = 0; p < parts; ++p) {
        CkPart part = ck_make_part(data, len, p, parts);
        for (uint32_t t = 0; t < T; ++t) {
            CkPartial v = ck_thread_body<true, true>(part, t, T, tf, x2n.v, x32);
            acc_crc ^= v.crc; acc_a += v.a; acc_b += v.b;
        }
        CkPartial e = ck_edge_bytes<true, true>(part, x2n.v);
        acc_crc ^= e.crc; acc_a += e.a; acc_b += e.b;
    }
    ck_finish(acc_crc, acc_a, acc_b, len, init_crc, init_adler, x2n.v, (which & 1) ? crc : nullptr, (which & 2) ? adler : nullptr);
    return 0;
}
