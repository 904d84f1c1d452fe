// zb_format.h — DEFLATE format constants shared by the inflate and deflate
// kernels (RFC 1951 §3.2.5-3.2.7; the reference keeps the same numbers in
// trees.c:62-72 extra_lbits/extra_dbits/bl_order, trees.h base_length/base_dist
// and inftrees.c:55-68 lbase/lext/dbase/dext).  Values are generated from the
// extra-bit tables at start-up rather than transcribed.
#pragma once
#include <stdint.h>
#include "zb_gf2.h"   // ZB_HD

namespace zb {

struct FormatTables {
    uint8_t  len_extra[32];    // extra bits per length code 0..28
    uint16_t len_base[32];     // smallest match length (3..258) per length code
    uint8_t  dist_extra[32];   // extra bits per distance code 0..29
    uint16_t dist_base[32];    // smallest distance (1..24577) per distance code
    uint8_t  len_code[256];    // (match length - 3) -> length code
    uint8_t  dist_code[512];   // d<256: code of distance d+1 ; 256+(d>>7): code for d>=256 (d = distance-1)
    uint8_t  cl_order[20];     // transmission order of the code-length code lengths
};

inline void format_fill(FormatTables &t) {
    static const uint8_t lx[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
    static const uint8_t order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
    unsigned l = 0;
    for (int c = 0; c < 32; ++c) { t.len_extra[c] = 0; t.len_base[c] = 0; t.dist_extra[c] = 0; t.dist_base[c] = 0; }
    for (int c = 0; c < 28; ++c) {
        t.len_extra[c] = lx[c];
        t.len_base[c] = (uint16_t)(l + 3);
        for (unsigned k = 0; k < (1u << lx[c]); ++k) t.len_code[l++] = (uint8_t)c;
    }
    t.len_extra[28] = 0; t.len_base[28] = 258; t.len_code[255] = 28;   // 258 has its own code
    unsigned d = 0;
    for (int c = 0; c < 30; ++c) {
        const unsigned x = c < 2 ? 0 : (unsigned)(c / 2 - 1);         // 0,0,0,0,1,1,2,2,...,13,13
        t.dist_extra[c] = (uint8_t)x;
        t.dist_base[c] = (uint16_t)(d + 1);
        for (unsigned k = 0; k < (1u << x); ++k, ++d) {
            if (d < 256) t.dist_code[d] = (uint8_t)c;
            else if ((d & 127) == 0) t.dist_code[256 + (d >> 7)] = (uint8_t)c;
        }
    }
    for (int i = 0; i < 19; ++i) t.cl_order[i] = order[i];
    t.cl_order[19] = 0;
}

// distance (1..32768) -> distance code, deflate.h:317 d_code
ZB_HD unsigned dist_to_code(const FormatTables &t, unsigned dist) {
    const unsigned d = dist - 1;
    return d < 256 ? t.dist_code[d] : t.dist_code[256 + (d >> 7)];
}

}  // namespace zb
