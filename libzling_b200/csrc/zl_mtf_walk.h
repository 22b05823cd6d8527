// The serial heart of the MTF rank pass: one context's literal list in stream order (ZlingMTFEncoder::Encode,
// src/libzling_lz.cpp:112-117: rank i = index[c]; swap the table entries i and mtfnext[i]).
//
// State of one context in shared memory:
//   R[b] (u16) = rank(b) | mtf_next(rank(b)) << 8     one load returns both ends of the swap
//   S[r] (u8)  = the byte at rank r
// Per literal the dependent chain is ONE shared-memory load (S[j], ~29 cycles) plus a few ALU operations: the R entry
// of a literal is loaded two literals ahead (before the stores of the two literals in front of it) and patched from
// registers when one of those two literals rewrote it (a literal rewrites exactly R[b] and R[o], o = the byte it
// swapped with).  Scalar ZL_HD code: tests/cxx/mtf_walk_test.cu fuzzes it on the host against the plain loop.
#pragma once
#include <stdint.h>
#include "zl_tables.h"

namespace zl {

ZL_HD void mtf_walk_init(uint16_t* R, const uint8_t* S) {      // (serial form; the kernel fills R with all lanes)
    for (int r = 0; r < 256; r++) R[S[r]] = (uint16_t) (r | (mtf_next(r) << 8));
}

// ranks of `cnt` literal bytes (bytes[] readable up to cnt + 3) -> out[]; N[r] = mtf_next(r) as a byte table
ZL_HD void mtf_walk(uint16_t* R, uint8_t* S, const uint8_t* N, const uint8_t* bytes, uint8_t* out, int cnt) {
    if (cnt <= 0) return;
    uint32_t b0 = bytes[0], b1 = cnt > 1 ? bytes[1] : 0u;
    uint32_t v0 = R[b0], v1 = R[b1];
    for (int q = 0; q < cnt; q++) {
        const uint32_t b2 = q + 2 < cnt ? bytes[q + 2] : 0u;
        const uint32_t i = v0 & 0xffu, j = v0 >> 8;
        const uint32_t o = S[j], nj = N[j];
        uint32_t v2 = R[b2];                                     // two ahead: read before this literal's stores
        const uint32_t vb = j | (nj << 8);
        S[i] = (uint8_t) o; S[j] = (uint8_t) b0;
        R[o] = (uint16_t) v0; R[b0] = (uint16_t) vb;            // (o == b0 when i == j == 0: both stores write 0)
        out[q] = (uint8_t) i;
        // the two literals in flight may be one of the bytes whose R entry was just rewritten
        v1 = b1 == b0 ? vb : (b1 == o ? v0 : v1);
        v2 = b2 == b0 ? vb : (b2 == o ? v0 : v2);
        b0 = b1; v0 = v1; b1 = b2; v1 = v2;
    }
}

}  // namespace zl
