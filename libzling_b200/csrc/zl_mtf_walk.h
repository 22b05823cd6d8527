// The serial heart of the MTF rank pass: one context's literal list in stream order (ZlingMTFEncoder::Encode,
// src/libzling_lz.cpp:112-117: rank i = index[c]; swap the table entries i and mtfnext[i]).
//
// State of one context in shared memory, laid out so that the dependent chain per literal is ONE load and the
// instruction count stays small (the walk is a single thread: it is bound by instruction issue, ~1 per 2-3 cycles):
//   R[b] (u16) = rank(b) | mtf_next(rank(b)) << 8     one load returns both ends of the swap
//   S[r] (u16) = byte at rank r | mtf_next(r) << 8    the high byte is a constant of the POSITION: the load that fetches
//                                                     the byte to swap with also brings mtf_next of the new rank
// The R entry of the next literal is loaded before the stores of the current one and patched from registers when the
// current literal rewrote it (a literal rewrites exactly R[b] and R[o], o = the byte it swapped with).
// Scalar ZL_HD code: tests/cxx/mtf_walk_test.cu fuzzes it on the host against the plain loop.
#pragma once
#include <stdint.h>
#include "zl_tables.h"

namespace zl {

// sym[r] = byte at rank r  ->  R, S   (serial form; the kernel fills them with all lanes)
ZL_HD void mtf_walk_init(uint16_t* R, uint16_t* S, const uint8_t* sym) {
    for (int r = 0; r < 256; r++) {
        R[sym[r]] = (uint16_t) (r | (mtf_next(r) << 8));
        S[r] = (uint16_t) (sym[r] | (mtf_next(r) << 8));
    }
}

// ranks of `cnt` literal bytes (bytes[] readable up to cnt + 3, 4-byte aligned) -> out[] (4-byte aligned, writable up to
// the next multiple of 4)
ZL_HD void mtf_walk(uint16_t* R, uint16_t* S, const uint8_t* bytes, uint8_t* out, int cnt) {
    if (cnt <= 0) return;
    const uint32_t* b4p = reinterpret_cast<const uint32_t*>(bytes);
    uint32_t* o4p = reinterpret_cast<uint32_t*>(out);
    uint32_t w = b4p[0];
    uint32_t b0 = w & 0xffu;
    uint32_t v0 = R[b0];
    uint32_t acc = 0;
    for (int q = 0; q < cnt; q++) {
        // next literal's byte (0 past the end: harmless, its R entry is only read)
        if ((q & 3) == 3) w = b4p[(q >> 2) + 1]; else w >>= 8;
        const uint32_t b1 = q + 1 < cnt ? (w & 0xffu) : 0u;
        uint32_t v1 = R[b1];                                     // read before this literal's stores, patched below
        const uint32_t i = v0 & 0xffu, j = v0 >> 8;
        const uint32_t sj = S[j];
        const uint32_t o = sj & 0xffu;
        const uint32_t vb = j | (sj & 0xff00u);                  // rank j | mtf_next(j) << 8
        reinterpret_cast<uint8_t*>(S)[2 * i] = (uint8_t) o;      // low bytes only: the high byte belongs to the position
        reinterpret_cast<uint8_t*>(S)[2 * j] = (uint8_t) b0;
        R[o] = (uint16_t) v0; R[b0] = (uint16_t) vb;            // (o == b0 when i == j == 0: both stores write the same value)
        acc |= i << (8 * (q & 3));
        if ((q & 3) == 3) { o4p[q >> 2] = acc; acc = 0; }
        v1 = b1 == b0 ? vb : (b1 == o ? v0 : v1);
        b0 = b1; v0 = v1;
    }
    if (cnt & 3) o4p[cnt >> 2] = acc;
}

}  // namespace zl
