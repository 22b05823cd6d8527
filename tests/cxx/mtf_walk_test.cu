// Host fuzz of mtf_walk (libzling_b200/csrc/zl_mtf_walk.h) against the plain MTF loop of the reference's
// ZlingMTFEncoder (src/libzling_lz.cpp:112-117).  TEST INFRASTRUCTURE.  exit 0 = identical ranks and final tables.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../libzling_b200/csrc/zl_mtf_walk.h"
using namespace zl;
int main() {
    uint32_t seed = 12345;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
    for (int iter = 0; iter < 400; iter++) {
        const int n = 1 + rnd() % 5000, alpha = 1 + rnd() % (iter % 3 == 0 ? 4 : 256);
        std::vector<uint8_t> bytes(n + 8, 0), out(n + 8), want(n);
        for (int i = 0; i < n; i++) { uint32_t r = rnd(); bytes[i] = (uint8_t) ((iter & 1) ? (r % alpha) : ((r % alpha) * (r >> 12 & 1) + 32)); }
        uint8_t S[256], T[256], idx[256]; uint16_t R[256], S16[256];
        memcpy(S, kMtfInit, 256); memcpy(T, kMtfInit, 256);
        for (int k = 0; k < (int) (rnd() % 50); k++) { int a = rnd() % 256, b = rnd() % 256; uint8_t t = S[a]; S[a] = S[b]; S[b] = t; t = T[a]; T[a] = T[b]; T[b] = t; }
        for (int r = 0; r < 256; r++) idx[T[r]] = (uint8_t) r;
        mtf_walk_init(R, S16, S);
        // split into chunks like the kernel does
        int at = 0;
        while (at < n) { int c = 4 * (1 + rnd() % 75); if (c > n - at) c = n - at; mtf_walk(R, S16, bytes.data() + at, out.data() + at, c); at += c; }   // chunks start 4-byte aligned, like the kernel's
        for (int r = 0; r < 256; r++) { S[r] = (uint8_t) S16[r]; if ((S16[r] >> 8) != mtf_next(r)) { printf("S high byte damaged at iter %d\n", iter); return 1; } }
        for (int q = 0; q < n; q++) {
            const int c = bytes[q], i = idx[c], j = mtf_next(i);
            const uint8_t other = T[j];
            T[i] = other; T[j] = (uint8_t) c; idx[other] = (uint8_t) i; idx[c] = (uint8_t) j;
            want[q] = (uint8_t) i;
        }
        if (memcmp(out.data(), want.data(), n) || memcmp(S, T, 256)) { printf("MISMATCH at iter %d\n", iter); return 1; }
        for (int r = 0; r < 256; r++) if (R[S[r]] != (uint16_t) (r | (mtf_next(r) << 8))) { printf("R table inconsistent at iter %d\n", iter); return 1; }
    }
    printf("mtf_walk OK\n");
    return 0;
}
