// Format-defining constants of the zling bitstream, product copy (host + device).
// Values are facts of the on-disk format; provenance in the reference: src/tables/gen.py
//   - kMtfInit: initial MTF order (permutation of 0..255 "auto-generated from enwik8", gen.py:31-49)
//   - mtfnext(i) = floor(0.95 i) for i<128, floor(0.55 i) otherwise (gen.py:51-56) -> zl_mtf_next()
//   - match-index buckets (gen.py:10-19): extra bits 0,0,0,0,1,1,2,2,..,7,7 then 8 x 14 -> zl_idx_bucket()
// tests/test_tables.py checks these against the reference's generated .inc files.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZL_HD __host__ __device__ __forceinline__
#else
#define ZL_HD inline
#endif

namespace zl {

constexpr int kBlockBytes   = 16777216;  // src/libzling.cpp:70
constexpr int kSubSymbols   = 262144;    // src/libzling.cpp:71
constexpr int kSubBytesMax  = 393216;    // src/libzling.cpp:72
constexpr int kRing         = 4096;      // src/libzling_lz.h:44
constexpr int kSlots        = 8192;      // src/libzling_lz.h:45
constexpr int kLazyBelow    = 128;       // src/libzling_lz.h:46
constexpr int kMinLen       = 4;         // src/libzling_lz.h:47
constexpr int kMaxLen       = 259;       // src/libzling_lz.h:48
constexpr int kSyms1        = 514;       // src/libzling.cpp:63
constexpr int kSyms2        = 32;        // src/libzling.cpp:64
constexpr int kCap1         = 15;        // src/libzling.cpp:65
constexpr int kCap2         = 8;         // src/libzling.cpp:66
constexpr int kFastBits     = 10;        // src/libzling.cpp:67
constexpr int kGuard        = 275;       // kMatchMaxLen + 16, src/libzling.cpp:68
constexpr int kTableBytes   = 273;       // 257 + 16 nibble-packed length bytes, src/libzling.cpp:232-237
constexpr int kNil          = 65535;
constexpr int kMaxSubPerBlock = 72;      // >= ceil(16 MiB / 262143) + slack: a full sub-block consumes >= 262143 bytes

// (match depth, lazy depth at pos+1, lazy depth at pos+2) per level, src/libzling_lz.cpp:129-135
ZL_HD int depth_main(int level)  { return level == 0 ? 2 : level == 1 ? 4 : level == 2 ? 6 : level == 3 ? 8 : 16; }
ZL_HD int depth_lazy1(int level) { return level <= 1 ? 1 : level; }
ZL_HD int depth_lazy2(int level) { return level <= 2 ? 0 : level - 2; }

ZL_HD int mtf_next(int i) { return i < 128 ? (i * 95) / 100 : (i * 55) / 100; }

// idx (1..4095) -> bucket code 0..31, number of extra bits, bucket base
ZL_HD int idx_bucket(int idx) {
    if (idx < 4) return idx;
    if (idx >= 512) return 18 + ((idx - 512) >> 8);
    int k = 0;
    for (int v = idx; v > 1; v >>= 1) k++;          // floor(log2(idx)), 2..8
    return 2 * k + ((idx >> (k - 1)) & 1);
}
ZL_HD int idx_extra_bits(int bucket) { return bucket < 4 ? 0 : (bucket < 18 ? (bucket - 2) / 2 : 8); }
ZL_HD int idx_base(int bucket) {
    if (bucket < 4) return bucket;
    if (bucket >= 18) return 512 + ((bucket - 18) << 8);
    int k = bucket >> 1;
    return (1 << k) + ((bucket & 1) << (k - 1));
}

static const uint8_t kMtfInit[256] = {
     32, 101, 116,  97, 105, 111, 110, 114, 115, 108, 104, 100,  99, 117,  93,  91,
    109, 112, 103, 102,  10, 121,  98,  39, 119,  46,  44, 118,  59,  38, 124,  47,
     49, 107,  61,  48,  67,  65,  58,  45,  84,  83,  60,  62,  50, 113,  73,  57,
     42, 120,  41,  40,  66,  77,  80,  69,  68,  53,  51,  72,  70,  56,  52,  71,
     82,  54,  76,  55,  78,  87, 122, 125, 123,  79, 106,  85,  74,  75, 208,  95,
    195,  35,  86, 215,  90,  34,  89, 209, 128, 224, 184, 131,  92, 227,  37,  33,
    176, 169, 206, 226, 130,  63,  88,  81, 161, 153,  43, 129, 188, 179, 216, 164,
    181, 189, 148, 190, 173, 187, 186, 229, 225, 167, 217, 177, 178, 168, 149, 185,
    197, 144, 147, 196, 207, 194, 180, 156, 132, 170, 166, 136, 182, 191,   9, 230,
    141, 160, 175,  36, 152, 140, 165, 145,  94, 133, 163, 183, 171, 157, 137, 174,
    134, 135, 236, 151, 231, 155, 201, 158, 138, 143, 150, 162, 159, 139, 172, 154,
    126, 232, 235, 146, 233, 228, 202, 203, 142, 214, 237, 204, 219, 234, 213,  96,
    218, 199,  64, 210, 239, 198, 211, 205, 212, 240, 222, 220, 200,   0,   1,   2,
      3,   4,   5,   6,   7,   8,  11,  12,  13,  14,  15,  16,  17,  18,  19,  20,
     21,  22,  23,  24,  25,  26,  27,  28,  29,  30,  31, 127, 192, 193, 221, 223,
    238, 241, 242, 243, 244, 245, 246, 247, 248, 249, 250, 251, 252, 253, 254, 255
};

}  // namespace zl
