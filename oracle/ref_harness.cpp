// TEST INFRASTRUCTURE ONLY — C-ABI harness around the *unmodified* reference
// sources under /root/reference/src (compiled where they lie by oracle/Makefile
// into oracle/_ref/libzling_ref.so).  Nothing under libzling_b200/ may link or
// load this; it exists so tests/ and bench.py's cpu_baseline leg can ask the real
// reference for whole-stream bytes and for intermediates (per-sub-block symbol
// buffers, Huffman length/encode tables).
//
// Reference entry points wrapped here:
//   baidu::zling::Encode / Decode          src/libzling.h:44-45
//   lz::ZlingRolzEncoder::Encode / Reset   src/libzling_lz.h:81-82
//   huffman::ZlingMakeLengthTable/EncodeTable  src/libzling_huffman.h:51-54
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <new>

#include "libzling.h"
#include "libzling_lz.h"
#include "libzling_huffman.h"

namespace {

struct MemIn : baidu::zling::Inputter {
    const unsigned char* p; size_t n, at;
    MemIn(const unsigned char* p_, size_t n_) : p(p_), n(n_), at(0) {}
    size_t GetData(unsigned char* buf, size_t len) override {
        size_t k = n - at < len ? n - at : len;
        memcpy(buf, p + at, k); at += k; return k;
    }
    bool IsEnd() override { return at >= n; }
    bool IsErr() override { return false; }
};

struct MemOut : baidu::zling::Outputter {
    unsigned char* p; size_t cap, at; bool overflow;
    MemOut(unsigned char* p_, size_t cap_) : p(p_), cap(cap_), at(0), overflow(false) {}
    size_t PutData(unsigned char* buf, size_t len) override {
        if (at + len > cap) { overflow = true; at += len; return len; }
        memcpy(p + at, buf, len); at += len; return len;
    }
    bool IsErr() override { return false; }
};

}  // namespace

extern "C" {

// returns compressed size, or -1 on error / -2 if `cap` was too small (size still counted)
long long zref_encode(const unsigned char* in, size_t n, int level, unsigned char* out, size_t cap) {
    if (level < 0 || level > 4) return -1;
    MemIn i(in, n); MemOut o(out, cap);
    try {
        if (baidu::zling::Encode(&i, &o, NULL, level) != 0) return -1;
    } catch (...) { return -1; }
    return o.overflow ? -2 : (long long) o.at;
}

// returns decoded size, -1 on I/O error, -3 on malformed stream (reference threw), -2 overflow
long long zref_decode(const unsigned char* in, size_t n, unsigned char* out, size_t cap) {
    MemIn i(in, n); MemOut o(out, cap);
    try {
        if (baidu::zling::Decode(&i, &o, NULL) != 0) return -1;
    } catch (const std::runtime_error&) { return -3;
    } catch (...) { return -1; }
    return o.overflow ? -2 : (long long) o.at;
}

// --- intermediates -------------------------------------------------------
void* zref_rolz_new(void) { return new (std::nothrow) baidu::zling::lz::ZlingRolzEncoder(); }
void  zref_rolz_free(void* h) { delete static_cast<baidu::zling::lz::ZlingRolzEncoder*>(h); }
void  zref_rolz_reset(void* h) { static_cast<baidu::zling::lz::ZlingRolzEncoder*>(h)->Reset(); }
// one sub-block: returns rlen, advances *encpos (src/libzling.cpp:206)
int zref_rolz_encode(void* h, int level, const unsigned char* ibuf, uint16_t* tbuf, int ilen, int olen, int* encpos) {
    return static_cast<baidu::zling::lz::ZlingRolzEncoder*>(h)->Encode(
        level, const_cast<unsigned char*>(ibuf), tbuf, ilen, olen, encpos);
}

void zref_make_length_table(const uint32_t* freq, uint32_t* len, int n, int maxlen) {
    baidu::zling::huffman::ZlingMakeLengthTable(freq, len, n, maxlen);
}
void zref_make_encode_table(const uint32_t* len, uint16_t* enc, int n, int maxlen) {
    baidu::zling::huffman::ZlingMakeEncodeTable(len, enc, n, maxlen);
}
void zref_make_decode_table(const uint32_t* len, uint16_t* enc, uint16_t* dec, int n, int maxlen) {
    baidu::zling::huffman::ZlingMakeDecodeTable(len, enc, dec, n, maxlen);
}

}  // extern "C"
