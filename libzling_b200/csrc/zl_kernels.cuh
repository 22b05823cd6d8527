// Device-side layout shared by the kernels (zl_kernels.cu) and the host engine (zl_engine.cu).
#pragma once
#include <stdint.h>
#include "zl_tables.h"

namespace zl {

// ---- token word -------------------------------------------------------------------------------------
// One u32 per token, written by the parse kernel, literal ranks patched in place by the MTF kernel:
//   bits  0..9   symbol: 0..255 literal (raw byte after parse, MTF rank after zl_mtf_rank), 256/257 word-MRU
//                hit, 258+k match of length 4+k                                   
//   bits 10..21  aux: match idx (1..4095) for matches, context byte (previous input byte) for literals
//   bits 22..29  the literal's byte (kept so that the MTF pass can be replayed after a level-feedback re-parse)
//   bit  31      raw literal (first two bytes of a block are not MTF coded, src/libzling_lz.cpp:150-151)
constexpr uint32_t kTokSymMask = 0x3ffu;
constexpr uint32_t kTokRaw     = 0x80000000u;
ZL_HD uint32_t tok_match(uint32_t len, uint32_t idx) { return (258u + len - kMinLen) | (idx << 10); }
ZL_HD uint32_t tok_word(uint32_t which) { return 256u + which; }
ZL_HD uint32_t tok_literal(uint32_t byte, uint32_t ctx, bool raw) { return byte | (ctx << 10) | (byte << 22) | (raw ? kTokRaw : 0u); }
ZL_HD uint32_t tok_sym(uint32_t t)  { return t & kTokSymMask; }
ZL_HD uint32_t tok_aux(uint32_t t)  { return (t >> 10) & 0xfffu; }
ZL_HD uint32_t tok_byte(uint32_t t) { return (t >> 22) & 0xffu; }

// ---- per-sub-block record (mirrors zlb_subblock in include/zlb.h) --------------------------------------
struct SubBlock {
    uint32_t tok_begin, tok_end;   // token range inside the block's token array
    uint32_t enc_begin, enc_end;   // input byte range inside the block; enc_end is the framed `encpos`
    uint32_t rlen;                 // u16 symbol count the reference would have produced (match = 2)
    uint32_t level;                // level this sub-block was parsed with
    uint32_t olen;                 // payload bytes: 273 + ceil(bits / 8)
    uint32_t bits_lo;              // total code bits (low 32; < 2^23 anyway)
};

// ---- ring entry of a context bucket ---------------------------------------------------------------------
// The reference keeps suffix[4096] (u16), offset[4096] (u32 = pos | check<<24) and hash[8192] (u16) per context
// (src/libzling_lz.h:98-103).  Here suffix and offset share one 8-byte word so that one 8-byte load returns the
// whole chain node: bits 0..23 pos, 24..31 check byte, 32..47 suffix (ring index of the next older node).
ZL_HD uint64_t ring_make(uint32_t pos, uint32_t check, uint32_t suffix) {
    return (uint64_t) (pos | (check << 24)) | ((uint64_t) suffix << 32);
}
ZL_HD uint32_t ring_pos(uint64_t e)    { return (uint32_t) e & 0xffffffu; }
ZL_HD uint32_t ring_check(uint64_t e)  { return ((uint32_t) e) >> 24; }
ZL_HD uint32_t ring_suffix(uint64_t e) { return (uint32_t) (e >> 32) & 0xffffu; }
constexpr uint64_t kRingEmpty = (uint64_t) kNil << 32;      // Reset(): offset = 0, suffix = 65535 (lz.cpp:197-209)

// ---- Huffman tables of one sub-block -------------------------------------------------------------------
struct HuffTables {
    uint16_t code1[kSyms1 + 2];
    uint16_t code2[kSyms2];
    uint8_t  len1[kSyms1 + 2];     // padded to even for nibble packing (src/libzling.cpp:214)
    uint8_t  len2[kSyms2];
};

// ---- decode: one record per framed sub-block (built by the host from the container headers) -------------
struct DecSub {
    unsigned long long payload_off;   // offset of the 273-byte table + bits inside the compressed buffer
    uint32_t encpos, rlen, olen;      // BE32 fields of the frame (src/libzling.cpp:322-324)
    uint32_t block;                   // block index inside this call
    uint32_t sym_off;                 // u16 offset of this sub-block's symbols inside the block's symbol area
    uint32_t pad;
};

// one stream's sub-block records for the ROLZ decode of a call (zl_rolz_decode_kernel, grid.x = stream)
struct DecRange {
    int s0, s1;                // sub-block records [s0, s1) of the call's DecSub array, in stream order
    uint8_t* state;            // 256 x 256 rank -> byte tables: carried in, written back when the stream's blocks decoded cleanly
};

// per-block strides of the device arrays (all blocks of a batch use the same stride)
constexpr size_t kTokStride  = (size_t) kBlockBytes;                 // u32 tokens per block (worst case 1 per byte)
constexpr size_t kLitStride  = (size_t) kBlockBytes;                 // u32 literal -> token index per block
constexpr size_t kRingStride = (size_t) 256 * kRing;                 // u64 ring entries per block
constexpr size_t kHashStride = (size_t) 256 * kSlots;                // u16 slot heads per block

// one stream's block range for the MTF rank pass of a call (zl_mtf_ctx_kernel, grid.y = stream)
struct MtfRange {
    int first, end;            // blocks [first, end) get new ranks in this pass; first == end: nothing to do
    int b0, pad;               // first block of the stream inside the call (first == b0: start from state_in, else from the checkpoint)
    const uint8_t* state_in;   // 256 x 256 rank -> byte tables carried into the stream's first block
    uint8_t* state_out;        // tables after the stream's last block
};

struct ParseArgs {
    const uint8_t*  in;        // block b at in + b * kBlockBytes
    const uint32_t* ilen;      // [nblocks] bytes in block
    const uint8_t*  plan;      // [nblocks][kMaxSubPerBlock] level to use for sub-block j
    const uint8_t*  active;    // [nblocks] 1 = (re)parse this block in this launch
    uint64_t*       ring;      // [nblocks] x kRingStride
    uint16_t*       hash;      // [nblocks] x kHashStride
    uint32_t*       tok;       // [nblocks] x kTokStride
    uint32_t*       lit;       // [nblocks] x kLitStride
    SubBlock*       sub;       // [nblocks][kMaxSubPerBlock]
    uint32_t*       nsub;      // [nblocks]
    uint32_t*       ntok;      // [nblocks]
    uint32_t*       nlit;      // [nblocks]
    const uint8_t*  pre_tail;  // null, or 65536 bytes that precede block 0 in its stream + a u32 'valid' flag (ranges of a sharded stream)
};

}  // namespace zl
