// sm_100a kernels of the zling block pipeline (encode side).
//
//   zl_reset_buckets   Reset()                                   src/libzling_lz.cpp:197-209
//   zl_rolz_parse      EncodeImpl + MatchAndUpdate + MatchLazy    src/libzling_lz.cpp:139-316
//   zl_mtf_rank        ZlingMTFEncoder::Encode in stream order    src/libzling_lz.cpp:112-117
//   zl_huff_build      freq count + MakeLengthTable + MakeEncodeTable   src/libzling.cpp:219-229,
//                                                                 src/libzling_huffman.cpp:41-138
//   zl_huff_pack       nibble header + ZlingCodebuf packing + framing  src/libzling.cpp:232-257,269-278
//
// Integer / byte work throughout: no tensor cores.  The bound that matters is the serial token chain of the
// parse; the other kernels are parallel over sub-blocks and tokens.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zl_kernels.cuh"
#include "zl_mtf_walk.h"

namespace zl {

// =====================================================================================================
// bucket reset
// =====================================================================================================
__global__ void zl_reset_buckets_kernel(uint64_t* ring, uint16_t* hash, const uint8_t* active, int nblocks) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    for (int b = 0; b < nblocks; b++) {
        if (!active[b]) continue;
        ulonglong2* r = reinterpret_cast<ulonglong2*>(ring + (size_t) b * kRingStride);
        for (size_t i = tid; i < kRingStride / 2; i += stride) r[i] = make_ulonglong2(kRingEmpty, kRingEmpty);
        uint4* h = reinterpret_cast<uint4*>(hash + (size_t) b * kHashStride);
        for (size_t i = tid; i < kHashStride / 8; i += stride) h[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
}

}  // namespace zl

#include "zl_parse_v4.cuh"
namespace zl {

// =====================================================================================================
// MTF rank pass, context-parallel form (zl_lit_count / zl_lit_scan / zl_lit_scatter / zl_mtf_ctx)
// =====================================================================================================
// The 256 MTF tables are independent (ZlingMTFEncoder m_mtf[256], src/libzling_lz.h:105), so the stream-order pass
// splits into 256 chains: literals are bucketed by context with a stable counting sort (order inside a context is
// stream order), then one CTA per context walks its list.  The longest list (context 0x20 on text) is the
// critical path; everything else runs beside it.
constexpr int kLitUnit = 8192;                                    // literals per warp in the bucketing kernels
constexpr int kLitUnitsMax = kBlockBytes / kLitUnit;              // 2048 units per block at most
constexpr int kLitWarps = 8;

// grid (ceil(units / kLitWarps), nblocks); hist[b][unit][256]
__global__ void __launch_bounds__(kLitWarps * 32) zl_lit_count_kernel(const uint32_t* tok_all, const uint32_t* lit_all, const uint32_t* nlit,
                                                                    int first_block, uint32_t* hist) {
    const int b = blockIdx.y + first_block, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kLitWarps + warp;
    const int n = (int) nlit[b];
    __shared__ uint32_t s_h[kLitWarps][256];
    if (unit * kLitUnit >= n) return;
    for (int i = lane; i < 256; i += 32) s_h[warp][i] = 0;
    __syncwarp();
    const uint32_t* tok = tok_all + (size_t) b * kTokStride;
    const uint32_t* lit = lit_all + (size_t) b * kLitStride;
    const int lo = unit * kLitUnit, hi = min(lo + kLitUnit, n);
    for (int i = lo + lane; i < hi; i += 32) atomicAdd(&s_h[warp][tok_aux(tok[lit[i]]) & 0xff], 1u);
    __syncwarp();
    uint32_t* out = hist + ((size_t) b * kLitUnitsMax + unit) * 256;
    for (int i = lane; i < 256; i += 32) out[i] = s_h[warp][i];
}

// grid nblocks, 256 threads: hist[b][unit][ctx] -> start offset of (unit, ctx) inside the block's bucketed list;
// ctx_off[b][ctx] (257 entries) = start of each context's list
__global__ void __launch_bounds__(256) zl_lit_scan_kernel(const uint32_t* nlit, int first_block, uint32_t* hist, uint32_t* ctx_off) {
    const int b = blockIdx.x + first_block, ctx = threadIdx.x;
    const int n = (int) nlit[b];
    const int units = (n + kLitUnit - 1) / kLitUnit;
    uint32_t* h = hist + (size_t) b * kLitUnitsMax * 256;
    uint32_t total = 0;
    for (int u = 0; u < units; u++) total += h[(size_t) u * 256 + ctx];
    __shared__ uint32_t s_scan[256];
    s_scan[ctx] = total;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const uint32_t v = ctx >= d ? s_scan[ctx - d] : 0;
        __syncthreads();
        s_scan[ctx] += v;
        __syncthreads();
    }
    uint32_t run = s_scan[ctx] - total;
    ctx_off[(size_t) b * 257 + ctx] = run;
    if (ctx == 255) ctx_off[(size_t) b * 257 + 256] = s_scan[255];
    for (int u = 0; u < units; u++) { const uint32_t v = h[(size_t) u * 256 + ctx]; h[(size_t) u * 256 + ctx] = run; run += v; }
}

// same grid as the count kernel; lbuf[b][k] = token index << 8 | literal byte, grouped by context, stream order inside
__global__ void __launch_bounds__(kLitWarps * 32) zl_lit_scatter_kernel(const uint32_t* tok_all, const uint32_t* lit_all, const uint32_t* nlit,
                                                                      int first_block, const uint32_t* hist, uint32_t* lbuf_all) {
    const int b = blockIdx.y + first_block, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * kLitWarps + warp;
    const int n = (int) nlit[b];
    __shared__ uint32_t s_at[kLitWarps][256];
    if (unit * kLitUnit >= n) return;
    const uint32_t* base = hist + ((size_t) b * kLitUnitsMax + unit) * 256;
    for (int i = lane; i < 256; i += 32) s_at[warp][i] = base[i];
    __syncwarp();
    const uint32_t* tok = tok_all + (size_t) b * kTokStride;
    const uint32_t* lit = lit_all + (size_t) b * kLitStride;
    uint32_t* lbuf = lbuf_all + (size_t) b * kLitStride;
    const int lo = unit * kLitUnit, hi = min(lo + kLitUnit, n);
    for (int i0 = lo; i0 < hi; i0 += 32) {
        const int i = i0 + lane;
        const bool live = i < hi;
        uint32_t ti = 0, t = 0;
        if (live) { ti = lit[i]; t = tok[ti]; }
        const int ctx = live ? (int) (tok_aux(t) & 0xff) : 256 + lane;
        const uint32_t same = __match_any_sync(0xffffffffu, ctx);
        const int order = __popc(same & ((1u << lane) - 1u));
        uint32_t at = 0;
        if (live) at = s_at[warp][ctx] + order;
        __syncwarp();
        if (live && (same >> lane) == 1u) s_at[warp][ctx] = at + 1;       // last lane of the group advances the cursor
        if (live) lbuf[at] = (ti << 8) | tok_byte(t);
        __syncwarp();
    }
}

constexpr int kMtfChunk = 256;                                    // literal records staged per round

// grid 256 (one CTA = one warp per context).  Lane 0 walks the context's literal list of every block in stream
// order (ZlingMTFEncoder::Encode, lz.cpp:112-117; the walk itself is mtf_walk, zl_mtf_walk.h: one dependent
// shared-memory load per literal); all lanes prefetch the next records into shared memory and write the token words.
__global__ void __launch_bounds__(32) zl_mtf_ctx_kernel(uint32_t* tok_all, const uint32_t* lbuf_all, const uint32_t* ctx_off, const MtfRange* ranges,
                                                        uint8_t* checkpoints /* [nblocks][65536] */) {
    // grid (256 contexts, streams of the call): every stream has its own tables; a stream whose blocks need no new ranks in this
    // pass (first == end) is skipped
    const MtfRange rg = ranges[blockIdx.y];
    const int first_block = rg.first, nblocks = rg.end;
    if (first_block >= nblocks) return;
    const uint8_t* state_in = rg.state_in; uint8_t* state_out = rg.state_out;
    const int ctx = blockIdx.x, lane = threadIdx.x;
    __shared__ __align__(16) uint16_t s_S[256];           // rank -> byte | mtf_next(rank) << 8   (zl_mtf_walk.h)
    __shared__ __align__(16) uint16_t s_R[256];           // byte -> rank | mtf_next(rank) << 8
    __shared__ __align__(16) uint32_t s_rec[2][kMtfChunk];
    __shared__ __align__(16) uint8_t s_out[kMtfChunk + 16];
    __shared__ __align__(16) uint8_t s_byte[2][kMtfChunk + 16];
    {
        // the stream's first block starts from the carried state; a replay from a later block (level-feedback re-parse)
        // from the checkpoint the previous pass left for that block
        const uint8_t* src = (first_block == rg.b0 ? state_in : checkpoints + (size_t) first_block * 65536) + ctx * 256;
        for (int i = lane; i < 256; i += 32) {
            const uint32_t sy = src[i], nx = (uint32_t) mtf_next(i);
            s_S[i] = (uint16_t) (sy | (nx << 8));
            s_R[sy] = (uint16_t) ((uint32_t) i | (nx << 8));
        }
        __syncwarp();
    }
    for (int b = first_block; b < nblocks; b++) {
        {   // MTF state at the start of block b (replay point for level-feedback re-parses)
            uint8_t* dst = checkpoints + (size_t) b * 65536 + ctx * 256;
            for (int i = lane; i < 256; i += 32) dst[i] = (uint8_t) s_S[i];
        }
        const uint32_t lo = ctx_off[(size_t) b * 257 + ctx], hi = ctx_off[(size_t) b * 257 + ctx + 1];
        const uint32_t* list = lbuf_all + (size_t) b * kLitStride + lo;
        const int n = (int) (hi - lo);
        uint32_t* tok = tok_all + (size_t) b * kTokStride;
        uint32_t pre[kMtfChunk / 32];
        #pragma unroll
        for (int q = 0; q < kMtfChunk / 32; q++) { const int i = q * 32 + lane; pre[q] = i < n ? list[i] : 0; }
        for (int base = 0, buf = 0; base < n; base += kMtfChunk, buf ^= 1) {
            #pragma unroll
            for (int q = 0; q < kMtfChunk / 32; q++) { s_rec[buf][q * 32 + lane] = pre[q]; s_byte[buf][q * 32 + lane] = (uint8_t) pre[q]; }
            #pragma unroll
            for (int q = 0; q < kMtfChunk / 32; q++) { const int i = base + kMtfChunk + q * 32 + lane; pre[q] = i < n ? list[i] : 0; }
            __syncwarp();
            const int cnt = min(kMtfChunk, n - base);
            if (lane == 0) mtf_walk(s_R, s_S, s_byte[buf], s_out, cnt);           // the serial chain: zl_mtf_walk.h
            __syncwarp();
            for (int q = lane; q < cnt; q += 32) {
                const uint32_t rec = s_rec[buf][q];
                tok[rec >> 8] = (uint32_t) s_out[q] | ((uint32_t) ctx << 10) | ((rec & 0xffu) << 22);
            }
            __syncwarp();
        }
    }
    __syncwarp();
    for (int i = lane; i < 256; i += 32) state_out[ctx * 256 + i] = (uint8_t) s_S[i];
}

// =====================================================================================================
// Huffman tables
// =====================================================================================================
// libstdc++ binary-heap emulation on index arrays (stl_heap.h __push_heap/__adjust_heap/
// __pop_heap/__make_heap with a weight-only "greater" comparator, src/libzling_huffman.cpp:63-67,82-92).
struct HeapWork {
    int32_t* w;      // node weights [2*nsym]
    int16_t* kid;    // children     [2*nsym][2]
    int16_t* heap;   // heap of node ids [nsym]
    int      n;
};
__device__ __forceinline__ void heap_sift_up(HeapWork& h, int hole, int top, int v) {
    int parent = (hole - 1) / 2;
    const int wv = h.w[v];
    while (hole > top && h.w[h.heap[parent]] > wv) {
        h.heap[hole] = h.heap[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h.heap[hole] = (int16_t) v;
}
__device__ __forceinline__ void heap_adjust(HeapWork& h, int hole, int len, int v) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (h.w[h.heap[child]] > h.w[h.heap[child - 1]]) child--;
        h.heap[hole] = h.heap[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h.heap[hole] = h.heap[child - 1];
        hole = child - 1;
    }
    heap_sift_up(h, hole, top, v);
}
__device__ __forceinline__ int heap_pop(HeapWork& h) {
    const int top = h.heap[0];
    if (h.n > 1) {
        const int last = h.heap[h.n - 1];
        h.heap[h.n - 1] = (int16_t) top;
        heap_adjust(h, 0, h.n - 1, last);
    }
    h.n--;
    return top;
}
__device__ __forceinline__ void heap_push(HeapWork& h, int v) {
    h.heap[h.n] = (int16_t) v;
    h.n++;
    heap_sift_up(h, h.n - 1, 0, v);
}

// ZlingMakeLengthTable (huffman.cpp:41-112), run by ONE thread on shared-memory work arrays
__device__ void make_length_table(const uint32_t* freq, uint8_t* len, int nsym, int cap, HeapWork h, int16_t* leafsym, uint8_t* depth) {
    for (int s = 0; s < nsym; s++) len[s] = 0;
    for (int shift = 0;; shift++) {
        int nleaf = 0;
        for (int s = 0; s < nsym; s++) {
            if (freq[s] > 0) {
                h.w[nleaf] = (int32_t) ((freq[s] + ((1u << shift) - 1u)) >> shift);
                leafsym[nleaf] = (int16_t) s;
                h.heap[nleaf] = (int16_t) nleaf;
                nleaf++;
            }
        }
        if (nleaf == 0) return;
        h.n = nleaf;
        if (nleaf >= 2) {
            for (int parent = (nleaf - 2) / 2; parent >= 0; parent--) heap_adjust(h, parent, nleaf, h.heap[parent]);
        }
        int nnode = nleaf;
        while (h.n > 1) {
            const int x = heap_pop(h);
            const int y = heap_pop(h);
            h.w[nnode] = h.w[x] + h.w[y];
            h.kid[2 * nnode] = (int16_t) x;
            h.kid[2 * nnode + 1] = (int16_t) y;
            heap_push(h, nnode);
            nnode++;
        }
        // children always have smaller ids than their parent: one backward sweep assigns depths
        int longest = 0;
        depth[nnode - 1] = 0;
        for (int node = nnode - 1; node >= nleaf; node--) {
            const uint8_t d = (uint8_t) (depth[node] + 1);
            depth[h.kid[2 * node]] = d;
            depth[h.kid[2 * node + 1]] = d;
        }
        for (int i = 0; i < nleaf; i++) {
            const int l = depth[i] > 1 ? depth[i] : 1;                   // huffman.cpp:97
            len[leafsym[i]] = (uint8_t) l;
            longest = max(longest, l);
        }
        if (longest <= cap) return;                                      // else halve weights and rebuild, :107-110
    }
}

// ZlingMakeEncodeTable (huffman.cpp:114-138), all threads of the CTA: canonical code = first code of its
// length + number of smaller symbols with the same length, then bit-reversed for LSB-first emission
__device__ void make_encode_table(const uint8_t* len, uint16_t* code, int nsym, int cap, uint32_t* s_first /*[16]*/) {
    if (threadIdx.x == 0) {
        uint32_t count[16];
        for (int l = 0; l < 16; l++) count[l] = 0;
        for (int s = 0; s < nsym; s++) count[len[s]]++;
        uint32_t next = 0;
        for (int l = 1; l <= cap; l++) { s_first[l] = next; next = (next + count[l]) << 1; }
        s_first[0] = 0;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < nsym; s += blockDim.x) {
        const int l = len[s];
        uint32_t c = 0;
        if (l) {
            int before = 0;
            for (int q = 0; q < s; q++) before += (len[q] == l);
            c = (s_first[l] + before) & 0xffffu;
            c = __brev(c) >> (32 - l);
        }
        code[s] = (uint16_t) c;
    }
    __syncthreads();
}

constexpr int kBuildThreads = 256;

// grid (kMaxSubPerBlock, nblocks): one CTA per sub-block
__global__ void __launch_bounds__(kBuildThreads) zl_huff_build_kernel(const uint32_t* tok_all, SubBlock* sub_all, const uint32_t* nsub,
                                                                    const uint8_t* active, HuffTables* tab_all) {
    const int b = blockIdx.y, j = blockIdx.x;
    if (!active[b] || j >= (int) nsub[b]) return;
    __shared__ uint32_t s_f1[kSyms1 + 2], s_f2[kSyms2];
    __shared__ uint8_t  s_l1[kSyms1 + 2], s_l2[kSyms2];
    __shared__ uint16_t s_c1[kSyms1 + 2], s_c2[kSyms2];
    __shared__ int32_t  s_w1[2 * kSyms1], s_w2[2 * kSyms2];
    __shared__ int16_t  s_kid1[4 * kSyms1], s_kid2[4 * kSyms2];
    __shared__ int16_t  s_heap1[kSyms1], s_heap2[kSyms2], s_leaf1[kSyms1], s_leaf2[kSyms2];
    __shared__ uint8_t  s_depth1[2 * kSyms1], s_depth2[2 * kSyms2];
    __shared__ uint32_t s_first1[16], s_first2[16];
    __shared__ unsigned long long s_bits;

    SubBlock* sb = sub_all + (size_t) b * kMaxSubPerBlock + j;
    const uint32_t* tok = tok_all + (size_t) b * kTokStride;
    const int t0 = sb->tok_begin, t1 = sb->tok_end;
    for (int i = threadIdx.x; i < kSyms1 + 2; i += blockDim.x) s_f1[i] = 0;
    if (threadIdx.x < kSyms2) s_f2[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_bits = 0;
    __syncthreads();
    for (int i = t0 + threadIdx.x; i < t1; i += blockDim.x) {            // libzling.cpp:219-224
        const uint32_t t = tok[i];
        const uint32_t s = tok_sym(t);
        atomicAdd(&s_f1[s], 1u);
        if (s >= 258) atomicAdd(&s_f2[idx_bucket((int) tok_aux(t))], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        HeapWork h; h.w = s_w1; h.kid = s_kid1; h.heap = s_heap1; h.n = 0;
        make_length_table(s_f1, s_l1, kSyms1, kCap1, h, s_leaf1, s_depth1);
        s_l1[kSyms1] = 0; s_l1[kSyms1 + 1] = 0;
    } else if (threadIdx.x == 32) {
        HeapWork h; h.w = s_w2; h.kid = s_kid2; h.heap = s_heap2; h.n = 0;
        make_length_table(s_f2, s_l2, kSyms2, kCap2, h, s_leaf2, s_depth2);
    }
    __syncthreads();
    make_encode_table(s_l1, s_c1, kSyms1, kCap1, s_first1);
    make_encode_table(s_l2, s_c2, kSyms2, kCap2, s_first2);
    // payload size: sum of freq * code length (+ extra bits of the idx buckets)
    unsigned long long bits = 0;
    for (int s = threadIdx.x; s < kSyms1; s += blockDim.x) bits += (unsigned long long) s_f1[s] * s_l1[s];
    if (threadIdx.x < kSyms2) bits += (unsigned long long) s_f2[threadIdx.x] * (s_l2[threadIdx.x] + idx_extra_bits(threadIdx.x));
    for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_bits, bits);
    __syncthreads();
    HuffTables* tab = tab_all + (size_t) b * kMaxSubPerBlock + j;
    for (int s = threadIdx.x; s < kSyms1 + 2; s += blockDim.x) { tab->len1[s] = s < kSyms1 ? s_l1[s] : 0; tab->code1[s] = s < kSyms1 ? s_c1[s] : 0; }
    if (threadIdx.x < kSyms2) { tab->len2[threadIdx.x] = s_l2[threadIdx.x]; tab->code2[threadIdx.x] = s_c2[threadIdx.x]; }
    if (threadIdx.x == 0) {
        sb->bits_lo = (uint32_t) s_bits;
        sb->olen = kTableBytes + (uint32_t) ((s_bits + 7) >> 3);
    }
}

// test hook: tables only, freq supplied by the host (grid = ntables)
__global__ void __launch_bounds__(kBuildThreads) zl_huff_tables_only_kernel(const uint32_t* freq_all, int nsym, int cap, uint8_t* len_all, uint16_t* code_all) {
    __shared__ uint32_t s_f[kSyms1];
    __shared__ uint8_t  s_l[kSyms1];
    __shared__ uint16_t s_c[kSyms1];
    __shared__ int32_t  s_w[2 * kSyms1];
    __shared__ int16_t  s_kid[4 * kSyms1], s_heap[kSyms1], s_leaf[kSyms1];
    __shared__ uint8_t  s_depth[2 * kSyms1];
    __shared__ uint32_t s_first[16];
    const uint32_t* freq = freq_all + (size_t) blockIdx.x * nsym;
    for (int i = threadIdx.x; i < nsym; i += blockDim.x) s_f[i] = freq[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        HeapWork h; h.w = s_w; h.kid = s_kid; h.heap = s_heap; h.n = 0;
        make_length_table(s_f, s_l, nsym, cap, h, s_leaf, s_depth);
    }
    __syncthreads();
    make_encode_table(s_l, s_c, nsym, cap, s_first);
    for (int i = threadIdx.x; i < nsym; i += blockDim.x) {
        len_all[(size_t) blockIdx.x * nsym + i] = s_l[i];
        code_all[(size_t) blockIdx.x * nsym + i] = s_c[i];
    }
}

// =====================================================================================================
// bit packing + framing
// =====================================================================================================
constexpr int kPackThreads = 512;
constexpr int kPackPerThread = 8;
constexpr int kPackChunk = kPackThreads * kPackPerThread;           // tokens per round
constexpr int kPackWords = kPackChunk + 8;                          // <= 31 bits per token < 1 word per token

// grid (kMaxSubPerBlock, nblocks).  Writes, at out + out_off[b][j]:
//   01 | BE32 encpos | BE32 rlen | BE32 olen | 273 table bytes | LSB-first code bits (zero padded), and the
//   block's 00 stop flag after its last sub-block (libzling.cpp:200,232-257,269-278).
__global__ void __launch_bounds__(kPackThreads) zl_huff_pack_kernel(const uint32_t* tok_all, const SubBlock* sub_all, const uint32_t* nsub,
                                                                  const HuffTables* tab_all, const unsigned long long* out_off, uint8_t* out) {
    const int b = blockIdx.y, j = blockIdx.x;
    const int ns = (int) nsub[b];
    if (j >= ns) return;
    __shared__ uint16_t s_c1[kSyms1 + 2], s_c2[kSyms2];
    __shared__ uint8_t  s_l1[kSyms1 + 2], s_l2[kSyms2];
    __shared__ uint32_t s_buf[kPackWords];
    __shared__ uint32_t s_warp[kPackThreads / 32];
    __shared__ uint32_t s_total;

    const SubBlock sb = sub_all[(size_t) b * kMaxSubPerBlock + j];
    const HuffTables* tab = tab_all + (size_t) b * kMaxSubPerBlock + j;
    const uint32_t* tok = tok_all + (size_t) b * kTokStride;
    uint8_t* dst = out + out_off[(size_t) b * kMaxSubPerBlock + j];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int s = tid; s < kSyms1 + 2; s += blockDim.x) { s_c1[s] = tab->code1[s]; s_l1[s] = tab->len1[s]; }
    if (tid < kSyms2) { s_c2[tid] = tab->code2[tid]; s_l2[tid] = tab->len2[tid]; }
    for (int i = tid; i < kPackWords; i += blockDim.x) s_buf[i] = 0;
    if (tid == 0) {
        dst[0] = 1;
        const uint32_t v[3] = { sb.enc_end, sb.rlen, sb.olen };
        for (int q = 0; q < 3; q++) { dst[1 + 4 * q] = v[q] >> 24; dst[2 + 4 * q] = v[q] >> 16; dst[3 + 4 * q] = v[q] >> 8; dst[4 + 4 * q] = v[q]; }
        if (j == ns - 1) dst[13 + sb.olen] = 0;
    }
    __syncthreads();
    for (int i = tid; i < kTableBytes; i += blockDim.x) {                // libzling.cpp:232-237
        dst[13 + i] = i < 257 ? (uint8_t) (s_l1[2 * i] << 4 | s_l1[2 * i + 1]) : (uint8_t) (s_l2[2 * (i - 257)] << 4 | s_l2[2 * (i - 257) + 1]);
    }
    uint8_t* bitsdst = dst + 13 + kTableBytes;
    uint32_t carry_bits = 0;        // bits already sitting in s_buf[0] from the previous round (0..31)
    size_t   flushed = 0;           // bytes already written to bitsdst

    for (int base = sb.tok_begin; base < (int) sb.tok_end; base += kPackChunk) {
        uint32_t val[kPackPerThread]; uint8_t nb[kPackPerThread];
        uint32_t mine = 0;
        #pragma unroll
        for (int q = 0; q < kPackPerThread; q++) {
            const int i = base + tid * kPackPerThread + q;
            uint32_t v = 0, n = 0;
            if (i < (int) sb.tok_end) {
                const uint32_t t = tok[i];
                const uint32_t s = tok_sym(t);
                v = s_c1[s]; n = s_l1[s];
                if (s >= 258) {                                          // libzling.cpp:243-246
                    const int idx = (int) tok_aux(t), bk = idx_bucket(idx);
                    v |= (uint32_t) s_c2[bk] << n; n += s_l2[bk];
                    v |= (uint32_t) (idx - idx_base(bk)) << n; n += idx_extra_bits(bk);
                }
            }
            val[q] = v; nb[q] = (uint8_t) n; mine += n;
        }
        // exclusive scan of per-thread bit counts over the CTA
        uint32_t incl = mine;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t wv = lane < kPackThreads / 32 ? s_warp[lane] : 0, wi = wv;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
            if (lane < kPackThreads / 32) s_warp[lane] = wi - wv;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        uint32_t pos = carry_bits + s_warp[wid] + incl - mine;
        #pragma unroll
        for (int q = 0; q < kPackPerThread; q++) {
            if (nb[q]) {
                const uint32_t wd = pos >> 5, sh = pos & 31;
                atomicOr(&s_buf[wd], val[q] << sh);
                if (sh + nb[q] > 32) atomicOr(&s_buf[wd + 1], val[q] >> (32 - sh));
                pos += nb[q];
            }
        }
        __syncthreads();
        const uint32_t have = carry_bits + s_total;       // valid bits in s_buf
        const uint32_t full = have >> 5;                   // complete words to flush
        for (uint32_t i = tid; i < full * 4; i += blockDim.x) bitsdst[flushed + i] = (uint8_t) (s_buf[i >> 2] >> ((i & 3) * 8));
        __syncthreads();
        const uint32_t tail = s_buf[full];
        __syncthreads();
        for (uint32_t i = tid; i <= full; i += blockDim.x) s_buf[i] = 0;
        __syncthreads();
        if (tid == 0) s_buf[0] = tail;
        flushed += (size_t) full * 4;
        carry_bits = have & 31;
        __syncthreads();
    }
    // tail: remaining bits, zero padded to a byte (libzling.cpp:255-257)
    const uint32_t rest = (carry_bits + 7) >> 3;
    if (tid < rest) bitsdst[flushed + tid] = (uint8_t) (s_buf[0] >> (tid * 8));
}

}  // namespace zl

// =====================================================================================================
// decode side
//   zl_huff_decode   nibble tables -> canonical codes -> LUT decode      src/libzling.cpp:347-402,
//                                                                        src/libzling_huffman.cpp:114-153
//   zl_rolz_decode   ZlingRolzDecoder::Decode/GetMatchAndUpdate + MTF     src/libzling_lz.cpp:119-126,318-399
// =====================================================================================================
namespace zl {

constexpr int kDecodeThreads = 256;

// grid = number of sub-blocks in the call; sub-blocks are independently decodable (explicit olen/rlen).
// One thread walks the bit stream (the code lengths make it a serial chain); the CTA builds the tables.
// status[s] = 0 ok, else the reference's throw site (1 bad code1, 2 bad code2, 3 bad extra bits / truncated).
__global__ void __launch_bounds__(kDecodeThreads) zl_huff_decode_kernel(const uint8_t* comp, const DecSub* subs, uint16_t* sym_all, int* status) {
    extern __shared__ uint16_t s_lut1[];                      // 1 << 15 entries
    __shared__ uint16_t s_lut2[1 << kCap2];
    __shared__ uint16_t s_c1[kSyms1 + 2], s_c2[kSyms2];
    __shared__ uint8_t  s_l1[kSyms1 + 2], s_l2[kSyms2];
    __shared__ uint32_t s_first1[16], s_first2[16];
    const DecSub sb = subs[blockIdx.x];
    const uint8_t* pl = comp + sb.payload_off;
    const int tid = threadIdx.x;

    for (int i = tid; i < 257; i += blockDim.x) { s_l1[2 * i] = pl[i] >> 4; s_l1[2 * i + 1] = pl[i] & 15; }   // libzling.cpp:347-351
    if (tid < 16) { s_l2[2 * tid] = pl[257 + tid] >> 4; s_l2[2 * tid + 1] = pl[257 + tid] & 15; }
    for (int i = tid; i < (1 << kCap1); i += blockDim.x) s_lut1[i] = 0xffff;
    for (int i = tid; i < (1 << kCap2); i += blockDim.x) s_lut2[i] = 0xffff;
    __syncthreads();
    make_encode_table(s_l1, s_c1, kSyms1, kCap1, s_first1);
    make_encode_table(s_l2, s_c2, kSyms2, kCap2, s_first2);
    for (int s = tid; s < kSyms1; s += blockDim.x) {          // ZlingMakeDecodeTable, huffman.cpp:140-153
        const int l = s_l1[s];
        if (l > 0) for (int i = s_c1[s]; i < (1 << kCap1); i += 1 << l) s_lut1[i] = (uint16_t) s;
    }
    if (tid < kSyms2) {
        const int l = s_l2[tid];
        if (l > 0 && l <= kCap2) for (int i = s_c2[tid]; i < (1 << kCap2); i += 1 << l) s_lut2[i] = (uint16_t) tid;
    }
    __syncthreads();
    if (tid != 0) return;

    uint16_t* sym = sym_all + (size_t) sb.block * (2 * kTokStride) + sb.sym_off;
    unsigned long long acc = 0;
    int nbits = 0, err = 0;
    uint32_t rp = kTableBytes;
    for (uint32_t i = 0; i < sb.rlen; i++) {                   // libzling.cpp:368-402
        if (nbits < 32) {
            uint32_t w = 0;
            #pragma unroll
            for (int q = 0; q < 4; q++) { if (rp + q < sb.olen) w |= (uint32_t) pl[rp + q] << (8 * q); }
            rp += 4;
            acc |= (unsigned long long) w << nbits;
            nbits += 32;
        }
        const uint32_t s = s_lut1[(uint32_t) acc & ((1u << kCap1) - 1)];
        if (s >= (uint32_t) kSyms1) { err = 1; break; }
        acc >>= s_l1[s]; nbits -= s_l1[s];
        sym[i] = (uint16_t) s;
        if (s >= 258) {
            const uint32_t bk = s_lut2[(uint32_t) acc & 0xffu];
            if (bk >= (uint32_t) kSyms2) { err = 2; break; }
            acc >>= s_l2[bk]; nbits -= s_l2[bk];
            const int eb = idx_extra_bits((int) bk);
            const uint32_t idx = (uint32_t) idx_base((int) bk) + ((uint32_t) acc & ((1u << eb) - 1u));
            acc >>= eb; nbits -= eb;
            if (idx >= (uint32_t) kRing || i + 1 >= sb.rlen) { err = 3; break; }
            sym[++i] = (uint16_t) idx;
        }
    }
    status[blockIdx.x] = err;
}

// symbol i of a sub-block through a 32-entry window held one-per-lane (coalesced refill every 32 symbols)
__device__ __forceinline__ int warp_sym(const uint16_t* sym, int rlen, int i, int& wbase, uint32_t& window, int lane) {
    if (i - wbase >= 32) {
        wbase = i & ~31;
        window = wbase + lane < rlen ? sym[wbase + lane] : 0;
    }
    return (int) __shfl_sync(0xffffffffu, window, i & 31);
}

// One warp per stream; control flow is warp-uniform, lane 0 owns the scalar stores, all lanes copy matches.
// Serial across blocks as well (the MTF tables are stream-lifetime state, src/libzling_lz.h:137).
// result[0] = 0 ok / 4 = lz decode failed (libzling.cpp:406-408); block_len[b] = decoded bytes of block b.
// grid = streams of the call (DecRange per stream: its sub-block records, MTF tables, result slot); every stream has its own
// offset ring at ring_all + stream * 256 * kRing.
__global__ void __launch_bounds__(32) zl_rolz_decode_kernel(const DecSub* subs_all, const DecRange* ranges, const uint16_t* sym_all, uint8_t* out,
                                                            uint32_t* ring_all, uint32_t* block_len, int* result_all) {
    extern __shared__ uint8_t s_sym[];                         // [ctx][rank] -> byte, 64 KiB
    __shared__ uint16_t s_head[256];
    __shared__ uint32_t s_mru[256];
    const int lane = threadIdx.x;
    const DecRange rg = ranges[blockIdx.x];
    const DecSub* subs = subs_all + rg.s0;
    const int nsubs = rg.s1 - rg.s0;
    uint32_t* ring = ring_all + (size_t) blockIdx.x * 256 * kRing;
    uint8_t* mtf_state = rg.state;
    int* result = result_all + blockIdx.x;
    {
        const uint4* src = reinterpret_cast<const uint4*>(mtf_state);
        uint4* dst = reinterpret_cast<uint4*>(s_sym);
        for (int i = lane; i < 65536 / 16; i += 32) dst[i] = src[i];
    }
    int cur_block = -1, op = 0, fail = 0;
    uint8_t* blk = out;
    uint32_t p1 = 0, p2 = 0, p3 = 0;                           // last three decoded bytes (p1 = most recent)
    for (int si = 0; si < nsubs && !fail; si++) {
        const DecSub sb = subs[si];
        if ((int) sb.block != cur_block) {                     // Reset(), lz.cpp:378-386
            if (cur_block >= 0 && lane == 0) block_len[cur_block] = (uint32_t) op;
            cur_block = (int) sb.block;
            blk = out + (size_t) cur_block * kBlockBytes;
            uint4* r = reinterpret_cast<uint4*>(ring);
            for (int i = lane; i < 256 * kRing / 4; i += 32) r[i] = make_uint4(0, 0, 0, 0);
            for (int i = lane; i < 256; i += 32) s_head[i] = 0;
            op = 0;
        }
        for (int i = lane; i < 256; i += 32) s_mru[i] = 0;     // lz.cpp:324
        __syncwarp();
        const uint16_t* sym = sym_all + (size_t) sb.block * (2 * kTokStride) + sb.sym_off;
        const int rlen = (int) sb.rlen, encpos = (int) sb.encpos;
        int ip = 0;
        uint32_t window = 0;                                   // 32 symbols held across the warp
        int wbase = -32;
        #define ZL_SYM(i_) warp_sym(sym, rlen, (i_), wbase, window, lane)
        for (int first = 0; first < 2; first++) {              // lz.cpp:327-328
            if (op == first && ip < rlen) {
                const int v = ZL_SYM(ip) & 0xff; ip++;
                if (lane == 0) blk[op] = (uint8_t) v;
                p3 = p2; p2 = p1; p1 = (uint32_t) v; op++;
            }
        }
        while (ip < rlen) {
            const int s = ZL_SYM(ip); ip++;
            if (s < 258 && op + (s >= 256 ? 2 : 1) > encpos) { fail = 4; break; }   // would overrun: lz.cpp:366-368
            const int c = (int) p1;
            const int head = (s_head[c] + 1) & (kRing - 1);    // every token start is inserted, lz.cpp:388-399
            __syncwarp();
            if (lane == 0) { s_head[c] = (uint16_t) head; ring[(size_t) c * kRing + head] = (uint32_t) op; }
            if (s < 256) {                                     // literal: MTF decode, lz.cpp:122-126,332-337
                uint8_t* sy = s_sym + c * 256;
                const int byte = sy[s], jn = mtf_next(s);
                const int other = sy[jn];
                __syncwarp();
                if (lane == 0) { sy[s] = (uint8_t) other; sy[jn] = (uint8_t) byte; blk[op] = (uint8_t) byte; }
                op++;
                p3 = p2; p2 = p1; p1 = (uint32_t) byte;
                const uint32_t m = s_mru[p3];
                __syncwarp();
                if (lane == 0) s_mru[p3] = ((p2 << 8) | p1) | (m << 16);
                __syncwarp();
            } else if (s < 258) {                              // word MRU hit, lz.cpp:339-352
                const uint32_t m = s_mru[c];
                const uint32_t w = s == 256 ? (m & 0xffffu) : (m >> 16);
                if (lane == 0) { blk[op] = (uint8_t) (w >> 8); blk[op + 1] = (uint8_t) w; }
                op += 2;
                p3 = p1; p2 = w >> 8; p1 = w & 0xffu;
                __syncwarp();
                if (s == 257 && lane == 0) s_mru[c] = w | (m << 16);
                __syncwarp();
            } else {                                           // match, lz.cpp:354-366
                const int len = s - 258 + kMinLen;
                if (ip >= rlen) { fail = 4; break; }
                const int idx = ZL_SYM(ip); ip++;
                __syncwarp();
                const uint32_t from = ring[(size_t) c * kRing + ((head - idx) & (kRing - 1))];
                if (op + len > encpos || from >= (uint32_t) op) { fail = 4; break; }
                const uint32_t dist = (uint32_t) op - from;
                // byte-serial LZ copy semantics (IncrementalCopyFastPath, lz.cpp:91-104): byte k comes from
                // from + (k mod dist), all of which were written by earlier tokens
                uint32_t t1 = 0, t2 = 0, t3 = 0;
                for (int kq = lane; kq < len; kq += 32) {
                    const uint8_t v = blk[from + ((uint32_t) kq % dist)];
                    blk[op + kq] = v;
                    if (kq == len - 1) t1 = v;
                    if (kq == len - 2) t2 = v;
                    if (kq == len - 3) t3 = v;
                }
                p1 = __shfl_sync(0xffffffffu, t1, (len - 1) & 31);
                p2 = __shfl_sync(0xffffffffu, t2, (len - 2) & 31);
                p3 = __shfl_sync(0xffffffffu, t3, (len - 3) & 31);
                op += len;
                const uint32_t w = (p2 << 8) | p1;
                const uint32_t m = s_mru[p3];
                __syncwarp();
                if (lane == 0 && (m & 0xffffu) != w) s_mru[p3] = w | (m << 16);
                __syncwarp();
            }
            if (op > encpos) { fail = 4; break; }              // lz.cpp:366-368
        }
        #undef ZL_SYM
        if (!fail && op != encpos) fail = 4;                   // lz.cpp:371-373
    }
    __syncwarp();
    if (lane == 0) {
        if (cur_block >= 0) block_len[cur_block] = (uint32_t) op;
        result[0] = fail;
    }
    if (!fail) {
        uint4* dst = reinterpret_cast<uint4*>(mtf_state);
        const uint4* src = reinterpret_cast<const uint4*>(s_sym);
        for (int i = lane; i < 65536 / 16; i += 32) dst[i] = src[i];
    }
}

}  // namespace zl
