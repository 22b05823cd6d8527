// zl_rolz_parse_v4 — the ROLZ parse as a data-parallel fixed-point iteration over windows of 1024 positions.
//
// The reference's parse (EncodeImpl + MatchAndUpdate + MatchLazy, src/libzling_lz.cpp:139-316) is one serial chain
// per 16 MiB block: whether position p is a token start, and what the bucket holds when p is probed, depend on
// every earlier decision.  v3 walked that chain with ONE thread per block (470 cycles per token).  v4 keeps v3's
// exactness argument — records against a frozen bucket state G plus the inserts still pending in the window — and
// replaces the walker with a Jacobi iteration in which every step is parallel over the window's positions:
//
//   SPEC      every position x of the window (one thread each): key, link to the nearest earlier same-key position
//             of the window, the hash chain of x in G (G = all inserts before the window; up to dmax nodes with match
//             lengths), and the decision the reference takes if no insert of this window interferes ("frozen").
//   ROUNDS    state = one decision per position (kind, length).  Each round: (1) the token starts M are the orbit of
//             the window's entry position under "next = x + step(decision)" (pointer doubling); (2) per-context
//             ranks of the marked positions give every pending insert its ring slot; (3) every position re-derives
//             its decision from M and the decisions: in-window candidates (marked same-key positions, newest first)
//             merged with the frozen record in the reference's visiting order, ring slots overwritten since the
//             freeze, the lazy probes at x+1 / x+2, the word-MRU state at x (folded from the pushes of the marked
//             token ends), the sub-block roll-over.  The iteration stops when no marked position changed its
//             decision: by induction over the token order the state is then exactly the reference's parse (the first
//             token's decision depends on carried state only, every later one on earlier tokens only).
//   FINALIZE  ring/hash writes of the window's inserts into G, token words and literal list from the marks (prefix
//             sums), carried word-MRU / insert counters / symbol count.
//
// Every function of the algorithm is scalar ZL_HD code; tests/cxx/parse_v4_sim.cu replays the phases on the host
// against the CPU checker (the GPU then has to confirm the kernel's synchronisation and its parallel forms of the
// ordered passes: link building, ranks, orbit, prefix sums).
#pragma once
#include "zl_kernels.cuh"

namespace zl {

#ifndef ZL_V4_T
#define ZL_V4_T 1024
#endif
constexpr int kV4T       = ZL_V4_T;             // threads per CTA = positions with a record per window (a multiple of 128)
constexpr int kV4N       = kV4T;                // positions with a record: the main ones + two lazy look-ahead positions
constexpr int kV4W       = kV4N - 2;            // main positions per window (the window stride)
constexpr int kV4R       = 2048;                // byte ring: >= 4 + kV4N + 264 + 16
constexpr int kV4Tail    = 288;                 // bytes staged past the last look-ahead position
constexpr int kV4Buckets = 4096;                // buckets of the link builder
constexpr int kV4Words   = (kV4N + 2 + 31) / 32 + 1;  // bitset words over window positions (+2: contexts of the last positions; +1 word: funnel reads)
constexpr int kV4PfWords = 128;                 // 4096-bit Bloom bitmap of the pushed (context, word) pairs
constexpr uint32_t kV4KeyInvalid = 0x80000000u; // position cannot be probed (first two bytes / last 273 bytes of the block)
constexpr uint32_t kV4KeyMask    = 0x1fffffu;   // (context << 13) | hash slot
constexpr uint32_t kV4Auto       = 0xffu;       // plan entry: predict the level (see v4_next_level)

// decision word: len(9) | kind(3) << 9 | ref(13) << 12; ref = the best candidate of a match: a ring slot of G, or
// kV4RefWin | rel of a position of this window whose insert is still pending (its slot is known once the ranks are)
constexpr uint32_t kV4Match = 1, kV4Lit = 2, kV4Word0 = 3, kV4Word1 = 4;
constexpr uint32_t kV4DecCmp = 0xfffu;          // the part of a decision that defines the parse (the match idx is derived)
constexpr uint32_t kV4RefWin = 0x1000u;
ZL_HD uint32_t v4_dec_len(uint32_t d)  { return d & 511u; }
ZL_HD uint32_t v4_dec_kind(uint32_t d) { return (d >> 9) & 7u; }
ZL_HD uint32_t v4_dec_ref(uint32_t d)  { return (d >> 12) & 0x1fffu; }
ZL_HD uint32_t v4_dec_step(uint32_t d) { const uint32_t k = v4_dec_kind(d); return k == kV4Match ? v4_dec_len(d) : (k == kV4Lit ? 1u : 2u); }
ZL_HD uint32_t v4_dec_syms(uint32_t d) { return v4_dec_kind(d) == kV4Match ? 2u : 1u; }
// fx word: frozen slot head (16) | flags << 16
constexpr uint32_t kV4F_SELF = 1u << 16, kV4F_L1 = 1u << 17, kV4F_L2 = 1u << 18, kV4F_ST = 1u << 19, kV4F_STL = 1u << 20, kV4F_HAZ = 0xfu << 16;   // ST: a slot read by the records of x / x+1 / x+2 may be overwritten inside the window

// host-side statistics of the replay (tests/cxx/parse_v4_sim.cu with -DZL_V4_STATS): loop trip counts of the serial helpers
#if defined(ZL_V4_STATS)
struct V4Stats { unsigned long long calls[8], steps[8], maxsteps[8], hist[8][16]; };
static V4Stats g_v4stats;
static inline void v4_stat(int k, unsigned long long n) {
    g_v4stats.calls[k]++; g_v4stats.steps[k] += n; if (n > g_v4stats.maxsteps[k]) g_v4stats.maxsteps[k] = n;
    int b = 0; while ((1ull << b) <= n && b < 15) b++; g_v4stats.hist[k][b]++;
}
#endif
#if defined(ZL_V4_STATS) && !defined(__CUDA_ARCH__)
#define V4_STAT(k, n) v4_stat(k, (unsigned long long) (n))
#else
#define V4_STAT(k, n) do { } while (0)
#endif
// ---- portable intrinsics ---------------------------------------------------------------------------------------------
ZL_HD uint32_t z4_funnel(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
ZL_HD int z4_ffs(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int) v);
#else
    return __builtin_ffs((int) v);
#endif
}
ZL_HD int z4_clz(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __clz((int) v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
ZL_HD int z4_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
ZL_HD uint64_t z4_ld_ring(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}
ZL_HD uint32_t z4_ld_hash(const uint16_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
ZL_HD uint32_t z4_ld_in32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
ZL_HD uint4 z4_ld_in128(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// unaligned little-endian 32-bit load from the input block (global memory); `in` is 16-byte aligned
ZL_HD uint32_t z4_in32(const uint8_t* in, uint32_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in) + (off >> 2);
    return z4_funnel(z4_ld_in32(w), z4_ld_in32(w + 1), (off & 3u) * 8u);
}
ZL_HD uint32_t z4_hash(uint32_t w) { return w + ((w >> 16) & 0xffu) * 137u + (w >> 24) * 13337u; }   // lz.cpp:55-57

// ---- shared-memory layout ----------------------------------------------------------------------------------------------
struct V4Layout {
    int dmax, lmax;
    int rb, key, link, blink, pcnt, occw, mcnt, pushw, pf, hdr, node, nodeq, fdec, fx, rank, mark, plit, sup, dec, ndec, occ, mbits, cnt, mru, mru2, scratch, total;
    int scratch_bytes;
};
// scratch is a union: SPEC uses it for the link builder's bucket tables, the rounds for the orbit / rank tables
constexpr int kV4Groups      = 8;                                          // link builder: position groups with their own bucket table
constexpr int kV4ScratchSpec = kV4Groups * kV4Buckets * 2;
constexpr int kV4ScratchRnd  = 4096 + 33 * 256 * 2;
__host__ __device__ inline V4Layout v4_layout(int dmax, int lmax) {
    V4Layout L; L.dmax = dmax; L.lmax = lmax;
    int at = 0;
    auto take = [&](int bytes) { int o = at; at += (bytes + 15) & ~15; return o; };
    L.rb    = take(kV4R);
    L.key   = take(4 * kV4N);
    L.link  = take(2 * kV4N);
    L.blink = take(2 * kV4N);
    L.pcnt  = take(4 * 256);
    L.occw  = take(4 * 256);
    L.mcnt  = take(4 * 256);
    L.pushw = take(4 * 256);
    L.pf    = take(4 * kV4PfWords);
    L.hdr   = take(4 * kV4N);
    L.node  = take(4 * kV4N * dmax);
    L.nodeq = take(4 * kV4N * lmax);
    L.fdec  = take(4 * kV4N);
    L.fx    = take(4 * kV4N);
    L.rank  = take(2 * kV4N);
    L.mark  = take(kV4N + 2);
    L.plit  = take(kV4N + 2);
    L.sup   = take(kV4N + 2);
    L.dec   = take(4 * kV4N);
    L.ndec  = take(4 * kV4N);
    L.occ   = take(4 * 256 * kV4Words);
    L.mbits = take(4 * kV4Words);
    L.cnt   = take(4 * 256);
    L.mru   = take(4 * 256);
    L.mru2  = take(4 * 256);
    L.scratch_bytes = kV4ScratchSpec > kV4ScratchRnd ? kV4ScratchSpec : kV4ScratchRnd;
    L.scratch = take(L.scratch_bytes);
    L.total = at;
    return L;
}

struct V4Ctx {
    // block
    const uint8_t* in; int ilen;
    uint64_t* ring; uint16_t* hash;             // G: bucket state in global memory
    uint32_t* tok; uint32_t* lit; SubBlock* sub; const uint8_t* plan; int base_level;
    // shared memory (indexed by rel = x - lo unless noted)
    uint32_t* rbw;                              // input bytes: ring of kV4R bytes viewed as words, indexed by block position
    uint32_t* key; uint16_t* link; uint16_t* blink; uint32_t* pcnt; uint32_t* occw; uint32_t* mcnt; uint32_t* pushw; uint32_t* pf; uint32_t* hdr; uint32_t* node; uint32_t* nodeq;
    uint32_t* fdec; uint32_t* fx; uint16_t* rank; uint8_t* mark; uint8_t* plit; uint8_t* sup; uint32_t* dec; uint32_t* ndec;
    uint32_t* occ;                              // [256][kV4Words]: bit i of occ[c] <=> in[lo + i - 3] == c (a token END at lo + i pushes into context c;
                                                // position lo + i - 2 has context c)
    // pcnt[c] = bytes of value c among in[lo - 3 .. lo + N - 2] (the set bits of occ[c]): an upper bound of the window positions
    // whose context byte is c, i.e. of the inserts this window can make into context c
    // occw[c]: bit w set <=> word w of occ[c] is not empty (w < 32);  mcnt[c] = MARKED positions whose context byte is c (per round)
    // pushw[c] (per round): bit w set <=> a MARKED position in bitset word w pushes into context c: the only words a word-MRU scan visits
    // pf: Bloom bitmap (per round) of the (context, word) pairs pushed by the marked positions: a word test whose pair is neither
    //     in it nor in the carried MRU entry cannot hit (v4_decide_word), which spares the scan of the pushes
    uint32_t* mbits;                            // [kV4Words]: bit i <=> position lo + i is a token start
    uint32_t* cnt; uint32_t* mru; uint32_t* mru2;   // carried: inserts per context before the window, word MRU at the window's entry
    uint32_t* last;                             // host replay only: bucket table of the serial link builder
    int dmax, lmax;
    int coop;                                   // device only: 1 = the calling WARP evaluates ONE position together (all 32 lanes pass the same
                                                // arguments and follow the same control flow); the loops over bitset words / compare words of the
                                                // helpers below then run one word per lane.  0 = plain scalar code (the host replay, one position per thread)
};
__host__ __device__ inline void v4_bind(V4Ctx& c, uint8_t* smem, const V4Layout& L) {
    c.rbw = (uint32_t*) (smem + L.rb); c.key = (uint32_t*) (smem + L.key); c.link = (uint16_t*) (smem + L.link);
    c.blink = (uint16_t*) (smem + L.blink); c.pcnt = (uint32_t*) (smem + L.pcnt); c.occw = (uint32_t*) (smem + L.occw); c.mcnt = (uint32_t*) (smem + L.mcnt); c.pushw = (uint32_t*) (smem + L.pushw); c.pf = (uint32_t*) (smem + L.pf); c.hdr = (uint32_t*) (smem + L.hdr);
    c.node = (uint32_t*) (smem + L.node); c.nodeq = (uint32_t*) (smem + L.nodeq); c.fdec = (uint32_t*) (smem + L.fdec);
    c.fx = (uint32_t*) (smem + L.fx); c.rank = (uint16_t*) (smem + L.rank); c.mark = smem + L.mark; c.plit = smem + L.plit;
    c.sup = smem + L.sup; c.dec = (uint32_t*) (smem + L.dec); c.ndec = (uint32_t*) (smem + L.ndec);
    c.occ = (uint32_t*) (smem + L.occ); c.mbits = (uint32_t*) (smem + L.mbits);
    c.cnt = (uint32_t*) (smem + L.cnt); c.mru = (uint32_t*) (smem + L.mru); c.mru2 = (uint32_t*) (smem + L.mru2);
    c.last = nullptr;
    c.dmax = L.dmax; c.lmax = L.lmax;
    c.coop = 0;
}

// per-window constants (uniform over the CTA)
struct V4Win {
    int lo, wend;            // window = positions [lo, lo + kV4W); tokens start at x < wend = min(lo + kV4W, ilen - kGuard)
    int entry;               // first token start of the window (>= lo)
    int level;               // level in force at the entry = level the frozen decisions assume
    int rpos, level2;        // sub-block roll-over: the token at rpos opens a new sub-block parsed at level2 (rpos < 0: none)
    int skip_push;           // the entry has no word-MRU push (block start: the two raw bytes push nothing)
    int prev_lit;            // the token that ends at the entry is a literal
};
ZL_HD int v4_stage_hi(int k) { return (((k + 1) * kV4W + 2 + kV4Tail) + 15) & ~15; }   // bytes [.., hi) are staged once window k is prepared

// ---- byte ring ---------------------------------------------------------------------------------------------------------
ZL_HD uint32_t v4_rb32(const uint32_t* rbw, uint32_t pos) {          // unaligned LE 32-bit load at block position pos
    const uint32_t i = (pos >> 2) & (kV4R / 4 - 1);
    return z4_funnel(rbw[i], rbw[(i + 1) & (kV4R / 4 - 1)], (pos & 3u) * 8u);
}
ZL_HD uint32_t v4_rb8(const uint32_t* rbw, uint32_t pos) {
    return (rbw[(pos >> 2) & (kV4R / 4 - 1)] >> ((pos & 3u) * 8u)) & 0xffu;
}
// stage 16 input bytes at block offset src (multiple of 16, may be negative or past the block: zeros)
ZL_HD void v4_stage16(const V4Ctx& c, int src) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (src >= 0 && src < c.ilen) {
        v = z4_ld_in128(reinterpret_cast<const uint4*>(c.in + src));
        const int over = src + 16 - c.ilen;                              // bytes past the block end are staged as zeros
        if (over > 0) {
            uint32_t w[4] = { v.x, v.y, v.z, v.w };
            for (int b = 16 - over; b < 16; b++) w[b >> 2] &= ~(0xffu << ((b & 3) * 8));
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    uint32_t* d = c.rbw + (((uint32_t) src >> 2) & (kV4R / 4 - 1));
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}

// ---- SPEC A: key of a position --------------------------------------------------------------------------------------------
ZL_HD uint32_t v4_key_of(const V4Ctx& c, int x) {
    if (x < 2 || x + 273 >= c.ilen) return kV4KeyInvalid;
    const uint32_t h = z4_hash(v4_rb32(c.rbw, (uint32_t) x));
    const uint32_t ctx = v4_rb8(c.rbw, (uint32_t) x - 1);
    return (ctx << 13) | (h & (kSlots - 1)) | (((h >> 13) & 0xffu) << 21);
}
ZL_HD uint32_t v4_ctx_of(uint32_t key) { return (key >> 13) & 0xffu; }
ZL_HD uint32_t v4_bucket_of(uint32_t key) { return ((key & kV4KeyMask) * 2654435761u) >> 20; }   // 12 bits

// ---- SPEC B: nearest earlier position of the window in the same bucket (host form; the kernel builds the same chains
// with one bucket table per group of 128 positions) -----------------------------------------------------------------------
inline void v4_bucket_pass_serial(const V4Ctx& c) {
    for (int i = 0; i < kV4Buckets; i++) c.last[i] = 0;
    for (int rel = 0; rel < kV4N; rel++) {
        const uint32_t k = c.key[rel];
        uint32_t d = 0;
        if (!(k & kV4KeyInvalid)) {
            const uint32_t b = v4_bucket_of(k);
            const uint32_t prev = c.last[b];
            if (prev != 0) d = (uint32_t) rel - (prev - 1);
            c.last[b] = (uint32_t) rel + 1;
        }
        c.blink[rel] = (uint16_t) d;
    }
}

// exact GetCommonLength (lz.cpp:66-89) with both operands inside the byte ring: aligned words of both sides, eight at a time
ZL_HD int v4_common_len_ring(const uint32_t* rbw, uint32_t p, uint32_t q, int coop = 0) {
    if (v4_rb32(rbw, p) != v4_rb32(rbw, q)) return 0;
#if defined(__CUDA_ARCH__)
    if (coop) {                                                          // one word per lane, 128 bytes per step
        const uint32_t lane = threadIdx.x & 31u;
        for (uint32_t n = 4; n < 272; n += 128) {
            const uint32_t o = n + 4u * lane;
            const uint32_t d = o < 272u ? v4_rb32(rbw, p + o) ^ v4_rb32(rbw, q + o) : 0u;
            const uint32_t bal = __ballot_sync(0xffffffffu, d != 0);
            if (bal) {
                const int f = __ffs((int) bal) - 1;
                const uint32_t df = __shfl_sync(0xffffffffu, d, f);
                const int l = (int) n + 4 * f + ((__ffs((int) df) - 1) >> 3);
                return l < kMaxLen ? l : kMaxLen;
            }
        }
        return kMaxLen;
    }
#endif
    (void) coop;
    for (int n = 4; n < 272; n += 32) {
        const uint32_t pa = p + (uint32_t) n, qa = q + (uint32_t) n;
        const uint32_t shp = (pa & 3u) * 8u, shq = (qa & 3u) * 8u, pi = pa >> 2, qi = qa >> 2;
        uint32_t pw[9], qw[9];
        #pragma unroll
        for (int i = 0; i <= 8; i++) { pw[i] = rbw[(pi + (uint32_t) i) & (kV4R / 4 - 1)]; qw[i] = rbw[(qi + (uint32_t) i) & (kV4R / 4 - 1)]; }
        uint32_t fd = 0; int fi = 8;
        #pragma unroll
        for (int i = 7; i >= 0; i--) {
            const uint32_t d = z4_funnel(pw[i], pw[i + 1], shp) ^ z4_funnel(qw[i], qw[i + 1], shq);
            if (d) { fd = d; fi = i; }
        }
        if (fd) { const int l = n + 4 * fi + ((z4_ffs(fd) - 1) >> 3); return l < kMaxLen ? l : kMaxLen; }
    }
    return kMaxLen;
}
// The same with the candidate q read from global memory (q < x, anywhere in the block).  The candidate's bytes come
// in batches of aligned 32-bit loads issued together (one L2 round trip per batch instead of one per 4 bytes): 16
// bytes first (most matches end there), then 32 at a time.
template <int NW>     // compare NW 32-bit words at offset n; returns the number of equal leading bytes (4 * NW if all agree)
ZL_HD int v4_cmp_words(const V4Ctx& c, uint32_t x, uint32_t q, int n) {
    const uint32_t qa = q + (uint32_t) n, xa = x + (uint32_t) n;
    const uint32_t* qp = reinterpret_cast<const uint32_t*>(c.in) + (qa >> 2);
    const uint32_t shq = (qa & 3u) * 8u, shx = (xa & 3u) * 8u, xi = xa >> 2;
    uint32_t qw[NW + 1], xw[NW + 1];
    #pragma unroll
    for (int i = 0; i <= NW; i++) { qw[i] = z4_ld_in32(qp + i); xw[i] = c.rbw[(xi + (uint32_t) i) & (kV4R / 4 - 1)]; }
    uint32_t fd = 0; int fi = NW;                                        // first differing word (scanning from the last one down)
    #pragma unroll
    for (int i = NW - 1; i >= 0; i--) {
        const uint32_t d = z4_funnel(qw[i], qw[i + 1], shq) ^ z4_funnel(xw[i], xw[i + 1], shx);
        if (d) { fd = d; fi = i; }
    }
    return fd ? 4 * fi + ((z4_ffs(fd) - 1) >> 3) : 4 * NW;
}
ZL_HD int v4_common_len_mixed(const V4Ctx& c, uint32_t x, uint32_t q) {
    int e = v4_cmp_words<4>(c, x, q, 0);
    if (e < 4) return 0;
    if (e < 16) return e;
    for (int n = 16; n < 272; n += 32) {
        e = v4_cmp_words<8>(c, x, q, n);
        if (e < 32) { const int l = n + e; return l < kMaxLen ? l : kMaxLen; }
    }
    return kMaxLen;                                                      // 272 bytes agree; the length is capped at 259
}

// ---- SPEC C: nearest earlier position of the window with the same key -----------------------------------------------------
ZL_HD void v4_link_position(const V4Ctx& c, int rel) {
    const uint32_t k = c.key[rel];
    uint32_t out = 0;
    if (!(k & kV4KeyInvalid)) {
        int y = rel;
        uint32_t d = c.blink[rel];
        while (d != 0) {
            y -= (int) d;
            if (((c.key[y] ^ k) & kV4KeyMask) == 0) { out = (uint32_t) (rel - y); break; }
            d = c.blink[y];
        }
    }
    c.link[rel] = (uint16_t) out;
}

// ---- SPEC D: chain record of a position against the frozen bucket state G -----------------------------------------------
ZL_HD uint32_t v4_ring_dist(uint32_t slot, uint32_t head_at_freeze) {   // inserts into the context until `slot` is overwritten (1..4096)
    return ((slot - head_at_freeze - 1u) & (kRing - 1)) + 1u;
}
// hdr[rel] = nodes recorded (5) | (smallest ring distance of a slot the walk read - 1) << 5
// node[rel * dmax + i] = match length (0 if the check byte differs, lz.cpp:245) | ring slot << 9
// nodeq[rel * lmax + i] = candidate position (the lazy probes compare 4 bytes at an offset only known later)
ZL_HD void v4_spec_position(const V4Ctx& c, int lo, int rel) {
    const int x = lo + rel;
    const uint32_t k = c.key[rel];
    if (k & kV4KeyInvalid) { c.hdr[rel] = (uint32_t) (kRing - 1) << 5; c.fx[rel] = (uint32_t) kNil; return; }
    const uint32_t ctx = v4_ctx_of(k), slot = k & (kSlots - 1), chk = k >> 21;
    const uint32_t head_b = c.cnt[ctx] & (kRing - 1);
    const uint64_t* rc = c.ring + (size_t) ctx * kRing;
    uint32_t node = z4_ld_hash(c.hash + (size_t) ctx * kSlots + slot);
    c.fx[rel] = node;
    uint32_t nvis = 0, dmin = kRing;
    if (node != (uint32_t) kNil) {
        dmin = v4_ring_dist(node, head_b);
        uint64_t e = z4_ld_ring(rc + node);
        for (int i = 0; i < c.dmax; i++) {
            const uint32_t q = ring_pos(e);
            const uint32_t nxt = ring_suffix(e);
            // the next chain node is fetched while this one's bytes are compared (independent loads overlap)
            const uint64_t e2 = nxt != (uint32_t) kNil ? z4_ld_ring(rc + nxt) : 0ull;
            int len = 0;
            if (ring_check(e) == chk) len = v4_common_len_mixed(c, (uint32_t) x, q);
            c.node[rel * c.dmax + i] = (uint32_t) len | (node << 9);
            if (i < c.lmax) c.nodeq[rel * c.lmax + i] = q;
            nvis = i + 1;
            if (nxt == (uint32_t) kNil) break;
            dmin = min(dmin, v4_ring_dist(nxt, head_b));
            if (q <= ring_pos(e2)) break;                                // lz.cpp:264
            node = nxt; e = e2;
        }
    }
    c.hdr[rel] = nvis | ((dmin - 1u) << 5);
}

// The same record in two steps, for helper CTAs that have more threads than positions: the chain walk (dependent ring loads, one
// thread per position) leaves every candidate position in qall[rel * dmax + i] (bit 31: its check byte matches), the compares
// (the bulk of the work at the deep levels) are then spread over several threads per position.
ZL_HD void v4_spec_walk(const V4Ctx& c, int lo, int rel, uint32_t* qall) {
    const uint32_t k = c.key[rel];
    if (k & kV4KeyInvalid) { c.hdr[rel] = (uint32_t) (kRing - 1) << 5; c.fx[rel] = (uint32_t) kNil; return; }
    const uint32_t ctx = v4_ctx_of(k), slot = k & (kSlots - 1), chk = k >> 21;
    const uint32_t head_b = c.cnt[ctx] & (kRing - 1);
    const uint64_t* rc = c.ring + (size_t) ctx * kRing;
    uint32_t node = z4_ld_hash(c.hash + (size_t) ctx * kSlots + slot);
    c.fx[rel] = node;
    uint32_t nvis = 0, dmin = kRing;
    if (node != (uint32_t) kNil) {
        dmin = v4_ring_dist(node, head_b);
        uint64_t e = z4_ld_ring(rc + node);
        for (int i = 0; i < c.dmax; i++) {
            const uint32_t q = ring_pos(e);
            const uint32_t nxt = ring_suffix(e);
            const uint64_t e2 = nxt != (uint32_t) kNil ? z4_ld_ring(rc + nxt) : 0ull;
            c.node[rel * c.dmax + i] = node << 9;
            qall[rel * c.dmax + i] = q | (ring_check(e) == chk ? 0x80000000u : 0u);
            if (i < c.lmax) c.nodeq[rel * c.lmax + i] = q;
            nvis = i + 1;
            if (nxt == (uint32_t) kNil) break;
            dmin = min(dmin, v4_ring_dist(nxt, head_b));
            if (q <= ring_pos(e2)) break;                                // lz.cpp:264
            node = nxt; e = e2;
        }
    }
    c.hdr[rel] = nvis | ((dmin - 1u) << 5);
}
ZL_HD void v4_spec_compare(const V4Ctx& c, int lo, int rel, int i, const uint32_t* qall) {
    const uint32_t v = qall[rel * c.dmax + i];
    if (v >> 31) c.node[rel * c.dmax + i] |= (uint32_t) v4_common_len_mixed(c, (uint32_t) (lo + rel), v & 0xffffffu);
}

// MatchLazy's test against recorded node i of position zrel (lz.cpp:303-305): do the 4 bytes at offset `at` agree?  The node's
// exact match length l against zrel (when its check byte matched) usually answers without touching global memory: the bytes
// agree up to l and differ at l, so l >= at + 4 => yes, at <= l < at + 4 => no; only l < at (or an unknown l) needs the bytes.
ZL_HD bool v4_lazy_node_hit(const V4Ctx& c, int zrel, int i, uint32_t at, uint32_t mine) {
    const uint32_t l = c.node[zrel * c.dmax + i] & 511u;
    if (l >= at + 4u) return true;
    if (l >= at && l != 0u) return false;
    return z4_in32(c.in, c.nodeq[zrel * c.lmax + i] + at) == mine;
}

// ---- SPEC E: the frozen decision of a main position ----------------------------------------------------------------------
// fdec[rel] = flen(9) | fbest(9) << 9 | fslot(12) << 18: match length after the lazy veto / before it / ring slot of the best
// node, all against G alone; fx[rel] gets the flags saying which of x, x+1, x+2 have an earlier same-key position in the window
// part 1 (needs only the records: the helper CTAs of a cluster run it): fdec
ZL_HD void v4_frozen_core(const V4Ctx& c, int lo, int rel, int level) {
    const int x = lo + rel;
    const int D = depth_main(level), L1 = depth_lazy1(level), L2 = depth_lazy2(level);
    const uint32_t hdr = c.hdr[rel];
    const int nvis = (int) (hdr & 31u);
    uint32_t fbest = 0, fslot = 0;
    const int take = nvis < D ? nvis : D;
    for (int i = 0; i < take; i++) {
        const uint32_t nd = c.node[rel * c.dmax + i];
        if ((nd & 511u) > fbest) { fbest = nd & 511u; fslot = nd >> 9; }
    }
    uint32_t flen = fbest >= (uint32_t) kMinLen ? fbest : 0u;
    if (fbest < (uint32_t) kMinLen) fbest = 0;
    if (fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow) {  // lz.cpp:270-281 against G
        const uint32_t at = fbest - 3u;
        for (int which = 1; which <= 2 && flen; which++) {
            const int depth = which == 1 ? L1 : L2;
            const int nvx = (int) (c.hdr[rel + which] & 31u);
            int tk = nvx < depth ? nvx : depth;
            if (tk > c.lmax) tk = c.lmax;
            const uint32_t mine = v4_rb32(c.rbw, (uint32_t) (x + which) + at);
            for (int i = 0; i < tk; i++)
                if (v4_lazy_node_hit(c, rel + which, i, at, mine)) flen = 0;
        }
    }
    c.fdec[rel] = flen | (fbest << 9) | (fslot << 18);
}
// part 2 (needs the window's links and byte counts): the flags of fx
ZL_HD void v4_frozen_flags(const V4Ctx& c, int rel) {
    const uint32_t hdr = c.hdr[rel];
    const int nvis = (int) (hdr & 31u);
    const uint32_t fbest = (c.fdec[rel] >> 9) & 511u;
    const bool lazy_matters = fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow;
    uint32_t fl = 0;
    if (c.link[rel]) fl |= kV4F_SELF;
    if (lazy_matters) {
        if (c.link[rel + 1]) fl |= kV4F_L1;
        if (c.link[rel + 2]) fl |= kV4F_L2;
    }
    // static bound of the staleness tests (v4_maybe_stale without the per-round marked counts): positions with the context byte
    {
        bool st = nvis > 0 && (hdr >> 5) + 1u <= c.pcnt[v4_ctx_of(c.key[rel])];
        if (lazy_matters) {
            for (int q = 1; q <= 2; q++) {
                const uint32_t hq = c.hdr[rel + q];
                if ((hq & 31u) && (hq >> 5) + 1u <= c.pcnt[v4_ctx_of(c.key[rel + q])] + 1u) st = true;
            }
        }
        if (st) fl |= kV4F_ST;
        // the same bound for the records of rel + 1 / rel + 2 whatever the frozen length is (the full probe may run its lazy tests with
        // another length): information for v4_probe_general only, it does not send the position to the hazard checks
        for (int q = 1; q <= 2; q++) {
            const uint32_t hq = c.hdr[rel + q];
            if ((hq & 31u) && (hq >> 5) + 1u <= c.pcnt[v4_ctx_of(c.key[rel + q])] + 1u) fl |= kV4F_STL;
        }
    }
    c.fx[rel] = (c.fx[rel] & 0xffffu) | fl;
}
ZL_HD void v4_frozen_position(const V4Ctx& c, int lo, int rel, int level) {
    v4_frozen_core(c, lo, rel, level);
    v4_frozen_flags(c, rel);
}

// ---- ROUNDS: the pending view -----------------------------------------------------------------------------------------------
// While position x (rel) is being decided the reference's bucket state is G plus the inserts of the marked positions
// before x plus x's own insert (the insert precedes the search, lz.cpp:227-230).
ZL_HD bool v4_pending(const V4Ctx& c, int xrel, int y) { return y == xrel || (y < xrel && c.mark[y] != 0); }
// window positions [32 w, 32 w + 32) whose context byte is cq (position rel has context cq <=> bit rel + 2 of occ[cq])
ZL_HD uint32_t v4_ctxbits(const V4Ctx& c, uint32_t cq, int w) {
    const uint32_t* o = c.occ + cq * kV4Words;
    return z4_funnel(o[w], o[w + 1], 2);
}
// inserts into context cq pending before rel: marked positions y < rel with that context.  Exact, on demand (the
// per-context ranks of ALL positions are only built once per window, after the rounds: v4_head_final)
ZL_HD uint32_t v4_rank_live(const V4Ctx& c, int rel, uint32_t cq) {
    V4_STAT(1, rel >> 5);
#if defined(__CUDA_ARCH__)
    if (c.coop) {                                                        // one bitset word per lane
        const int lane = threadIdx.x & 31, wl = rel >> 5;
        uint32_t m = 0;
        if (lane <= wl) { m = v4_ctxbits(c, cq, lane) & c.mbits[lane]; if (lane == wl) m &= (1u << (rel & 31)) - 1u; }
        return __reduce_add_sync(0xffffffffu, (uint32_t) __popc(m));
    }
#endif
    uint32_t n = 0;
    const int wl = rel >> 5;
    for (int w = 0; w < wl; w++) n += (uint32_t) z4_popc(v4_ctxbits(c, cq, w) & c.mbits[w]);
    if (rel & 31) n += (uint32_t) z4_popc(v4_ctxbits(c, cq, wl) & c.mbits[wl] & ((1u << (rel & 31)) - 1u));
    return n;
}
ZL_HD uint32_t v4_head_live(const V4Ctx& c, int rel) {                  // ring slot of the insert made at rel, from the marks
    const uint32_t cq = v4_ctx_of(c.key[rel]);
    return (c.cnt[cq] + v4_rank_live(c, rel, cq) + 1u) & (kRing - 1);
}
ZL_HD uint32_t v4_head_final(const V4Ctx& c, int rel) {                 // the same from the rank table (FINALIZE)
    return (c.cnt[v4_ctx_of(c.key[rel])] + c.rank[rel] + 1u) & (kRing - 1);
}
// inserts into the context of position rel + q (q = 1, 2) made before and at rel
ZL_HD uint32_t v4_cnt_lazy(const V4Ctx& c, int rel, int q) {
    const uint32_t cw = v4_ctx_of(c.key[rel + q]);
    return v4_rank_live(c, rel, cw) + (v4_ctx_of(c.key[rel]) == cw ? 1u : 0u);
}
// is some same-key position before z, not after xrel, pending?
ZL_HD bool v4_link_hazard(const V4Ctx& c, int z, int xrel, bool self_counts) {
    int y = z;
    int steps = 0;
    while (true) {
        const uint32_t d = c.link[y];
        if (!d) { V4_STAT(0, steps); return false; }
        y -= (int) d;
        steps++;
        if (y > xrel) continue;
        if (y == xrel) { if (self_counts) { V4_STAT(0, steps); return true; } continue; }
        if (c.mark[y]) { V4_STAT(0, steps); return true; }
    }
}
// Leading nodes of a record whose ring slots have not been overwritten after `kc` further inserts into the context.
// A record node that HAS been overwritten ends the reference's walk right there when it is not the first one (the slot
// now holds a newer, i.e. larger, position: the "offset <= next offset" test of lz.cpp:264 fires), so a record cut at
// its first overwritten node is still exact; only an overwritten FIRST node needs the literal replay.
ZL_HD int v4_valid_nodes(const V4Ctx& c, int rel, int nvis, uint32_t head_b, uint32_t kc) {
    int i = 0;
    while (i < nvis && v4_ring_dist(c.node[rel * c.dmax + i] >> 9, head_b) > kc) i++;
    return i;
}
// Could a slot read by the record of rel have been overwritten by this window's inserts?  (static bound: pcnt)
ZL_HD bool v4_maybe_stale(const V4Ctx& c, int rel, uint32_t hdr, uint32_t extra) {
    const uint32_t cq = v4_ctx_of(c.key[rel]);
    const uint32_t a = c.pcnt[cq], b = c.mcnt[cq] + 1u;                  // inserts into cq so far <= positions with that context, <= marked ones + the one being decided
    return (hdr & 31u) && (hdr >> 5) + 1u <= (a < b ? a : b) + extra;
}
// slot head seen by the insert at y: the nearest marked same-key position before it, else the frozen head
ZL_HD uint32_t v4_suffix_live(const V4Ctx& c, int y) {
    int t = y;
    while (true) {
        const uint32_t d = c.link[t];
        if (!d) break;
        t -= (int) d;
        if (c.mark[t]) return v4_head_live(c, t);
    }
    return c.fx[y] & 0xffffu;
}
// the reference's ring[cq][n] as seen while xrel is decided
ZL_HD uint64_t v4_live_entry(const V4Ctx& c, int lo, int xrel, uint32_t cq, uint32_t n) {
    const uint32_t ord = (n - c.cnt[cq]) & (kRing - 1);                  // n is pending iff it is insert number 1.. of this window
#if defined(__CUDA_ARCH__)
    if (c.coop && ord != 0 && ord <= (uint32_t) kV4N) {                  // the same selection, one bitset word per lane
        const int lane = threadIdx.x & 31, wl = xrel >> 5;
        uint32_t m = 0;
        if (lane <= wl) { m = v4_ctxbits(c, cq, lane) & c.mbits[lane]; if (lane == wl) m &= (1u << (xrel & 31)) - 1u; }
        const uint32_t pc = (uint32_t) __popc(m);
        uint32_t incl = pc;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const uint32_t bal = __ballot_sync(0xffffffffu, incl >= ord);
        if (bal) {
            const int L = __ffs((int) bal) - 1;
            const uint32_t before = __shfl_sync(0xffffffffu, incl - pc, L);
            uint32_t mL = __shfl_sync(0xffffffffu, m, L);
            for (uint32_t i = 1; i < ord - before; i++) mL &= mL - 1u;
            const int y = L * 32 + __ffs((int) mL) - 1;
            return ring_make((uint32_t) (lo + y), c.key[y] >> 21, v4_suffix_live(c, y));
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (ord - total == 1u && v4_ctx_of(c.key[xrel]) == cq) return ring_make((uint32_t) (lo + xrel), c.key[xrel] >> 21, v4_suffix_live(c, xrel));
        return z4_ld_ring(c.ring + (size_t) cq * kRing + n);
    }
#endif
    if (ord != 0 && ord <= (uint32_t) kV4N) {
        // the ord-th position with context cq among the pending ones (marked before xrel; xrel itself comes last)
        uint32_t left = ord;
        const int wl = xrel >> 5;
        for (int w = 0; w <= wl; w++) {
            uint32_t m = v4_ctxbits(c, cq, w) & c.mbits[w];
            if (w == wl) m &= (1u << (xrel & 31)) - 1u;
            const uint32_t pc = (uint32_t) z4_popc(m);
            if (left <= pc) {
                for (uint32_t i = 1; i < left; i++) m &= m - 1u;         // drop the left-1 lowest set bits
                const int y = w * 32 + z4_ffs(m) - 1;
                return ring_make((uint32_t) (lo + y), c.key[y] >> 21, v4_suffix_live(c, y));
            }
            left -= pc;
        }
        if (left == 1 && v4_ctx_of(c.key[xrel]) == cq) return ring_make((uint32_t) (lo + xrel), c.key[xrel] >> 21, v4_suffix_live(c, xrel));
    }
    return z4_ld_ring(c.ring + (size_t) cq * kRing + n);
}
ZL_HD int v4_common_len_any(const V4Ctx& c, uint32_t x, uint32_t q) {     // both operands from global memory
    if (z4_in32(c.in, x) != z4_in32(c.in, q)) return 0;
    for (int n = 4; n < 256; n += 4) {
        const uint32_t d = z4_in32(c.in, x + n) ^ z4_in32(c.in, q + n);
        if (d) return n + ((z4_ffs(d) - 1) >> 3);
    }
    const uint32_t d = z4_in32(c.in, x + 256) ^ z4_in32(c.in, q + 256);
    const int t = d ? ((z4_ffs(d) - 1) >> 3) : 4;
    return 256 + (t < 3 ? t : 3);
}
// literal replay of the candidate walk of MatchAndUpdate (lz.cpp:234-267) on the pending view
ZL_HD int v4_main_live(const V4Ctx& c, int lo, int xrel, uint32_t node, uint32_t head, uint32_t chk, uint32_t cq, int D, uint32_t* bestslot) {
    if (node == (uint32_t) kNil || node == head) return 0;
    int best = kMinLen - 1;
    uint64_t e = v4_live_entry(c, lo, xrel, cq, node);
    for (int hop = 0; hop < D; hop++) {
        const uint32_t cand = ring_pos(e);
        if (ring_check(e) == chk) {
            const int l = v4_common_len_any(c, (uint32_t) (lo + xrel), cand);
            if (l > best) { best = l; *bestslot = node; if (best == kMaxLen) break; }
        }
        const uint32_t nxt = ring_suffix(e);
        if (nxt == (uint32_t) kNil) break;
        const uint64_t e2 = v4_live_entry(c, lo, xrel, cq, nxt);
        if (cand <= ring_pos(e2)) break;
        node = nxt; e = e2;
    }
    return best;
}
// literal replay of MatchLazy (lz.cpp:291-316) at position zrel on the pending view of xrel
ZL_HD bool v4_lazy_live(const V4Ctx& c, int lo, int xrel, int zrel, int best, int depth) {
    const uint32_t k = c.key[zrel];
    const uint32_t cq = v4_ctx_of(k);
    uint32_t node = c.fx[zrel] & 0xffffu;                                // hash[cq][slot]: newest pending same-key insert, else G's
    {
        int y = zrel;
        while (true) {
            const uint32_t d = c.link[y];
            if (!d) break;
            y -= (int) d;
            if (y <= xrel && v4_pending(c, xrel, y)) { node = v4_head_live(c, y); break; }
        }
    }
    if (node == (uint32_t) kNil) return false;
    const uint32_t at = (uint32_t) best - 3u;
    const uint32_t mine = z4_in32(c.in, (uint32_t) (lo + zrel) + at);
    uint64_t e = v4_live_entry(c, lo, xrel, cq, node);
    for (int hop = 0; hop < depth; hop++) {
        const uint32_t cand = ring_pos(e);
        if (z4_in32(c.in, cand + at) == mine) return true;
        const uint32_t nxt = ring_suffix(e);
        if (nxt == (uint32_t) kNil) break;
        const uint64_t e2 = v4_live_entry(c, lo, xrel, cq, nxt);
        if (cand <= ring_pos(e2)) break;
        e = e2;
    }
    return false;
}

// Does the frozen decision at rel NOT stand because of the inserts pending in the window?
ZL_HD bool v4_hazard(const V4Ctx& c, int rel, uint32_t fd, uint32_t fxw, int L2) {
    const uint32_t fbest = (fd >> 9) & 511u;
    const int nlazy = (fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow) ? (L2 > 0 ? 2 : 1) : 0;
    if ((fxw & kV4F_SELF) && v4_link_hazard(c, rel, rel, false)) return true;
    if (nlazy >= 1 && (fxw & kV4F_L1) && v4_link_hazard(c, rel + 1, rel, true)) return true;
    if (nlazy >= 2 && (fxw & kV4F_L2) && v4_link_hazard(c, rel + 2, rel, true)) return true;
    if (!(fxw & kV4F_ST)) return false;                                  // the static bound already rules the staleness tests out
    const uint32_t h0 = c.hdr[rel];
    if (v4_maybe_stale(c, rel, h0, 0)) {
        const uint32_t cq = v4_ctx_of(c.key[rel]);
        const uint32_t kc0 = v4_rank_live(c, rel, cq) + 1u;
        if ((h0 >> 5) + 1u <= kc0 && v4_valid_nodes(c, rel, (int) (h0 & 31u), c.cnt[cq] & (kRing - 1), kc0) < (int) (h0 & 31u)) return true;
    }
    for (int q = 1; q <= nlazy; q++) {
        const uint32_t hw = c.hdr[rel + q];
        if (!v4_maybe_stale(c, rel + q, hw, 1)) continue;
        const uint32_t cw = v4_ctx_of(c.key[rel + q]);
        const uint32_t kc = v4_cnt_lazy(c, rel, q);
        if ((hw >> 5) + 1u <= kc && v4_valid_nodes(c, rel + q, (int) (hw & 31u), c.cnt[cw] & (kRing - 1), kc) < (int) (hw & 31u)) return true;
    }
    return false;
}

#if defined(ZL_V4_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long g_v4prof[16];
#endif
#if defined(ZL_V4_PROFILE) && defined(__CUDA_ARCH__)
#define V4_GP_T() clock64()
#define V4_GP(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_v4prof[i], (unsigned long long) (v)); } while (0)
#else
#define V4_GP_T() 0ll
#define V4_GP(i, v) do { } while (0)
#endif
// The full MatchAndUpdate (lz.cpp:211-289) at rel on the pending view: in-window candidates (newest first), then the
// frozen record.  Returns the match length (0 = none) and the reference to the best candidate (a ring slot, or
// kV4RefWin | rel of a pending position).
ZL_HD int v4_probe_general(const V4Ctx& c, int lo, int rel, int level, uint32_t* ref_out) {
    const int x = lo + rel;
    const int D = depth_main(level), L1 = depth_lazy1(level), L2 = depth_lazy2(level);
    const uint32_t kx = c.key[rel], fxw = c.fx[rel];
    const uint32_t chk = kx >> 21, cq = v4_ctx_of(kx);
    V4_STAT(4, 1);
    int best = kMinLen - 1, visited = 0, first_pending = -1;
    uint32_t ref = 0;
    bool done = false;
    const long long g0_ = V4_GP_T();
    int hops_ = 0; (void) hops_;
    {
        int y = rel;
        while (true) {
            const uint32_t dl = c.link[y];
            if (!dl) break;
            y -= (int) dl;
            hops_++;
            if (!c.mark[y]) continue;
            if (first_pending < 0) first_pending = y;
            if (visited < D && !done) {
                visited++;
                if ((c.key[y] >> 21) == chk) {
                    const int l = v4_common_len_ring(c.rbw, (uint32_t) x, (uint32_t) (lo + y), c.coop);
                    if (l > best) { best = l; ref = kV4RefWin | (uint32_t) y; if (best == kMaxLen) done = true; }
                }
            } else break;
        }
    }
    const long long g1_ = V4_GP_T();
    V4_GP(0, g1_ - g0_); V4_GP(1, hops_); V4_GP(2, visited);
    const uint32_t hdr = c.hdr[rel];
    int nvis = (int) (hdr & 31u);
    bool stale0 = false;
    uint32_t kc0 = 0;
    if ((fxw & kV4F_ST) && v4_maybe_stale(c, rel, hdr, 0)) {
        kc0 = v4_rank_live(c, rel, cq) + 1u;
        if ((hdr >> 5) + 1u <= kc0) {
            const int nv = v4_valid_nodes(c, rel, nvis, c.cnt[cq] & (kRing - 1), kc0);
            stale0 = nv == 0;
            nvis = nv;
        }
    }
    if (!done && visited < D && (nvis > 0 || stale0)) {
        if (stale0) {                                                    // the record's first slot has been overwritten: replay literally
            V4_STAT(5, 1);
            const uint32_t head = (c.cnt[cq] + kc0) & (kRing - 1);
            const uint32_t suffix = first_pending >= 0 ? v4_head_live(c, first_pending) : (c.fx[rel] & 0xffffu);
            uint32_t bn = 0;
            best = v4_main_live(c, lo, rel, suffix, head, chk, cq, D, &bn);
            ref = bn;
        } else {
            const int take = nvis < D - visited ? nvis : D - visited;
            for (int i = 0; i < take; i++) {
                const uint32_t nd = c.node[rel * c.dmax + i];
                const int l = (int) (nd & 511u);
                if (l > best) { best = l; ref = nd >> 9; if (best == kMaxLen) break; }
            }
        }
    }
    const long long g2_ = V4_GP_T();
    V4_GP(3, g2_ - g1_); V4_GP(4, stale0 ? 1 : 0);
    if (best < kMinLen) return 0;
    V4_GP(5, 1);
    if (best < kLazyBelow) {                                             // lz.cpp:270-281
        const uint32_t at = (uint32_t) best - 3u;
        for (int which = 1; which <= 2; which++) {
            const int depth = which == 1 ? L1 : L2;
            if (depth == 0) break;
            const int relz = rel + which;
            const uint32_t hz = c.hdr[relz];
            int nvz = (int) (hz & 31u);
            if ((fxw & kV4F_STL) && v4_maybe_stale(c, relz, hz, 1)) {    // stale lazy record: cut it, or replay when its head is gone
                const uint32_t cz = v4_ctx_of(c.key[relz]);
                const uint32_t kcz = v4_cnt_lazy(c, rel, which);
                if ((hz >> 5) + 1u <= kcz) {
                    nvz = v4_valid_nodes(c, relz, nvz, c.cnt[cz] & (kRing - 1), kcz);
                    if (nvz == 0) {
                        if (v4_lazy_live(c, lo, rel, relz, best, depth)) return 0;
                        continue;
                    }
                }
            }
            const uint32_t mine = v4_rb32(c.rbw, (uint32_t) (lo + relz) + at);
            int vis = 0;
            int y = relz;
            while (vis < depth) {                                        // pending same-key positions, newest first (rel itself included)
                const uint32_t dl = c.link[y];
                if (!dl) break;
                y -= (int) dl;
                if (!v4_pending(c, rel, y)) continue;
                vis++;
                if (v4_rb32(c.rbw, (uint32_t) (lo + y) + at) == mine) return 0;
            }
            int tk = nvz < depth - vis ? nvz : depth - vis;
            if (tk > c.lmax) tk = c.lmax;
            for (int i = 0; i < tk; i++)
                if (v4_lazy_node_hit(c, relz, i, at, mine)) return 0;
        }
    }
    V4_GP(6, V4_GP_T() - g2_);
    *ref_out = ref;
    return best;
}

// ---- ROUNDS: word MRU ---------------------------------------------------------------------------------------------------------
// Every token that ENDS at e pushes the word in[e-2..e-1] into the MRU of context in[e-3] (lz.cpp:163-166,183-185,
// 190-191): unconditionally after a literal, otherwise only when the front differs (a 256 hit changes nothing, a 257
// hit always differs).  The state seen at x for context cq is the fold of the pushes at the marked positions e <= x
// with in[e-3] == cq (bitset occ[cq] & mbits) over the carried state.  Returns w0 | w1 << 16.
ZL_HD uint32_t v4_mru_state(const V4Ctx& c, const V4Win& w, int xrel, uint32_t cq) {
    uint32_t base = c.mru[cq];
    int elo = w.entry - w.lo + (w.skip_push ? 1 : 0);                     // pushes happen at marked e >= elo
    if (w.rpos >= 0 && w.lo + xrel >= w.rpos) { base = 0; elo = w.rpos - w.lo + 1; }   // the MRU is zeroed when a sub-block opens (lz.cpp:147)
    if (xrel < elo) return base;
    const uint32_t* occ = c.occ + cq * kV4Words;
    // pushes newest first: (a1,u1), (a2,u2), ...; the state is that after the newest EFFECTIVE push j: (a_j, front before j)
    bool have = false;                // a push whose effectiveness is still unknown (needs the front before it)
    uint32_t a = 0; bool u = false;
    const int whi = xrel >> 5, wlo = elo >> 5;
#if defined(__CUDA_ARCH__)
    if (c.coop) {                                                        // one bitset word per lane; the newest push is found with a ballot
        const int lane = threadIdx.x & 31;
        uint32_t m = (lane >= wlo && lane <= whi) ? occ[lane] & c.mbits[lane] : 0u;
        if (lane == whi) m &= 0xffffffffu >> (31 - (xrel & 31));
        if (lane == wlo) m &= 0xffffffffu << (elo & 31);
        while (true) {
            const uint32_t bal = __ballot_sync(0xffffffffu, m != 0);
            if (!bal) break;
            const int hw = 31 - __clz((int) bal);
            const uint32_t mw = __shfl_sync(0xffffffffu, m, hw);
            const int bq = 31 - __clz((int) mw);
            if (lane == hw) m &= ~(1u << bq);
            const int e = hw * 32 + bq;
            const uint32_t xe = (uint32_t) (w.lo + e);
            const uint32_t pw = (v4_rb8(c.rbw, xe - 2) << 8) | v4_rb8(c.rbw, xe - 1);
            if (have && (u || pw != a)) return a | (pw << 16);
            have = true; a = pw; u = e == w.entry - w.lo ? w.prev_lit != 0 : c.plit[e] != 0;
        }
        if (!have) return base;
        if (u || (base & 0xffffu) != a) return a | (base << 16);
        return base;
    }
#endif
    // words of the window in which a marked position pushes into cq, between the first possible push and xrel, newest first
    uint32_t words = c.pushw[cq] & (0xffffffffu >> (31 - whi)) & (0xffffffffu << wlo);
    int st_words = 0, st_push = 0;
    (void) st_words; (void) st_push;
    while (words) {
        st_words++;
        const int wi = 31 - z4_clz(words);
        words &= ~(1u << wi);
        uint32_t m = occ[wi] & c.mbits[wi];
        if (wi == whi) m &= 0xffffffffu >> (31 - (xrel & 31));
        if (wi == wlo) m &= 0xffffffffu << (elo & 31);
        while (m) {
            st_push++;
            const int b = 31 - z4_clz(m);
            m &= ~(1u << b);
            const int e = wi * 32 + b;
            const uint32_t xe = (uint32_t) (w.lo + e);
            const uint32_t pw = (v4_rb8(c.rbw, xe - 2) << 8) | v4_rb8(c.rbw, xe - 1);
            if (have) {                                                  // pw is the front before the pending push a
                if (u || pw != a) { V4_STAT(2, st_words); V4_STAT(3, st_push); return a | (pw << 16); }
                // the pending push was a no-op: the state is that after this (older) push
            }
            have = true; a = pw; u = e == w.entry - w.lo ? w.prev_lit != 0 : c.plit[e] != 0;
        }
    }
    V4_STAT(2, st_words); V4_STAT(3, st_push);
    if (!have) return base;
    if (u || (base & 0xffffu) != a) return a | (base << 16);
    return base;
}

// ---- ROUNDS: the decision of a position given the marks --------------------------------------------------------------------
// the pieces of v4_decide the kernel runs as separate, queue-fed stages (same results: the static flags of fx are a superset
// of the conditions under which v4_hazard can say yes)
ZL_HD uint32_t v4_dec_match(uint32_t len, uint32_t ref) { return len | (kV4Match << 9) | (ref << 12); }
ZL_HD uint32_t v4_pf_hash(uint32_t ctx, uint32_t word) { return (((ctx << 16) | word) * 2654435761u) >> 20; }   // 12 bits
// the push a token END at position xe makes: context in[xe-3], word in[xe-2..xe-1]
ZL_HD void v4_push_of(const V4Ctx& c, uint32_t xe, uint32_t* ctx3, uint32_t* word) {
    const uint32_t t = v4_rb32(c.rbw, xe - 3u);
    *ctx3 = t & 0xffu; *word = ((t >> 8) & 0xffu) << 8 | ((t >> 16) & 0xffu);
}
ZL_HD uint32_t v4_decide_word(const V4Ctx& c, const V4Win& w, int rel) {   // no match is taken at rel: word-MRU test, lz.cpp:172-185
    const int x = w.lo + rel;
    const uint32_t cq = v4_ctx_of(c.key[rel]);
    const uint32_t wd = (v4_rb8(c.rbw, (uint32_t) x) << 8) | v4_rb8(c.rbw, (uint32_t) x + 1);
    {   // quick reject: every MRU entry is a word pushed in this window or a carried one (zero after a roll-over)
        const uint32_t base = (w.rpos >= 0 && x >= w.rpos) ? 0u : c.mru[cq];
        const uint32_t h = v4_pf_hash(cq, wd);
        const bool inpf = ((c.pf[h >> 5] >> (h & 31u)) & 1u) != 0, inbase = (base & 0xffffu) == wd || (base >> 16) == wd;
        V4_STAT(6, inpf ? 1 : 0); V4_STAT(7, inbase ? 1 : 0);
        if (!inpf && !inbase) return kV4Lit << 9;
    }
    const uint32_t m = v4_mru_state(c, w, rel, cq);
    const uint32_t kind = (m & 0xffffu) == wd ? kV4Word0 : ((m >> 16) == wd ? kV4Word1 : kV4Lit);
    return kind << 9;
}
#if defined(ZL_V4_PROFILE) && defined(__CUDA_ARCH__)
#define V4_PROF_T() ((uint32_t) clock64())
#else
#define V4_PROF_T() 0u
#endif
ZL_HD uint32_t v4_decide(const V4Ctx& c, const V4Win& w, int rel, uint32_t* prof = nullptr) {
    const int x = w.lo + rel;
    const int level = (w.rpos >= 0 && x >= w.rpos) ? w.level2 : w.level;
    const uint32_t fd = c.fdec[rel], fxw = c.fx[rel];
    uint32_t len = fd & 511u, ref = (fd >> 18) & (kRing - 1);
    const uint32_t p0 = V4_PROF_T();
    const bool hz = level != w.level || v4_hazard(c, rel, fd, fxw, depth_lazy2(level));
    const uint32_t p1 = V4_PROF_T();
    if (hz) {
        uint32_t rf = 0;
        len = (uint32_t) v4_probe_general(c, w.lo, rel, level, &rf);
        ref = rf;
    }
    const uint32_t p2 = V4_PROF_T();
    if (prof) { prof[0] = p1 - p0; prof[1] = p2 - p1; prof[2] = 0; }
    if (len) return len | (kV4Match << 9) | (ref << 12);
    const uint32_t kd = v4_decide_word(c, w, rel);                      // lz.cpp:172-185 (x + 1 < ilen holds in the probe region)
    if (prof) prof[2] = V4_PROF_T() - p2;
    return kd;
}

// ---- FINALIZE (the rank table is valid now) ------------------------------------------------------------------------------------
// marked position rel: find the slot head its insert replaces, and tell the previous owner of the slot that it has been superseded
ZL_HD uint32_t v4_claim_slot(const V4Ctx& c, int rel) {
    int t = rel;
    while (true) {
        const uint32_t d = c.link[t];
        if (!d) break;
        t -= (int) d;
        if (c.mark[t]) { c.sup[t] = 1; return v4_head_final(c, t); }
    }
    return c.fx[rel] & 0xffffu;
}
ZL_HD void v4_apply_position(const V4Ctx& c, int lo, int rel, uint32_t suffix) {
    const uint32_t k = c.key[rel];
    const uint32_t cq = v4_ctx_of(k), slot = k & (kSlots - 1), chk = k >> 21, head = v4_head_final(c, rel);
    c.ring[(size_t) cq * kRing + head] = ring_make((uint32_t) (lo + rel), chk, suffix);
    if (!c.sup[rel]) c.hash[(size_t) cq * kSlots + slot] = (uint16_t) head;
}
ZL_HD uint32_t v4_token_of(const V4Ctx& c, int lo, int rel) {
    const uint32_t d = c.dec[rel], kind = v4_dec_kind(d);
    if (kind == kV4Lit) return tok_literal(v4_rb8(c.rbw, (uint32_t) (lo + rel)), v4_rb8(c.rbw, (uint32_t) (lo + rel) - 1), false);
    if (kind == kV4Word0) return tok_word(0);
    if (kind == kV4Word1) return tok_word(1);
    const uint32_t ref = v4_dec_ref(d);
    const uint32_t slot = (ref & kV4RefWin) ? v4_head_final(c, (int) (ref & (kV4RefWin - 1))) : ref;
    return tok_match(v4_dec_len(d), (v4_head_final(c, rel) - slot) & (kRing - 1));     // lz.cpp:283
}

// ---- resolver state carried across windows ------------------------------------------------------------------------------------
struct V4Run {
    int ip, op, j, level, tok_begin, enc_begin;
    int prev_lit;                    // the previous token was a literal (its word-MRU push is unconditional)
    int skip_push;                   // no push is pending on arrival (block start: the two raw bytes push nothing)
    int tail;                        // ip reached the last 275 bytes: the rest is done by v4_resolve_tail
};
// Level of sub-block j.  The reference derives it from the PREVIOUS sub-block's Huffman size (src/libzling.cpp:261-266:
// olen / (consumed + 1) > 0.95 => level 0), which is not known during the parse (the literal ranks depend on MTF state
// carried across blocks).  plan[j] is the host's word: a level it has verified, or kV4Auto = "predict": apply the reference's
// rule with one byte per symbol as the size estimate (what incompressible data costs: its symbols are literals of ~8 bits;
// measured on the mixed corpus: 1 wrong guess in 1716 sub-blocks).  The host verifies every level afterwards from the
// real sizes and re-parses from the first wrong one, so a wrong guess only costs time.
ZL_HD int v4_next_level(const V4Ctx& c, int j, int consumed, int op) {
    const uint32_t p = c.plan[j < kMaxSubPerBlock ? j : kMaxSubPerBlock - 1];
    if (p != kV4Auto) return (int) p;
    if (j == 0) return c.base_level;
    return (unsigned long long) op * 20ull > (unsigned long long) (consumed + 1) * 19ull ? 0 : c.base_level;
}
ZL_HD void v4_close_subblock(const V4Ctx& c, const V4Run& r, int ip, int op, int nt) {
    if (r.j < kMaxSubPerBlock) {
        SubBlock sb; sb.tok_begin = (uint32_t) r.tok_begin; sb.tok_end = (uint32_t) nt; sb.enc_begin = (uint32_t) r.enc_begin;
        sb.enc_end = (uint32_t) ip; sb.rlen = (uint32_t) op; sb.level = (uint32_t) r.level; sb.olen = 0; sb.bits_lo = 0;
        c.sub[r.j] = sb;
    }
}
// The last 275 bytes of the block (no probe, no insert: lz.cpp:158) and blocks shorter than that: plain serial
// code, tokens written directly.  nt / nl = tokens / literals emitted so far.
ZL_HD void v4_resolve_tail(const V4Ctx& c, V4Run& r, int* nt_io, int* nl_io) {
    int nt = *nt_io, nl = *nl_io;
    int ip = r.ip, op = r.op;
    const uint8_t* in = c.in;
    bool pending_push = !r.skip_push && ip >= 3;
    while (ip < c.ilen) {
        if (pending_push) {
            const uint32_t c3 = in[ip - 3], w = ((uint32_t) in[ip - 2] << 8) | in[ip - 1];
            const uint32_t m = c.mru[c3];
            if (r.prev_lit || (m & 0xffffu) != w) c.mru[c3] = w | (m << 16);
        }
        pending_push = true;
        if (op + 1 >= kSubSymbols) {                                     // sub-block full (lz.cpp:153): close it, open the next
            v4_close_subblock(c, r, ip, op, nt);
            const int next = v4_next_level(c, r.j + 1, ip - r.enc_begin, op);
            r.j++; r.level = next;
            for (int i = 0; i < 256; i++) c.mru[i] = 0;                  // lz.cpp:147
            op = 0; r.tok_begin = nt; r.enc_begin = ip;
        }
        const uint32_t c1 = in[ip - 1], cur = in[ip];
        if (ip + 1 < c.ilen) {
            const uint32_t w = (cur << 8) | in[ip + 1];
            const uint32_t m = c.mru[c1];
            if ((m & 0xffffu) == w) { c.tok[nt++] = tok_word(0); op++; ip += 2; r.prev_lit = 0; continue; }
            if ((m >> 16) == w) { c.tok[nt++] = tok_word(1); op++; ip += 2; r.prev_lit = 0; continue; }
        }
        c.tok[nt] = tok_literal(cur, c1, false);
        c.lit[nl++] = (uint32_t) nt;
        nt++; op++; ip++; r.prev_lit = 1;
    }
    r.ip = ip; r.op = op;
    *nt_io = nt; *nl_io = nl;
}

}  // namespace zl

#if defined(__CUDACC__)
#include <cooperative_groups.h>
namespace zl {
namespace cg = cooperative_groups;

struct V4Counters { unsigned long long tokens, windows, rounds, decides_general, cyc_spec, cyc_rounds, cyc_final, cyc_total, cyc_orbit, cyc_rank, cyc_decide, ph[40]; };

__device__ __forceinline__ uint32_t v4_lt_mask(int lane) { return (1u << lane) - 1u; }
// warp 0: per-warp totals arr[0..31] -> exclusive prefix in place, grand total in arr[32] (callers synchronise around it)
__device__ __forceinline__ void v4_warp0_prefix(int* arr, int lane) {
    const int v = arr[lane];
    int incl = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    arr[lane] = incl - v;
    if (lane == 31) arr[32] = incl;
}

// ---- SPEC B + C for a whole CTA: link[t] = distance to the nearest earlier position of the window with the same key.  First the
// nearest earlier position in the same BUCKET (blink): groups of 4 warps own a bucket table (gtab, zeroed by the caller); inside a
// group the warps take turns in position order, then positions without a predecessor in their own group look at the final tables
// of the groups before theirs.  Then every position follows its bucket chain to the first equal key.  key[] must be complete.
__device__ __forceinline__ void v4_build_links_cta(const V4Ctx& c, uint16_t* gtab, int tid, int lane, int warp) {
    const int g = warp >> 2, turn = warp & 3;
    uint16_t* tg = gtab + g * kV4Buckets;
    const uint32_t kx = c.key[tid];
    const bool valid = !(kx & kV4KeyInvalid);
    const uint32_t bk = valid ? v4_bucket_of(kx) : 0xffff0000u + (uint32_t) lane;
    uint32_t dist = 0;
    const uint32_t grp = __match_any_sync(0xffffffffu, bk);              // same-bucket lanes of this warp (all warps at once; only the
    const uint32_t lower = grp & v4_lt_mask(lane);                       // table look-ups below have to go in position order)
    if (valid && lower) dist = (uint32_t) lane - (31u - (uint32_t) __clz(lower));
    for (int i = 0; i < 4; i++) {
        if (turn == i) {
            if (valid && !lower) { const uint32_t prev = tg[bk]; if (prev) dist = (uint32_t) tid - (prev - 1u); }
            __syncwarp();
            if (valid && (grp >> lane) == 1u) tg[bk] = (uint16_t) (tid + 1);
        }
        asm volatile("bar.sync %0, 128;" :: "r"(1 + g) : "memory");
    }
    __syncthreads();
    if (valid && dist == 0) {
        for (int g2 = g - 1; g2 >= 0; g2--) {
            const uint32_t prev = gtab[g2 * kV4Buckets + bk];
            if (prev) { dist = (uint32_t) tid - (prev - 1u); break; }
        }
    }
    c.blink[tid] = (uint16_t) dist;
    __syncthreads();
    v4_link_position(c, tid);
}

// ---- the kernel: grid = blocks of the batch, kV4T threads, thread t owns position lo + t of the current window -----------
// DMAX / LMAX (chain nodes recorded per position / of them with a stored candidate position) are compile-time: every
// shared-memory offset is then a constant
template <int DMAX, int LMAX>
__global__ void __launch_bounds__(kV4T, 1) zl_rolz_parse_v4_kernel(ParseArgs a, int base_level, V4Counters* counters) {
    constexpr int dmax = DMAX, lmax = LMAX;
    // A block may be given a thread-block CLUSTER of CL CTAs (launch attribute): CTA 0 of the cluster (the leader) runs the
    // pipeline below; the other CTAs (helpers, each on its own SM) do the one part that is pure throughput — the chain walk of
    // every position against G — and write the records straight into the leader's shared memory (DSMEM).  CL = 1: no helpers.
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int) cluster.num_blocks(), crank = (int) cluster.block_rank();
    const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!a.active[b]) return;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ V4Run s_run;
    __shared__ V4Win s_win;
    __shared__ int s_nt, s_nl, s_exit, s_lastrel, s_rpos_rel, s_rpos_nt, s_op_at_rpos;
    __shared__ int s_wtok[33], s_wlit[33], s_wsym[33], s_wsya[33];
    __shared__ unsigned s_wmax[4];                                               // profiling build: per-round maxima over the warps
    __shared__ int s_take[4];                                                    // next entry to take from the stage queues
    __shared__ int s_nq[4];                                                      // queue lengths of the decide stages
    __shared__ unsigned long long s_ph[40];                                      // phase timers (thread 0's clock between barriers)
    long long tprev = 0;
#define V4_TICK(i) do { if (tid == 0) { const long long now_ = clock64(); s_ph[i] += (unsigned long long) (now_ - tprev); tprev = now_; } } while (0)
    if (tid < 40) s_ph[tid] = 0;
    if (tid < 4) s_wmax[tid] = 0;
#if defined(ZL_V4_PROFILE)
    long long wt_ = 0;
#define V4_WT0() do { wt_ = clock64(); } while (0)
#define V4_WTICK(i) do { if (tid == 992) { const long long n_ = clock64(); s_ph[i] += (unsigned long long) (n_ - wt_); wt_ = n_; } } while (0)
#define V4_WMAX(i, v) do { if (lane == 0) atomicMax(&s_wmax[i], (unsigned) (v)); } while (0)
#define V4_WADD(i, v) do { if (lane == 0) atomicAdd(&s_wmax[i], (unsigned) (v)); } while (0)
#define V4_PADD(i, v) do { if (lane == 0) atomicAdd(&s_ph[i], (unsigned long long) (v)); } while (0)
#else
#define V4_WT0() do { } while (0)
#define V4_WTICK(i) do { } while (0)
#define V4_WMAX(i, v) do { } while (0)
#define V4_WADD(i, v) do { } while (0)
#define V4_PADD(i, v) do { } while (0)
#endif
    static_assert(kV4T == 1024, "the kernel assumes 32 warps (prefix helpers, wcnt rows, link groups)");
    const V4Layout L = v4_layout(dmax, lmax);
    V4Ctx c;
    v4_bind(c, smem_raw, L);
    c.in = a.in + (size_t) b * kBlockBytes; c.ilen = (int) a.ilen[b];
    c.ring = a.ring + (size_t) b * kRingStride; c.hash = a.hash + (size_t) b * kHashStride;
    c.tok = a.tok + (size_t) b * kTokStride; c.lit = a.lit + (size_t) b * kLitStride;
    c.sub = a.sub + (size_t) b * kMaxSubPerBlock; c.plan = a.plan + (size_t) b * kMaxSubPerBlock; c.base_level = base_level;
    const int ilen = c.ilen;
    if (crank != 0) {
        // ---------------------------------------------------------------- helper CTA: SPEC D for a share of the positions
        __shared__ int s_hip, s_hlevel;
        V4Ctx hc = c;                                                    // own byte ring, own copy of the records; the insert counters are the leader's
        hc.cnt = cluster.map_shared_rank(c.cnt, 0);
        uint32_t* l_hdr = cluster.map_shared_rank(c.hdr, 0); uint32_t* l_node = cluster.map_shared_rank(c.node, 0);
        uint32_t* l_nodeq = cluster.map_shared_rank(c.nodeq, 0); uint32_t* l_fx = cluster.map_shared_rank(c.fx, 0);
        uint32_t* l_fdec = cluster.map_shared_rank(c.fdec, 0);
        const V4Run* leader_run = cluster.map_shared_rank(&s_run, 0);
        const bool link_helper = CL >= 4;                                // the last helper builds the window's links instead of walking chains
        const bool i_link = link_helper && crank == CL - 1;
        uint16_t* l_link = cluster.map_shared_rank(c.link, 0);
        const int H = link_helper ? CL - 2 : CL - 1, chunk = (kV4N + H - 1) / H;
        const int r0 = (crank - 1) * chunk, r1 = r0 + chunk < kV4N ? r0 + chunk : kV4N;
        const int r2 = r1 + 2 < kV4N ? r1 + 2 : kV4N;                    // two more records: the lazy tests of the last positions look at them
        const int span = r2 - r0 > 0 ? r2 - r0 : 1;
        const int T = kV4T / span < 1 ? 1 : (kV4T / span > DMAX ? DMAX : kV4T / span);   // threads per position in the compare step
        uint32_t* qall = reinterpret_cast<uint32_t*>(smem_raw + L.scratch);   // [kV4N * DMAX] candidate positions (the helpers do not use the scratch area otherwise)
        const int hlim = ilen - kGuard;
        const int hnwin = hlim > 2 ? (hlim + kV4W - 1) / kV4W : 0;
        int hstaged = -16;
        for (int k = 0; k < hnwin; k++) {
            const int lo = k * kV4W;
            const int wend = lo + kV4W < hlim ? lo + kV4W : hlim;
            const int hi = v4_stage_hi(k);
            for (int src = hstaged + tid * 16; src < hi; src += kV4T * 16) v4_stage16(c, src);
            hstaged = hi;
            __syncthreads();
#if defined(ZL_V4_PROFILE)
            long long ht_ = clock64();
#define V4_HTICK(i) do { if (tid == 0 && (crank == 1 || i_link)) { const long long n_ = clock64(); s_ph[(i_link ? 36 : 32) + (i)] += (unsigned long long) (n_ - ht_); ht_ = n_; } } while (0)
#else
#define V4_HTICK(i) do { } while (0)
#endif
            cluster.sync();                                              // B1: the leader has finished the previous window (G, counters, run state)
            V4_HTICK(0);
            if (tid == 0) { s_hip = leader_run->ip; s_hlevel = leader_run->level; }
            __syncthreads();
            if (s_hip >= wend) continue;                                 // the leader skips this window too
            if (i_link) {
                uint16_t* gtab = reinterpret_cast<uint16_t*>(smem_raw + L.scratch);
                c.key[tid] = v4_key_of(c, lo + tid);
                { uint4* z = reinterpret_cast<uint4*>(gtab); for (int i = tid; i < kV4ScratchSpec / 16; i += kV4T) z[i] = make_uint4(0, 0, 0, 0); }
                __syncthreads();
                v4_build_links_cta(c, gtab, tid, lane, warp);
                l_link[tid] = c.link[tid];
                V4_HTICK(1);
                cluster.sync();                                          // B2
                V4_HTICK(2);
                continue;
            }
            const int rel = r0 + tid;
            if (rel < r2) {
                c.key[rel] = v4_key_of(c, lo + rel);
                v4_spec_walk(hc, lo, rel, qall);
            }
            __syncthreads();
            V4_HTICK(1);
            {   // compares: T threads per position, each takes every T-th recorded node
                const int prel = r0 + tid / T, sub = tid % T;
                if (prel < r2) {
                    const int nv = (int) (c.hdr[prel] & 31u);
                    for (int i = sub; i < nv; i += T) v4_spec_compare(c, lo, prel, i, qall);
                }
            }
            __syncthreads();
            if (rel < r1) {
                if (rel < kV4W) { v4_frozen_core(c, lo, rel, s_hlevel); l_fdec[rel] = c.fdec[rel]; }
                const uint32_t hd = c.hdr[rel];
                const int nv = (int) (hd & 31u);                          // only the recorded nodes travel
                l_hdr[rel] = hd; l_fx[rel] = c.fx[rel];
                for (int i = 0; i < nv; i++) l_node[rel * DMAX + i] = c.node[rel * DMAX + i];
                for (int i = 0; i < nv && i < LMAX; i++) l_nodeq[rel * LMAX + i] = c.nodeq[rel * LMAX + i];
            }
            V4_HTICK(2);
            cluster.sync();                                              // B2: records and frozen decisions are in the leader's shared memory
            V4_HTICK(3);
        }
#if defined(ZL_V4_PROFILE)
        if (tid == 0 && counters && (crank == 1 || i_link)) for (int i = 32; i < 40; i++) atomicAdd(&counters->ph[i], s_ph[i]);
#endif
        return;
    }
    const bool link_helper = CL >= 4;                                    // the cluster's last CTA builds the links of every window
    V4Ctx cc = c;                                                        // the same state, evaluated by a whole warp per position
    cc.coop = 1;
    uint8_t* scratch = smem_raw + L.scratch;
    uint16_t* gtab = reinterpret_cast<uint16_t*>(scratch);                         // SPEC: [kV4Groups][kV4Buckets] last position + 1 per bucket
    uint16_t* E = reinterpret_cast<uint16_t*>(scratch);                            // ROUNDS: first position past its own 32-position segment on the orbit of each position
    uint16_t* E4 = reinterpret_cast<uint16_t*>(scratch + 2048);                     // ROUNDS: the same for the 128-position group of each position
    uint16_t* qhaz = reinterpret_cast<uint16_t*>(scratch + 24576);                  // ROUNDS: positions waiting for the hazard check,
    uint16_t* qgen = qhaz + kV4T;                                                   //         for the full probe on the pending view,
    uint16_t* qmru = qgen + kV4T;                                                   //         for the word-MRU test

    for (int i = tid; i < 256; i += kV4T) { c.cnt[i] = 0; c.mru[i] = 0; c.pcnt[i] = 0; }
    // Level of the block's FIRST sub-block when the host left it open (blocks after the first of a call): the reference
    // carries it over from the last sub-block of the previous block (current_level outlives the block loop,
    // src/libzling.cpp:185,261-266), which is being parsed by another CTA right now.  Predict it from the order-0 entropy of
    // the previous block's last 64 KiB: data that Huffman cannot shrink below 0.95 of its size (the reference's rule) has
    // a near-uniform byte histogram.  The host verifies the level afterwards and re-parses on a wrong guess.
    int level0_pred = base_level;
    const bool has_pre = b > 0 || (a.pre_tail && a.pre_tail[65536]);     // block 0 of a sharded range: the bytes in front of it came from the rank before
    if (has_pre && a.plan[(size_t) b * kMaxSubPerBlock] == kV4Auto && base_level != 0) {
        __syncthreads();
        const uint8_t* tail = b > 0 ? c.in - 65536 : a.pre_tail;         // the previous block's tail (blocks are contiguous, full except the last)
        for (int i = tid * 16; i < 65536; i += kV4T * 16) {
            const uint4 v = z4_ld_in128(reinterpret_cast<const uint4*>(tail + i));
            const uint32_t wv[4] = { v.x, v.y, v.z, v.w };
            #pragma unroll
            for (int q = 0; q < 16; q++) atomicAdd(&c.pcnt[(wv[q >> 2] >> ((q & 3) * 8)) & 0xffu], 1u);
        }
        __syncthreads();
        float part = 0.f;
        if (tid < 256) { const float n = (float) c.pcnt[tid]; part = n > 0.f ? n * __log2f(n) : 0.f; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __shared__ float s_ent[8];
        if (tid < 256 && lane == 0) s_ent[warp] = part;
        __syncthreads();
        float sum = 0.f;
        for (int i = 0; i < 8; i++) sum += s_ent[i];
        const float H = 16.f - sum / 65536.f;                            // bits per byte
        if (H > 7.6f) level0_pred = 0;
    }
    long long cyc_spec = 0, cyc_rounds = 0, cyc_final = 0, cyc_orbit = 0, cyc_rank = 0, cyc_decide = 0;
    unsigned long long n_rounds = 0, n_windows = 0;
    const long long t_begin = clock64();
    if (tid == 0) {
        V4Run r;
        r.ip = 0; r.op = 0; r.j = 0; r.tok_begin = 0; r.enc_begin = 0; r.prev_lit = 0; r.skip_push = 1; r.tail = 0;
        r.level = c.plan[0] == kV4Auto ? level0_pred : v4_next_level(c, 0, 0, 0);
        int nt = 0;
        for (int first = 0; first < 2; first++) {                        // first two bytes raw, lz.cpp:150-151
            if (r.ip == first && r.ip < ilen) { c.tok[nt++] = tok_literal(c.in[r.ip], 0, true); r.op++; r.ip++; }
        }
        s_run = r; s_nt = nt; s_nl = 0;
    }
    __syncthreads();
    const int lim = ilen - kGuard;
    const int nwin = lim > 2 ? (lim + kV4W - 1) / kV4W : 0;              // windows that contain probe positions
    int staged_hi = -16;
    const int segbase = warp * 32, segend = segbase + 32;
    uint32_t* const tabA = c.mcnt;                                       // per-round tables, set 0: mcnt, pushw, pf, plit as laid out
    uint32_t* const tabB = reinterpret_cast<uint32_t*>(scratch + 32768); // set 1 lives in the scratch area (free during the rounds)
    uint32_t* const pushwA = c.pushw; uint32_t* const pfA = c.pf; uint8_t* const plitA = c.plit;

    for (int k = 0; k < nwin; k++) {
        const int lo = k * kV4W;
        const int wend = lo + kV4W < lim ? lo + kV4W : lim;
        const int Wn = wend - lo;
        const int hi = v4_stage_hi(k);
        for (int src = staged_hi + tid * 16; src < hi; src += kV4T * 16) v4_stage16(c, src);
        staged_hi = hi;
        __syncthreads();
        if (CL > 1) cluster.sync();                                      // B1 (see the helper loop above)
        if (s_run.ip >= wend) continue;                                  // no token starts in this window (uniform)
        const long long t0 = clock64();
        tprev = t0;
        // ================================================= SPEC =================================================
        c.key[tid] = v4_key_of(c, lo + tid);
        for (int i = tid; i < 256 * kV4Words; i += kV4T) c.occ[i] = 0;
        if (tid < 256) { c.pcnt[tid] = 0; c.occw[tid] = 0; }
        if (!link_helper) { uint4* z = reinterpret_cast<uint4*>(gtab); for (int i = tid; i < kV4ScratchSpec / 16; i += kV4T) z[i] = make_uint4(0, 0, 0, 0); }
        __syncthreads();
        V4_TICK(0);
        {   // occ: bit i of occ[v] <=> in[lo + i - 3] == v; thread t owns bit t (word = warp), threads 0..1 also bits N, N+1
            const int p = lo + tid - 3;
            const uint32_t v = p >= 0 ? v4_rb8(c.rbw, (uint32_t) p) : 256u + (uint32_t) lane;
            const uint32_t grp = __match_any_sync(0xffffffffu, v);
            if (p >= 0 && (grp >> lane) == 1u) { c.occ[v * kV4Words + warp] = grp; atomicAdd(&c.pcnt[v], (uint32_t) __popc(grp)); atomicOr(&c.occw[v], 1u << warp); }
            if (tid < 2) { const int p2 = lo + kV4N + tid - 3; const uint32_t v2 = v4_rb8(c.rbw, (uint32_t) p2); atomicOr(&c.occ[v2 * kV4Words + (kV4N >> 5)], 1u << tid); atomicAdd(&c.pcnt[v2], 1u); }
        }
        if (CL == 1) v4_spec_position(c, lo, tid);                       // chain records against G (with a cluster: done by the helper CTAs)
        V4_TICK(1);
        if (!link_helper) {
            v4_build_links_cta(c, gtab, tid, lane, warp);                // (with a cluster of >= 4 CTAs: done by the last helper CTA)
            __syncthreads();
        }
        V4_TICK(3);
        if (CL > 1) cluster.sync();                                      // B2: the helpers' records have arrived
        V4_TICK(4);
        const int wlevel = s_run.level;
        if (tid < kV4W) { if (CL == 1) v4_frozen_core(c, lo, tid, wlevel); v4_frozen_flags(c, tid); }
        if (tid == 0) {
            V4Win w; w.lo = lo; w.wend = wend; w.entry = s_run.ip; w.level = wlevel; w.rpos = -1; w.level2 = wlevel;
            w.skip_push = s_run.skip_push; w.prev_lit = s_run.prev_lit;
            s_win = w;
        }
        { const uint32_t fl = tid < kV4W ? c.fdec[tid] & 511u : 0u; c.dec[tid] = fl ? (fl | (kV4Match << 9)) : (kV4Lit << 9); }
        __syncthreads();
        V4_TICK(5);
        const long long t1 = clock64();
        cyc_spec += t1 - t0;
        // ================================================= ROUNDS ===============================================
        // Three barriers per round: (1) orbit tables ready — it also carries the vote "did the previous round change a decision";
        // (2) marks + per-round tables + stage queues ready; (3) decisions of the round written.  The per-round tables
        // (mcnt, pushw, pf, plit) exist twice, so that the set of the next round is cleared while this round's is still read
        // (and FINALIZE finds the set of the last round untouched).
        const int entry_rel = s_win.entry - lo;
        const bool may_roll = s_run.op + 2 * kV4N + 1 >= kSubSymbols;
        bool marked = false;
        uint32_t mydec = c.dec[tid];
        int par = 0, ch = 1;
        while (true) {
            const long long r0 = clock64();
            // ---- orbit of the entry under next = x + step(decision).  Inside a 32-position segment: pointer doubling with
            // shuffles; across segments: every warp chases the segment exits from the window's entry up to its own segment
            int cj[6];
            { int t = tid < Wn ? tid + (int) v4_dec_step(mydec) : Wn; cj[0] = t < Wn ? t : Wn; }
            #pragma unroll
            for (int l = 0; l < 5; l++) { const int nx = __shfl_sync(0xffffffffu, cj[l], cj[l] & 31); cj[l + 1] = cj[l] < segend ? nx : cj[l]; }
            E[tid] = (uint16_t) cj[5];
            {   // second level: first position past the own 128-position group (the four warps of the group synchronise among themselves)
                asm volatile("bar.sync %0, 128;" :: "r"(1 + (warp >> 2)) : "memory");
                const int gend = ((warp >> 2) + 1) * 128;
                int e = cj[5];
                while (e < gend && e < Wn) e = E[e];
                E4[tid] = (uint16_t) e;
            }
            uint32_t* const n_mcnt = par ? tabB : tabA;
            uint32_t* const n_pushw = par ? tabB + 256 : pushwA;
            uint32_t* const n_pf = par ? tabB + 512 : pfA;
            uint8_t* const n_plit = par ? reinterpret_cast<uint8_t*>(tabB + 512 + kV4PfWords) : plitA;
            n_plit[tid] = 0;
            if (tid < 256) { n_mcnt[tid] = 0; n_pushw[tid] = 0; }
            if (tid >= 256 && tid < 256 + kV4PfWords) n_pf[tid - 256] = 0;
            if (tid == 0) s_rpos_rel = 0x7fffffff;
            if (tid < 3) { s_nq[tid] = 0; s_take[tid] = 0; }
            const int changed = __syncthreads_or(ch);                    // barrier 1
            if (!changed) break;
            n_rounds++;
            c.mcnt = n_mcnt; c.pushw = n_pushw; c.pf = n_pf; c.plit = n_plit;
            cc.mcnt = n_mcnt; cc.pushw = n_pushw; cc.pf = n_pf; cc.plit = n_plit;
            V4_WT0();
            int cur = entry_rel;
            while ((cur >> 7) < (warp >> 2) && cur < Wn) cur = E4[cur];
            while ((cur >> 5) < warp && cur < Wn) cur = E[cur];
            uint32_t M = ((cur >> 5) == warp && cur < Wn) ? 1u << (cur & 31) : 0u;
            #pragma unroll
            for (int l = 4; l >= 0; l--) {
                const int tg = cj[l];
                const uint32_t contrib = (((M >> lane) & 1u) && tg < segend && tg < Wn) ? 1u << (tg & 31) : 0u;
                M |= __reduce_or_sync(0xffffffffu, contrib);
            }
            V4_WTICK(9);
            marked = (M >> lane) & 1u;
            c.mark[tid] = (uint8_t) marked;
            if (lane == 0) c.mbits[warp] = M;
            V4_WTICK(10);
            {   // mcnt[ctx] = marked positions per context (bounds the inserts a record can have missed)
                const uint32_t cv = marked ? v4_ctx_of(c.key[tid]) : 256u;   // (one group for all unmarked lanes: the match costs per distinct value)
                const uint32_t grp = __match_any_sync(0xffffffffu, cv);
                if (marked && (grp >> lane) == 1u) atomicAdd(&c.mcnt[cv], (uint32_t) __popc(grp));
            }
            {   // pushw: one atomic per warp and context pushed into
                const uint32_t c3v = marked ? v4_rb8(c.rbw, (uint32_t) (lo + tid) - 3u) : 256u;
                const uint32_t grp = __match_any_sync(0xffffffffu, c3v);
                if (marked && (grp >> lane) == 1u) atomicOr(&c.pushw[c3v], 1u << warp);
            }
            if (marked) {
                uint32_t c3, pw;
                v4_push_of(c, (uint32_t) (lo + tid), &c3, &pw);
                const uint32_t h = v4_pf_hash(c3, pw);
                atomicOr(&c.pf[h >> 5], 1u << (h & 31u));
                const int t = tid + (int) v4_dec_step(mydec);
                if (t < Wn) c.plit[t] = v4_dec_kind(mydec) == kV4Lit;
                else { s_exit = lo + t; s_lastrel = tid; }
            }
            // ---- stage A: every MARKED position classifies itself (unmarked ones keep their frozen decision until the orbit
            // reaches them) into the queues of the stages below: 0 hazard check, 1 full probe, 2 word test
            auto classify = [&](const V4Win& w) {
                uint32_t nd = mydec;
                const int level_here = (w.rpos >= 0 && lo + tid >= w.rpos) ? w.level2 : w.level;
                int which = -1;
                if (marked && tid >= entry_rel && tid < Wn) {
                    const uint32_t fd = c.fdec[tid], fxw = c.fx[tid];
                    if (level_here != w.level) which = 1;
                    else if (fxw & kV4F_HAZ) which = 0;
                    else if (fd & 511u) nd = v4_dec_match(fd & 511u, (fd >> 18) & (kRing - 1));
                    else which = 2;
                }
                #pragma unroll
                for (int qi = 0; qi < 3; qi++) {                         // one shared-memory atomic per warp and queue
                    const uint32_t bal = __ballot_sync(0xffffffffu, which == qi);
                    if (bal) {
                        int base = 0;
                        if (lane == __ffs(bal) - 1) base = atomicAdd(&s_nq[qi], __popc(bal));
                        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
                        if (which == qi) (qi == 0 ? qhaz : qi == 1 ? qgen : qmru)[base + __popc(bal & v4_lt_mask(lane))] = (uint16_t) tid;
                    }
                }
                c.ndec[tid] = nd;
            };
            const long long r1 = clock64();
            V4_WTICK(11);
            if (!may_roll) classify(s_win);
            V4_WTICK(15);
            __syncthreads();                                             // barrier 2
            V4_TICK(13);
            // ---- sub-block roll-over (rare): symbols before each marked position
            if (may_roll) {
                const uint32_t mysym = marked ? v4_dec_syms(mydec) : 0u;
                uint32_t incl = mysym;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (lane == 31) s_wsym[warp] = (int) incl;
                __syncthreads();
                if (warp == 0) v4_warp0_prefix(s_wsym, lane);
                __syncthreads();
                const int before = s_run.op + s_wsym[warp] + (int) (incl - mysym);
                if (marked && before + 1 >= kSubSymbols) atomicMin(&s_rpos_rel, tid);
                __syncthreads();
                if (s_rpos_rel == tid) s_op_at_rpos = before;
                __syncthreads();
                if (tid == 0) {
                    V4Win w = s_win;
                    w.rpos = -1; w.level2 = w.level;
                    if (s_rpos_rel != 0x7fffffff) {
                        w.rpos = lo + s_rpos_rel;
                        w.level2 = v4_next_level(c, s_run.j + 1, w.rpos - s_run.enc_begin, s_op_at_rpos);
                    }
                    s_win = w;
                }
                __syncthreads();
                classify(s_win);
                __syncthreads();
            }
            const long long r2 = clock64();
            // ---- stages B + C + D without barriers between them: the warps take hazard-check entries one at a time (cc.coop = 1:
            // the 32 lanes evaluate the entry together), run the full probe right away when the hazard holds and the word test when
            // the probe finds no match; a warp that finds no entry left turns to the word tests stage A queued (one per lane)
            const V4Win w = s_win;
            const int n_haz = s_nq[0], n_gen = s_nq[1], n_mru = s_nq[2];     // stable: read behind barrier 2
            int n_items = 0; (void) n_items;
            while (true) {
                int i = 0;
                if (lane == 0) i = atomicAdd(&s_take[0], 1);
                i = __shfl_sync(0xffffffffu, i, 0);
                if (i >= n_haz + n_gen) break;
                const int rel = i < n_haz ? qhaz[i] : qgen[i - n_haz];
                const uint32_t fd = c.fdec[rel];
                uint32_t len = fd & 511u, rf = (fd >> 18) & (kRing - 1);
                n_items++;
#if defined(ZL_V4_PROFILE)
                const long long i0_ = clock64();
                const bool hz_ = i >= n_haz || v4_hazard(cc, rel, fd, c.fx[rel], depth_lazy2(w.level));
                const long long i1_ = clock64();
                V4_PADD(24, i1_ - i0_); V4_PADD(25, 1);
                if (hz_) {
                    len = (uint32_t) v4_probe_general(cc, lo, rel, (w.rpos >= 0 && lo + rel >= w.rpos) ? w.level2 : w.level, &rf);
                    V4_WADD(3, 1);
                    V4_PADD(26, clock64() - i1_); V4_PADD(27, 1);
                }
                const long long i2_ = clock64();
                const uint32_t nd = len ? v4_dec_match(len, rf) : v4_decide_word(cc, w, rel);
                if (!len) { V4_PADD(28, clock64() - i2_); V4_PADD(29, 1); }
                V4_WMAX(2, clock64() - i0_);                              // slowest single entry of the round
#else
                if (i >= n_haz || v4_hazard(cc, rel, fd, c.fx[rel], depth_lazy2(w.level)))
                    len = (uint32_t) v4_probe_general(cc, lo, rel, (w.rpos >= 0 && lo + rel >= w.rpos) ? w.level2 : w.level, &rf);
                const uint32_t nd = len ? v4_dec_match(len, rf) : v4_decide_word(cc, w, rel);
#endif
                if (lane == 0) c.ndec[rel] = nd;
            }
            V4_WMAX(0, clock64() - r2);
            while (true) {
                int i0 = 0;
                if (lane == 0) i0 = atomicAdd(&s_take[1], 32);
                i0 = __shfl_sync(0xffffffffu, i0, 0);
                if (i0 >= n_mru) break;
                if (i0 + lane < n_mru) { const int rel = qmru[i0 + lane]; c.ndec[rel] = v4_decide_word(c, w, rel); }
            }
            V4_WMAX(1, clock64() - r2);
            __syncthreads();                                             // barrier 3
            V4_TICK(14);
            if (tid == 0) {
                s_ph[20] += (unsigned long long) n_haz; s_ph[21] += (unsigned long long) n_gen + s_wmax[3]; s_ph[22] += (unsigned long long) n_mru;
                s_ph[16] += s_wmax[0]; s_ph[17] += s_wmax[1]; s_ph[18] += s_wmax[2];
                s_wmax[0] = s_wmax[1] = s_wmax[2] = s_wmax[3] = 0;
            }
            const uint32_t nd = c.ndec[tid];
            ch = marked && ((nd ^ mydec) & kV4DecCmp) != 0;
            c.dec[tid] = nd;
            mydec = nd;
            par ^= 1;
            const long long r3 = clock64();
            cyc_orbit += r1 - r0; cyc_rank += r2 - r1; cyc_decide += r3 - r2;
        }
        const long long t2 = clock64();
        tprev = t2;
        cyc_rounds += t2 - t1;
        n_windows++;
        // ================================================= FINALIZE =============================================
        {
            const V4Win w = s_win;
            const uint32_t d = c.dec[tid];
            // ring slot of every pending insert: per-context rank of the marked position (bitset popcount below it)
            c.sup[tid] = 0;
            if (marked) c.rank[tid] = (uint16_t) v4_rank_live(c, tid, v4_ctx_of(c.key[tid]));
            const bool islit = marked && v4_dec_kind(d) == kV4Lit;
            const uint32_t bt = __ballot_sync(0xffffffffu, marked), bl = __ballot_sync(0xffffffffu, islit);
            const bool after = w.rpos >= 0 && lo + tid >= w.rpos;
            uint32_t sy = marked ? v4_dec_syms(d) : 0u;
            uint32_t sya = after ? sy : 0u;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sy += __shfl_xor_sync(0xffffffffu, sy, o); sya += __shfl_xor_sync(0xffffffffu, sya, o); }
            if (lane == 0) { s_wtok[warp] = __popc(bt); s_wlit[warp] = __popc(bl); s_wsym[warp] = (int) sy; s_wsya[warp] = (int) sya; }
            __syncthreads();
            V4_TICK(6);
            if (warp == 30) { v4_warp0_prefix(s_wtok, lane); v4_warp0_prefix(s_wlit, lane); }
            else if (warp == 31) { v4_warp0_prefix(s_wsym, lane); v4_warp0_prefix(s_wsya, lane); }
            if ((tid & 3) == 3) {                                        // carried word MRU: context c on thread 4 c + 3 (8 contexts per warp);
                const int cq = tid >> 2;                                 // pushw (last round's marks) tells which contexts received a push at all
                c.mru2[cq] = c.pushw[cq] ? v4_mru_state(c, w, Wn - 1, (uint32_t) cq) : (w.rpos >= 0 ? 0u : c.mru[cq]);
            }
            uint32_t suffix = 0;
            if (marked) suffix = v4_claim_slot(c, tid);
            __syncthreads();
            V4_TICK(7);
            if (marked) {
                v4_apply_position(c, lo, tid, suffix);
                const int ti = s_nt + s_wtok[warp] + __popc(bt & v4_lt_mask(lane));
                c.tok[ti] = v4_token_of(c, lo, tid);
                if (islit) c.lit[s_nl + s_wlit[warp] + __popc(bl & v4_lt_mask(lane))] = (uint32_t) ti;
                if (w.rpos >= 0 && lo + tid == w.rpos) s_rpos_nt = ti;
            }
            __syncthreads();
            V4_TICK(8);
            if (tid < 256) { c.mru[tid] = c.mru2[tid]; c.cnt[tid] += c.mcnt[tid]; }   // mcnt: marked positions per context of the last round = this window's inserts
            if (tid == 0) {
                V4Run r = s_run;
                if (w.rpos >= 0) {                                       // sub-block full (lz.cpp:153): close it, open the next
                    v4_close_subblock(c, r, w.rpos, s_op_at_rpos, s_rpos_nt);
                    r.j++; r.level = w.level2; r.tok_begin = s_rpos_nt; r.enc_begin = w.rpos;
                    r.op = s_wsya[32];
                } else {
                    r.op += s_wsym[32];
                }
                r.ip = s_exit; r.skip_push = 0; r.prev_lit = v4_dec_kind(c.dec[s_lastrel]) == kV4Lit;
                s_run = r; s_nt += s_wtok[32]; s_nl += s_wlit[32];
            }
        }
        __syncthreads();
        V4_TICK(12);
        cyc_final += clock64() - t2;
    }
    if (tid == 0) {
        V4Run r = s_run;
        r.tail = 1;
        int nt = s_nt, nl = s_nl;
        v4_resolve_tail(c, r, &nt, &nl);
        if (ilen > 0) v4_close_subblock(c, r, r.ip, r.op, nt);
        a.nsub[b] = ilen > 0 ? r.j + 1 : 0; a.ntok[b] = nt; a.nlit[b] = nl;
        if (counters) {
            atomicAdd(&counters->tokens, (unsigned long long) nt);
            atomicAdd(&counters->windows, n_windows);
            atomicAdd(&counters->rounds, n_rounds);
            atomicAdd(&counters->cyc_spec, (unsigned long long) cyc_spec);
            atomicAdd(&counters->cyc_rounds, (unsigned long long) cyc_rounds);
            atomicAdd(&counters->cyc_final, (unsigned long long) cyc_final);
            atomicAdd(&counters->cyc_orbit, (unsigned long long) cyc_orbit);
            atomicAdd(&counters->cyc_rank, (unsigned long long) cyc_rank);
            atomicAdd(&counters->cyc_decide, (unsigned long long) cyc_decide);
            atomicAdd(&counters->cyc_total, (unsigned long long) (clock64() - t_begin));
            for (int i = 0; i < 32; i++) atomicAdd(&counters->ph[i], s_ph[i]);
        }
    }
}

}  // namespace zl
#endif
