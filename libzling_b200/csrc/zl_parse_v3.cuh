// zl_rolz_parse_v3 — the ROLZ parse as a two-stage pipeline inside one CTA per 16 MiB block.
//
// The reference's parse (EncodeImpl + MatchAndUpdate + MatchLazy, src/libzling_lz.cpp:139-316) is one serial
// chain per block.  v2 (zl_parse_v2.cuh) alternated "speculate a window" / "resolve a window"; its resolver spent
// ~1300 cycles per token on warp-wide hazard scans.  v3 keeps v2's exactness argument and changes the schedule:
//
//   producers (8 warps)    SPEC(k+1): for every position of window k+1 walk its hash chain in the bucket state G
//                          frozen at the START OF WINDOW k (G is only written between steps, see APPLY), record
//                          up to dmax nodes with match lengths, byte-equality maps for the lazy probes, the link
//                          to the nearest earlier position with the same (context, hash slot) key, and the
//                          decision the reference takes if no later insert interferes.
//   resolver (1 thread)    RESOLVE(k): walks the real token chain through window k in shared memory only.  A
//                          position is decided by its frozen record unless (a) an earlier token start inside the
//                          two live windows has the same key (static link + "was it a token start" flag), (b) a
//                          ring slot the record read has been overwritten since (insert counters), or (c) the
//                          level in force differs from the one the record assumed.  (a)/(c) merge the in-window
//                          candidates with the record in the reference's visiting order; (b) replays the
//                          reference's walk literally on G + the pending inserts.
//   all threads            APPLY(k): scatter the inserts of window k into G (ring entry + slot head), snapshot the
//                          per-context insert counters; token words / literal lists are emitted from the marks here.
//
// SPEC(k+1) and RESOLVE(k) touch disjoint state, so they run concurrently; every function of the algorithm is
// plain scalar code marked ZL_HD, and tests/cxx/parse_v3_sim.cu replays the same phases on the host against the
// CPU checker (the GPU box then only has to confirm the synchronisation).  Bit-exactness argument: DESIGN.md §4.
#pragma once
#include "zl_kernels.cuh"

namespace zl {

#ifndef ZL_V3_PROD
#define ZL_V3_PROD 256
#endif
constexpr int kV3Prod    = ZL_V3_PROD;          // producer threads: up to the 12 warps of SM sub-partitions 1..3 (a multiple of 32)
constexpr int kV3Threads = 512;                 // 16 warps (128 registers per thread); warp 0 = resolver, alone on sub-partition 0:
                                                // warps 4, 8, 12 only help in the short APPLY/EMIT phases, so the resolver keeps its
                                                // scheduler and instruction cache to itself
constexpr int kV3W       = kV3Prod - 2;         // main positions per window; each table also holds 2 lazy look-ahead positions
constexpr int kV3R       = 2048;                // per-position ring (bytes, keys, links, insert marks): >= 3 W + 320
constexpr int kV3Buckets = 4096;                // bucket table of the link builder
constexpr int kV3Tail    = 288;                 // bytes staged past a window's last look-ahead position
constexpr uint32_t kKeyInvalid = 0x80000000u;   // position cannot be probed (first two bytes / last 273 bytes of the block)
constexpr uint32_t kKeyMask    = 0x1fffffu;     // (context << 13) | hash slot

// dec word 1 flag bits
constexpr uint32_t kF_SELF = 1u << 16, kF_L1 = 1u << 17, kF_L2 = 1u << 18, kF_ST0 = 1u << 19, kF_ST1 = 1u << 20, kF_ST2 = 1u << 21, kF_NOPROBE = 1u << 22;
constexpr uint32_t kF_ANY = 0x7fu << 16, kF_FORCE = 1u << 23;
// per-position token mark (ins[]): ring head (12 bits) | kind << 12 | flags
constexpr uint32_t kKindMatch = 1, kKindLit = 2, kKindWord0 = 3, kKindWord1 = 4;
constexpr uint32_t kInsExplicit = 1u << 16, kInsSuperseded = 1u << 17;

// ---- portable intrinsics -----------------------------------------------------------------------------------------
ZL_HD uint32_t z3_funnel(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
ZL_HD int z3_ffs(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int) v);
#else
    return __builtin_ffs((int) v);
#endif
}
ZL_HD uint32_t z3_bytes_eq_mask(uint32_t a, uint32_t b) {     // 4-bit mask, bit i set when byte i agrees
#if defined(__CUDA_ARCH__)
    const uint32_t m = __vcmpeq4(a, b);
    return ((m & 0x01010101u) * 0x01020408u) >> 24;
#else
    uint32_t r = 0, x = a ^ b;
    for (int i = 0; i < 4; i++) if (((x >> (8 * i)) & 0xffu) == 0) r |= 1u << i;
    return r;
#endif
}
ZL_HD uint64_t z3_ld_ring(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(reinterpret_cast<const unsigned long long*>(p));
#else
    return *p;
#endif
}
ZL_HD uint32_t z3_ld_hash(const uint16_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
ZL_HD uint32_t z3_ld_in32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
ZL_HD uint4 z3_ld_in128(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// unaligned little-endian 32-bit load from the input block (global memory); `in` is 16-byte aligned
ZL_HD uint32_t z3_in32(const uint8_t* in, uint32_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in) + (off >> 2);
    return z3_funnel(z3_ld_in32(w), z3_ld_in32(w + 1), (off & 3u) * 8u);
}
ZL_HD uint32_t z3_hash(uint32_t w) { return w + ((w >> 16) & 0xffu) * 137u + (w >> 24) * 13337u; }   // lz.cpp:55-57

// ---- shared-memory layout -------------------------------------------------------------------------------------------
struct V3Layout {
    int dmax, lmax;
    int rb, key, link, blink, llen, ins, suf, tw, last, cnt, snap, mru, pcnt, tab[2], total;
    int t_hdr, t_node, t_eq, t_dec, t_size;          // offsets inside one table
};
__host__ __device__ inline V3Layout v3_layout(int dmax, int lmax) {
    V3Layout L; L.dmax = dmax; L.lmax = lmax;
    int at = 0;
    auto take = [&](int bytes) { int o = at; at += (bytes + 15) & ~15; return o; };
    L.rb    = take(kV3R);
    L.key   = take(4 * kV3R);
    L.link  = take(2 * kV3R);
    L.blink = take(2 * kV3R);
    L.llen  = take(2 * kV3R);
    L.ins   = take(4 * kV3R);
    L.suf   = take(2 * kV3R);
    L.tw    = take(4 * kV3R);
    L.last  = take(4 * kV3Buckets);
    L.cnt   = take(4 * 256);
    L.snap  = take(4 * 256 * 3);
    L.mru   = take(4 * 256);
    L.pcnt  = take(4 * 256 * 2);
    const int n = kV3W + 2;
    int t = 0;
    auto ttake = [&](int bytes) { int o = t; t += (bytes + 15) & ~15; return o; };
    L.t_hdr = ttake(4 * n); L.t_node = ttake(4 * n * dmax); L.t_eq = ttake(20 * n * lmax); L.t_dec = ttake(16 * n);
    L.t_size = t;
    L.tab[0] = take(t); L.tab[1] = take(t);
    L.total = at;
    return L;
}

struct V3Table { uint32_t* hdr; uint32_t* node; uint32_t* eq; uint4* dec; };

struct V3Ctx {
    // block
    const uint8_t* in; int ilen;
    uint64_t* ring; uint16_t* hash;             // G: bucket state in global memory
    uint32_t* tok; uint32_t* lit; SubBlock* sub; const uint8_t* plan; int base_level;
    // shared memory
    uint32_t* rbw;                              // input bytes, ring of kV3R bytes viewed as words
    uint32_t* key; uint16_t* link; uint16_t* blink; uint16_t* llen; uint32_t* ins; uint16_t* suf; uint32_t* tw;
    uint32_t* last; uint32_t* cnt; uint32_t* snap; uint32_t* mru; uint32_t* pcnt;
    uint8_t* tab0; int tab_stride, t_hdr, t_node, t_eq, t_dec;   // two tables, selected arithmetically (no dynamic struct indexing)
    int dmax, lmax;
};
ZL_HD V3Table v3_table(const V3Ctx& c, int j) {
    uint8_t* b = c.tab0 + (j & 1) * c.tab_stride;
    V3Table t;
    t.hdr = (uint32_t*) (b + c.t_hdr); t.node = (uint32_t*) (b + c.t_node); t.eq = (uint32_t*) (b + c.t_eq); t.dec = (uint4*) (b + c.t_dec);
    return t;
}

__host__ __device__ inline void v3_bind(V3Ctx& c, uint8_t* smem, const V3Layout& L) {
    c.rbw = (uint32_t*) (smem + L.rb); c.key = (uint32_t*) (smem + L.key); c.link = (uint16_t*) (smem + L.link);
    c.blink = (uint16_t*) (smem + L.blink); c.llen = (uint16_t*) (smem + L.llen); c.ins = (uint32_t*) (smem + L.ins); c.suf = (uint16_t*) (smem + L.suf); c.tw = (uint32_t*) (smem + L.tw);
    c.last = (uint32_t*) (smem + L.last); c.cnt = (uint32_t*) (smem + L.cnt); c.snap = (uint32_t*) (smem + L.snap);
    c.mru = (uint32_t*) (smem + L.mru); c.pcnt = (uint32_t*) (smem + L.pcnt);
    c.tab0 = smem + L.tab[0]; c.tab_stride = L.tab[1] - L.tab[0];
    c.t_hdr = L.t_hdr; c.t_node = L.t_node; c.t_eq = L.t_eq; c.t_dec = L.t_dec;
    c.dmax = L.dmax; c.lmax = L.lmax;
}

// window geometry: table j holds positions [j W, (j+1) W + 2); its records are relative to the bucket state that
// contains exactly the inserts made at positions < base(j)
ZL_HD int v3_base(int j) { return j >= 1 ? (j - 1) * kV3W : 0; }
ZL_HD int v3_stage_hi(int j) { return (((j + 1) * kV3W + 2 + kV3Tail) + 15) & ~15; }   // bytes [.., hi) staged once SPEC(j) ran
ZL_HD int v3_new_lo(int j) { return j == 0 ? 0 : j * kV3W + 2; }                      // positions first seen by SPEC(j)
ZL_HD int v3_new_hi(int j) { return (j + 1) * kV3W + 2; }

// ---- byte ring ------------------------------------------------------------------------------------------------------
ZL_HD uint32_t v3_rb32(const uint32_t* rbw, uint32_t pos) {          // unaligned LE 32-bit load at block position pos
    const uint32_t i = (pos >> 2) & (kV3R / 4 - 1);
    return z3_funnel(rbw[i], rbw[(i + 1) & (kV3R / 4 - 1)], (pos & 3u) * 8u);
}
ZL_HD uint32_t v3_rb8(const uint32_t* rbw, uint32_t pos) {
    return (rbw[(pos >> 2) & (kV3R / 4 - 1)] >> ((pos & 3u) * 8u)) & 0xffu;
}
// stage 16 input bytes at block offset src (multiple of 16, may be negative or past the block: zeros)
ZL_HD void v3_stage16(const V3Ctx& c, int src) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (src >= 0 && src < c.ilen) {
        v = z3_ld_in128(reinterpret_cast<const uint4*>(c.in + src));
        const int over = src + 16 - c.ilen;                              // bytes past the block end are staged as zeros
        if (over > 0) {
            uint32_t w[4] = { v.x, v.y, v.z, v.w };
            for (int b = 16 - over; b < 16; b++) w[b >> 2] &= ~(0xffu << ((b & 3) * 8));
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    uint32_t* d = c.rbw + (((uint32_t) src >> 2) & (kV3R / 4 - 1));
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}

// ---- SPEC phase A: key of a position -------------------------------------------------------------------------------
// pcnt[j & 1][ctx] = positions first seen by SPEC(j) whose context byte is ctx: an upper bound of the inserts they can
// make into that context (cleared by the caller before the pass)
ZL_HD void v3_key_position(const V3Ctx& c, int x, int j) {
    uint32_t k = kKeyInvalid;
    if (x >= 2 && x + 273 < c.ilen) {
        const uint32_t h = z3_hash(v3_rb32(c.rbw, (uint32_t) x));
        const uint32_t ctx = v3_rb8(c.rbw, (uint32_t) x - 1);
        k = (ctx << 13) | (h & (kSlots - 1)) | (((h >> 13) & 0xffu) << 21);
#if defined(__CUDA_ARCH__)
        atomicAdd(&c.pcnt[256 * (j & 1) + ctx], 1u);
#else
        c.pcnt[256 * (j & 1) + ctx]++;
#endif
    }
    c.key[x & (kV3R - 1)] = k;
    c.ins[x & (kV3R - 1)] = 0;
}
ZL_HD uint32_t v3_bucket_of(uint32_t key) { return ((key & kKeyMask) * 2654435761u) >> 20; }   // 12 bits

// ---- SPEC phase B: bucket chains, positions in increasing order (host / reference form; the kernel runs the same
// recurrence 32 positions at a time with __match_any_sync) ------------------------------------------------------------
inline void v3_bucket_pass_serial(const V3Ctx& c, int lo, int hi) {
    for (int x = lo; x < hi; x++) {
        const uint32_t k = c.key[x & (kV3R - 1)];
        uint32_t d = 0;
        if (!(k & kKeyInvalid)) {
            const uint32_t b = v3_bucket_of(k);
            const uint32_t prev = c.last[b];
            if (prev != 0 && (uint32_t) x - (prev - 1) < (uint32_t) kV3R) d = (uint32_t) x - (prev - 1);
            c.last[b] = (uint32_t) x + 1;
        }
        c.blink[x & (kV3R - 1)] = (uint16_t) d;
    }
}

ZL_HD int v3_common_len_ring(const uint32_t* rbw, uint32_t p, uint32_t q);

// ---- SPEC phase C: nearest earlier position with the same key, not before `lb`, and the match length against it ------
ZL_HD void v3_link_position(const V3Ctx& c, int x, int lb) {
    const uint32_t k = c.key[x & (kV3R - 1)];
    uint32_t out = 0;
    if (!(k & kKeyInvalid)) {
        int y = x;
        uint32_t d = c.blink[x & (kV3R - 1)];
        while (d != 0) {
            y -= (int) d;
            if (y < lb) break;
            if (((c.key[y & (kV3R - 1)] ^ k) & kKeyMask) == 0) { out = (uint32_t) (x - y); break; }
            d = c.blink[y & (kV3R - 1)];
        }
    }
    c.link[x & (kV3R - 1)] = (uint16_t) out;
    // GetCommonLength against that position: the candidate the resolver's general path needs most often
    c.llen[x & (kV3R - 1)] = out ? (uint16_t) v3_common_len_ring(c.rbw, (uint32_t) x, (uint32_t) x - out) : (uint16_t) 0;
}

// ---- SPEC phase D: chain record of a position against the frozen bucket state ---------------------------------------
ZL_HD uint32_t v3_ring_dist(uint32_t slot, uint32_t head_at_base) {      // inserts into the context until `slot` is overwritten (1..4096)
    return ((slot - head_at_base - 1u) & (kRing - 1)) + 1u;
}
// exact GetCommonLength (lz.cpp:66-89); own side from the byte ring up to 132, then both sides from global memory
ZL_HD int v3_common_len_global(const uint8_t* in, uint32_t x, uint32_t q, int from) {
    for (int n = from; n < 256; n += 4) {
        const uint32_t d = z3_in32(in, x + n) ^ z3_in32(in, q + n);
        if (d) return n + ((z3_ffs(d) - 1) >> 3);
    }
    const uint32_t d = z3_in32(in, x + 256) ^ z3_in32(in, q + 256);
    const int t = d ? ((z3_ffs(d) - 1) >> 3) : 4;
    return 256 + (t < 3 ? t : 3);
}
template <int WO>
ZL_HD void v3_eq_body(const uint32_t (&qw)[41], uint32_t sh, const uint32_t* rbw, uint32_t x, uint32_t* out5) {
    #pragma unroll
    for (int g = 0; g < 5; g++) {
        uint32_t bits = 0;
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const int wj = g * 8 + j;
            if (wj < 33) {
                const uint32_t qword = z3_funnel(qw[wj + WO], qw[wj + WO + 1], sh);
                bits |= z3_bytes_eq_mask(qword, v3_rb32(rbw, x + 4 * wj)) << (4 * j);
            }
        }
        out5[g] = bits;
    }
}
// bit j of the map = (in[x+j] == in[q+j]), j < 132
ZL_HD void v3_eq_bits(const uint8_t* in, const uint32_t* rbw, uint32_t x, uint32_t q, uint32_t* out5) {
    uint32_t qw[41];
    const uint4* qa = reinterpret_cast<const uint4*>(in + (q & ~15u));
    #pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint4 v = z3_ld_in128(qa + i);
        qw[4 * i] = v.x; qw[4 * i + 1] = v.y; qw[4 * i + 2] = v.z; qw[4 * i + 3] = v.w;
    }
    qw[40] = 0;
    const uint32_t sh = (q & 3u) * 8u;
    switch ((q >> 2) & 3u) {
        case 0:  v3_eq_body<0>(qw, sh, rbw, x, out5); break;
        case 1:  v3_eq_body<1>(qw, sh, rbw, x, out5); break;
        case 2:  v3_eq_body<2>(qw, sh, rbw, x, out5); break;
        default: v3_eq_body<3>(qw, sh, rbw, x, out5); break;
    }
}
ZL_HD int v3_len_from_eq(const uint8_t* in, uint32_t x, uint32_t q, const uint32_t* eq5) {
    #pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t z = ~eq5[g];
        if (z) { const int n = g * 32 + z3_ffs(z) - 1; return n < kMinLen ? 0 : n; }
    }
    const uint32_t z = ~eq5[4] & 0xfu;
    if (z) return 128 + z3_ffs(z) - 1;
    return v3_common_len_global(in, x, q, 132);
}
ZL_HD void v3_spec_position(const V3Ctx& c, int j, int rel) {
    const V3Table t = v3_table(c, j);
    const int x = j * kV3W + rel;
    const uint32_t k = c.key[x & (kV3R - 1)];
    if (k & kKeyInvalid) { t.hdr[rel] = (uint32_t) (kRing - 1) << 5; return; }
    const uint32_t ctx = (k >> 13) & 0xffu, slot = k & (kSlots - 1), chk = k >> 21;
    const uint32_t* snap = c.snap + 256 * ((j + 2) % 3);               // counters at base(j): snapshot j-1 (all zero for j = 0)
    const uint32_t head_b = snap[ctx] & (kRing - 1);
    const uint64_t* rc = c.ring + (size_t) ctx * kRing;
    uint32_t node = z3_ld_hash(c.hash + (size_t) ctx * kSlots + slot);
    uint32_t nvis = 0, dmin = kRing;
    if (node != (uint32_t) kNil) {
        dmin = v3_ring_dist(node, head_b);
        uint64_t e = z3_ld_ring(rc + node);
        for (int i = 0; i < c.dmax; i++) {
            const uint32_t q = ring_pos(e);
            const uint32_t nxt = ring_suffix(e);
            // the next chain node is fetched while this one's bytes are compared (independent loads overlap)
            const uint64_t e2 = nxt != (uint32_t) kNil ? z3_ld_ring(rc + nxt) : 0ull;
            int len = 0;
            if (i < c.lmax) {
                uint32_t* eq5 = t.eq + (size_t) (rel * c.lmax + i) * 5;
                v3_eq_bits(c.in, c.rbw, (uint32_t) x, q, eq5);
                if (ring_check(e) == chk) len = v3_len_from_eq(c.in, (uint32_t) x, q, eq5);
            } else if (ring_check(e) == chk) {
                len = z3_in32(c.in, (uint32_t) x) == z3_in32(c.in, q) ? v3_common_len_global(c.in, (uint32_t) x, q, 4) : 0;
            }
            t.node[rel * c.dmax + i] = (uint32_t) len | (node << 9);
            nvis = i + 1;
            if (nxt == (uint32_t) kNil) break;
            dmin = min(dmin, v3_ring_dist(nxt, head_b));
            if (q <= ring_pos(e2)) break;
            node = nxt; e = e2;
        }
    }
    t.hdr[rel] = nvis | ((dmin - 1u) << 5);
}

// ---- SPEC phase E: the frozen decision of a main position -----------------------------------------------------------
// dec[rel] (16 bytes, ONE shared-memory load per token for the resolver):
//   .x  flen(9) | fbest(9) << 9 | fslot(12) << 18     frozen match length after the lazy veto / before it / ring slot of the best node
//   .y  fhead(16) | flags << 16 | ctx << 24            frozen slot head (the insert's suffix), hazard flags, context byte in[x-1]
//   .z  in[x-3] | (in[x-2] << 8 | in[x-1]) << 8        the word-MRU push made when a token ENDS at x (lz.cpp:163-166,183-185,190-191)
//   .w  in[x] << 8 | in[x+1]                           the word tested at x when no match is taken (lz.cpp:172-185)
ZL_HD bool v3_eq_hit(const uint32_t* eqw, uint32_t at) {                 // bytes [at, at+4) agree
    return (z3_funnel(eqw[at >> 5], eqw[(at >> 5) + 1], at & 31u) & 0xfu) == 0xfu;
}
ZL_HD void v3_decide_position(const V3Ctx& c, int j, int rel, int level) {
    const V3Table t = v3_table(c, j);
    const int x = j * kV3W + rel;
    const int D = depth_main(level), L1 = depth_lazy1(level), L2 = depth_lazy2(level);
    const uint32_t hdr = t.hdr[rel];
    const int nvis = (int) (hdr & 31u);
    uint32_t fbest = 0, fslot = 0;
    const uint32_t fhead = nvis > 0 ? (t.node[rel * c.dmax] >> 9) : (uint32_t) kNil;
    const int take = nvis < D ? nvis : D;
    for (int i = 0; i < take; i++) {
        const uint32_t nd = t.node[rel * c.dmax + i];
        if ((nd & 511u) > fbest) { fbest = nd & 511u; fslot = nd >> 9; }
    }
    uint32_t flen = fbest;
    if (fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow) {
        const uint32_t at = fbest - 3u;
        for (int which = 1; which <= 2 && flen; which++) {
            const int depth = which == 1 ? L1 : L2;
            const int nvx = (int) (t.hdr[rel + which] & 31u);
            int tk = nvx < depth ? nvx : depth;
            if (tk > c.lmax) tk = c.lmax;
            for (int i = 0; i < tk; i++)
                if (v3_eq_hit(t.eq + (size_t) ((rel + which) * c.lmax + i) * 5, at)) flen = 0;
        }
    }
    // static hazard flags: which of x, x+1, x+2 have an earlier same-key position inside the live windows, and which
    // records read a ring slot that could be overwritten by the inserts of two windows
    // inserts into a context since base(j) <= positions of the live windows with that context byte (+2: the two
    // look-ahead positions of window j-2, which SPEC(j-2) counted)
    const uint32_t b3 = v3_rb32(c.rbw, (uint32_t) (x - 3));              // bytes x-3, x-2, x-1, x (zeros before the block)
    const uint32_t nxt = v3_rb8(c.rbw, (uint32_t) x + 1);
    const uint32_t ctx = (b3 >> 16) & 0xffu, ctx1 = b3 >> 24, ctx2 = nxt;
    uint32_t fl = 0;
    const bool lazy_matters = fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow;
    if (c.link[x & (kV3R - 1)]) fl |= kF_SELF;
    if (nvis > 0 && (hdr >> 5) + 1u <= c.pcnt[ctx] + c.pcnt[256 + ctx] + 2u) fl |= kF_ST0;
    if (lazy_matters) {
        if (c.link[(x + 1) & (kV3R - 1)]) fl |= kF_L1;
        { const uint32_t h1 = t.hdr[rel + 1]; if ((h1 & 31u) && (h1 >> 5) + 1u <= c.pcnt[ctx1] + c.pcnt[256 + ctx1] + 2u) fl |= kF_ST1; }
        if (L2 > 0) {
            if (c.link[(x + 2) & (kV3R - 1)]) fl |= kF_L2;
            { const uint32_t h2 = t.hdr[rel + 2]; if ((h2 & 31u) && (h2 >> 5) + 1u <= c.pcnt[ctx2] + c.pcnt[256 + ctx2] + 2u) fl |= kF_ST2; }
        }
    }
    const uint32_t kx = c.key[x & (kV3R - 1)];
    if (kx & kKeyInvalid) fl |= kF_NOPROBE;                              // first two / last 273 bytes of the block: generic path
    const uint32_t push = (b3 & 0xffu) | ((((b3 >> 8) & 0xffu) << 8 | ctx) << 8);
    t.dec[rel] = make_uint4(flen | (fbest << 9) | (fslot << 18), fhead | fl | (ctx << 24), push, ((b3 >> 24) << 8) | nxt);
}

// ---- per-position token marks ------------------------------------------------------------------------------------------
// ins[x] (u32): head(12) | kind << 12 | kInsExplicit | kInsSuperseded.  Every token that starts in the probe region
// inserts (lz.cpp:227-230), so "kind != 0" is also the "x was inserted" mark the hazard checks and APPLY look at.
ZL_HD uint32_t v3_kind(uint32_t m) { return (m >> 12) & 7u; }
ZL_HD uint32_t v3_suffix_of(const V3Ctx& c, int y, uint32_t m) {         // suffix of the pending insert at y
    if (m & kInsExplicit) return c.suf[y & (kV3R - 1)];
    const V3Table t = v3_table(c, y / kV3W);
    return t.dec[y - (y / kV3W) * kV3W].y & 0xffffu;                     // clean token: the frozen slot head
}

// ---- APPLY: scatter one position's insert into G --------------------------------------------------------------------
ZL_HD void v3_apply_position(const V3Ctx& c, int y) {
    const uint32_t m = c.ins[y & (kV3R - 1)];
    if (!v3_kind(m)) return;
    const uint32_t k = c.key[y & (kV3R - 1)];
    const uint32_t ctx = (k >> 13) & 0xffu, slot = k & (kSlots - 1), chk = k >> 21, head = m & (kRing - 1);
    c.ring[(size_t) ctx * kRing + head] = ring_make((uint32_t) y, chk, v3_suffix_of(c, y, m));
    if (!(m & kInsSuperseded)) c.hash[(size_t) ctx * kSlots + slot] = (uint16_t) head;
}

// ---- EMIT: the token word of a marked position (all threads, after the window is resolved) ----------------------------
ZL_HD uint32_t v3_token_of(const V3Ctx& c, int y, uint32_t m) {
    const uint32_t kind = v3_kind(m);
    if (kind == kKindLit) return tok_literal(v3_rb8(c.rbw, (uint32_t) y), v3_rb8(c.rbw, (uint32_t) y - 1), false);
    if (kind == kKindWord0) return tok_word(0);
    if (kind == kKindWord1) return tok_word(1);
    if (m & kInsExplicit) return c.tw[y & (kV3R - 1)];
    const V3Table t = v3_table(c, y / kV3W);
    const uint32_t d0 = t.dec[y - (y / kV3W) * kV3W].x;
    return tok_match(d0 & 511u, ((m & (kRing - 1)) - ((d0 >> 18) & (kRing - 1))) & (kRing - 1));
}

#if defined(ZL_V3_FLAG_HIST)
static unsigned long long g_flag_hist[256];
#endif
// ---- RESOLVE --------------------------------------------------------------------------------------------------------
struct V3Run {                       // resolver state carried across windows (one thread)
    int ip, op, j, level, tok_begin, enc_begin;
    int prev_lit;                    // the previous token was a literal (its word-MRU push is unconditional)
    int skip_push;                   // no push is pending on arrival (block start: the two raw bytes push nothing)
    int tail;                        // ip reached the last 275 bytes: the rest is done by v3_resolve_tail
    uint32_t n_general, n_slow, n_linkwalk, n_flagged;
    long long cyc_special;          // device: cycles spent on flagged tokens (hazard check + general path)
};

// GetCommonLength with both operands inside the byte ring
ZL_HD int v3_common_len_ring(const uint32_t* rbw, uint32_t p, uint32_t q) {
    if (v3_rb32(rbw, p) != v3_rb32(rbw, q)) return 0;
    for (int n = 4; n < 256; n += 4) {
        const uint32_t d = v3_rb32(rbw, p + n) ^ v3_rb32(rbw, q + n);
        if (d) return n + ((z3_ffs(d) - 1) >> 3);
    }
    const uint32_t d = v3_rb32(rbw, p + 256) ^ v3_rb32(rbw, q + 256);
    const int t = d ? ((z3_ffs(d) - 1) >> 3) : 4;
    return 256 + (t < 3 ? t : 3);
}
// is some earlier same-key position of z (not before base) a token start that inserted (or the position `self`)?
ZL_HD bool v3_link_hazard(const V3Ctx& c, int z, int base, int self) {
    int y = z;
    while (true) {
        const uint32_t d = c.link[y & (kV3R - 1)];
        if (!d) return false;
        y -= (int) d;
        if (y < base) return false;
        if (y == self || v3_kind(c.ins[y & (kV3R - 1)])) return true;
    }
}

// The reference's view of ring[ctx][n] / hash[ctx][slot] right now: G holds every insert before window k's start,
// the inserts of window k so far are still pending in the per-position arrays.
struct V3Live { const V3Ctx* c; int k; int upto; };      // pending = starts in [k W, upto]
ZL_HD uint64_t v3_live_entry(const V3Live& lv, uint32_t ctx, uint32_t n) {
    const V3Ctx& c = *lv.c;
    const uint32_t done_k = c.snap[256 * (lv.k % 3) + ctx];              // inserts into ctx before window k (all in G)
    const uint32_t ord = (n - done_k) & (kRing - 1);                     // n is pending iff it is insert number 1..pend of this window
    if (ord == 0 || ord > c.cnt[ctx] - done_k) return z3_ld_ring(c.ring + (size_t) ctx * kRing + n);
    for (int y = lv.upto; y >= lv.k * kV3W; y--) {          // newest first; a ring slot is written at most once per window
        const uint32_t m = c.ins[y & (kV3R - 1)];
        if (v3_kind(m) && (m & (kRing - 1)) == n && ((c.key[y & (kV3R - 1)] >> 13) & 0xffu) == ctx)
            return ring_make((uint32_t) y, c.key[y & (kV3R - 1)] >> 21, v3_suffix_of(c, y, m));
    }
    return z3_ld_ring(c.ring + (size_t) ctx * kRing + n);
}
ZL_HD uint32_t v3_live_head(const V3Live& lv, uint32_t key21, int before) {      // hash[ctx][slot] as seen just before position `before` inserts
    const V3Ctx& c = *lv.c;
    for (int y = before - 1; y >= lv.k * kV3W; y--) {
        const uint32_t m = c.ins[y & (kV3R - 1)];
        if (v3_kind(m) && ((c.key[y & (kV3R - 1)] ^ key21) & kKeyMask) == 0) return m & (kRing - 1);
    }
    return z3_ld_hash(c.hash + (size_t) (key21 >> 13) * kSlots + (key21 & (kSlots - 1)));
}
ZL_HD int v3_common_len_any(const V3Ctx& c, uint32_t x, uint32_t q) {         // x in the byte ring window, q anywhere earlier
    if (z3_in32(c.in, x) != z3_in32(c.in, q)) return 0;
    return v3_common_len_global(c.in, x, q, 4);
}
// literal replay of MatchLazy (lz.cpp:291-316) on the live view
ZL_HD bool v3_lazy_live(const V3Live& lv, int z, int best, int depth) {
    const V3Ctx& c = *lv.c;
    const uint32_t k = c.key[z & (kV3R - 1)];
    const uint32_t ctx = (k >> 13) & 0xffu;
    uint32_t node = v3_live_head(lv, k & kKeyMask, lv.upto + 1);
    if (node == (uint32_t) kNil) return false;
    const uint32_t at = (uint32_t) best - 3u;
    const uint32_t mine = z3_in32(c.in, (uint32_t) z + at);
    uint64_t e = v3_live_entry(lv, ctx, node);
    for (int hop = 0; hop < depth; hop++) {
        const uint32_t cand = ring_pos(e);
        if (z3_in32(c.in, cand + at) == mine) return true;
        const uint32_t nxt = ring_suffix(e);
        if (nxt == (uint32_t) kNil) break;
        const uint64_t e2 = v3_live_entry(lv, ctx, nxt);
        if (cand <= ring_pos(e2)) break;
        e = e2;
    }
    return false;
}
// literal replay of the candidate walk of MatchAndUpdate (lz.cpp:234-267); x's own insert is already pending
ZL_HD int v3_main_live(const V3Live& lv, int x, uint32_t node, uint32_t head, uint32_t chk, uint32_t ctx, int D, uint32_t* bestslot) {
    const V3Ctx& c = *lv.c;
    if (node == (uint32_t) kNil || node == head) return 0;
    int best = kMinLen - 1;
    uint64_t e = v3_live_entry(lv, ctx, node);
    for (int hop = 0; hop < D; hop++) {
        const uint32_t cand = ring_pos(e);
        if (ring_check(e) == chk) {
            const int l = v3_common_len_any(c, (uint32_t) x, cand);
            if (l > best) { best = l; *bestslot = node; if (best == kMaxLen) break; }
        }
        const uint32_t nxt = ring_suffix(e);
        if (nxt == (uint32_t) kNil) break;
        const uint64_t e2 = v3_live_entry(lv, ctx, nxt);
        if (cand <= ring_pos(e2)) break;
        node = nxt; e = e2;
    }
    return best;
}

// Leading nodes of a record whose ring slots have not been overwritten after `kc` further inserts into the context.
// A record node that HAS been overwritten ends the reference's walk right there when it is not the first one (the
// slot now holds a newer, i.e. larger, position: the "offset <= next offset" test of lz.cpp:264 fires), so a record
// cut at its first overwritten node is still exact; only an overwritten FIRST node needs the literal replay.
ZL_HD int v3_valid_nodes(const V3Table& t, int rel, int dmax, int nvis, uint32_t head_b, uint32_t kc) {
    int i = 0;
    while (i < nvis && v3_ring_dist(t.node[rel * dmax + i] >> 9, head_b) > kc) i++;
    return i;
}

// per-window constants of the resolver (hoisted out of the token loop)
struct V3Win {
    V3Table t; int k, base, D, L1, L2, tlevel;
    const uint32_t* snap_b;          // insert counters at base(k)
};
ZL_HD V3Win v3_window(const V3Ctx& c, int k, int level, int tlevel) {
    V3Win w; w.t = v3_table(c, k); w.k = k; w.base = v3_base(k);
    w.D = depth_main(level); w.L1 = depth_lazy1(level); w.L2 = depth_lazy2(level); w.tlevel = tlevel;
    w.snap_b = c.snap + 256 * ((k + 2) % 3);
    return w;
}

// Does any hazard flag of the decision at x actually hold?  (cn = insert counter of x's context INCLUDING x's insert)
ZL_HD bool v3_hazard(const V3Ctx& c, const V3Win& w, int x, const uint4& d, uint32_t cn) {
    const int rel = x - w.k * kV3W;
    const uint32_t fbest = (d.x >> 9) & 511u, ctx = d.y >> 24;
    const int nlazy = (fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow) ? (w.L2 > 0 ? 2 : 1) : 0;
    if ((d.y & kF_SELF) && v3_link_hazard(c, x, w.base, -1)) return true;
    if (nlazy >= 1 && (d.y & kF_L1) && v3_link_hazard(c, x + 1, w.base, x)) return true;
    if (nlazy >= 2 && (d.y & kF_L2) && v3_link_hazard(c, x + 2, w.base, x)) return true;
    if (d.y & kF_ST0) {
        const uint32_t h0 = w.t.hdr[rel], kc0 = cn - w.snap_b[ctx];
        if ((h0 >> 5) + 1u <= kc0 && v3_valid_nodes(w.t, rel, c.dmax, (int) (h0 & 31u), w.snap_b[ctx] & (kRing - 1), kc0) < (int) (h0 & 31u)) return true;
    }
    if (d.y & (kF_ST1 | kF_ST2)) {
        for (int q = 1; q <= nlazy; q++) {
            if (d.y & (q == 1 ? kF_ST1 : kF_ST2)) {
                const uint32_t cw = q == 1 ? (d.w >> 8) : (d.w & 0xffu);      // context of x+1 is in[x], of x+2 is in[x+1]
                const uint32_t kc = (cw == ctx ? cn : c.cnt[cw]) - w.snap_b[cw];
                const uint32_t hw = w.t.hdr[rel + q];
                if ((hw >> 5) + 1u <= kc && v3_valid_nodes(w.t, rel + q, c.dmax, (int) (hw & 31u), w.snap_b[cw] & (kRing - 1), kc) < (int) (hw & 31u)) return true;
            }
        }
    }
    return false;
}

// Token start x of window k whose frozen decision does not stand (a hazard holds, or the level differs): the full
// MatchAndUpdate (lz.cpp:211-289) with the in-window candidates.  Books the insert (cnt / ins / suf, kInsExplicit)
// and returns the match length (0 = none) and *midx.
ZL_HD int v3_probe_general(const V3Ctx& c, V3Run& r, const V3Win& w, int x, const uint4& d, uint32_t* midx) {
    const V3Table& t = w.t;
    const int k = w.k, rel = x - k * kV3W, base = w.base;
    const uint32_t* snap_b = w.snap_b;
    const uint32_t fhead = d.y & 0xffffu, ctx = d.y >> 24;
    const int D = w.D, L1 = w.L1, L2 = w.L2;
    const uint32_t cntc = c.cnt[ctx] + 1u, head = cntc & (kRing - 1);
    const uint32_t kc0 = cntc - snap_b[ctx];

    // ---------------- general path: in-window candidates (newest first), then the frozen record ----------------
    r.n_general++;
    const uint32_t kx = c.key[x & (kV3R - 1)];
    const uint32_t chk = kx >> 21;
    int best = kMinLen - 1, visited = 0;
    uint32_t bestslot = 0, suffix = fhead;
    bool done = false, have_suffix = false;
    {
        int y = x;
        while (true) {
            const uint32_t dl = c.link[y & (kV3R - 1)];
            if (!dl) break;
            y -= (int) dl;
            if (y < base) break;
            const uint32_t m = c.ins[y & (kV3R - 1)];
            if (!v3_kind(m)) continue;
            if (!have_suffix) {
                suffix = m & (kRing - 1); have_suffix = true;
                if (y >= k * kV3W) c.ins[y & (kV3R - 1)] = m | kInsSuperseded;   // same APPLY pass: x owns the slot head
            }
            if (visited < D && !done) {
                visited++;
                if ((c.key[y & (kV3R - 1)] >> 21) == chk) {
                    const int l = y == x - (int) c.link[x & (kV3R - 1)] ? (int) c.llen[x & (kV3R - 1)]      // precomputed by SPEC
                                                                        : v3_common_len_ring(c.rbw, (uint32_t) x, (uint32_t) y);
                    if (l > best) { best = l; bestslot = m & (kRing - 1); if (best == kMaxLen) done = true; }
                }
            } else break;
        }
    }
    // insert (lz.cpp:227-230): pending until APPLY
    c.cnt[ctx] = cntc;
    c.ins[x & (kV3R - 1)] = head | kInsExplicit | (kKindLit << 12);      // provisional kind: the live replay must see x as inserted
    c.suf[x & (kV3R - 1)] = (uint16_t) suffix;
    V3Live lv; lv.c = &c; lv.k = k; lv.upto = x;
    const uint32_t hdr = t.hdr[rel];
    int nvis = (int) (hdr & 31u);
    bool stale0 = false;
    if (nvis > 0 && (hdr >> 5) + 1u <= kc0) {
        const int nv = v3_valid_nodes(t, rel, c.dmax, nvis, snap_b[ctx] & (kRing - 1), kc0);
        stale0 = nv == 0;
        nvis = nv;
    }
    if (!done && visited < D && (nvis > 0 || stale0)) {
        if (stale0) {                                                    // the record's first slot has been overwritten: replay literally
            r.n_slow++;
            uint32_t bn = 0;
            best = v3_main_live(lv, x, suffix, head, chk, ctx, D, &bn);
            bestslot = bn;
        } else {
            const int take = nvis < D - visited ? nvis : D - visited;
            for (int i = 0; i < take; i++) {
                const uint32_t nd = t.node[rel * c.dmax + i];
                const int l = (int) (nd & 511u);
                if (l > best) { best = l; bestslot = nd >> 9; if (best == kMaxLen) break; }
            }
        }
    }
    if (best < kMinLen) return 0;
    if (best < kLazyBelow) {                                             // lz.cpp:270-281
        const uint32_t at = (uint32_t) best - 3u;
        for (int which = 1; which <= 2; which++) {
            const int depth = which == 1 ? L1 : L2;
            if (depth == 0) break;
            const int z = x + which, relz = rel + which;
            const uint32_t cz = v3_rb8(c.rbw, (uint32_t) z - 1);
            const uint32_t hz = t.hdr[relz];
            int nvz = (int) (hz & 31u);
            const uint32_t kcz = c.cnt[cz] - snap_b[cz];
            if (nvz > 0 && (hz >> 5) + 1u <= kcz) {                      // stale lazy record: cut it, or replay when its head is gone
                nvz = v3_valid_nodes(t, relz, c.dmax, nvz, snap_b[cz] & (kRing - 1), kcz);
                if (nvz == 0) {
                    r.n_slow++;
                    if (v3_lazy_live(lv, z, best, depth)) return 0;
                    continue;
                }
            }
            const uint32_t mine = v3_rb32(c.rbw, (uint32_t) z + at);
            int vis = 0;
            int y = z;
            while (vis < depth) {                                        // in-window same-key starts, newest first (x itself included)
                const uint32_t dl = c.link[y & (kV3R - 1)];
                if (!dl) break;
                y -= (int) dl;
                if (y < base) break;
                if (!v3_kind(c.ins[y & (kV3R - 1)])) continue;
                vis++;
                if (v3_rb32(c.rbw, (uint32_t) y + at) == mine) return 0;
            }
            int tk = nvz < depth - vis ? nvz : depth - vis;
            if (tk > c.lmax) tk = c.lmax;
            for (int i = 0; i < tk; i++)
                if (v3_eq_hit(t.eq + (size_t) (relz * c.lmax + i) * 5, at)) return 0;
        }
    }
    *midx = (head - bestslot) & (kRing - 1);
    return best;
}

// tokens of window k marked so far (the resolver needs a token index only when a sub-block closes)
ZL_HD int v3_count_marks(const V3Ctx& c, int lo, int hi) {
    int n = 0;
    for (int y = lo; y < hi; y++) n += v3_kind(c.ins[y & (kV3R - 1)]) != 0;
    return n;
}
ZL_HD void v3_close_subblock(const V3Ctx& c, V3Run& r, int nt) {
    if (r.j < kMaxSubPerBlock) {
        SubBlock sb; sb.tok_begin = (uint32_t) r.tok_begin; sb.tok_end = (uint32_t) nt; sb.enc_begin = (uint32_t) r.enc_begin;
        sb.enc_end = (uint32_t) r.ip; sb.rlen = (uint32_t) r.op; sb.level = (uint32_t) r.level; sb.olen = 0; sb.bits_lo = 0;
        c.sub[r.j] = sb;
    }
}
// Level of sub-block j.  The reference derives it from the PREVIOUS sub-block's Huffman size (src/libzling.cpp:261-266:
// olen / (consumed + 1) > 0.95 => level 0), which is not known during the parse (the literal ranks depend on MTF state
// carried across blocks).  plan[j] is the host's word: a level it has verified, or kPlanAuto = "predict": a sub-block
// that is almost all single-byte symbols (consumed <= 1.125 x symbols) will not compress.  The host verifies every
// level afterwards from the real sizes and re-parses from the first wrong one, so a wrong guess only costs time.
constexpr uint32_t kPlanAuto = 0xffu;
ZL_HD int v3_next_level(const V3Ctx& c, const V3Run& r, int j) {
    const uint32_t p = c.plan[j < kMaxSubPerBlock ? j : kMaxSubPerBlock - 1];
    if (p != kPlanAuto) return (int) p;
    if (j == 0) return c.base_level;
    const int consumed = r.ip - r.enc_begin;
    return consumed <= r.op + (r.op >> 3) ? 0 : c.base_level;
}
ZL_HD void v3_rollover(const V3Ctx& c, V3Run& r, int nt) {                // sub-block full (lz.cpp:153): close it, open the next
    v3_close_subblock(c, r, nt);
    const int next = v3_next_level(c, r, r.j + 1);
    r.j++;
    r.level = next;
    for (int i = 0; i < 256; i++) c.mru[i] = 0;                          // lz.cpp:147
    r.op = 0; r.tok_begin = nt; r.enc_begin = r.ip;
}

// Tokens starting in window k (EncodeImpl, lz.cpp:139-195), probe region only (x + 275 < ilen).  The walker does
// the minimum that is serial: one 16-byte decision load, the word-MRU push of the token that ended here, the
// context's insert counter, and a per-position mark; token words, literal lists and bucket writes are produced
// from the marks by all threads afterwards (v3_token_of / v3_apply_position).  nt0 = tokens emitted before window k.
// The common path is straight-line code (selects, no branches) so that its independent strands overlap; the
// decisions of both possible next positions are loaded speculatively.
ZL_HD void v3_resolve_window(const V3Ctx& c, V3Run& r, int k, int tlevel, int nt0) {
    if (r.tail) return;
    V3Win w = v3_window(c, k, r.level, tlevel);
    const uint4* dec = w.t.dec - k * kV3W;                               // dec[x] for x in the table
    const int lim = c.ilen - kGuard;                                     // probes happen at x < lim (lz.cpp:158)
    const int wend = (k + 1) * kV3W < lim ? (k + 1) * kV3W : lim;
    const int xmax = (k + 1) * kV3W + 1;                                 // last position the table holds
    int x = r.ip, op = r.op;
    uint32_t prev_lit = (uint32_t) r.prev_lit, skip_push = (uint32_t) r.skip_push;
    uint32_t force = r.level != tlevel ? kF_FORCE : 0u;
    if (x >= wend) { if (x >= lim) r.tail = 1; return; }
    uint4 d = dec[x];
    while (true) {
        uint32_t flen = d.x & 511u;
        // the decisions of both possible next positions are requested first: shared-memory loads are not moved across
        // the stores below by the compiler, and the chain x -> decision -> next x is the critical path
        const int xa0 = x + (flen ? (int) flen : 1);
        const uint4 dA0 = dec[xa0 < xmax ? xa0 : xmax], dB = dec[x + 2];
        const uint32_t c3 = d.z & 0xffu, pw = d.z >> 8, ctx = d.y >> 24;
        const uint32_t m = c.mru[c3];
        const uint32_t cn = c.cnt[ctx] + 1u;
        // word-MRU push of the token that ended at x: unconditional after a literal, else only if the top differs
        if (!skip_push && (prev_lit || (m & 0xffffu) != pw)) c.mru[c3] = pw | (m << 16);
        skip_push = 0;
        const bool special = ((d.y & kF_ANY) | force) != 0 || op + 1 >= kSubSymbols;
        if (!special && flen) {                                          // clean match (55 % of the tokens): the short way round
            c.cnt[ctx] = cn;
            c.ins[x & (kV3R - 1)] = (cn & (kRing - 1)) | (kKindMatch << 12);
            op += 2; prev_lit = 0; x = xa0; d = dA0;
            if (x >= wend) break;
            continue;
        }
        if (__builtin_expect(op + 1 >= kSubSymbols, 0)) {                // rare: the sub-block is full
            r.ip = x; r.op = op;
            v3_rollover(c, r, nt0 + v3_count_marks(c, k * kV3W, x));
            op = 0;
            force = r.level != tlevel ? kF_FORCE : 0u;
            w = v3_window(c, k, r.level, tlevel);
        }
        uint32_t mark = cn & (kRing - 1);
        if (__builtin_expect(((d.y & kF_ANY) | force) != 0, 0)) {        // flagged: most flags turn out not to hold
            r.n_flagged++;
#if defined(__CUDA_ARCH__)
            const long long t_sp = clock64();
#endif
#if defined(ZL_V3_FLAG_HIST) && !defined(__CUDA_ARCH__)
            g_flag_hist[((d.y >> 16) & 0x7fu) | (force ? 0x80u : 0u)]++;
#endif
            if (force || v3_hazard(c, w, x, d, cn)) {
                uint32_t midx = 0;
                flen = (uint32_t) v3_probe_general(c, r, w, x, d, &midx);
                mark |= kInsExplicit;
                if (flen) c.tw[x & (kV3R - 1)] = tok_match(flen, midx);
            } else {
                c.cnt[ctx] = cn;
            }
#if defined(__CUDA_ARCH__)
            r.cyc_special += clock64() - t_sp;
#endif
        } else {
            c.cnt[ctx] = cn;
        }
        // next position: x + flen after a match, else x + 2 after a word hit, x + 1 after a literal
 const int xa = x + (flen ? (int) flen : 1), xb = x + 2;
        const uint4 dA = xa == xa0 ? dA0 : dec[xa < xmax ? xa : xmax];   // the general path may have changed the length
        const uint32_t m1 = c.mru[ctx];                                  // lz.cpp:172-185 (x + 1 < ilen holds in the probe region)
        const bool w0 = (m1 & 0xffffu) == d.w, w1 = (m1 >> 16) == d.w;
        const bool word = !flen && (w0 || w1);
        const uint32_t kind = flen ? kKindMatch : (w0 ? kKindWord0 : (w1 ? kKindWord1 : kKindLit));
        c.ins[x & (kV3R - 1)] = mark | (kind << 12);
        prev_lit = kind == kKindLit;
        op += flen ? 2 : 1;
        x = word ? xb : xa;
        d = word ? dB : dA;
        if (x >= wend) break;
    }
    r.ip = x; r.op = op; r.prev_lit = (int) prev_lit; r.skip_push = (int) skip_push;
    if (x >= lim) r.tail = 1;
}

// The last 275 bytes of the block (no probe, no insert: lz.cpp:158) and blocks shorter than that: plain serial
// code, tokens written directly.  nt / nl = tokens / literals emitted so far.
ZL_HD void v3_resolve_tail(const V3Ctx& c, V3Run& r, int* nt_io, int* nl_io) {
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
        if (op + 1 >= kSubSymbols) { r.ip = ip; r.op = op; v3_rollover(c, r, nt); op = 0; }
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

#if defined(__CUDACC__)
struct V3Counters { unsigned long long tokens, general, slow, linkwalk, windows, cyc_resolve, cyc_spec, cyc_total, flagged, cyc_special; };

__device__ __forceinline__ void v3_bar_producers() { asm volatile("bar.sync 1, %0;" :: "n"(kV3Prod) : "memory"); }

// ---- the kernel: grid = blocks of the batch, kV3Threads threads; warp 0 = resolver, warps w with w % 4 != 0 = producers
__global__ void __launch_bounds__(kV3Threads, 1) zl_rolz_parse_v3_kernel(ParseArgs a, int dmax, int lmax, int base_level, int serialize, V3Counters* counters) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!a.active[b]) return;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_level, s_tlevel[2], s_nt, s_nl, s_wtok[kV3Threads / 32], s_wlit[kV3Threads / 32];
    const V3Layout L = v3_layout(dmax, lmax);
    V3Ctx c;
    v3_bind(c, smem_raw, L);
    c.in = a.in + (size_t) b * kBlockBytes; c.ilen = (int) a.ilen[b];
    c.ring = a.ring + (size_t) b * kRingStride; c.hash = a.hash + (size_t) b * kHashStride;
    c.tok = a.tok + (size_t) b * kTokStride; c.lit = a.lit + (size_t) b * kLitStride;
    c.sub = a.sub + (size_t) b * kMaxSubPerBlock; c.plan = a.plan + (size_t) b * kMaxSubPerBlock; c.base_level = base_level;
    const int ilen = c.ilen;
    const int pwarp = warp - 1 - (warp >> 2);                            // producer warp index over the warps with warp % 4 != 0
    const bool producer = (warp & 3) != 0 && pwarp < kV3Prod / 32;       // a warp runs on sub-partition warp % 4
    const int ptid = pwarp * 32 + lane;                                  // 0 .. kV3Prod-1 over the producer warps

    for (int i = tid; i < 256; i += kV3Threads) { c.cnt[i] = 0; c.mru[i] = 0; c.snap[i] = 0; c.snap[256 + i] = 0; c.snap[512 + i] = 0; c.pcnt[i] = 0; c.pcnt[256 + i] = 0; }
    for (int i = tid; i < kV3Buckets; i += kV3Threads) c.last[i] = 0;
    for (int i = tid; i < kV3R; i += kV3Threads) { c.ins[i] = 0; c.link[i] = 0; c.blink[i] = 0; c.key[i] = kKeyInvalid; }

    V3Run r;
    r.ip = 0; r.op = 0; r.j = 0; r.level = 0; r.tok_begin = 0; r.enc_begin = 0; r.prev_lit = 0; r.skip_push = 1; r.tail = 0;
    r.enc_begin = 0; r.level = v3_next_level(c, r, 0);
    r.n_general = 0; r.n_slow = 0; r.n_linkwalk = 0; r.n_flagged = 0; r.cyc_special = 0;
    long long cyc_res = 0, cyc_spec = 0;
    const long long t_begin = clock64();
    if (tid == 0) {
        int nt = 0;
        for (int first = 0; first < 2; first++) {                        // first two bytes raw, lz.cpp:150-151
            if (r.ip == first && r.ip < ilen) { c.tok[nt++] = tok_literal(c.in[r.ip], 0, true); r.op++; r.ip++; }
        }
        s_level = r.level; s_nt = nt; s_nl = 0;
    }
    __syncthreads();
    const int lim = ilen - kGuard;
    const int nwin = lim > 2 ? (lim + kV3W - 1) / kV3W : 0;              // windows that contain probe positions
    int staged_hi = -16;                                                 // bytes [.., staged_hi) are in the ring (first window also stages 16 lead bytes)

    for (int k = -1; k < nwin; k++) {
        const int tlevel_next = s_level;                                 // level assumed by the decisions of table k+1
        if (producer) {
            const int j = k + 1;
            if (j < nwin) {
                const long long t0 = clock64();
                const int hi = v3_stage_hi(j);
                for (int src = staged_hi + ptid * 16; src < hi; src += kV3Prod * 16) v3_stage16(c, src);
                if (ptid < 256) c.pcnt[256 * (j & 1) + ptid] = 0;
                v3_bar_producers();
                const int nlo = v3_new_lo(j), nhi = v3_new_hi(j);
                for (int x = nlo + ptid; x < nhi; x += kV3Prod) v3_key_position(c, x, j);
                v3_bar_producers();
                if (ptid < 32) {                                         // bucket chains, 32 positions per round in increasing order
                    for (int x0 = nlo; x0 < nhi; x0 += 32) {
                        const int x = x0 + ptid;
                        const uint32_t kx = x < nhi ? c.key[x & (kV3R - 1)] : kKeyInvalid;
                        const bool valid = !(kx & kKeyInvalid);
                        const uint32_t bk = valid ? v3_bucket_of(kx) : (uint32_t) kV3Buckets + ptid;
                        const uint32_t grp = __match_any_sync(0xffffffffu, bk);
                        const uint32_t lower = grp & ((1u << ptid) - 1u);
                        uint32_t dist = 0;
                        if (valid) {
                            if (lower) dist = (uint32_t) ptid - (31u - __clz(lower));
                            else { const uint32_t prev = c.last[bk]; if (prev != 0 && (uint32_t) x - (prev - 1) < (uint32_t) kV3R) dist = (uint32_t) x - (prev - 1); }
                        }
                        __syncwarp();
                        if (valid && (grp >> ptid) == 1u) c.last[bk] = (uint32_t) x + 1;
                        if (x < nhi) c.blink[x & (kV3R - 1)] = (uint16_t) dist;
                        __syncwarp();
                    }
                }
                v3_spec_position(c, j, ptid);
                v3_bar_producers();
                const int lb = v3_base(j);
                for (int x = nlo + ptid; x < nhi; x += kV3Prod) v3_link_position(c, x, lb);
                v3_bar_producers();
                if (ptid < kV3W) v3_decide_position(c, j, ptid, tlevel_next);
                if (ptid == 0) s_tlevel[j & 1] = tlevel_next;
                cyc_spec += clock64() - t0;
            }
        }
        if (serialize) __syncthreads();                                  // experiment: RESOLVE(k) after SPEC(k+1) instead of beside it
        if (tid == 0 && k >= 0) {
            const long long t0 = clock64();
            v3_resolve_window(c, r, k, s_tlevel[k & 1], s_nt);
            cyc_res += clock64() - t0;
        }
        __syncthreads();
        if (k >= 0) {                                                    // APPLY(k) + EMIT(k) + counter snapshot k+1
            const int lo = k * kV3W;
            const int y = lo + tid;                                      // kV3Threads >= kV3W: one position per thread
            uint32_t m = 0;
            if (tid < kV3W && y < ilen) { m = c.ins[y & (kV3R - 1)]; if (v3_kind(m)) v3_apply_position(c, y); }
            const uint32_t kind = v3_kind(m);
            const uint32_t bt = __ballot_sync(0xffffffffu, kind != 0), bl = __ballot_sync(0xffffffffu, kind == kKindLit);
            if (lane == 0) { s_wtok[warp] = __popc(bt); s_wlit[warp] = __popc(bl); }
            __syncthreads();
            int tb = s_nt, lb = s_nl, ttot = 0, ltot = 0;
            for (int w = 0; w < kV3Threads / 32; w++) {
                const int tw = s_wtok[w], lw = s_wlit[w];
                if (w < warp) { tb += tw; lb += lw; }
                ttot += tw; ltot += lw;
            }
            if (kind) {
                const int ti = tb + __popc(bt & ((1u << lane) - 1u));
                c.tok[ti] = v3_token_of(c, y, m);
                if (kind == kKindLit) c.lit[lb + __popc(bl & ((1u << lane) - 1u))] = (uint32_t) ti;
            }
            uint32_t* snap = c.snap + 256 * ((k + 1) % 3);
            for (int i = tid; i < 256; i += kV3Threads) snap[i] = c.cnt[i];
            __syncthreads();
            if (tid == 0) { s_level = r.level; s_nt += ttot; s_nl += ltot; }
        }
        staged_hi = v3_stage_hi(k + 1);
        __syncthreads();
    }
    if (tid == 0) {
        int nt = s_nt, nl = s_nl;
        v3_resolve_tail(c, r, &nt, &nl);
        if (ilen > 0) v3_close_subblock(c, r, nt);
        a.nsub[b] = ilen > 0 ? r.j + 1 : 0; a.ntok[b] = nt; a.nlit[b] = nl;
        if (counters) {
            atomicAdd(&counters->tokens, (unsigned long long) nt);
            atomicAdd(&counters->general, r.n_general);
            atomicAdd(&counters->slow, r.n_slow);
            atomicAdd(&counters->linkwalk, r.n_linkwalk);
            atomicAdd(&counters->flagged, r.n_flagged);
            atomicAdd(&counters->windows, (unsigned long long) nwin);
            atomicAdd(&counters->cyc_resolve, (unsigned long long) cyc_res);
            atomicAdd(&counters->cyc_total, (unsigned long long) (clock64() - t_begin));
            atomicAdd(&counters->cyc_special, (unsigned long long) r.cyc_special);
        }
    }
    if (tid == 33 && counters) atomicAdd(&counters->cyc_spec, (unsigned long long) cyc_spec);
}
#endif

}  // namespace zl
