// zl_rolz_parse_v2 — the ROLZ parse as "speculate wide, resolve narrow".
//
// The reference's parse (EncodeImpl + MatchAndUpdate + MatchLazy, src/libzling_lz.cpp:139-316) is one serial
// chain per 16 MiB block: whether position p is a token start, and what the dictionary holds when p is probed,
// depends on every earlier decision.  A GPU thread that follows that chain literally pays 4-6 dependent L2/HBM
// round trips per token (zl_rolz_parse_kernel, kept as the exact fallback).  Here one CTA owns a block and
// alternates two phases over windows of W input positions:
//
//   SPEC     every thread takes one or two positions x of the window and walks x's hash chain in the bucket state
//            G as it is at the window start ("frozen"): up to DMAX chain nodes with their match lengths against x,
//            plus, for the first LMAX nodes, a 132-bit byte-equality map (all a lazy probe can ask about).  The
//            records land in shared memory.  All the dependent global-memory latency of the parse lives here, and it
//            is overlapped across W positions.
//   RESOLVE  warp 0 walks the real token chain through the window using only shared memory: the record of the
//            position, plus a "mini dictionary" of the inserts made since the window start (those are invisible
//            to the frozen records).  The reference's candidate order is reproduced exactly:
//                in-window inserts with the same (context, hash slot), newest first, then the frozen chain.
//            A frozen record is only trusted if none of the ring slots it read has been overwritten since the
//            window start (every record carries the smallest "inserts until overwritten" distance of the slots it
//            read); otherwise that probe is redone exactly, on the live structure, like the v1 kernel does.
//            Warp 0 also applies every insert to G immediately (plain stores, nobody else reads G in this phase),
//            so G is always the reference's bucket state and the exact fallback is always available.
//
// Bit-exactness argument: DESIGN.md §4.2.
#pragma once
#include "zl_kernels.cuh"

namespace zl {

constexpr int kV2Threads = 512;
constexpr int kMdBuckets = 1024;
constexpr int kWinPad    = 8;        // bytes staged before the window start (mru looks back 3, ctx 1)
constexpr int kWinTail   = 2 + 272;  // lazy probes at +1/+2, compares reach 259 + word loads

struct V2Layout {       // offsets into dynamic shared memory, in bytes
    int W, dmax, lmax;
    int win, hdr, key, node, eq, dec, cnt, cntT, mru, mdhead, mdset, mdpos, mdkey, mdring, total;
};
__host__ __device__ inline V2Layout v2_layout(int W, int dmax, int lmax) {
    V2Layout L; L.W = W; L.dmax = dmax; L.lmax = lmax;
    int at = 0;
    auto take = [&](int bytes) { int o = at; at += (bytes + 15) & ~15; return o; };
    L.win    = take(W + kWinTail + 96);           // staged: up to 23 lead bytes + W + tail + 16, written in 16 B chunks
    L.hdr    = take(4 * (W + 2));
    L.key    = take(4 * (W + 2));
    L.node   = take(4 * (W + 2) * dmax);
    L.eq     = take(4 * (W + 2) * lmax * 5);
    L.dec    = take(8 * (W + 2));
    L.cnt    = take(4 * 256);
    L.cntT   = take(4 * 256);
    L.mru    = take(4 * 256);
    L.mdhead = take(2 * kMdBuckets);
    L.mdset  = take(4 * 2048);
    L.mdpos  = take(4 * W);
    L.mdkey  = take(4 * W);
    L.mdring = take(2 * W);
    L.total  = at;
    return L;
}

// ---- helpers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds32u(const uint8_t* sbase, uint32_t off) {       // unaligned LE load, shared memory
    const uint32_t* w = reinterpret_cast<const uint32_t*>(sbase) + (off >> 2);
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}
__device__ __forceinline__ uint32_t ring_dist(uint32_t slot, uint32_t head_at_horizon) {
    // number of further inserts into this context after which `slot` is overwritten (1..4096)
    return ((slot - head_at_horizon - 1u) & (kRing - 1)) + 1u;
}
__device__ __forceinline__ uint32_t md_bucket(uint32_t key21) { return (key21 * 2654435761u) >> 22; }   // 10 bits

// exact GetCommonLength (lz.cpp:66-89) by one thread, early exit; both operands in global memory
__device__ __forceinline__ int thread_common_len(const uint8_t* in, uint32_t x, uint32_t q) {
    if (ld32u(in, x) != ld32u(in, q)) return 0;
    for (int n = 4; n < 256; n += 4) {
        const uint32_t d = ld32u(in, x + n) ^ ld32u(in, q + n);
        if (d) return n + ((__ffs(d) - 1) >> 3);
    }
    const uint32_t d = ld32u(in, x + 256) ^ ld32u(in, q + 256);
    const int t = d ? ((__ffs(d) - 1) >> 3) : 4;
    return 256 + (t < 3 ? t : 3);
}
// bit j of the 160-bit map = (in[x+j] == in[q+j]), j < 132.  The candidate side is fetched with ten aligned
// 16-byte loads (its 132 bytes sit in at most 160 aligned bytes), the position's own side comes from the staged
// shared-memory window; WO = word offset of q inside its 16-byte line.
template <int WO>
__device__ __forceinline__ void eq_bits_body(const uint32_t (&qw)[41], uint32_t sh, const uint8_t* win, uint32_t x, uint32_t* out5) {
    #pragma unroll
    for (int g = 0; g < 5; g++) {
        uint32_t bits = 0;
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const int wj = g * 8 + j;
            if (wj < 33) {
                const uint32_t qword = __funnelshift_r(qw[wj + WO], qw[wj + WO + 1], sh);
                const uint32_t m = __vcmpeq4(qword, lds32u(win, x + 4 * wj));
                bits |= (((m & 0x01010101u) * 0x01020408u) >> 24) << (4 * j);
            }
        }
        out5[g] = bits;
    }
}
__device__ __forceinline__ void thread_eq_bits(const uint8_t* in, const uint8_t* win, uint32_t x, uint32_t q, uint32_t* out5) {
    uint32_t qw[41];
    const uint4* qa = reinterpret_cast<const uint4*>(in + (q & ~15u));
    #pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint4 v = __ldg(qa + i);
        qw[4 * i] = v.x; qw[4 * i + 1] = v.y; qw[4 * i + 2] = v.z; qw[4 * i + 3] = v.w;
    }
    qw[40] = 0;
    const uint32_t sh = (q & 3u) * 8u;
    switch ((q >> 2) & 3u) {
        case 0:  eq_bits_body<0>(qw, sh, win, x, out5); break;
        case 1:  eq_bits_body<1>(qw, sh, win, x, out5); break;
        case 2:  eq_bits_body<2>(qw, sh, win, x, out5); break;
        default: eq_bits_body<3>(qw, sh, win, x, out5); break;
    }
}
// common length from an equality map when the mismatch lies inside it, else finish with the word loop
__device__ __forceinline__ int len_from_eq_bits(const uint8_t* in, uint32_t x, uint32_t q, const uint32_t* eq5) {
    #pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t z = ~eq5[g];
        if (z) { const int n = g * 32 + __ffs(z) - 1; return n < kMinLen ? 0 : n; }
    }
    const uint32_t z = ~eq5[4] & 0xfu;
    if (z) return 128 + __ffs(z) - 1;
    for (int n = 132; n < 256; n += 4) {
        const uint32_t d = ld32u(in, x + n) ^ ld32u(in, q + n);
        if (d) return n + ((__ffs(d) - 1) >> 3);
    }
    const uint32_t d = ld32u(in, x + 256) ^ ld32u(in, q + 256);
    const int t = d ? ((__ffs(d) - 1) >> 3) : 4;
    return 256 + (t < 3 ? t : 3);
}

// ---- SPEC: record of one position against the frozen bucket state ------------------------------------------
struct V2Smem {
    uint8_t*  win; uint32_t* hdr; uint32_t* key; uint32_t* node; uint32_t* eq;
    uint32_t* cnt; uint32_t* cntT; uint32_t* mru;
    uint16_t* mdhead; uint32_t* mdpos; uint32_t* mdkey; uint16_t* mdring;
};

__device__ __forceinline__ void v2_spec_position(const uint8_t* in, const uint8_t* win, int ilen, uint32_t x, int rel, const uint64_t* ring,
                                                 const uint16_t* hash, const V2Smem& s, int dmax, int lmax) {
    if (x < 2 || (int) x + 273 >= ilen) { s.hdr[rel] = (uint32_t) (kRing - 1) << 5; s.key[rel] = 0; return; }
    const uint32_t c = win[x - 1];
    const uint32_t h = ctx_hash(lds32u(win, x));
    const uint32_t slot = h & (kSlots - 1), chk = (h >> 13) & 0xffu;
    s.key[rel] = slot | (chk << 13);
    const uint32_t headT = s.cntT[c] & (kRing - 1);
    const uint64_t* rc = ring + (size_t) c * kRing;
    uint32_t node = __ldcg(hash + (size_t) c * kSlots + slot);
    uint32_t nvis = 0, dmin = kRing;
    if (node != (uint32_t) kNil) {
        dmin = ring_dist(node, headT);
        uint64_t e = __ldcg(rc + node);
        for (int i = 0; i < dmax; i++) {
            const uint32_t q = ring_pos(e);
            int len = 0;
            if (i < lmax) {
                uint32_t* eq5 = s.eq + (size_t) (rel * lmax + i) * 5;
                thread_eq_bits(in, win, x, q, eq5);
                if (ring_check(e) == chk) len = len_from_eq_bits(in, x, q, eq5);
            } else if (ring_check(e) == chk) {
                len = thread_common_len(in, x, q);
            }
            s.node[rel * dmax + i] = (uint32_t) len | (node << 9);
            nvis = i + 1;
            const uint32_t nxt = ring_suffix(e);
            if (nxt == (uint32_t) kNil) break;
            dmin = min(dmin, ring_dist(nxt, headT));
            const uint64_t e2 = __ldcg(rc + nxt);
            if (q <= ring_pos(e2)) break;
            node = nxt; e = e2;
        }
    }
    s.hdr[rel] = nvis | ((dmin - 1u) << 5);
}

// ---- RESOLVE helpers (warp 0, warp-uniform control flow) -------------------------------------------------------
// GetCommonLength with both operands inside the shared-memory window
__device__ __forceinline__ int warp_common_len_smem(const uint8_t* win, uint32_t p, uint32_t q, int lane) {
    const uint32_t o = (uint32_t) lane * 8u;
    const uint32_t x0 = lds32u(win, p + o) ^ lds32u(win, q + o);
    const uint32_t x1 = lds32u(win, p + o + 4) ^ lds32u(win, q + o + 4);
    const int n = x0 ? ((__ffs(x0) - 1) >> 3) : (x1 ? 4 + ((__ffs(x1) - 1) >> 3) : 8);
    const uint32_t miss = __ballot_sync(0xffffffffu, n < 8);
    int len;
    if (miss == 0) {
        const uint32_t xt = lds32u(win, p + 256) ^ lds32u(win, q + 256);
        const int t = xt ? ((__ffs(xt) - 1) >> 3) : 4;
        len = 256 + (t < 3 ? t : 3);
    } else {
        const int first = __ffs(miss) - 1;
        len = first * 8 + __shfl_sync(0xffffffffu, n, first);
    }
    return len < kMinLen ? 0 : len;
}

// exact probes on the live structure (same walk as zl_rolz_parse_kernel, L2-coherent loads)
__device__ __forceinline__ bool lazy_probe_live(const uint8_t* in, const uint64_t* ring, const uint16_t* hash, uint32_t pos, int best, int depth) {
    const uint32_t c = in[pos - 1];
    const uint32_t slot = ctx_hash(ld32u(in, pos)) & (kSlots - 1);
    uint32_t node = __ldcg(hash + (size_t) c * kSlots + slot);
    if (node == (uint32_t) kNil) return false;
    const uint64_t* rc = ring + (size_t) c * kRing;
    const uint32_t at = (uint32_t) best - 3u;
    const uint32_t mine = ld32u(in, pos + at);
    uint64_t e = __ldcg(rc + node);
    for (int hop = 0; hop < depth; hop++) {
        const uint32_t cand = ring_pos(e);
        if (ld32u(in, cand + at) == mine) return true;
        const uint32_t nxt = ring_suffix(e);
        if (nxt == (uint32_t) kNil) break;
        const uint64_t e2 = __ldcg(rc + nxt);
        if (cand <= ring_pos(e2)) break;
        e = e2;
    }
    return false;
}
__device__ __forceinline__ int main_probe_live(const uint8_t* in, const uint64_t* rc, uint32_t pos, uint32_t node, uint32_t head, uint32_t chk,
                                               int depth, int lane, uint32_t* bestnode_out) {
    if (node == (uint32_t) kNil || node == head) return 0;
    int best = kMinLen - 1;
    uint32_t bestnode = 0;
    uint64_t e = __ldcg(rc + node);
    for (int hop = 0; hop < depth; hop++) {
        const uint32_t cand = ring_pos(e);
        if (ring_check(e) == chk) {
            const int l = warp_common_len(in, pos, cand, lane);
            if (l > best) { best = l; bestnode = node; if (best == kMaxLen) break; }
        }
        const uint32_t nxt = ring_suffix(e);
        if (nxt == (uint32_t) kNil) break;
        const uint64_t e2 = __ldcg(rc + nxt);
        if (cand <= ring_pos(e2)) break;
        node = nxt; e = e2;
    }
    *bestnode_out = bestnode;
    return best;
}

struct V2Counters { unsigned long long tokens, slow_main, slow_lazy, md_hits, windows, cyc_spec, cyc_resolve, general; };

// frozen decision of a position, packed for one 8-byte shared-memory load by the resolver:
//   bits 0..8 flen (match length the reference takes if nothing in the window interferes, 0 = no match)
//   bits 9..17 fbest (best frozen length before the lazy veto), 18..29 fslot (ring slot of that node),
//   bits 32..47 fhead (frozen head of the position's hash slot, kNil if empty)
__device__ __forceinline__ unsigned long long dec_pack(uint32_t flen, uint32_t fbest, uint32_t fslot, uint32_t fhead) {
    return (unsigned long long) (flen | (fbest << 9) | (fslot << 18)) | ((unsigned long long) fhead << 32);
}

// open-addressed set of the (context, slot) keys inserted since the window start: exact membership in ~1 probe
constexpr int kMdSet = 2048;
__device__ __forceinline__ uint32_t mdset_hash(uint32_t key21) { return (key21 * 2654435761u) >> 21; }   // 11 bits
__device__ __forceinline__ bool mdset_contains(const uint32_t* set, uint32_t key21) {
    uint32_t h = mdset_hash(key21);
    while (true) {
        const uint32_t v = set[h];
        if (v == 0) return false;
        if (v == key21 + 1u) return true;
        h = (h + 1u) & (kMdSet - 1);
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kV2Threads) zl_rolz_parse_v2_kernel(ParseArgs a, int W, int dmax, int lmax, int base_level, V2Counters* counters) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!a.active[b]) return;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const V2Layout L = v2_layout(W, dmax, lmax);
    V2Smem s;
    s.win = smem_raw + L.win; s.hdr = (uint32_t*) (smem_raw + L.hdr); s.key = (uint32_t*) (smem_raw + L.key);
    s.node = (uint32_t*) (smem_raw + L.node); s.eq = (uint32_t*) (smem_raw + L.eq);
    s.cnt = (uint32_t*) (smem_raw + L.cnt); s.cntT = (uint32_t*) (smem_raw + L.cntT); s.mru = (uint32_t*) (smem_raw + L.mru);
    s.mdhead = (uint16_t*) (smem_raw + L.mdhead); s.mdpos = (uint32_t*) (smem_raw + L.mdpos); s.mdkey = (uint32_t*) (smem_raw + L.mdkey);
    s.mdring = (uint16_t*) (smem_raw + L.mdring);
    unsigned long long* s_dec = (unsigned long long*) (smem_raw + L.dec);
    uint32_t* s_mdset = (uint32_t*) (smem_raw + L.mdset);
    __shared__ int s_ip, s_level;

    const uint8_t* in = a.in + (size_t) b * kBlockBytes;
    const int ilen = (int) a.ilen[b];
    uint64_t* ring = a.ring + (size_t) b * kRingStride;
    uint16_t* hash = a.hash + (size_t) b * kHashStride;
    uint32_t* tok = a.tok + (size_t) b * kTokStride;
    uint32_t* lit = a.lit + (size_t) b * kLitStride;
    SubBlock* sub = a.sub + (size_t) b * kMaxSubPerBlock;
    const uint8_t* plan = a.plan + (size_t) b * kMaxSubPerBlock;

    for (int i = tid; i < 256; i += kV2Threads) { s.cnt[i] = 0; s.mru[i] = 0; }

    // resolver state (meaningful in warp 0 only; identical in all its lanes)
    int ip = 0, nt = 0, nl = 0, j = 0, op = 0;
    int level = plan[0];
    int tok_begin = 0, enc_begin = 0;
    unsigned long long c_slow_main = 0, c_slow_lazy = 0, c_md = 0, c_win = 0, c_general = 0;
    long long cyc_spec = 0, cyc_res = 0;

    if (warp == 0) {                                                     // first two bytes raw, lz.cpp:150-151
        for (int first = 0; first < 2; first++) {
            if (ip == first && ip < ilen) {
                if (lane == 0) tok[nt] = tok_literal(in[ip], 0, true);
                nt++; op++; ip++;
            }
        }
        if (lane == 0) { s_ip = ip; s_level = level; }
    }
    __syncthreads();

    while (true) {
        const int wstart = s_ip;
        if (wstart >= ilen) break;
        const int wend = min(wstart + W, ilen);
        const int wlevel = s_level;                                      // level in force at the window start
        // ---------------- stage the input window: win[i] = in[wbase + i]
        const int wbase = (wstart - kWinPad) & ~15;                     // may be negative for the first window
        {
            const int nbytes = (wstart - wbase) + W + kWinTail + 16;
            for (int i = tid * 16; i < nbytes; i += kV2Threads * 16) {
                const int src = wbase + i;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (src >= 0 && src < ilen + 16) v = __ldg(reinterpret_cast<const uint4*>(in + src));
                *reinterpret_cast<uint4*>(s.win + i) = v;
            }
            for (int i = tid; i < 256; i += kV2Threads) s.cntT[i] = s.cnt[i];
            for (int i = tid; i < kMdBuckets; i += kV2Threads) s.mdhead[i] = 0;
            for (int i = tid; i < kMdSet; i += kV2Threads) s_mdset[i] = 0;
        }
        __syncthreads();
        // ---------------- SPEC 1: chain records against the frozen state
        const uint8_t* win = s.win - wbase;                              // win[p] valid for p in [wbase, wbase + staged)
        const long long t0 = clock64();
        for (int rel = tid; rel < W + 2; rel += kV2Threads) v2_spec_position(in, win, ilen, (uint32_t) (wstart + rel), rel, ring, hash, s, dmax, lmax);
        __syncthreads();
        // ---------------- SPEC 2: the decision the reference takes at each position if no insert of this window
        // interferes: best frozen node within the depth of the level in force, then the two lazy probes
        {
            const int D = depth_main(wlevel), L1 = depth_lazy1(wlevel), L2 = depth_lazy2(wlevel);
            for (int rel = tid; rel < W; rel += kV2Threads) {
                const uint32_t hdr = s.hdr[rel];
                const int nvis = (int) (hdr & 31u);
                uint32_t fbest = 0, fslot = 0, flen = 0;
                const uint32_t fhead = nvis > 0 ? (s.node[rel * dmax] >> 9) : (uint32_t) kNil;
                const int take = min(nvis, D);
                for (int i = 0; i < take; i++) {
                    const uint32_t nd = s.node[rel * dmax + i];
                    if ((nd & 511u) > fbest) { fbest = nd & 511u; fslot = nd >> 9; }
                }
                flen = fbest;
                if (fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow) {
                    const uint32_t at = fbest - 3u;
                    for (int which = 1; which <= 2 && flen; which++) {
                        const int depth = which == 1 ? L1 : L2;
                        const int nvx = (int) (s.hdr[rel + which] & 31u);
                        const int tk = min(min(nvx, depth), lmax);
                        for (int i = 0; i < tk; i++) {
                            const uint32_t* eqw = s.eq + (size_t) ((rel + which) * lmax + i) * 5;
                            if ((__funnelshift_r(eqw[at >> 5], eqw[(at >> 5) + 1], at & 31u) & 0xfu) == 0xfu) flen = 0;
                        }
                    }
                }
                s_dec[rel] = dec_pack(flen, fbest, fslot, fhead);
            }
        }
        __syncthreads();
        const long long t1 = clock64();
        cyc_spec += t1 - t0;
        // ---------------- RESOLVE (warp 0)
        if (warp == 0) {
            c_win++;
            int mdcount = 0;
            int D = depth_main(level), L1 = depth_lazy1(level), L2 = depth_lazy2(level);
            bool frozen_ok = true;                                       // SPEC 2 used the level in force at the window start
            while (ip < wend) {
                if (op + 1 >= kSubSymbols) {                             // sub-block full (lz.cpp:153): close it, open the next
                    if (lane == 0 && j < kMaxSubPerBlock) {
                        SubBlock sb; sb.tok_begin = tok_begin; sb.tok_end = nt; sb.enc_begin = enc_begin; sb.enc_end = ip;
                        sb.rlen = op; sb.level = level; sb.olen = 0; sb.bits_lo = 0;
                        sub[j] = sb;
                    }
                    j++;
                    level = plan[j < kMaxSubPerBlock ? j : kMaxSubPerBlock - 1];
                    D = depth_main(level); L1 = depth_lazy1(level); L2 = depth_lazy2(level);
                    frozen_ok = level == wlevel;
                    for (int i = lane; i < 256; i += 32) s.mru[i] = 0;   // lz.cpp:147
                    __syncwarp();
                    op = 0; tok_begin = nt; enc_begin = ip;
                }
                int mlen = 0;
                uint32_t midx = 0;
                if (ip + kGuard < ilen) {                                // lz.cpp:158 — probe + insert
                    const int rel = ip - wstart;
                    // hazard scan, three positions at once: lane 0 -> ip (main probe), lanes 1, 2 -> the lazy probes
                    const int li = lane < 2 ? lane : 2;
                    const uint32_t hdr_i = s.hdr[rel + li], key_i = s.key[rel + li];
                    const uint32_t c_i = win[ip + li - 1];
                    const uint32_t c = __shfl_sync(0xffffffffu, c_i, 0);
                    const uint32_t cnt_i = s.cnt[c_i], cntT_i = s.cntT[c_i];
                    const uint32_t mdk_i = (c_i << 13) | (key_i & (kSlots - 1));
                    const uint32_t mdk = __shfl_sync(0xffffffffu, mdk_i, 0);
                    const uint32_t kc_i = cnt_i - cntT_i + ((li == 0 || c_i == c) ? 1u : 0u);
                    const bool stale_i = (hdr_i & 31u) != 0 && (hdr_i >> 5) + 1u <= kc_i;
                    const bool hit_i = mdset_contains(s_mdset, mdk_i) || (li > 0 && mdk_i == mdk);
                    const unsigned long long dec = s_dec[rel];
                    const uint32_t fbest = ((uint32_t) dec >> 9) & 511u;
                    const int nlazy = (fbest >= (uint32_t) kMinLen && fbest < (uint32_t) kLazyBelow) ? (L2 > 0 ? 2 : 1) : 0;
                    const uint32_t haz = __ballot_sync(0xffffffffu, (stale_i || hit_i) && li <= nlazy) & 7u;
                    const bool fast = haz == 0 && frozen_ok;

                    const uint32_t key = __shfl_sync(0xffffffffu, key_i, 0);
                    const uint32_t slot = key & (kSlots - 1), chk = key >> 13;
                    const uint32_t cntc = __shfl_sync(0xffffffffu, cnt_i, 0) + 1u, head = cntc & (kRing - 1);
                    const uint32_t bucket = md_bucket(mdk);
                    const uint32_t bhead = s.mdhead[bucket];
                    int best = kMinLen - 1, visited = 0;
                    uint32_t bestslot = 0, suffix = (uint32_t) (dec >> 32);
                    bool done = false, veto = false;
                    if (fast) {
                        best = (int) fbest; bestslot = ((uint32_t) dec >> 18) & (kRing - 1);
                        veto = fbest != 0 && ((uint32_t) dec & 511u) == 0;
                    } else {
                        c_general++;
                        // in-window inserts with the same key, newest first
                        bool have_suffix = false;
                        for (uint32_t e = bhead; e != 0; ) {
                            const uint32_t kk = s.mdkey[e - 1];
                            if ((kk & 0x1fffffu) == mdk) {
                                const uint32_t pc = s.mdpos[e - 1], rs = s.mdring[e - 1];
                                if (!have_suffix) { suffix = rs; have_suffix = true; }
                                if (visited < D && !done) {
                                    visited++; c_md++;
                                    if ((pc >> 24) == chk) {
                                        const int l = warp_common_len_smem(win, (uint32_t) ip, pc & 0xffffffu, lane);
                                        if (l > best) { best = l; bestslot = rs; if (best == kMaxLen) done = true; }
                                    }
                                }
                            }
                            e = kk >> 21;
                        }
                    }
                    // insert (lz.cpp:227-230): live structure, mini dictionary, counters
                    __syncwarp();
                    if (lane == 0) {
                        s.cnt[c] = cntc;
                        ring[(size_t) c * kRing + head] = ring_make((uint32_t) ip, chk, suffix);
                        hash[(size_t) c * kSlots + slot] = (uint16_t) head;
                        s.mdpos[mdcount] = (uint32_t) ip | (chk << 24);
                        s.mdkey[mdcount] = mdk | (bhead << 21);
                        s.mdring[mdcount] = (uint16_t) head;
                        s.mdhead[bucket] = (uint16_t) (mdcount + 1);
                        uint32_t h = mdset_hash(mdk);
                        while (s_mdset[h] != 0 && s_mdset[h] != mdk + 1u) h = (h + 1u) & (kMdSet - 1);
                        s_mdset[h] = mdk + 1u;
                    }
                    mdcount++;
                    __syncwarp();
                    if (!fast) {
                        const uint32_t hdr = __shfl_sync(0xffffffffu, hdr_i, 0);
                        const int nvis = (int) (hdr & 31u);
                        const uint32_t dmin = (hdr >> 5) + 1u, kc = cntc - __shfl_sync(0xffffffffu, cntT_i, 0);
                        // frozen part of the chain
                        if (!done && visited < D && nvis > 0) {
                            if (dmin <= kc) {                            // a slot the record read has been overwritten: redo exactly
                                c_slow_main++;
                                uint32_t bn = 0;
                                best = main_probe_live(in, ring + (size_t) c * kRing, (uint32_t) ip, suffix, head, chk, D, lane, &bn);
                                bestslot = bn;
                            } else {
                                const int take = min(nvis, D - visited);
                                const uint32_t nd = lane < take ? s.node[rel * dmax + lane] : 0u;
                                const int l = (int) (nd & 511u);
                                const int mx = __reduce_max_sync(0xffffffffu, l);
                                if (mx > best) {
                                    const int who = __ffs(__ballot_sync(0xffffffffu, l == mx)) - 1;
                                    best = mx; bestslot = __shfl_sync(0xffffffffu, nd, who) >> 9;
                                }
                            }
                        }
                        if (best >= kMinLen && best < kLazyBelow) {      // lz.cpp:270-281
                            const uint32_t at = (uint32_t) best - 3u;
                            for (int which = 1; which <= 2 && !veto; which++) {
                                const int depth = which == 1 ? L1 : L2;
                                if (depth == 0) break;
                                const uint32_t x = (uint32_t) ip + which;
                                const int relx = rel + which;
                                const uint32_t cx = win[x - 1];
                                const uint32_t hdrx = s.hdr[relx], keyx = s.key[relx];
                                const uint32_t mdkx = (cx << 13) | (keyx & (kSlots - 1));
                                const uint32_t mine = lds32u(win, x + at);
                                int vis = 0;
                                for (uint32_t e = s.mdhead[md_bucket(mdkx)]; e != 0 && vis < depth && !veto; ) {
                                    const uint32_t kk = s.mdkey[e - 1];
                                    if ((kk & 0x1fffffu) == mdkx) {
                                        vis++;
                                        if (lds32u(win, (s.mdpos[e - 1] & 0xffffffu) + at) == mine) veto = true;
                                    }
                                    e = kk >> 21;
                                }
                                const int nvx = (int) (hdrx & 31u);
                                if (!veto && vis < depth && nvx > 0) {
                                    const uint32_t kcx = s.cnt[cx] - s.cntT[cx];
                                    if ((hdrx >> 5) + 1u <= kcx) {
                                        c_slow_lazy++;
                                        veto = lazy_probe_live(in, ring, hash, x, best, depth);
                                    } else {
                                        const int take = min(min(nvx, depth - vis), lmax);
                                        bool hit = false;
                                        if (lane < take) {
                                            const uint32_t* eqw = s.eq + (size_t) (relx * lmax + lane) * 5;
                                            const uint32_t lo = eqw[at >> 5], hi = eqw[(at >> 5) + 1];
                                            hit = (__funnelshift_r(lo, hi, at & 31u) & 0xfu) == 0xfu;
                                        }
                                        veto = __any_sync(0xffffffffu, hit);
                                    }
                                }
                            }
                        }
                    }
                    if (best >= kMinLen && !veto) { mlen = best; midx = (head - bestslot) & (kRing - 1); }
                }
                if (mlen) {
                    if (lane == 0) tok[nt] = tok_match((uint32_t) mlen, midx);
                    nt++; op += 2; ip += mlen;
                    const uint32_t c3 = win[ip - 3];
                    const uint32_t w = ((uint32_t) win[ip - 2] << 8) | win[ip - 1];
                    const uint32_t m = s.mru[c3];
                    __syncwarp();
                    if (lane == 0 && (m & 0xffffu) != w) s.mru[c3] = w | (m << 16);      // lz.cpp:163-166
                    __syncwarp();
                    continue;
                }
                if (ip + 1 < ilen) {                                     // lz.cpp:172-185
                    const uint32_t c1 = win[ip - 1];
                    const uint32_t w = ((uint32_t) win[ip] << 8) | win[ip + 1];
                    const uint32_t m = s.mru[c1];
                    if ((m & 0xffffu) == w) {
                        if (lane == 0) tok[nt] = tok_word(0);
                        nt++; op++; ip += 2;
                        continue;
                    }
                    if ((m >> 16) == w) {
                        if (lane == 0) tok[nt] = tok_word(1);
                        nt++; op++; ip += 2;
                        __syncwarp();
                        if (lane == 0) s.mru[c1] = w | (m << 16);
                        __syncwarp();
                        continue;
                    }
                }
                if (lane == 0) {                                         // literal, lz.cpp:188-191
                    tok[nt] = tok_literal(win[ip], win[ip - 1], false);
                    lit[nl] = (uint32_t) nt;
                }
                nt++; nl++; op++; ip++;
                {
                    const uint32_t c3 = win[ip - 3];
                    const uint32_t w = ((uint32_t) win[ip - 2] << 8) | win[ip - 1];
                    const uint32_t m = s.mru[c3];
                    __syncwarp();
                    if (lane == 0) s.mru[c3] = w | (m << 16);
                    __syncwarp();
                }
            }
            if (lane == 0) { s_ip = ip; s_level = level; }
            cyc_res += clock64() - t1;
        }
        __syncthreads();
    }
    if (warp == 0 && lane == 0) {
        if (ilen > 0 && j < kMaxSubPerBlock) {
            SubBlock sb; sb.tok_begin = tok_begin; sb.tok_end = nt; sb.enc_begin = enc_begin; sb.enc_end = ip;
            sb.rlen = op; sb.level = level; sb.olen = 0; sb.bits_lo = 0;
            sub[j] = sb;
        }
        a.nsub[b] = ilen > 0 ? j + 1 : 0; a.ntok[b] = nt; a.nlit[b] = nl;
        if (counters) {
            atomicAdd(&counters->tokens, (unsigned long long) nt);
            atomicAdd(&counters->slow_main, c_slow_main);
            atomicAdd(&counters->slow_lazy, c_slow_lazy);
            atomicAdd(&counters->md_hits, c_md);
            atomicAdd(&counters->windows, c_win);
            atomicAdd(&counters->cyc_spec, (unsigned long long) cyc_spec);
            atomicAdd(&counters->cyc_resolve, (unsigned long long) cyc_res);
            atomicAdd(&counters->general, c_general);
        }
    }
}

}  // namespace zl
