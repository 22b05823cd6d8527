/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the zling ROLZ+Huffman path in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (libzling_b200/) never links, loads or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the pin is the
 * unmodified reference itself compiled in this container (oracle/Makefile -> oracle/_ref/libzling_ref.so):
 * tests/test_oracle_vs_ref.py compares this file with it byte-for-byte (whole streams e0-e4, per-sub-block
 * symbol buffers, Huffman tables on 20k random tables), and tests/golden/ holds vectors produced by the
 * reference (tests/golden/make_golden.py).
 *
 * Each function names the reference lines it restates (paths relative to /root/reference).  This is a
 * re-derivation on flat arrays (no classes, no STL, no recursion), written from the behavioural spec in
 * SURVEY.md App. A, not a copy of the reference code.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "zling_tables.h"

enum {
    ZO_BLOCK_IN      = 16777216, /* src/libzling.cpp:70 */
    ZO_SUB_SYMS      = 262144,   /* src/libzling.cpp:71 */
    ZO_SUB_BYTES_MAX = 393216,   /* src/libzling.cpp:72 */
    ZO_RING          = 4096,     /* src/libzling_lz.h:44 */
    ZO_SLOTS         = 8192,     /* src/libzling_lz.h:45 */
    ZO_LAZY_BELOW    = 128,      /* src/libzling_lz.h:46 */
    ZO_MINLEN        = 4,        /* src/libzling_lz.h:47 */
    ZO_MAXLEN        = 259,      /* src/libzling_lz.h:48 */
    ZO_NSYM1         = 514,      /* src/libzling.cpp:63 */
    ZO_NSYM2         = 32,       /* src/libzling.cpp:64 */
    ZO_CAP1          = 15,       /* src/libzling.cpp:65 */
    ZO_CAP2          = 8,        /* src/libzling.cpp:66 */
    ZO_FAST          = 10,       /* src/libzling.cpp:67 */
    ZO_GUARD         = 275,      /* kMatchMaxLen + 16, src/libzling.cpp:68 */
    ZO_NIL           = 65535
};

/* ------------------------------------------------------------------ tables */
static uint8_t  g_mtfinit[256];
static uint8_t  g_mtfnext[256];
static uint8_t  g_idx_code[ZO_RING];
static uint16_t g_idx_base[ZO_NSYM2];
static uint8_t  g_idx_bits[ZO_NSYM2];
static int      g_tables_ready;

static int hexval(char c) { return c <= '9' ? c - '0' : c - 'a' + 10; }

/* src/tables/gen.py:10-19,31-56 (rules), values checked against the src/tables .inc files by tests/test_tables.py */
void zo_tables_init(void) {
    if (g_tables_ready) return;
    for (int i = 0; i < 256; i++) {
        g_mtfinit[i] = (uint8_t) (hexval(zo_mtfinit_hex[2 * i]) * 16 + hexval(zo_mtfinit_hex[2 * i + 1]));
        g_mtfnext[i] = (uint8_t) (i < 128 ? (i * 95) / 100 : (i * 55) / 100);
    }
    int filled = 0;
    for (int b = 0; b < ZO_NSYM2; b++) {
        int eb = b < 4 ? 0 : (b < 18 ? (b - 2) / 2 : 8);
        g_idx_bits[b] = (uint8_t) eb;
        g_idx_base[b] = (uint16_t) filled;
        for (int k = 0; k < (1 << eb); k++) g_idx_code[filled++] = (uint8_t) b;
    }
    g_tables_ready = 1;
}
const uint8_t*  zo_table_mtfinit(void)  { zo_tables_init(); return g_mtfinit; }
const uint8_t*  zo_table_mtfnext(void)  { zo_tables_init(); return g_mtfnext; }
const uint8_t*  zo_table_idx_code(void) { zo_tables_init(); return g_idx_code; }
const uint16_t* zo_table_idx_base(void) { zo_tables_init(); return g_idx_base; }
const uint8_t*  zo_table_idx_bits(void) { zo_tables_init(); return g_idx_bits; }

/* ------------------------------------------------------------------ Huffman */

/* Length-limited code lengths.  Restates ZlingMakeLengthTable, src/libzling_huffman.cpp:41-112; the
 * tie-breaking of the reference comes from libstdc++'s binary heap (bits/stl_heap.h, GCC 13.3:
 * __push_heap :135-148, __adjust_heap :224-249, __pop_heap :254-266, __make_heap :340-361) used through
 * std::priority_queue with a weight-only "greater" comparator (huffman.cpp:63-67,82-92); this is an
 * index-array emulation of exactly those sift orders. */
typedef struct {
    int32_t w[2 * ZO_NSYM1];    /* node weights: leaves first, internal nodes appended */
    int16_t kid[2 * ZO_NSYM1][2];
    int16_t heap[ZO_NSYM1];
    int     hn;
} zo_hufwork;

static void heap_sift_up(zo_hufwork* h, int hole, int top, int v) {      /* stl_heap.h __push_heap */
    int parent = (hole - 1) / 2;
    while (hole > top && h->w[h->heap[parent]] > h->w[v]) {
        h->heap[hole] = h->heap[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h->heap[hole] = (int16_t) v;
}
static void heap_adjust(zo_hufwork* h, int hole, int len, int v) {       /* stl_heap.h __adjust_heap */
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (h->w[h->heap[child]] > h->w[h->heap[child - 1]]) child--;
        h->heap[hole] = h->heap[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h->heap[hole] = h->heap[child - 1];
        hole = child - 1;
    }
    heap_sift_up(h, hole, top, v);
}
static int heap_pop(zo_hufwork* h) {                                     /* pop_heap + pop_back */
    int top = h->heap[0];
    if (h->hn > 1) {
        int last = h->heap[h->hn - 1];
        h->heap[h->hn - 1] = (int16_t) top;
        heap_adjust(h, 0, h->hn - 1, last);
    }
    h->hn--;
    return top;
}
static void heap_push(zo_hufwork* h, int v) {                            /* push_back + push_heap */
    h->heap[h->hn] = (int16_t) v;
    h->hn++;
    heap_sift_up(h, h->hn - 1, 0, v);
}

void zo_make_length_table(const uint32_t* freq, uint32_t* len, int nsym, int cap) {
    zo_hufwork* h = (zo_hufwork*) malloc(sizeof(zo_hufwork));
    int16_t leafsym[ZO_NSYM1];
    int16_t stack[2 * ZO_NSYM1], depth[2 * ZO_NSYM1];
    memset(len, 0, sizeof(uint32_t) * (size_t) nsym);                    /* huffman.cpp:42 */

    for (int shift = 0;; shift++) {                                      /* huffman.cpp:71,107-110 */
        int nleaf = 0;
        for (int s = 0; s < nsym; s++) {                                 /* huffman.cpp:74-78 */
            if (freq[s] > 0) {
                h->w[nleaf] = (int32_t) ((freq[s] + ((1u << shift) - 1)) >> shift);
                leafsym[nleaf] = (int16_t) s;
                h->heap[nleaf] = (int16_t) nleaf;
                nleaf++;
            }
        }
        if (nleaf == 0) break;                                           /* huffman.cpp:79-81 */
        h->hn = nleaf;
        if (nleaf >= 2) {                                                /* make_heap, huffman.cpp:82-85 */
            for (int parent = (nleaf - 2) / 2; parent >= 0; parent--) heap_adjust(h, parent, nleaf, h->heap[parent]);
        }
        int nnode = nleaf;
        while (h->hn > 1) {                                              /* huffman.cpp:88-92 */
            int a = heap_pop(h);
            int b = heap_pop(h);
            h->w[nnode] = h->w[a] + h->w[b];
            h->kid[nnode][0] = (int16_t) a;
            h->kid[nnode][1] = (int16_t) b;
            heap_push(h, nnode);
            nnode++;
        }
        /* depth walk, huffman.cpp:95-103 (iterative; leaf length = max(depth,1)) */
        int sp = 0, longest = 0;
        stack[sp] = h->heap[0]; depth[sp] = 0; sp++;
        while (sp > 0) {
            sp--;
            int node = stack[sp], d = depth[sp];
            if (node < nleaf) {
                int l = d > 1 ? d : 1;
                len[leafsym[node]] = (uint32_t) l;
                if (l > longest) longest = l;
            } else {
                stack[sp] = h->kid[node][0]; depth[sp] = (int16_t) (d + 1); sp++;
                stack[sp] = h->kid[node][1]; depth[sp] = (int16_t) (d + 1); sp++;
            }
        }
        if (longest <= cap) break;                                       /* huffman.cpp:107 */
    }
    free(h);
}

static uint16_t reverse16(uint16_t v) {
    v = (uint16_t) ((v >> 8) | (v << 8));
    v = (uint16_t) (((v & 0xf0f0) >> 4) | ((v & 0x0f0f) << 4));
    v = (uint16_t) (((v & 0xcccc) >> 2) | ((v & 0x3333) << 2));
    v = (uint16_t) (((v & 0xaaaa) >> 1) | ((v & 0x5555) << 1));
    return v;
}

/* canonical codes, stored bit-reversed so they can be emitted LSB-first.
 * Restates ZlingMakeEncodeTable, src/libzling_huffman.cpp:114-138 */
void zo_make_encode_table(const uint32_t* len, uint16_t* enc, int nsym, int cap) {
    unsigned next = 0;
    memset(enc, 0, sizeof(uint16_t) * (size_t) nsym);
    for (int l = 1; l <= cap; l++) {
        for (int s = 0; s < nsym; s++) {
            if (len[s] == (uint32_t) l) enc[s] = (uint16_t) next++;
        }
        next <<= 1;
    }
    for (int s = 0; s < nsym; s++) {
        unsigned sh = 16u - len[s];          /* len 0 -> shift 16 -> 0 (int promotion), huffman.cpp:135 */
        enc[s] = (uint16_t) ((unsigned) reverse16(enc[s]) >> sh);
    }
}

/* flat LUT: every index whose low len[s] bits equal enc[s] maps to s.
 * Restates ZlingMakeDecodeTable, src/libzling_huffman.cpp:140-153 */
void zo_make_decode_table(const uint32_t* len, const uint16_t* enc, uint16_t* dec, int nsym, int bits) {
    for (int i = 0; i < (1 << bits); i++) dec[i] = 0xffff;
    for (int s = 0; s < nsym; s++) {
        if (len[s] > 0 && len[s] <= (uint32_t) bits) {
            for (int i = enc[s]; i < (1 << bits); i += 1 << len[s]) dec[i] = (uint16_t) s;
        }
    }
}

/* ------------------------------------------------------------------ ROLZ encoder */
typedef struct {
    uint16_t suffix[ZO_RING];
    uint32_t entry[ZO_RING];     /* pos | check<<24 */
    uint16_t slot_head[ZO_SLOTS];
    uint16_t head;
} zo_bucket;                     /* src/libzling_lz.h:98-103 */

typedef struct {
    zo_bucket bucket[256];
    uint8_t   mtf_sym[256][256];     /* rank -> byte   (src/libzling_lz.h:55) */
    uint8_t   mtf_rank[256][256];    /* byte -> rank   (src/libzling_lz.h:56) */
    /* optional side channel for kernel tests: one record per token of the last zo_rolz_encode call */
    uint32_t* tok_pos;               /* token start position in block */
    uint8_t*  tok_raw;               /* raw byte for literal tokens, 0 otherwise */
    int       tok_cap, tok_n;
} zo_rolz;

static const int zo_depth[5][3] = { {2, 1, 0}, {4, 1, 0}, {6, 2, 0}, {8, 3, 1}, {16, 4, 2} };  /* lz.cpp:129-135 */

void zo_rolz_reset(zo_rolz* z) {                                         /* lz.cpp:197-209 */
    for (int c = 0; c < 256; c++) {
        memset(z->bucket[c].entry, 0, sizeof z->bucket[c].entry);
        memset(z->bucket[c].suffix, 0xff, sizeof z->bucket[c].suffix);
        memset(z->bucket[c].slot_head, 0xff, sizeof z->bucket[c].slot_head);
        z->bucket[c].head = 0;
    }
}
zo_rolz* zo_rolz_new(void) {                                             /* lz.cpp:106-111, lz.h:69-71 */
    zo_tables_init();
    zo_rolz* z = (zo_rolz*) calloc(1, sizeof(zo_rolz));
    if (!z) return NULL;
    for (int c = 0; c < 256; c++) {
        for (int r = 0; r < 256; r++) {
            z->mtf_sym[c][r] = g_mtfinit[r];
            z->mtf_rank[c][g_mtfinit[r]] = (uint8_t) r;
        }
    }
    zo_rolz_reset(z);
    return z;
}
void zo_rolz_free(zo_rolz* z) { free(z); }
void zo_rolz_trace(zo_rolz* z, uint32_t* pos, uint8_t* raw, int cap) { z->tok_pos = pos; z->tok_raw = raw; z->tok_cap = cap; z->tok_n = 0; }
int  zo_rolz_trace_count(const zo_rolz* z) { return z->tok_n; }
/* expose / load the stream-lifetime MTF state (64 KiB rank->byte tables), for carry tests */
void zo_rolz_get_mtf(const zo_rolz* z, uint8_t* out) { memcpy(out, z->mtf_sym, 65536); }

static uint32_t le32(const uint8_t* p) { return (uint32_t) p[0] | (uint32_t) p[1] << 8 | (uint32_t) p[2] << 16 | (uint32_t) p[3] << 24; }
static uint32_t ctx_hash(const uint8_t* p) { return le32(p) + p[2] * 137u + p[3] * 13337u; }   /* lz.cpp:55-57 */

/* lz.cpp:66-89: 0 unless the first four bytes agree, else exact common prefix capped at 259 */
static int common_len(const uint8_t* a, const uint8_t* b) {
    if (le32(a) != le32(b)) return 0;
    int n = 4;
    while (n < ZO_MAXLEN && a[n] == b[n]) n++;
    return n;
}

static int mtf_rank_and_update(zo_rolz* z, int ctx, int byte) {          /* lz.cpp:112-117 */
    uint8_t* sym = z->mtf_sym[ctx];
    uint8_t* rank = z->mtf_rank[ctx];
    int i = rank[byte], j = g_mtfnext[i];
    int other = sym[j];
    sym[i] = (uint8_t) other; sym[j] = (uint8_t) byte;
    rank[other] = (uint8_t) i; rank[byte] = (uint8_t) j;
    return i;
}

/* lz.cpp:291-316 */
static int lazy_probe(const zo_rolz* z, const uint8_t* buf, int pos, int bestlen, int depth) {
    const zo_bucket* b = &z->bucket[buf[pos - 1]];
    int node = b->slot_head[ctx_hash(buf + pos) % ZO_SLOTS];
    if (node == ZO_NIL) return 0;
    int at = bestlen - 3;
    for (int hop = 0; hop < depth; hop++) {
        uint32_t cand = b->entry[node] & 0xffffff;
        if (le32(buf + pos + at) == le32(buf + cand + at)) return 1;
        node = b->suffix[node];
        if (node == ZO_NIL || cand <= (b->entry[node] & 0xffffff)) break;
    }
    return 0;
}

/* lz.cpp:211-289.  Returns match length (0 = none) and *idx. */
static int probe_and_insert(zo_rolz* z, const uint8_t* buf, int pos, const int* d, int* idx) {
    uint32_t h = ctx_hash(buf + pos);
    uint32_t check = (h / ZO_SLOTS) % 256, slot = h % ZO_SLOTS;
    zo_bucket* b = &z->bucket[buf[pos - 1]];
    int node = b->slot_head[slot];

    b->head = (uint16_t) ((b->head + 1) & (ZO_RING - 1));               /* insert first, lz.cpp:227-230 */
    b->suffix[b->head] = b->slot_head[slot];
    b->entry[b->head] = (uint32_t) pos | check << 24;
    b->slot_head[slot] = b->head;

    if (node == ZO_NIL || node == b->head) return 0;                     /* lz.cpp:234-237 */

    int best = ZO_MINLEN - 1, bestnode = 0;
    for (int hop = 0; hop < d[0]; hop++) {                               /* lz.cpp:240-267 */
        uint32_t e = b->entry[node];
        uint32_t cand = e & 0xffffff;
        if ((e >> 24) == check && buf[pos + best] == buf[cand + best]) {
            int l = common_len(buf + pos, buf + cand);
            if (l > best) {
                best = l; bestnode = node;
                if (best == ZO_MAXLEN) break;
            }
        }
        node = b->suffix[node];
        if (node == ZO_NIL || cand <= (b->entry[node] & 0xffffff)) break;
    }
    if (best < ZO_MINLEN) return 0;
    if (best < ZO_LAZY_BELOW) {                                          /* lz.cpp:270-281 */
        if (d[1] > 0 && lazy_probe(z, buf, pos + 1, best, d[1])) return 0;
        if (d[2] > 0 && lazy_probe(z, buf, pos + 2, best, d[2])) return 0;
    }
    *idx = (b->head - bestnode) & (ZO_RING - 1);                         /* lz.cpp:283 */
    return best;
}

static void trace_token(zo_rolz* z, int pos, int raw) {
    if (z->tok_pos && z->tok_n < z->tok_cap) { z->tok_pos[z->tok_n] = (uint32_t) pos; z->tok_raw[z->tok_n] = (uint8_t) raw; }
    z->tok_n++;
}

/* One sub-block.  Restates ZlingRolzEncoder::Encode/EncodeImpl, src/libzling_lz.cpp:128-195.
 * Returns number of u16 symbols written (rlen), advances *encpos; -1 on bad level (lz.cpp:136). */
int zo_rolz_encode(zo_rolz* z, int level, const uint8_t* buf, uint16_t* sym, int ilen, int symcap, int* encpos) {
    if (level < 0 || level > 4) return -1;
    const int* d = zo_depth[level];
    uint16_t mru[256][2];
    int ip = *encpos, op = 0;
    memset(mru, 0, sizeof mru);                                          /* lz.cpp:147 */
    z->tok_n = 0;

    for (int k = 0; k < 2; k++) {                                        /* lz.cpp:150-151 */
        if (ip == k && op < symcap && ip < ilen) { trace_token(z, ip, buf[ip]); sym[op++] = buf[ip++]; }
    }
    while (op + 1 < symcap && ip < ilen) {                               /* lz.cpp:153 */
        if (ip + ZO_GUARD < ilen) {                                      /* lz.cpp:158 */
            int idx, l = probe_and_insert(z, buf, ip, d, &idx);
            if (l) {
                trace_token(z, ip, 0);
                sym[op++] = (uint16_t) (258 + l - ZO_MINLEN);
                sym[op++] = (uint16_t) idx;
                ip += l;
                int c = buf[ip - 3], w = buf[ip - 2] << 8 | buf[ip - 1];
                if (mru[c][0] != w) { mru[c][1] = mru[c][0]; mru[c][0] = (uint16_t) w; }   /* lz.cpp:163-166 */
                continue;
            }
        }
        if (ip + 1 < ilen) {                                             /* lz.cpp:172-185 */
            int c = buf[ip - 1], w = buf[ip] << 8 | buf[ip + 1];
            if (mru[c][0] == w) { trace_token(z, ip, 0); sym[op++] = 256; ip += 2; continue; }
            if (mru[c][1] == w) {
                trace_token(z, ip, 0);
                sym[op++] = 257; ip += 2;
                mru[c][1] = mru[c][0]; mru[c][0] = (uint16_t) w;
                continue;
            }
        }
        trace_token(z, ip, buf[ip]);
        sym[op++] = (uint16_t) mtf_rank_and_update(z, buf[ip - 1], buf[ip]);   /* lz.cpp:188 */
        ip++;
        int c = buf[ip - 3], w = buf[ip - 2] << 8 | buf[ip - 1];          /* lz.cpp:190-191 */
        mru[c][1] = mru[c][0]; mru[c][0] = (uint16_t) w;
    }
    *encpos = ip;
    return op;
}

/* ------------------------------------------------------------------ byte sink / bit writer */
typedef struct { uint8_t* p; size_t cap, n; } zo_sink;
static void sink_byte(zo_sink* s, int v) { if (s->n < s->cap) s->p[s->n] = (uint8_t) v; s->n++; }
static void sink_be32(zo_sink* s, uint32_t v) {                          /* src/libzling_utils.cpp:59-65 */
    sink_byte(s, v >> 24); sink_byte(s, v >> 16); sink_byte(s, v >> 8); sink_byte(s, v);
}

/* Entropy-code one sub-block's symbols into payload[] (273 table bytes + LSB-first bits).
 * Restates src/libzling.cpp:212-258.  Returns olen. */
int zo_huff_encode_subblock(const uint16_t* sym, int rlen, uint8_t* payload) {
    uint32_t f1[ZO_NSYM1] = {0}, f2[ZO_NSYM2] = {0}, l1[ZO_NSYM1], l2[ZO_NSYM2];
    uint16_t e1[ZO_NSYM1], e2[ZO_NSYM2];
    zo_tables_init();
    for (int i = 0; i < rlen; i++) {                                     /* libzling.cpp:219-224 */
        f1[sym[i]]++;
        if (sym[i] >= 258) { i++; f2[g_idx_code[sym[i]]]++; }
    }
    zo_make_length_table(f1, l1, ZO_NSYM1, ZO_CAP1);
    zo_make_length_table(f2, l2, ZO_NSYM2, ZO_CAP2);
    zo_make_encode_table(l1, e1, ZO_NSYM1, ZO_CAP1);
    zo_make_encode_table(l2, e2, ZO_NSYM2, ZO_CAP2);

    int op = 0;
    for (int i = 0; i < ZO_NSYM1; i += 2) payload[op++] = (uint8_t) (l1[i] << 4 | l1[i + 1]);   /* :232-234 */
    for (int i = 0; i < ZO_NSYM2; i += 2) payload[op++] = (uint8_t) (l2[i] << 4 | l2[i + 1]);   /* :235-237 */

    uint64_t acc = 0; int nbits = 0;                                     /* ZlingCodebuf, :80-105,240-257 */
    for (int i = 0; i < rlen; i++) {
        int s = sym[i];
        acc |= (uint64_t) e1[s] << nbits; nbits += (int) l1[s];
        if (s >= 258) {
            int idx = sym[++i], b = g_idx_code[idx];
            acc |= (uint64_t) e2[b] << nbits; nbits += (int) l2[b];
            acc |= (uint64_t) (idx - g_idx_base[b]) << nbits; nbits += g_idx_bits[b];
        }
        if (nbits >= 32) {
            for (int k = 0; k < 4; k++) { payload[op++] = (uint8_t) acc; acc >>= 8; }
            nbits -= 32;
        }
    }
    while (nbits > 0) { payload[op++] = (uint8_t) acc; acc >>= 8; nbits -= 8; }
    return op;
}

/* Whole stream.  Restates baidu::zling::Encode, src/libzling.cpp:174-291 (in-memory source/sink).
 * Returns compressed size (also when > cap: nothing is written past cap), -1 on bad level. */
/* `state` (optional, 65540 bytes: 256x256 rank->byte MTF tables + int32 LE current level) is what the reference
 * carries from block to block inside one Encode() call (m_mtf, src/libzling_lz.h:105; current_level,
 * src/libzling.cpp:185): loaded before the first block and stored back after the last, so that a stream can be
 * encoded range by range (tests of the multi-GPU carry hand-off). */
long long zo_encode_range(const uint8_t* in, size_t n, int level, uint8_t* out, size_t cap, uint8_t* state);
long long zo_encode(const uint8_t* in, size_t n, int level, uint8_t* out, size_t cap) {
    return zo_encode_range(in, n, level, out, cap, NULL);
}
long long zo_encode_range(const uint8_t* in, size_t n, int level, uint8_t* out, size_t cap, uint8_t* state) {
    if (level < 0 || level > 4) return -1;
    zo_rolz* z = zo_rolz_new();
    uint16_t* sym = (uint16_t*) malloc(sizeof(uint16_t) * (ZO_SUB_SYMS + ZO_GUARD));
    uint8_t* payload = (uint8_t*) malloc(ZO_SUB_BYTES_MAX + ZO_GUARD + 8);
    zo_sink sink = { out, cap, 0 };
    int cur_level = level;                                               /* libzling.cpp:185 — outlives blocks */
    if (state) {
        memcpy(z->mtf_sym, state, 65536);
        for (int c = 0; c < 256; c++) for (int r = 0; r < 256; r++) z->mtf_rank[c][z->mtf_sym[c][r]] = (uint8_t) r;
        int32_t lv; memcpy(&lv, state + 65536, 4);
        cur_level = lv;
    }

    for (size_t off = 0; off < n; off += ZO_BLOCK_IN) {
        int ilen = (int) (n - off < ZO_BLOCK_IN ? n - off : ZO_BLOCK_IN);
        const uint8_t* blk = in + off;
        int encpos = 0;
        zo_rolz_reset(z);                                                /* buckets only; MTF carried */
        while (encpos < ilen) {
            int before = encpos;
            sink_byte(&sink, 1);
            int rlen = zo_rolz_encode(z, cur_level, blk, sym, ilen, ZO_SUB_SYMS, &encpos);
            int olen = zo_huff_encode_subblock(sym, rlen, payload);
            /* libzling.cpp:261: 1.0*olen/(consumed+1) > 0.95  <=>  20*olen > 19*(consumed+1) in integers */
            cur_level = ((long long) olen * 20 > (long long) (encpos - before + 1) * 19) ? 0 : level;
            sink_be32(&sink, (uint32_t) encpos);
            sink_be32(&sink, (uint32_t) rlen);
            sink_be32(&sink, (uint32_t) olen);
            for (int i = 0; i < olen; i++) sink_byte(&sink, payload[i]);
        }
        sink_byte(&sink, 0);
    }
    if (state) {
        memcpy(state, z->mtf_sym, 65536);
        int32_t lv = cur_level; memcpy(state + 65536, &lv, 4);
    }
    free(payload); free(sym); zo_rolz_free(z);
    return (long long) sink.n;
}

/* ------------------------------------------------------------------ decoder */
typedef struct {
    uint32_t ring[256][ZO_RING];
    uint16_t head[256];
    uint8_t  mtf_sym[256][256];
} zo_unrolz;                                                             /* src/libzling_lz.h:131-137 */

static void unrolz_reset(zo_unrolz* u) { memset(u->ring, 0, sizeof u->ring); memset(u->head, 0, sizeof u->head); }   /* lz.cpp:378-386 */

static uint32_t unrolz_insert_lookup(zo_unrolz* u, const uint8_t* buf, int pos, int idx) {    /* lz.cpp:388-399 */
    int c = buf[pos - 1];
    u->head[c] = (uint16_t) ((u->head[c] + 1) & (ZO_RING - 1));
    u->ring[c][u->head[c]] = (uint32_t) pos;
    return u->ring[c][(u->head[c] - idx) & (ZO_RING - 1)];
}
static int unmtf(zo_unrolz* u, int ctx, int rank) {                      /* lz.cpp:122-126 */
    uint8_t* sym = u->mtf_sym[ctx];
    int byte = sym[rank], j = g_mtfnext[rank];
    sym[rank] = sym[j]; sym[j] = (uint8_t) byte;
    return byte;
}

/* One sub-block of symbols -> bytes.  Restates ZlingRolzDecoder::Decode, src/libzling_lz.cpp:318-376.
 * `room` = writable bytes after buf (the reference's 4-byte strided copy may overshoot by <=3 bytes into its
 * 275-byte sentinel, lz.cpp:91-104; a byte-serial copy yields the same bytes in [0,encpos)). */
static int unrolz_subblock(zo_unrolz* u, const uint16_t* sym, uint8_t* buf, int rlen, int encpos, int* decpos) {
    uint16_t mru[256][2];
    int op = *decpos, ip = 0;
    memset(mru, 0, sizeof mru);
    for (int k = 0; k < 2; k++) if (op == k && ip < rlen) buf[op++] = (uint8_t) sym[ip++];
    while (ip < rlen) {
        int s = sym[ip++];
        if (s < 256) {
            buf[op] = (uint8_t) unmtf(u, buf[op - 1], s);
            unrolz_insert_lookup(u, buf, op, 0); op++;
            int c = buf[op - 3], w = buf[op - 2] << 8 | buf[op - 1];
            mru[c][1] = mru[c][0]; mru[c][0] = (uint16_t) w;
        } else if (s == 256 || s == 257) {
            int w = mru[buf[op - 1]][s - 256];
            buf[op] = (uint8_t) (w >> 8); unrolz_insert_lookup(u, buf, op, 0); op++;
            buf[op] = (uint8_t) w; op++;
            if (s == 257) { int c = buf[op - 3]; mru[c][1] = mru[c][0]; mru[c][0] = (uint16_t) w; }
        } else {
            int l = s - 258 + ZO_MINLEN;
            if (ip >= rlen) return -1;                                    /* idx slot missing */
            int idx = sym[ip++];
            uint32_t from = unrolz_insert_lookup(u, buf, op, idx);
            /* NB: the reference does not bound-check l against encpos before copying (it has a 275-byte
             * sentinel); refuse instead of writing past the block. idx==0 would self-reference (dst==src):
             * the reference spins forever there (lz.cpp:92-96); report failure. */
            if (op + l > encpos || from >= (uint32_t) op) return -1;
            for (int k = 0; k < l; k++) buf[op + k] = buf[from + k];
            op += l;
            int c = buf[op - 3], w = buf[op - 2] << 8 | buf[op - 1];
            if (mru[c][0] != w) { mru[c][1] = mru[c][0]; mru[c][0] = (uint16_t) w; }
        }
        if (op > encpos) return -1;
    }
    if (op != encpos) return -1;
    *decpos = op;
    return 0;
}

/* Whole stream.  Restates baidu::zling::Decode, src/libzling.cpp:293-427.
 * Returns decoded size; -3 malformed (where the reference throws), counts past cap without writing. */
long long zo_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap) {
    zo_tables_init();
    zo_unrolz* u = (zo_unrolz*) malloc(sizeof(zo_unrolz));
    uint8_t* blk = (uint8_t*) malloc(ZO_BLOCK_IN + ZO_GUARD);
    uint16_t* sym = (uint16_t*) malloc(sizeof(uint16_t) * (ZO_SUB_SYMS + ZO_GUARD));
    uint16_t* d1 = (uint16_t*) malloc(sizeof(uint16_t) << ZO_CAP1);
    uint16_t d1f[1 << ZO_FAST], d2[1 << ZO_CAP2];
    size_t at = 0, total = 0;
    long long rc = 0;
    for (int c = 0; c < 256; c++) memcpy(u->mtf_sym[c], g_mtfinit, 256);

    while (at < n && rc == 0) {                                          /* per block */
        int decpos = 0;
        unrolz_reset(u);
        while (at < n) {
            int flag = in[at++];
            if (flag != 0 && flag != 1) { rc = -3; break; }              /* libzling.cpp:315-317 */
            if (flag == 0) break;
            if (at + 12 > n) { rc = -3; break; }
            uint32_t encpos = (uint32_t) in[at] << 24 | in[at + 1] << 16 | in[at + 2] << 8 | in[at + 3];
            uint32_t rlen   = (uint32_t) in[at + 4] << 24 | in[at + 5] << 16 | in[at + 6] << 8 | in[at + 7];
            uint32_t olen   = (uint32_t) in[at + 8] << 24 | in[at + 9] << 16 | in[at + 10] << 8 | in[at + 11];
            at += 12;
            if (rlen > ZO_SUB_SYMS || olen > ZO_SUB_BYTES_MAX) { rc = -3; break; }          /* :326-328 */
            if (at + olen > n || olen < 273 || encpos > ZO_BLOCK_IN) { rc = -3; break; }
            const uint8_t* pl = in + at; at += olen;

            uint32_t l1[ZO_NSYM1], l2[ZO_NSYM2];
            uint16_t e1[ZO_NSYM1], e2[ZO_NSYM2];
            for (int i = 0; i < ZO_NSYM1; i += 2) { l1[i] = pl[i / 2] >> 4; l1[i + 1] = pl[i / 2] & 15; }      /* :347-351 */
            for (int i = 0; i < ZO_NSYM2; i += 2) { l2[i] = pl[257 + i / 2] >> 4; l2[i + 1] = pl[257 + i / 2] & 15; }
            zo_make_encode_table(l1, e1, ZO_NSYM1, ZO_CAP1);
            zo_make_encode_table(l2, e2, ZO_NSYM2, ZO_CAP2);
            zo_make_decode_table(l1, e1, d1, ZO_NSYM1, ZO_CAP1);
            zo_make_decode_table(l1, e1, d1f, ZO_NSYM1, ZO_FAST);
            zo_make_decode_table(l2, e2, d2, ZO_NSYM2, ZO_CAP2);

            uint64_t acc = 0; int nbits = 0; uint32_t rp = 273;
            for (uint32_t i = 0; i < rlen && rc == 0; i++) {             /* :368-402 */
                while (nbits < 32) { acc |= (uint64_t) (rp < olen ? pl[rp] : 0) << nbits; rp++; nbits += 8; }
                int s = d1f[acc & ((1u << ZO_FAST) - 1)];
                if (s == 0xffff) s = d1[acc & ((1u << ZO_CAP1) - 1)];
                if (s >= ZO_NSYM1) { rc = -3; break; }
                acc >>= l1[s]; nbits -= (int) l1[s];
                sym[i] = (uint16_t) s;
                if (s >= 258) {
                    int b = d2[acc & 0xff];
                    if (b >= ZO_NSYM2) { rc = -3; break; }
                    acc >>= l2[b]; nbits -= (int) l2[b];
                    uint32_t extra = (uint32_t) (acc & ((1u << g_idx_bits[b]) - 1));
                    acc >>= g_idx_bits[b]; nbits -= g_idx_bits[b];
                    uint32_t idx = g_idx_base[b] + extra;
                    if (idx >= ZO_RING) { rc = -3; break; }
                    if (i + 1 >= rlen) { rc = -3; break; }                /* idx slot missing */
                    sym[++i] = (uint16_t) idx;
                }
            }
            if (rc) break;
            if (unrolz_subblock(u, sym, blk, (int) rlen, (int) encpos, &decpos) != 0) { rc = -3; break; }   /* :406-408 */
        }
        if (rc) break;
        for (int i = 0; i < decpos; i++) { if (total < cap) out[total] = blk[i]; total++; }
    }
    free(d1); free(sym); free(blk); free(u);
    return rc ? rc : (long long) total;
}
