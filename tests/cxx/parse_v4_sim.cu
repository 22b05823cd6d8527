// Host replay of zl_rolz_parse_v4's phases (libzling_b200/csrc/zl_parse_v4.cuh) against the oracle's tokeniser.
// TEST INFRASTRUCTURE: every algorithmic function of the v4 kernel is scalar ZL_HD code; this program runs the same
// functions in the kernel's phase order with loops instead of threads (the ordered passes — link building, ranks,
// orbit, prefix sums — in their obvious serial form), so the speculate / iterate / finalize logic is checked
// bit-for-bit on the CPU before any GPU time is spent.  What it cannot check is the kernel's synchronisation and its
// parallel forms of the ordered passes — that is what the -m gpu parity tests are for.
//   parse_v4_sim <file> <level> [plan-digits|auto]
// exit code 0 = token streams identical.  Prints iteration statistics (rounds per window, general-path evaluations).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../libzling_b200/csrc/zl_parse_v4.cuh"

extern "C" {
typedef struct zo_rolz zo_rolz;
zo_rolz* zo_rolz_new(void);
void zo_rolz_free(zo_rolz*);
void zo_rolz_trace(zo_rolz*, uint32_t* pos, uint8_t* raw, int cap);
int zo_rolz_trace_count(const zo_rolz*);
int zo_rolz_encode(zo_rolz*, int level, const uint8_t* buf, uint16_t* sym, int ilen, int symcap, int* encpos);
}

using namespace zl;

struct Tok { uint32_t pos, sym, aux, byte; };

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: parse_v4_sim file level [plan-digits|auto]\n"); return 2; }
    const int level = atoi(argv[2]);
    uint8_t plan[kMaxSubPerBlock];
    memset(plan, level, sizeof plan);
    if (argc > 3 && !strcmp(argv[3], "auto")) memset(plan, 0xff, sizeof plan);
    else if (argc > 3) for (size_t i = 0; i < strlen(argv[3]) && i < (size_t) kMaxSubPerBlock; i++) plan[i] = argv[3][i] == 'a' ? 0xff : (uint8_t) (argv[3][i] - '0');
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> data;
    { uint8_t buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n); }
    fclose(f);
    if (data.size() > (size_t) kBlockBytes) data.resize(kBlockBytes);
    const int ilen = (int) data.size();

    uint8_t* in = (uint8_t*) aligned_alloc(64, (size_t) ilen + 512);
    memset(in, 0, (size_t) ilen + 512);
    memcpy(in, data.data(), ilen);
    std::vector<uint64_t> ring((size_t) 256 * kRing, kRingEmpty);
    std::vector<uint16_t> hash((size_t) 256 * kSlots, 0xffff);
    std::vector<uint32_t> tok((size_t) ilen + 16), lit((size_t) ilen + 16);
    std::vector<SubBlock> sub(kMaxSubPerBlock);
    const int dmax = depth_main(level), lmax = depth_lazy1(level);
    const V4Layout L = v4_layout(dmax, lmax);
    uint8_t* smem = (uint8_t*) aligned_alloc(64, (size_t) L.total + 64);
    memset(smem, 0, L.total);
    std::vector<uint32_t> last(kV4Buckets);
    V4Ctx c;
    v4_bind(c, smem, L);
    c.last = last.data();
    c.in = in; c.ilen = ilen; c.ring = ring.data(); c.hash = hash.data(); c.tok = tok.data(); c.lit = lit.data();
    c.sub = sub.data(); c.plan = plan; c.base_level = level;

    V4Run r; memset(&r, 0, sizeof r);
    r.level = v4_next_level(c, 0, 0, 0); r.skip_push = 1;
    int nt = 0, nl = 0;
    for (int first = 0; first < 2; first++)
        if (r.ip == first && r.ip < ilen) { c.tok[nt++] = tok_literal(in[r.ip], 0, true); r.op++; r.ip++; }
    const int lim = ilen - kGuard;
    const int nwin = lim > 2 ? (lim + kV4W - 1) / kV4W : 0;
    int staged_hi = -16;
    unsigned long long st_rounds = 0, st_windows = 0, st_general = 0, st_decides = 0, st_maxrounds = 0, st_hist[9] = {0};
    std::vector<uint32_t> cntw(256);
    // the kernel re-derives the decisions of the MARKED positions only (unmarked ones keep their frozen decision until the
    // orbit reaches them); ZL_V4_ALL=1 evaluates every position each round (fewer rounds, 4x the work)
    const bool marked_only = !(getenv("ZL_V4_ALL") && atoi(getenv("ZL_V4_ALL")) != 0);

    for (int k = 0; k < nwin; k++) {
        const int lo = k * kV4W;
        const int wend = lo + kV4W < lim ? lo + kV4W : lim;
        const int hi = v4_stage_hi(k);
        for (int src = staged_hi; src < hi; src += 16) v4_stage16(c, src);
        staged_hi = hi;
        if (r.ip >= wend) continue;
        // ---- SPEC
        for (int rel = 0; rel < kV4N; rel++) c.key[rel] = v4_key_of(c, lo + rel);
        memset(c.occ, 0, sizeof(uint32_t) * 256 * kV4Words);
        for (int i = 0; i < kV4N + 2; i++) {                               // bit i of occ[b]: in[lo + i - 3] == b
            const int p = lo + i - 3;
            if (p < 0) continue;
            const uint32_t b = v4_rb8(c.rbw, (uint32_t) p);
            c.occ[b * kV4Words + (i >> 5)] |= 1u << (i & 31);
        }
        for (int cq = 0; cq < 256; cq++) {
            uint32_t n = 0, ow = 0;
            for (int wq = 0; wq < kV4Words; wq++) { n += (uint32_t) z4_popc(c.occ[cq * kV4Words + wq]); if (wq < 32 && c.occ[cq * kV4Words + wq]) ow |= 1u << wq; }
            c.pcnt[cq] = n; c.occw[cq] = ow;
        }
        v4_bucket_pass_serial(c);
        for (int rel = 0; rel < kV4N; rel++) v4_link_position(c, rel);
        for (int rel = 0; rel < kV4N; rel++) v4_spec_position(c, lo, rel);
        for (int rel = 0; rel < kV4W; rel++) v4_frozen_position(c, lo, rel, r.level);
        V4Win w; w.lo = lo; w.wend = wend; w.entry = r.ip; w.level = r.level; w.rpos = -1; w.level2 = r.level;
        w.skip_push = r.skip_push; w.prev_lit = r.prev_lit;
        for (int rel = 0; rel < kV4N; rel++) { const uint32_t fl = rel < kV4W ? (c.fdec[rel] & 511u) : 0u; c.dec[rel] = fl ? (fl | (kV4Match << 9)) : (kV4Lit << 9); }
        // ---- ROUNDS
        int exitpos = 0, rounds = 0, op_at_rpos = 0, last_rel = -1;
        while (true) {
            rounds++;
            memset(c.mark, 0, kV4N + 2); memset(c.plit, 0, kV4N + 2); memset(c.mbits, 0, 4 * kV4Words);
            { int x = w.entry, pl = w.prev_lit;
              while (x < wend) { const int rel = x - lo; c.mark[rel] = 1; c.plit[rel] = (uint8_t) pl; c.mbits[rel >> 5] |= 1u << (rel & 31);
                                 const uint32_t d = c.dec[rel]; pl = v4_dec_kind(d) == kV4Lit; last_rel = rel; x += (int) v4_dec_step(d); }
              exitpos = x; }
            w.rpos = -1; w.level2 = w.level;
            if (r.op + 2 * kV4N + 1 >= kSubSymbols) {
                int op = r.op;
                for (int rel = 0; rel < kV4W; rel++) {
                    if (!c.mark[rel]) continue;
                    if (op + 1 >= kSubSymbols) { w.rpos = lo + rel; op_at_rpos = op; break; }
                    op += (int) v4_dec_syms(c.dec[rel]);
                }
                if (w.rpos >= 0) w.level2 = v4_next_level(c, r.j + 1, w.rpos - r.enc_begin, op_at_rpos);
            }
            for (int i = 0; i < 256; i++) cntw[i] = 0;
            for (int rel = 0; rel < kV4N; rel++) {
                const uint32_t kx = c.key[rel];
                if (kx & kV4KeyInvalid) { c.rank[rel] = 0; continue; }
                c.rank[rel] = (uint16_t) cntw[v4_ctx_of(kx)];
                if (c.mark[rel]) cntw[v4_ctx_of(kx)]++;
            }
            for (int i = 0; i < 256; i++) c.mcnt[i] = cntw[i];
            for (int i = 0; i < kV4PfWords; i++) c.pf[i] = 0;
            for (int i = 0; i < 256; i++) c.pushw[i] = 0;
            for (int rel = 0; rel < kV4W; rel++) if (c.mark[rel]) { uint32_t c3, pw; v4_push_of(c, (uint32_t) (lo + rel), &c3, &pw); const uint32_t h = v4_pf_hash(c3, pw); c.pf[h >> 5] |= 1u << (h & 31u); c.pushw[c3] |= 1u << (rel >> 5); }
            bool changed = false;
            for (int rel = w.entry - lo; rel < wend - lo; rel++) {
                if (marked_only && !c.mark[rel]) { c.ndec[rel] = c.dec[rel]; continue; }
                const uint32_t nd = v4_decide(c, w, rel);
                st_decides++;
                c.ndec[rel] = nd;
                if (c.mark[rel] && ((nd ^ c.dec[rel]) & kV4DecCmp)) changed = true;
            }
            for (int rel = w.entry - lo; rel < wend - lo; rel++) c.dec[rel] = c.ndec[rel];
            if (!changed) break;
            if (rounds > 3 * kV4W) { fprintf(stderr, "window %d did not converge\n", k); return 3; }
        }
        st_rounds += rounds; st_windows++; if ((unsigned long long) rounds > st_maxrounds) st_maxrounds = rounds;
        st_hist[rounds < 8 ? rounds : 8]++;
        // ---- FINALIZE
        memset(c.sup, 0, kV4N + 2);
        std::vector<uint32_t> suffix(kV4N);
        for (int rel = 0; rel < kV4W; rel++) if (c.mark[rel]) suffix[rel] = v4_claim_slot(c, rel);
        int syms = 0, syms_after = 0;
        for (int rel = 0; rel < kV4W; rel++) {
            if (!c.mark[rel]) continue;
            if (w.rpos >= 0 && lo + rel == w.rpos) {                          // sub-block full (lz.cpp:153): close it, open the next
                v4_close_subblock(c, r, w.rpos, op_at_rpos, nt);
                r.j++; r.level = w.level2; r.tok_begin = nt; r.enc_begin = w.rpos;
            }
            v4_apply_position(c, lo, rel, suffix[rel]);
            const uint32_t t = v4_token_of(c, lo, rel);
            c.tok[nt] = t;
            if (v4_dec_kind(c.dec[rel]) == kV4Lit) c.lit[nl++] = (uint32_t) nt;
            nt++;
            const int s = (int) v4_dec_syms(c.dec[rel]);
            syms += s; if (w.rpos >= 0 && lo + rel >= w.rpos) syms_after += s;
        }
        for (int i = 0; i < 256; i++) c.mru2[i] = v4_mru_state(c, w, wend - lo - 1, (uint32_t) i);
        for (int i = 0; i < 256; i++) { c.mru[i] = c.mru2[i]; c.cnt[i] += cntw[i]; }
        r.op = w.rpos >= 0 ? syms_after : r.op + syms;
        r.ip = exitpos; r.skip_push = 0; r.prev_lit = v4_dec_kind(c.dec[last_rel]) == kV4Lit;
    }
    r.tail = 1;
    v4_resolve_tail(c, r, &nt, &nl);
    if (ilen > 0) v4_close_subblock(c, r, r.ip, r.op, nt);
    const int nsub = ilen > 0 ? r.j + 1 : 0;
    const int sim_nt = nt;

    // ---------------- oracle
    zo_rolz* z = zo_rolz_new();
    std::vector<uint16_t> sym(kSubSymbols + 300);
    std::vector<uint32_t> tp(kSubSymbols);
    std::vector<uint8_t> tr(kSubSymbols);
    int encpos = 0, j = 0;
    size_t t = 0;
    long long bad = -1;
    while (encpos < ilen && bad < 0) {
        const int lv = j < nsub ? (int) sub[j].level : level;             // the level the replay used (pinned or predicted)
        zo_rolz_trace(z, tp.data(), tr.data(), (int) tp.size());
        const int enc0 = encpos;
        const int rlen = zo_rolz_encode(z, lv, in, sym.data(), ilen, kSubSymbols, &encpos);
        const int ntk = zo_rolz_trace_count(z);
        if (j >= nsub) { fprintf(stderr, "sim produced %d sub-blocks, oracle has more\n", nsub); bad = (long long) t; break; }
        const SubBlock& sb = sub[j];
        if ((int) sb.rlen != rlen || (int) sb.enc_end != encpos || (int) sb.enc_begin != enc0 || (int) (sb.tok_end - sb.tok_begin) != ntk || (int) sb.level != lv) {
            fprintf(stderr, "sub-block %d differs: sim rlen %u enc [%u,%u) ntok %u level %u | oracle rlen %d enc [%d,%d) ntok %d level %d\n", j, sb.rlen, sb.enc_begin,
                    sb.enc_end, sb.tok_end - sb.tok_begin, sb.level, rlen, enc0, encpos, ntk, lv);
            if (bad < 0) bad = (long long) t;
        }
        int si = 0;
        for (int q = 0; q < ntk; q++, t++) {
            const uint32_t s = sym[si++];
            Tok want = { tp[q], s, 0, 0 };
            if (s >= 258) want.aux = sym[si++];
            if (s < 256) { want.sym = tr[q]; }                            // compare raw literal bytes (MTF is a later kernel)
            const uint32_t g = tok[t];
            Tok got = { 0, tok_sym(g), tok_sym(g) >= 258 ? tok_aux(g) : 0, 0 };
            if (tok_sym(g) < 256) got.sym = tok_byte(g);
            if ((long long) t >= sim_nt || got.sym != want.sym || got.aux != want.aux || (s < 256) != (tok_sym(g) < 256)) {
                fprintf(stderr, "token %zu (sub-block %d, pos %u) differs: sim sym %u aux %u | oracle sym %u aux %u\n", t, j, want.pos, got.sym, got.aux, want.sym, want.aux);
                bad = (long long) t;
                break;
            }
        }
        j++;
    }
    if (bad < 0 && ((long long) t != sim_nt || j != nsub)) { fprintf(stderr, "count mismatch: sim %d tokens %d subs, oracle %zu tokens %d subs\n", sim_nt, nsub, t, j); bad = 0; }
    zo_rolz_free(z);
    printf("%s level %d: %d bytes, %d tokens, %d sub-blocks, windows %llu, rounds/window %.2f (max %llu), decides %llu -> %s\n", argv[1], level, ilen, sim_nt, nsub,
           st_windows, st_windows ? (double) st_rounds / st_windows : 0.0, st_maxrounds, st_decides, bad < 0 ? "OK" : "MISMATCH");
    printf("  rounds histogram:");
    for (int i = 1; i <= 8; i++) printf(" %d%s:%llu", i, i == 8 ? "+" : "", st_hist[i]);
    printf("\n");
    (void) st_general;
#if defined(ZL_V4_STATS)
    const char* names[8] = { "link_hazard steps", "rank_live words", "mru words", "mru pushes", "general calls", "stale replays", "word: in bloom", "word: in carried" };
    for (int k = 0; k < 8; k++) {
        printf("  %-18s calls %llu (%.2f/window) mean %.2f max %llu  hist(<1,<2,<4,..):", names[k], g_v4stats.calls[k], (double) g_v4stats.calls[k] / (st_windows ? st_windows : 1),
               g_v4stats.calls[k] ? (double) g_v4stats.steps[k] / g_v4stats.calls[k] : 0.0, g_v4stats.maxsteps[k]);
        for (int b = 0; b < 12; b++) printf(" %llu", g_v4stats.hist[k][b]);
        printf("\n");
    }
#endif
    return bad < 0 ? 0 : 1;
}
