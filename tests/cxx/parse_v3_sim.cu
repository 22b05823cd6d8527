// Host replay of zl_rolz_parse_v3's phases (libzling_b200/csrc/zl_parse_v3.cuh) against the oracle's tokeniser.
// TEST INFRASTRUCTURE: every algorithmic function of the v3 kernel is scalar ZL_HD code; this program runs the same
// functions in the kernel's phase order with loops instead of threads, so the speculate / resolve / apply logic
// (hazard links, staleness, level changes, window edges) is checked bit-for-bit on the CPU before any GPU time is
// spent.  What it cannot check is the kernel's synchronisation — that is what the -m gpu parity tests are for.
//   parse_v3_sim <file> <level> [order] [plan-hex]      order 0: SPEC(k+1) before RESOLVE(k), 1: after
// exit code 0 = token streams identical.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../libzling_b200/csrc/zl_parse_v3.cuh"

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
    if (argc < 3) { fprintf(stderr, "usage: parse_v3_sim file level [order] [plan-digits]\n"); return 2; }
    const int level = atoi(argv[2]);
    const int order = argc > 3 ? atoi(argv[3]) : 0;
    uint8_t plan[kMaxSubPerBlock];
    memset(plan, level, sizeof plan);
    // plan digits pin levels per sub-block; 'a' = auto (the kernel predicts), "auto" = every sub-block auto
    if (argc > 4 && !strcmp(argv[4], "auto")) memset(plan, 0xff, sizeof plan);
    else if (argc > 4) for (size_t i = 0; i < strlen(argv[4]) && i < (size_t) kMaxSubPerBlock; i++) plan[i] = argv[4][i] == 'a' ? 0xff : (uint8_t) (argv[4][i] - '0');
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<uint8_t> data;
    { uint8_t buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n); }
    fclose(f);
    if (data.size() > (size_t) kBlockBytes) data.resize(kBlockBytes);
    const int ilen = (int) data.size();

    // ---------------- v3 replay
    uint8_t* in = (uint8_t*) aligned_alloc(64, (size_t) ilen + 512);
    memset(in, 0, (size_t) ilen + 512);
    memcpy(in, data.data(), ilen);
    std::vector<uint64_t> ring((size_t) 256 * kRing, kRingEmpty);
    std::vector<uint16_t> hash((size_t) 256 * kSlots, 0xffff);
    std::vector<uint32_t> tok((size_t) ilen + 16), lit((size_t) ilen + 16);
    std::vector<SubBlock> sub(kMaxSubPerBlock);
    const int maxlevel = level;        // records hold the requested level's depth; plan levels are <= it
    const int dmax = depth_main(maxlevel), lmax = depth_lazy1(maxlevel);
    const V3Layout L = v3_layout(dmax, lmax);
    uint8_t* smem = (uint8_t*) aligned_alloc(64, (size_t) L.total + 64);
    memset(smem, 0, L.total);
    V3Ctx c;
    v3_bind(c, smem, L);
    c.in = in; c.ilen = ilen; c.ring = ring.data(); c.hash = hash.data(); c.tok = tok.data(); c.lit = lit.data();
    c.sub = sub.data(); c.plan = plan; c.base_level = level;
    for (int i = 0; i < kV3R; i++) c.key[i] = kKeyInvalid;

    V3Run r; memset(&r, 0, sizeof r);
    r.level = v3_next_level(c, r, 0); r.skip_push = 1;
    int nt = 0, nl = 0;
    for (int first = 0; first < 2; first++)
        if (r.ip == first && r.ip < ilen) { c.tok[nt++] = tok_literal(in[r.ip], 0, true); r.op++; r.ip++; }
    std::vector<uint16_t> dump_flen((size_t) ilen + 600, 0);
    std::vector<uint8_t> dump_mark((size_t) ilen + 600, 0);
    int s_level = r.level, s_tlevel[2] = { 0, 0 };
    const int lim = ilen - kGuard;
    const int nwin = lim > 2 ? (lim + kV3W - 1) / kV3W : 0;
    int staged_hi = -16;
    auto spec = [&](int j, int tlevel) {
        const int hi = v3_stage_hi(j);
        for (int src = staged_hi; src < hi; src += 16) v3_stage16(c, src);
        const int nlo = v3_new_lo(j), nhi = v3_new_hi(j);
        for (int i = 0; i < 256; i++) c.pcnt[256 * (j & 1) + i] = 0;
        for (int x = nlo; x < nhi; x++) v3_key_position(c, x, j);
        v3_bucket_pass_serial(c, nlo, nhi);
        for (int rel = 0; rel < kV3W + 2; rel++) v3_spec_position(c, j, rel);
        for (int x = nlo; x < nhi; x++) v3_link_position(c, x, v3_base(j));
        for (int rel = 0; rel < kV3W; rel++) v3_decide_position(c, j, rel, tlevel);
        for (int rel = 0; rel < kV3W && j * kV3W + rel < ilen; rel++) dump_flen[(size_t) j * kV3W + rel] = (uint16_t) (v3_table(c, j).dec[rel].x & 511u);
        s_tlevel[j & 1] = tlevel;
    };
    for (int k = -1; k < nwin; k++) {
        const int tlevel_next = s_level;
        if (order == 0 && k + 1 < nwin) spec(k + 1, tlevel_next);
        if (k >= 0) v3_resolve_window(c, r, k, s_tlevel[k & 1], nt);
        if (order != 0 && k + 1 < nwin) spec(k + 1, tlevel_next);
        if (k >= 0) {                                                    // APPLY + EMIT, position order = token order
            const int lo = k * kV3W;
            for (int y = lo; y < lo + kV3W && y < ilen; y++) {
                const uint32_t m = c.ins[y & (kV3R - 1)];
                if (!v3_kind(m)) continue;
                v3_apply_position(c, y);
                dump_mark[y] = (uint8_t) (v3_kind(m) | ((m & kInsExplicit) ? 8u : 0u));
                c.tok[nt] = v3_token_of(c, y, m);
                if (v3_kind(m) == kKindLit) c.lit[nl++] = (uint32_t) nt;
                nt++;
            }
            uint32_t* snap = c.snap + 256 * ((k + 1) % 3);
            for (int i = 0; i < 256; i++) snap[i] = c.cnt[i];
            s_level = r.level;
        }
        staged_hi = v3_stage_hi(k + 1);
    }
    v3_resolve_tail(c, r, &nt, &nl);
    if (ilen > 0) v3_close_subblock(c, r, nt);
    const int nsub = ilen > 0 ? r.j + 1 : 0;
    const int sim_nt = nt;

    if (getenv("ZL_V3_DUMP")) {                                        // analysis aid (scripts/chain_sync_stats.py)
        FILE* df = fopen(getenv("ZL_V3_DUMP"), "wb");
        if (df) { fwrite(dump_flen.data(), 2, ilen, df); fwrite(dump_mark.data(), 1, ilen, df); fclose(df); }
    }

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
        const int nt = zo_rolz_trace_count(z);
        if (j >= nsub) { fprintf(stderr, "sim produced %d sub-blocks, oracle has more\n", nsub); bad = (long long) t; break; }
        const SubBlock& sb = sub[j];
        if ((int) sb.rlen != rlen || (int) sb.enc_end != encpos || (int) sb.enc_begin != enc0 || (int) (sb.tok_end - sb.tok_begin) != nt || (int) sb.level != lv) {
            fprintf(stderr, "sub-block %d differs: sim rlen %u enc [%u,%u) ntok %u level %u | oracle rlen %d enc [%d,%d) ntok %d level %d\n", j, sb.rlen, sb.enc_begin,
                    sb.enc_end, sb.tok_end - sb.tok_begin, sb.level, rlen, enc0, encpos, nt, lv);
        }
        int si = 0;
        for (int q = 0; q < nt; q++, t++) {
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
    printf("%s level %d order %d: %d bytes, %d tokens, %d sub-blocks, flagged %llu (%.2f%%), general %llu (%.2f%%), slow %llu, linkwalk %llu -> %s\n", argv[1], level, order, ilen,
           sim_nt, nsub, (unsigned long long) r.n_flagged, 100.0 * r.n_flagged / (sim_nt ? sim_nt : 1), (unsigned long long) r.n_general, 100.0 * r.n_general / (sim_nt ? sim_nt : 1),
           (unsigned long long) r.n_slow, (unsigned long long) r.n_linkwalk,
           bad < 0 ? "OK" : "MISMATCH");
#if defined(ZL_V3_FLAG_HIST)
    for (int i = 0; i < 256; i++) if (g_flag_hist[i]) printf("  flags %02x: %llu\n", i, g_flag_hist[i]);
#endif
    return bad < 0 ? 0 : 1;
}
