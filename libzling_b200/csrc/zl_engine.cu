// Host engine + C ABI (include/zlb.h) of the zling block pipeline.  One zlb_ctx per (process, GPU): device
// buffers for up to max_blocks 16 MiB blocks, one CUDA stream, pinned mirrors of the small tables.
//
// Encode of one call (N blocks of one stream):
//   H2D input -> [reset buckets -> zl_rolz_parse (one chain per block, all blocks concurrently)
//                 -> zl_mtf_rank (stream order, carried state) -> zl_huff_build (per sub-block)]*
//             -> host verifies the level-feedback plan (src/libzling.cpp:261-266) and, if a prediction was wrong,
//                re-parses the affected blocks (loop marked * above)
//             -> zl_huff_pack (frames written at their final offsets) -> D2H
// There is no CPU implementation of any stage in this library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

#include "../../include/zlb.h"
#include "zl_kernels.cuh"

#include "zl_kernels.cu"      // single translation unit: kernels + engine
#include "zl_shard.cuh"

using namespace zl;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return ZLB_E_CUDA; } } while (0)

enum { EV_START, EV_H2D, EV_PARSE0, EV_PARSE1, EV_MTF1, EV_BUILD1, EV_PACK0, EV_PACK1, EV_END, EV_COUNT };

struct zlb_ctx {
    int device = 0, max_blocks = 0;
    int max_clusters[5][5];         // [level][log2 size]: clusters of that size the GPU holds at once (-1 = not asked yet)
    int stats_cluster = 1;          // cluster size of the last parse launch
    uint32_t spec_reparsed = 0;     // blocks re-parsed by the last speculative completion (sharded streams)
    cudaStream_t stream = nullptr;
    // device
    uint8_t*  d_in = nullptr;       // max_blocks * 16 MiB + pad (host-input path)
    uint8_t*  d_out = nullptr;      // zlb_encode_bound(max_blocks * 16 MiB); decode: output blocks
    uint64_t* d_ring = nullptr;
    uint16_t* d_hash = nullptr;
    uint32_t* d_tok = nullptr;      // encode: tokens; decode: u16 symbols (2 per u32)
    uint32_t* d_lit = nullptr;
    SubBlock* d_sub = nullptr;
    HuffTables* d_tab = nullptr;
    uint32_t *d_nsub = nullptr, *d_ntok = nullptr, *d_nlit = nullptr, *d_ilen = nullptr;
    uint8_t  *d_plan = nullptr, *d_active = nullptr, *d_active2 = nullptr, *d_ckpt = nullptr;
    unsigned long long* d_outoff = nullptr;
    DecSub*   d_decsub = nullptr;
    int*      d_status = nullptr;
    uint32_t* d_decring = nullptr;
    uint8_t*  d_comp = nullptr;     // decode: compressed input
    size_t    comp_cap = 0, out_cap = 0, decsub_cap = 0;
    // pinned host mirrors
    SubBlock* h_sub = nullptr;
    uint32_t *h_nsub = nullptr, *h_ntok = nullptr, *h_nlit = nullptr, *h_ilen = nullptr;
    uint8_t  *h_plan = nullptr, *h_active = nullptr, *h_active2 = nullptr;
    unsigned long long* h_outoff = nullptr;
    int*      h_status = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    zlb_stats stats = {};
    int last_nblocks = 0;
    const void* pending = nullptr;  // encoder with a submitted, not yet completed range: it owns the context's buffers
    V4Counters* d_v4c = nullptr;
    V4Counters  h_v4c = {};
    int parse_cluster = 0;          // CTAs per block of the parse kernel; 0 = choose by the number of blocks (ZLB_PARSE_CLUSTER=1|2|4|8|16 pins it)
    uint32_t *d_lbuf = nullptr, *d_lhist = nullptr, *d_ctxoff = nullptr;
    MtfRange *d_mrng = nullptr, *h_mrng = nullptr;   // per-stream block ranges of the MTF pass (max_blocks entries)
    uint8_t* d_bstate = nullptr;                     // batch API: initial MTF tables + one scratch table per stream, allocated on first use
    int bstate_streams = 0;
    DecRange *d_drng = nullptr, *h_drng = nullptr;   // per-stream sub-block ranges of the ROLZ decode (max_blocks entries)
    uint8_t* d_dstate = nullptr;                     // batch decode: one MTF table set per stream
    int dstate_streams = 0, decring_streams = 0;
};

struct zlb_encoder {
    zlb_ctx* ctx;
    int level, cur_level;
    uint8_t* d_state[2];     // MTF rank->byte tables, ping-pong (input of a call stays intact until it succeeds)
    int cur;
    size_t submitted_n;      // bytes handed to zlb_encode_submit and not yet completed (0 = nothing pending)
};
enum { MODE_ALL = 0, MODE_SUBMIT = 1, MODE_COMPLETE = 2, MODE_SPECULATE = 3 };
struct zlb_decoder {
    zlb_ctx* ctx;
    uint8_t* d_state;
};

static size_t bound_bytes(size_t n) {
    // per sub-block: 13 frame bytes + 273 table bytes + <= 15 bits/symbol for <= 262144 symbols; a sub-block
    // covers >= 262143 input bytes except the last of each block; + 1 stop byte per block
    const size_t blocks = n / ZLB_BLOCK_BYTES + 1;
    const size_t subs = n / 262143 + blocks + 1;
    return n + n / 64 + subs * (13 + 273 + 8) + blocks + 1024;
}

extern "C" {

const char* zlb_last_error(void) { return g_err; }
const char* zlb_version(void) { return "libzling_b200 0.1 (sm_100a)"; }
size_t zlb_encode_bound(size_t n) { return bound_bytes(n); }

void* zlb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fail(ZLB_E_NOMEM, "zlb_host_alloc: cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
void zlb_host_free(void* p) { if (p) cudaFreeHost(p); }

int zlb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void zlb_destroy(zlb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void* dev[] = { c->d_in, c->d_out, c->d_ring, c->d_hash, c->d_tok, c->d_lit, c->d_sub, c->d_tab, c->d_nsub, c->d_ntok, c->d_nlit,
                    c->d_ilen, c->d_plan, c->d_active, c->d_active2, c->d_ckpt, c->d_outoff, c->d_decsub, c->d_status, c->d_decring, c->d_comp,
                    c->d_v4c, c->d_lbuf, c->d_lhist, c->d_ctxoff, c->d_mrng, c->d_bstate, c->d_drng, c->d_dstate };
    for (void* p : dev) if (p) cudaFree(p);
    if (c->h_mrng) cudaFreeHost(c->h_mrng);
    if (c->h_drng) cudaFreeHost(c->h_drng);
    void* host[] = { c->h_sub, c->h_nsub, c->h_ntok, c->h_nlit, c->h_ilen, c->h_plan, c->h_active, c->h_active2, c->h_outoff, c->h_status };
    for (void* p : host) if (p) cudaFreeHost(p);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int ctx_alloc(zlb_ctx* c) {
    const size_t nb = (size_t) c->max_blocks, nsb = nb * kMaxSubPerBlock;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& e : c->ev) CU(cudaEventCreate(&e));
    c->out_cap = bound_bytes(nb * kBlockBytes);
    c->comp_cap = c->out_cap;
    c->decsub_cap = nb * 160;       // >= 65 sub-blocks per block in valid streams; more are accepted up to this
    CU(cudaMalloc(&c->d_in, nb * kBlockBytes + 256));
    CU(cudaMalloc(&c->d_out, c->out_cap + 256));
    CU(cudaMalloc(&c->d_ring, nb * kRingStride * sizeof(uint64_t)));
    CU(cudaMalloc(&c->d_hash, nb * kHashStride * sizeof(uint16_t)));
    CU(cudaMalloc(&c->d_tok, nb * kTokStride * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_lit, nb * kLitStride * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_sub, nsb * sizeof(SubBlock)));
    CU(cudaMalloc(&c->d_tab, nsb * sizeof(HuffTables)));
    CU(cudaMalloc(&c->d_nsub, nb * 4)); CU(cudaMalloc(&c->d_ntok, nb * 4)); CU(cudaMalloc(&c->d_nlit, nb * 4)); CU(cudaMalloc(&c->d_ilen, nb * 4));
    CU(cudaMalloc(&c->d_plan, nsb)); CU(cudaMalloc(&c->d_active, nb)); CU(cudaMalloc(&c->d_active2, nb)); CU(cudaMalloc(&c->d_ckpt, (nb + 1) * 65536));
    CU(cudaMalloc(&c->d_outoff, nsb * sizeof(unsigned long long)));
    CU(cudaMalloc(&c->d_decsub, c->decsub_cap * sizeof(DecSub)));
    CU(cudaMalloc(&c->d_status, (c->decsub_cap + nb + 4) * sizeof(int)));
    CU(cudaMalloc(&c->d_decring, (size_t) 256 * kRing * sizeof(uint32_t)));
    c->decring_streams = 1;
    CU(cudaMalloc(&c->d_comp, c->comp_cap + 256));
    CU(cudaMemset(c->d_in, 0, nb * kBlockBytes + 256));
    CU(cudaHostAlloc(&c->h_sub, nsb * sizeof(SubBlock), cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->h_nsub, nb * 4, cudaHostAllocDefault)); CU(cudaHostAlloc(&c->h_ntok, nb * 4, cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->h_nlit, nb * 4, cudaHostAllocDefault)); CU(cudaHostAlloc(&c->h_ilen, nb * 4, cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->h_plan, nsb, cudaHostAllocDefault)); CU(cudaHostAlloc(&c->h_active, nb, cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->h_active2, nb, cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->h_outoff, nsb * sizeof(unsigned long long), cudaHostAllocDefault));
    CU(cudaHostAlloc(&c->h_status, (c->decsub_cap + nb + 4) * sizeof(int), cudaHostAllocDefault));
    CU(cudaMalloc(&c->d_v4c, sizeof(V4Counters)));
    CU(cudaMalloc(&c->d_lbuf, nb * kLitStride * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_lhist, nb * (size_t) kLitUnitsMax * 256 * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_ctxoff, nb * 257 * sizeof(uint32_t)));
    CU(cudaMalloc(&c->d_mrng, nb * sizeof(MtfRange)));
    CU(cudaHostAlloc(&c->h_mrng, nb * sizeof(MtfRange), cudaHostAllocDefault));
    CU(cudaMalloc(&c->d_drng, nb * sizeof(DecRange)));
    CU(cudaHostAlloc(&c->h_drng, nb * sizeof(DecRange), cudaHostAllocDefault));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, v4_layout(2, 1).total));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, v4_layout(4, 1).total));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, v4_layout(6, 2).total));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, v4_layout(8, 3).total));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, v4_layout(16, 4).total));
    // clusters of 16 CTAs are beyond the portable size (8): opt in per kernel
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<2, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<4, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<6, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<8, 3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CU(cudaFuncSetAttribute(zl_rolz_parse_v4_kernel<16, 4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    { const char* pv = getenv("ZLB_PARSE_CLUSTER"); if (pv) { const int v = atoi(pv); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) c->parse_cluster = v; } }
    CU(cudaFuncSetAttribute(zl_huff_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CU(cudaFuncSetAttribute(zl_rolz_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    return ZLB_OK;
}

zlb_ctx* zlb_create(int device, int max_blocks) {
    if (max_blocks < 1 || max_blocks > 1024) { fail(ZLB_E_ARG, "zlb_create: max_blocks must be in 1..1024"); return nullptr; }
    int n = zlb_device_count();
    if (n <= 0) { fail(ZLB_E_NODEVICE, "zlb_create: no CUDA device available (this library has no CPU path)"); return nullptr; }
    if (device < 0 || device >= n) { fail(ZLB_E_ARG, "zlb_create: device index out of range"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(ZLB_E_CUDA, "zlb_create: cudaSetDevice failed"); return nullptr; }
    zlb_ctx* c = new (std::nothrow) zlb_ctx();
    if (!c) { fail(ZLB_E_NOMEM, "zlb_create: out of host memory"); return nullptr; }
    c->device = device; c->max_blocks = max_blocks;
    for (auto& row : c->max_clusters) for (int& v : row) v = -1;
    if (ctx_alloc(c) != ZLB_OK) { zlb_destroy(c); return nullptr; }
    return c;
}
int zlb_max_blocks(const zlb_ctx* c) { return c ? c->max_blocks : ZLB_E_ARG; }

int zlb_get_stats(const zlb_ctx* c, zlb_stats* out) {
    if (!c || !out) return fail(ZLB_E_ARG, "zlb_get_stats: null argument");
    *out = c->stats;
    return ZLB_OK;
}

// ------------------------------------------------------------------------------------------ encoder
zlb_encoder* zlb_encoder_begin(zlb_ctx* c, int level) {
    if (!c) { fail(ZLB_E_ARG, "zlb_encoder_begin: null context"); return nullptr; }
    if (level < 0 || level > 4) { fail(ZLB_E_ARG, "zlb_encoder_begin: level must be 0..4"); return nullptr; }   // lz.cpp:136
    cudaSetDevice(c->device);
    zlb_encoder* e = new (std::nothrow) zlb_encoder();
    if (!e) { fail(ZLB_E_NOMEM, "zlb_encoder_begin: out of host memory"); return nullptr; }
    e->ctx = c; e->level = level; e->cur_level = level; e->cur = 0; e->d_state[0] = e->d_state[1] = nullptr; e->submitted_n = 0;
    uint8_t init[65536];
    for (int ctx = 0; ctx < 256; ctx++) memcpy(init + ctx * 256, kMtfInit, 256);          // lz.cpp:106-111
    if (cudaMalloc(&e->d_state[0], 65536) != cudaSuccess || cudaMalloc(&e->d_state[1], 65536) != cudaSuccess ||
        cudaMemcpy(e->d_state[0], init, 65536, cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(ZLB_E_CUDA, "zlb_encoder_begin: device allocation failed");
        if (e->d_state[0]) cudaFree(e->d_state[0]);
        if (e->d_state[1]) cudaFree(e->d_state[1]);
        delete e; return nullptr;
    }
    return e;
}
void zlb_encoder_end(zlb_encoder* e) {
    if (!e) return;
    cudaSetDevice(e->ctx->device);
    if (e->ctx->pending == e) { cudaStreamSynchronize(e->ctx->stream); e->ctx->pending = nullptr; }   // abandoned submit
    cudaFree(e->d_state[0]); cudaFree(e->d_state[1]);
    delete e;
}
int zlb_encoder_get_state(zlb_encoder* e, uint8_t* state) {
    if (!e || !state) return fail(ZLB_E_ARG, "zlb_encoder_get_state: null argument");
    CU(cudaSetDevice(e->ctx->device));
    CU(cudaMemcpy(state, e->d_state[e->cur], 65536, cudaMemcpyDeviceToHost));
    const int32_t lv = e->cur_level; memcpy(state + 65536, &lv, 4);
    return ZLB_OK;
}
int zlb_encoder_set_state(zlb_encoder* e, const uint8_t* state) {
    if (!e || !state) return fail(ZLB_E_ARG, "zlb_encoder_set_state: null argument");
    int32_t lv; memcpy(&lv, state + 65536, 4);
    if (lv < 0 || lv > 4) return fail(ZLB_E_ARG, "zlb_encoder_set_state: bad level");
    CU(cudaSetDevice(e->ctx->device));
    CU(cudaMemcpy(e->d_state[e->cur], state, 65536, cudaMemcpyHostToDevice));
    e->cur_level = lv;
    return ZLB_OK;
}

// the level-feedback rule: 1.0 * olen / (consumed + 1) > 0.95, libzling.cpp:261 (exact in integers: SURVEY §8 a11)
static inline bool incompressible(const SubBlock& s) {
    return (unsigned long long) s.olen * 20ull > (unsigned long long) (s.enc_end - s.enc_begin + 1) * 19ull;
}

// mode MODE_ALL: the whole pipeline.  MODE_SUBMIT: plan + parse launch only, returns without synchronising (the
// parse does not need the carried MTF state, and needs the carried level only for the first sub-block).
// MODE_COMPLETE: the rest, after the carried state may have been replaced (zlb_encoder_set_state); block 0 is
// parsed again only if the carried level turned out different from the one the submit assumed.
// one stream's contiguous block range inside a call
struct EncRange {
    int b0, b1;                    // blocks [b0, b1) of the call's block array
    int level_in, level_out;       // current_level carried in (src/libzling.cpp:185) / after the last sub-block
    uint8_t* st_in; uint8_t* st_out;   // MTF tables carried in / out (device, 65536 B each)
    int first_dirty;               // first block whose ranks must be recomputed in the next pass (b1 = none)
    unsigned long long out_off, out_len;   // where the range's framed bytes went inside d_out
};

// The pipeline over the blocks of a call: c->h_ilen[0..nb) is filled by the caller; R[0..nr) are the streams (ONE for the
// stream API, many for the batch API).  mode MODE_ALL: the whole pipeline.  MODE_SUBMIT (nr == 1): plan + parse launch only,
// returns without synchronising (the parse does not need the carried MTF state, and needs the carried level only for the first
// sub-block).  MODE_COMPLETE (nr == 1): the rest, after the carried state may have been replaced (zlb_encoder_set_state);
// block 0 is parsed again only if the carried level turned out different from the one the submit assumed.
static int encode_ranges(zlb_ctx* c, int level, const uint8_t* d_in, int nb, EncRange* R, int nr, uint8_t* d_out, size_t out_cap, size_t* out_len, int mode, const uint8_t* pre_tail = nullptr) {
    cudaStream_t st = c->stream;
    c->last_nblocks = nb;
    uint32_t launches = 0, parse_launches = 0, reparsed = 0;
    bool skip_first_parse = false;
    if (mode == MODE_SPECULATE) {
        skip_first_parse = true;                                  // the submitted parse is the first pass
    } else if (mode != MODE_COMPLETE) {
        for (int b = 0; b < nb; b++) {
            c->h_active[b] = 1;
            // the parse predicts the level of every sub-block the host has not pinned (kV4Auto, zl_parse_v4.cuh: v4_next_level)
            memset(c->h_plan + (size_t) b * kMaxSubPerBlock, (int) kV4Auto, kMaxSubPerBlock);
        }
        for (int r = 0; r < nr; r++) c->h_plan[(size_t) R[r].b0 * kMaxSubPerBlock] = (uint8_t) R[r].level_in;   // current_level outlives blocks, libzling.cpp:185
        // a range of a sharded stream whose carried level is still on its way: the kernel predicts it from the bytes that precede
        // the range (pre_tail), the walk below verifies it against the carried level like every other prediction
        if (pre_tail && nr == 1 && nb > 0) c->h_plan[(size_t) R[0].b0 * kMaxSubPerBlock] = (uint8_t) kV4Auto;
        CU(cudaMemcpyAsync(c->d_ilen, c->h_ilen, nb * 4, cudaMemcpyHostToDevice, st));
    } else if (c->h_plan[0] == (uint8_t) R[0].level_in || c->h_plan[0] == (uint8_t) kV4Auto) {
        skip_first_parse = true;                                  // the submit's guess of the carried level was right (or the kernel's
                                                                  // own prediction is verified by the walk below)
    } else {
        c->h_plan[0] = (uint8_t) R[0].level_in;
        memset(c->h_plan + 1, (int) kV4Auto, kMaxSubPerBlock - 1);  // levels pinned by a speculative completion rested on the wrong start
        for (int b = 0; b < nb; b++) c->h_active[b] = b == 0;
        reparsed++;
    }
    for (int r = 0; r < nr; r++) R[r].first_dirty = R[r].b0;

    ParseArgs pa;
    pa.in = d_in; pa.ilen = c->d_ilen; pa.plan = c->d_plan; pa.active = c->d_active; pa.ring = c->d_ring; pa.hash = c->d_hash;
    pa.tok = c->d_tok; pa.lit = c->d_lit; pa.sub = c->d_sub; pa.nsub = c->d_nsub; pa.ntok = c->d_ntok; pa.nlit = c->d_nlit;
    pa.pre_tail = pre_tail;

    float ms_parse = 0, ms_mtf = 0, ms_build = 0;
    for (int pass = 0;; pass++) {
        if (!(skip_first_parse && pass == 0)) {
            CU(cudaMemcpyAsync(c->d_plan, c->h_plan, (size_t) nb * kMaxSubPerBlock, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(c->d_active, c->h_active, nb, cudaMemcpyHostToDevice, st));
            CU(cudaEventRecord(c->ev[EV_PARSE0], st));
            zl_reset_buckets_kernel<<<296, 256, 0, st>>>(c->d_ring, c->d_hash, c->d_active, nb);
            const V4Layout lay = v4_layout(depth_main(level), depth_lazy1(level));
            if (pass == 0) CU(cudaMemsetAsync(c->d_v4c, 0, sizeof(V4Counters), st));
            {
                // one CTA per block, or a cluster of CTAs per block (helpers take the chain walks, zl_parse_v4.cuh) while the clusters
                // of all blocks of the call still fit the GPU at once
                // (the largest cluster of which the GPU can hold one per block that is parsed in this pass: asked from the occupancy
                // calculator once per level and size — a 16-CTA cluster needs 16 free SMs inside one GPC)
                int cl = c->parse_cluster;
                cudaLaunchConfig_t cfg = {};
                cfg.blockDim = dim3(kV4T); cfg.dynamicSmemBytes = (size_t) lay.total; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                if (cl <= 0) {
                    int n_active = 0;
                    for (int b = 0; b < nb; b++) n_active += c->h_active[b] != 0;
                    const int lv = level < 0 ? 0 : (level > 4 ? 4 : level);
                    cl = 1;
                    for (int k = 4; k >= 1 && cl == 1; k--) {             // 16, 8, 4, 2
                        int& cap = c->max_clusters[lv][k];
                        if (cap < 0) {
                            cap = 0;
                            attr[0].val.clusterDim.x = 1u << k; cfg.gridDim = dim3(1u << k);
                            int n = 0;
                            cudaError_t qe;
                            switch (lv) {
                                case 0:  qe = cudaOccupancyMaxActiveClusters(&n, zl_rolz_parse_v4_kernel<2, 1>, &cfg); break;
                                case 1:  qe = cudaOccupancyMaxActiveClusters(&n, zl_rolz_parse_v4_kernel<4, 1>, &cfg); break;
                                case 2:  qe = cudaOccupancyMaxActiveClusters(&n, zl_rolz_parse_v4_kernel<6, 2>, &cfg); break;
                                case 3:  qe = cudaOccupancyMaxActiveClusters(&n, zl_rolz_parse_v4_kernel<8, 3>, &cfg); break;
                                default: qe = cudaOccupancyMaxActiveClusters(&n, zl_rolz_parse_v4_kernel<16, 4>, &cfg); break;
                            }
                            if (qe == cudaSuccess) cap = n; else cudaGetLastError();
                        }
                        if (cap >= n_active && n_active > 0) cl = 1 << k;
                    }
                }
                attr[0].val.clusterDim.x = (unsigned) cl;
                cfg.gridDim = dim3((unsigned) (nb * cl));
                c->stats_cluster = cl;
                switch (level) {                                  // (depth, lazy depth) of the requested level, src/libzling_lz.cpp:129-135
                    case 0:  CU(cudaLaunchKernelEx(&cfg, zl_rolz_parse_v4_kernel<2, 1>, pa, level, c->d_v4c)); break;
                    case 1:  CU(cudaLaunchKernelEx(&cfg, zl_rolz_parse_v4_kernel<4, 1>, pa, level, c->d_v4c)); break;
                    case 2:  CU(cudaLaunchKernelEx(&cfg, zl_rolz_parse_v4_kernel<6, 2>, pa, level, c->d_v4c)); break;
                    case 3:  CU(cudaLaunchKernelEx(&cfg, zl_rolz_parse_v4_kernel<8, 3>, pa, level, c->d_v4c)); break;
                    default: CU(cudaLaunchKernelEx(&cfg, zl_rolz_parse_v4_kernel<16, 4>, pa, level, c->d_v4c)); break;
                }
            }
            CU(cudaEventRecord(c->ev[EV_PARSE1], st));
            launches += 2; parse_launches += 1;
        }
        if (mode == MODE_SUBMIT) { CU(cudaGetLastError()); return ZLB_OK; }
        int fd_min = nb;
        for (int r = 0; r < nr; r++) {
            MtfRange& m = c->h_mrng[r];
            m.first = R[r].first_dirty; m.end = R[r].b1; m.b0 = R[r].b0; m.pad = 0; m.state_in = R[r].st_in; m.state_out = R[r].st_out;
            if (R[r].first_dirty < fd_min) fd_min = R[r].first_dirty;
            for (int b = R[r].b0; b < R[r].b1; b++) c->h_active2[b] = b >= R[r].first_dirty;   // blocks whose ranks change: rebuild their tables
        }
        CU(cudaMemcpyAsync(c->d_mrng, c->h_mrng, (size_t) nr * sizeof(MtfRange), cudaMemcpyHostToDevice, st));
        {
            const dim3 lgrid(kLitUnitsMax / kLitWarps, nb - fd_min);
            zl_lit_count_kernel<<<lgrid, kLitWarps * 32, 0, st>>>(c->d_tok, c->d_lit, c->d_nlit, fd_min, c->d_lhist);
            zl_lit_scan_kernel<<<nb - fd_min, 256, 0, st>>>(c->d_nlit, fd_min, c->d_lhist, c->d_ctxoff);
            zl_lit_scatter_kernel<<<lgrid, kLitWarps * 32, 0, st>>>(c->d_tok, c->d_lit, c->d_nlit, fd_min, c->d_lhist, c->d_lbuf);
            zl_mtf_ctx_kernel<<<dim3(256, nr), 32, 0, st>>>(c->d_tok, c->d_lbuf, c->d_ctxoff, c->d_mrng, c->d_ckpt);
            launches += 4;
        }
        CU(cudaEventRecord(c->ev[EV_MTF1], st));
        CU(cudaMemcpyAsync(c->d_active2, c->h_active2, nb, cudaMemcpyHostToDevice, st));
        zl_huff_build_kernel<<<dim3(kMaxSubPerBlock, nb), 256, 0, st>>>(c->d_tok, c->d_sub, c->d_nsub, c->d_active2, c->d_tab);
        CU(cudaEventRecord(c->ev[EV_BUILD1], st));
        CU(cudaMemcpyAsync(c->h_sub, c->d_sub, (size_t) nb * kMaxSubPerBlock * sizeof(SubBlock), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->h_nsub, c->d_nsub, nb * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->h_ntok, c->d_ntok, nb * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaGetLastError());
        launches += 1;
        { float t; cudaEventElapsedTime(&t, c->ev[EV_PARSE0], c->ev[EV_PARSE1]); ms_parse += t;
          cudaEventElapsedTime(&t, c->ev[EV_PARSE1], c->ev[EV_MTF1]); ms_mtf += t;
          cudaEventElapsedTime(&t, c->ev[EV_MTF1], c->ev[EV_BUILD1]); ms_build += t; }

        // Verify the levels in stream order (libzling.cpp:261-266) and re-plan EVERY block that used a wrong one in
        // this pass: sub-blocks before the first wrong one are pinned to their (verified) levels, the wrong one gets
        // the level the reference uses, later ones are predicted again by the parse.  Behind a wrong block the
        // walk continues on the assumption that the compressibility of its last sub-block does not change with the
        // level (incompressible data stays incompressible), so all mispredicted blocks of a call are usually
        // re-parsed together, concurrently, in ONE extra pass.  A wrong assumption only costs another pass.
        int nbad = 0;
        for (int r = 0; r < nr; r++) {
            int cur = R[r].level_in, first_bad = -1;
            // MODE_SPECULATE: the carried level is not known yet; trust the one the parse predicted for the range's first sub-block
            if (mode == MODE_SPECULATE && R[r].b0 < R[r].b1 && c->h_nsub[R[r].b0] > 0) cur = (int) c->h_sub[(size_t) R[r].b0 * kMaxSubPerBlock].level;
            bool exact = true;                                        // `cur` is the reference's value, not an assumption
            for (int b = R[r].b0; b < R[r].b1; b++) {
                const int ns = (int) c->h_nsub[b];
                if (ns > kMaxSubPerBlock) return fail(ZLB_E_CUDA, "internal: sub-block table overflow");
                c->h_active[b] = 0;
                uint8_t* plan = c->h_plan + (size_t) b * kMaxSubPerBlock;
                int bad_j = -1;
                for (int j = 0; j < ns; j++) {
                    const SubBlock& sb = c->h_sub[(size_t) b * kMaxSubPerBlock + j];
                    if (bad_j < 0 && (int) sb.level != cur) {
                        bad_j = j;
                        plan[j] = (uint8_t) cur;
                        for (int q = j + 1; q < kMaxSubPerBlock; q++) plan[q] = (uint8_t) kV4Auto;
                    } else if (bad_j < 0 && exact) {
                        plan[j] = (uint8_t) sb.level;                 // verified: pin it for any later re-parse of this block
                    }
                    cur = incompressible(sb) ? 0 : level;
                }
                if (bad_j >= 0) {
                    c->h_active[b] = 1; nbad++; reparsed++;
                    if (first_bad < 0) first_bad = b;
                    exact = false;                                    // everything behind rests on an assumption
                }
            }
            R[r].first_dirty = first_bad < 0 ? R[r].b1 : first_bad;
            R[r].level_out = cur;
        }
        if (nbad == 0) break;
        if (pass > nb * kMaxSubPerBlock + 4) return fail(ZLB_E_CUDA, "internal: level-feedback replay did not converge");
    }
    if (mode == MODE_SPECULATE) { c->spec_reparsed = reparsed; return ZLB_OK; }

    // layout of the framed streams: per sub-block 1 + 12 + olen bytes, one stop byte per block; streams back to back
    unsigned long long off = 0, ntok = 0, nsub_total = 0;
    for (int r = 0; r < nr; r++) {
        R[r].out_off = off;
        for (int b = R[r].b0; b < R[r].b1; b++) {
            for (int j = 0; j < (int) c->h_nsub[b]; j++) {
                c->h_outoff[(size_t) b * kMaxSubPerBlock + j] = off;
                off += 13ull + c->h_sub[(size_t) b * kMaxSubPerBlock + j].olen;
            }
            off += 1;
            ntok += c->h_ntok[b]; nsub_total += c->h_nsub[b];
        }
        R[r].out_len = off - R[r].out_off;
    }
    if (off > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_encode_blocks: output buffer too small");
    if (nb > 0) {
        CU(cudaMemcpyAsync(c->d_outoff, c->h_outoff, (size_t) nb * kMaxSubPerBlock * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev[EV_PACK0], st));
        zl_huff_pack_kernel<<<dim3(kMaxSubPerBlock, nb), 512, 0, st>>>(c->d_tok, c->d_sub, c->d_nsub, c->d_tab, c->d_outoff, d_out);
        CU(cudaEventRecord(c->ev[EV_PACK1], st));
        CU(cudaGetLastError());
        launches += 1;
    }
    *out_len = (size_t) off;
    c->stats.ms_parse = ms_parse; c->stats.ms_mtf = ms_mtf; c->stats.ms_huff_build = ms_build;
    c->stats.launches = launches; c->stats.parse_launches = parse_launches; c->stats.reparsed_blocks = reparsed;
    c->stats.tokens = ntok; c->stats.subblocks = nsub_total;
    c->stats.slow_main = c->stats.slow_lazy = c->stats.window_hits = c->stats.windows = 0;
    c->stats.cyc_spec = c->stats.cyc_resolve = c->stats.general_path = 0;
    c->stats.cyc_total = 0; c->stats.flagged = 0;
    c->stats.rounds = 0; c->stats.cyc_final = c->stats.cyc_orbit = c->stats.cyc_rank = c->stats.cyc_decide = 0;
    {
        CU(cudaMemcpyAsync(&c->h_v4c, c->d_v4c, sizeof(V4Counters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        c->stats.windows = c->h_v4c.windows; c->stats.rounds = c->h_v4c.rounds;
        c->stats.cyc_spec = c->h_v4c.cyc_spec; c->stats.cyc_resolve = c->h_v4c.cyc_rounds; c->stats.cyc_final = c->h_v4c.cyc_final;
        c->stats.cyc_total = c->h_v4c.cyc_total; c->stats.cyc_orbit = c->h_v4c.cyc_orbit; c->stats.cyc_rank = c->h_v4c.cyc_rank;
        c->stats.cyc_decide = c->h_v4c.cyc_decide;
        if (getenv("ZLB_V4_TRACE")) {
            fprintf(stderr, "v4 cluster size %d (GPU holds", c->stats_cluster);
            for (int k = 1; k <= 4; k++) fprintf(stderr, " %d x %d", c->max_clusters[level < 0 ? 0 : (level > 4 ? 4 : level)][k], 1 << k);
            fprintf(stderr, ")\n");
            fprintf(stderr, "v4 phases (cycles per window, summed over blocks / windows):");
            for (int i = 0; i < 40; i++) fprintf(stderr, " ph%d=%.0f", i, (double) c->h_v4c.ph[i] / (double) (c->h_v4c.windows ? c->h_v4c.windows : 1));
            fprintf(stderr, "\n");
#if defined(ZL_V4_PROFILE)
            {
                unsigned long long gp[16];
                if (cudaMemcpyFromSymbol(gp, zl::g_v4prof, sizeof gp) == cudaSuccess) {
                    fprintf(stderr, "v4 probe profile (totals):");
                    for (int i = 0; i < 16; i++) fprintf(stderr, " g%d=%llu", i, gp[i]);
                    fprintf(stderr, "\n");
                    memset(gp, 0, sizeof gp);
                    cudaMemcpyToSymbol(zl::g_v4prof, gp, sizeof gp);
                }
            }
#endif
        }
    }
    return ZLB_OK;
}

// the stream API: ONE stream, n bytes = the next blocks of it
static int encode_device(zlb_encoder* e, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t out_cap, size_t* out_len, int* level_after, int mode = MODE_ALL, const uint8_t* pre_tail = nullptr) {
    zlb_ctx* c = e->ctx;
    const int nb = (int) ((n + kBlockBytes - 1) / kBlockBytes);
    for (int b = 0; b < nb; b++) c->h_ilen[b] = (uint32_t) (b + 1 < nb ? (size_t) kBlockBytes : n - (size_t) b * kBlockBytes);
    EncRange R;
    R.b0 = 0; R.b1 = nb; R.level_in = e->cur_level; R.level_out = e->cur_level;
    R.st_in = e->d_state[e->cur]; R.st_out = e->d_state[e->cur ^ 1]; R.first_dirty = 0; R.out_off = R.out_len = 0;
    const int rc = encode_ranges(c, e->level, d_in, nb, &R, 1, d_out, out_cap, out_len, mode, pre_tail);
    if (rc) return rc;
    *level_after = R.level_out;       // the caller commits (state ping-pong + level) once the call has succeeded
    return ZLB_OK;
}

static int check_encode_args(zlb_encoder* e, const void* in, size_t n, const void* out, size_t* out_len) {
    if (!e || !out_len || (n && (!in || !out))) return fail(ZLB_E_ARG, "zlb_encode_blocks: null argument");
    if (e->ctx->pending) return fail(ZLB_E_ARG, "zlb_encode_blocks: a submitted range is pending on this context (call zlb_encode_complete first)");
    if (n > (size_t) e->ctx->max_blocks * kBlockBytes) return fail(ZLB_E_ARG, "zlb_encode_blocks: more than max_blocks blocks in one call");
    return ZLB_OK;
}

static void finish_stats(zlb_ctx* c, bool with_pack) {
    float t = 0;
    c->stats.ms_pack = 0; c->stats.ms_total = 0; c->stats.ms_h2d = 0; c->stats.ms_d2h = 0;
    if (with_pack && cudaEventElapsedTime(&t, c->ev[EV_PACK0], c->ev[EV_PACK1]) == cudaSuccess) c->stats.ms_pack = t;
    if (cudaEventElapsedTime(&t, c->ev[EV_START], c->ev[EV_END]) == cudaSuccess) c->stats.ms_total = t;
    if (cudaEventElapsedTime(&t, c->ev[EV_START], c->ev[EV_H2D]) == cudaSuccess) c->stats.ms_h2d = t;
    cudaGetLastError();
}

int zlb_encode_blocks_device(zlb_encoder* e, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t out_cap, size_t* out_len) {
    int rc = check_encode_args(e, d_in, n, d_out, out_len);
    if (rc) return rc;
    zlb_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    *out_len = 0;
    if (n == 0) return ZLB_OK;
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    CU(cudaEventRecord(c->ev[EV_H2D], c->stream));
    int level_after = e->cur_level;
    rc = encode_device(e, d_in, n, d_out, out_cap, out_len, &level_after);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[EV_END], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    finish_stats(c, true);
    e->cur ^= 1; e->cur_level = level_after;
    return ZLB_OK;
}

int zlb_encode_blocks(zlb_encoder* e, const uint8_t* in, size_t n, uint8_t* out, size_t out_cap, size_t* out_len) {
    int rc = check_encode_args(e, in, n, out, out_len);
    if (rc) return rc;
    zlb_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    *out_len = 0;
    if (n == 0) return ZLB_OK;                                    // empty input => empty stream (libzling.cpp:187)
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    CU(cudaMemcpyAsync(c->d_in, in, n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_in + n, 0, 64, c->stream));
    CU(cudaEventRecord(c->ev[EV_H2D], c->stream));
    size_t produced = 0;
    int level_after = e->cur_level;
    rc = encode_device(e, c->d_in, n, c->d_out, c->out_cap, &produced, &level_after);
    if (rc) return rc;
    if (produced > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_encode_blocks: output buffer too small");
    CU(cudaMemcpyAsync(out, c->d_out, produced, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev[EV_END], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    finish_stats(c, true);
    { float t = 0; if (cudaEventElapsedTime(&t, c->ev[EV_PACK1], c->ev[EV_END]) == cudaSuccess) c->stats.ms_d2h = t; cudaGetLastError(); }
    *out_len = produced;
    e->cur ^= 1; e->cur_level = level_after;
    return ZLB_OK;
}

// ------------------------------------------------------------------------------------------ batch: many streams per call
// Independent streams (each a whole stream: fresh MTF tables, current_level = level) share ONE pass of the pipeline: their
// blocks are parsed side by side (one CTA per 16 MiB block: this is what fills the 148 SMs), the MTF rank pass runs one chain
// per (context, stream), Huffman build / pack run per sub-block as always.  The bytes of every stream equal Encode() on it alone.
int zlb_encode_batch(zlb_ctx* c, int level, zlb_stream_io* sv, int nstreams) {
    if (!c || !sv || nstreams < 1) return fail(ZLB_E_ARG, "zlb_encode_batch: bad argument");
    if (level < 0 || level > 4) return fail(ZLB_E_ARG, "zlb_encode_batch: level must be 0..4");
    if (c->pending) return fail(ZLB_E_ARG, "zlb_encode_batch: a submitted range is pending on this context");
    CU(cudaSetDevice(c->device));
    size_t nb_total = 0;
    for (int i = 0; i < nstreams; i++) {
        sv[i].out_len = 0;
        if (sv[i].n && (!sv[i].in || !sv[i].out)) return fail(ZLB_E_ARG, "zlb_encode_batch: null stream buffer");
        nb_total += (sv[i].n + kBlockBytes - 1) / kBlockBytes;
    }
    if (nb_total > (size_t) c->max_blocks) return fail(ZLB_E_ARG, "zlb_encode_batch: the streams hold more than max_blocks blocks");
    if (nb_total == 0) return ZLB_OK;
    if (c->bstate_streams < nstreams) {                                  // initial tables (shared, read-only) + one scratch table per stream
        cudaFree(c->d_bstate); c->d_bstate = nullptr; c->bstate_streams = 0;
        CU(cudaMalloc(&c->d_bstate, ((size_t) nstreams + 1) * 65536));
        c->bstate_streams = nstreams;
        uint8_t init[65536];
        for (int ctx = 0; ctx < 256; ctx++) memcpy(init + ctx * 256, kMtfInit, 256);          // lz.cpp:106-111
        CU(cudaMemcpy(c->d_bstate, init, 65536, cudaMemcpyHostToDevice));
    }
    cudaStream_t st = c->stream;
    std::vector<EncRange> R;
    CU(cudaEventRecord(c->ev[EV_START], st));
    int b = 0;
    for (int i = 0; i < nstreams; i++) {
        const int nbi = (int) ((sv[i].n + kBlockBytes - 1) / kBlockBytes);
        if (nbi == 0) continue;
        for (int q = 0; q < nbi; q++) c->h_ilen[b + q] = (uint32_t) (q + 1 < nbi ? (size_t) kBlockBytes : sv[i].n - (size_t) q * kBlockBytes);
        uint8_t* dst = c->d_in + (size_t) b * kBlockBytes;
        CU(cudaMemcpyAsync(dst, sv[i].in, sv[i].n, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(dst + sv[i].n, 0, 64, st));
        EncRange r;
        r.b0 = b; r.b1 = b + nbi; r.level_in = level; r.level_out = level; r.st_in = c->d_bstate; r.st_out = c->d_bstate + ((size_t) i + 1) * 65536;
        r.first_dirty = b; r.out_off = r.out_len = 0;
        R.push_back(r);
        b += nbi;
    }
    CU(cudaEventRecord(c->ev[EV_H2D], st));
    size_t produced = 0;
    const int rc = encode_ranges(c, level, c->d_in, b, R.data(), (int) R.size(), c->d_out, c->out_cap, &produced, MODE_ALL);
    if (rc) return rc;
    size_t ri = 0;
    for (int i = 0; i < nstreams; i++) {
        if (sv[i].n == 0) continue;
        const EncRange& r = R[ri++];
        if (r.out_len > sv[i].out_cap) return fail(ZLB_E_OVERFLOW, "zlb_encode_batch: a stream's output buffer is too small");
        CU(cudaMemcpyAsync(sv[i].out, c->d_out + r.out_off, (size_t) r.out_len, cudaMemcpyDeviceToHost, st));
        sv[i].out_len = (size_t) r.out_len;
    }
    CU(cudaEventRecord(c->ev[EV_END], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    finish_stats(c, true);
    return ZLB_OK;
}

// split form of zlb_encode_blocks for one stream sharded over several GPUs: the parse of a block range does not
// depend on the MTF state carried from the previous range, so it is launched first and the carried state is
// installed (zlb_encoder_set_state) while it runs
int zlb_encode_submit(zlb_encoder* e, const uint8_t* in, size_t n) {
    if (!e || (n && !in)) return fail(ZLB_E_ARG, "zlb_encode_submit: null argument");
    if (n == 0 || n > (size_t) e->ctx->max_blocks * kBlockBytes) return fail(ZLB_E_ARG, "zlb_encode_submit: size must be 1..max_blocks blocks");
    if (e->submitted_n || e->ctx->pending) return fail(ZLB_E_ARG, "zlb_encode_submit: a submit is already pending on this context");
    zlb_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    CU(cudaMemcpyAsync(c->d_in, in, n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_in + n, 0, 64, c->stream));
    CU(cudaEventRecord(c->ev[EV_H2D], c->stream));
    size_t produced = 0;
    int level_after = e->cur_level;
    const int rc = encode_device(e, c->d_in, n, c->d_out, c->out_cap, &produced, &level_after, MODE_SUBMIT);
    if (rc) return rc;
    e->submitted_n = n;
    c->pending = e;
    return ZLB_OK;
}
int zlb_encode_complete(zlb_encoder* e, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!e || !out || !out_len) return fail(ZLB_E_ARG, "zlb_encode_complete: null argument");
    if (!e->submitted_n) return fail(ZLB_E_ARG, "zlb_encode_complete: nothing was submitted");
    zlb_ctx* c = e->ctx;
    CU(cudaSetDevice(c->device));
    const size_t n = e->submitted_n;
    e->submitted_n = 0;
    c->pending = nullptr;
    *out_len = 0;
    size_t produced = 0;
    int level_after = e->cur_level;
    const int rc = encode_device(e, c->d_in, n, c->d_out, c->out_cap, &produced, &level_after, MODE_COMPLETE);
    if (rc) return rc;
    if (produced > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_encode_complete: output buffer too small");
    CU(cudaMemcpyAsync(out, c->d_out, produced, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev[EV_END], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    finish_stats(c, true);
    *out_len = produced;
    e->cur ^= 1; e->cur_level = level_after;
    return ZLB_OK;
}

// ------------------------------------------------------------------------------------------ multi-GPU
static constexpr size_t kTailBytes = 65536 + 16;
struct zlb_comm {
    zlb_ctx* ctx = nullptr;
    ncclComm_t nccl = nullptr;
    int rank = 0, world = 1;
    cudaStream_t stream = nullptr;          // exchange stream: the carry can arrive while the parse kernel runs on ctx->stream
    uint8_t* d_carry = nullptr;             // ZLB_STATE_BYTES (MTF tables + int32 level), padded
    uint8_t* d_tail = nullptr;              // [2][kTailBytes]: the 64 KiB before this rank's range + u32 valid flag (received), the same for the next rank (sent)
    unsigned long long* d_sizes = nullptr;  // [world + 1]: [world] = this rank's size (all-gather input)
    unsigned long long* h_sizes = nullptr;  // pinned mirror, [world + 1]
    int32_t* h_level = nullptr;             // pinned
    uint8_t* d_all = nullptr;               // rank 0: the gathered stream
    size_t all_cap = 0;
    cudaEvent_t ev[6] = {};
    zlb_shard_stats stats = {};
};
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString ? nccl_api().GetErrorString(r_) : "?", __FILE__, __LINE__); return ZLB_E_NCCL; } } while (0)

int zlb_comm_get_unique_id(uint8_t* id) {
    if (!id) return fail(ZLB_E_ARG, "zlb_comm_get_unique_id: null argument");
    NcclApi& n = nccl_api();
    if (!n.lib) return fail(ZLB_E_NCCL, "zlb_comm_get_unique_id: %s", n.error ? n.error : "NCCL unavailable");
    static_assert(sizeof(ncclUniqueId) == ZLB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NC(n.GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return ZLB_OK;
}
void zlb_comm_destroy(zlb_comm* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->nccl && nccl_api().CommDestroy) nccl_api().CommDestroy(m->nccl);
    cudaFree(m->d_carry); cudaFree(m->d_sizes); cudaFree(m->d_all); cudaFree(m->d_tail);
    if (m->h_sizes) cudaFreeHost(m->h_sizes);
    if (m->h_level) cudaFreeHost(m->h_level);
    for (auto& e : m->ev) if (e) cudaEventDestroy(e);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}
zlb_comm* zlb_comm_create(zlb_ctx* c, int rank, int world, const uint8_t* id) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) { fail(ZLB_E_ARG, "zlb_comm_create: bad argument"); return nullptr; }
    NcclApi& n = nccl_api();
    if (!n.lib) { fail(ZLB_E_NCCL, "zlb_comm_create: %s", n.error ? n.error : "NCCL unavailable"); return nullptr; }
    if (cudaSetDevice(c->device) != cudaSuccess) { fail(ZLB_E_CUDA, "zlb_comm_create: cudaSetDevice failed"); return nullptr; }
    zlb_comm* m = new (std::nothrow) zlb_comm();
    if (!m) { fail(ZLB_E_NOMEM, "zlb_comm_create: out of host memory"); return nullptr; }
    m->ctx = c; m->rank = rank; m->world = world;
    ncclUniqueId u; memcpy(&u, id, sizeof u);
    bool ok = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&m->d_carry, ZLB_STATE_BYTES + 60) == cudaSuccess &&
              cudaMalloc(&m->d_tail, 2 * kTailBytes) == cudaSuccess &&
              cudaMalloc(&m->d_sizes, ((size_t) world + 1) * 8) == cudaSuccess &&
              cudaHostAlloc(&m->h_sizes, ((size_t) world + 1) * 8, cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc(&m->h_level, 64, cudaHostAllocDefault) == cudaSuccess;
    for (auto& e : m->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    if (!ok) { fail(ZLB_E_CUDA, "zlb_comm_create: device allocation failed"); cudaGetLastError(); zlb_comm_destroy(m); return nullptr; }
    const ncclResult_t r = n.CommInitRank(&m->nccl, world, u, rank);
    if (r != ncclSuccess) { fail(ZLB_E_NCCL, "zlb_comm_create: ncclCommInitRank failed: %s", n.GetErrorString(r)); m->nccl = nullptr; zlb_comm_destroy(m); return nullptr; }
    return m;
}
int zlb_comm_get_stats(const zlb_comm* m, zlb_shard_stats* out) {
    if (!m || !out) return fail(ZLB_E_ARG, "zlb_comm_get_stats: null argument");
    *out = m->stats;
    return ZLB_OK;
}

// all ranks: sizes by ONE all-gather of u64, payloads by ONE group of sends / receives; rank 0 ends up with every
// rank's bytes back to back in m->d_all (block order = rank order) and the sizes in m->h_sizes[0..world)
static int gather_packed(zlb_comm* m, const uint8_t* d_local, size_t n_local, size_t* total) {
    NcclApi& n = nccl_api();
    const int W = m->world;
    m->h_sizes[W] = (unsigned long long) n_local;
    CU(cudaMemcpyAsync(m->d_sizes + W, m->h_sizes + W, 8, cudaMemcpyHostToDevice, m->stream));
    NC(n.AllGather(m->d_sizes + W, m->d_sizes, 1, ncclUint64, m->nccl, m->stream));
    CU(cudaMemcpyAsync(m->h_sizes, m->d_sizes, (size_t) W * 8, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    size_t sum = 0;
    for (int r = 0; r < W; r++) sum += (size_t) m->h_sizes[r];
    *total = sum;
    if (m->rank == 0) {
        if (sum + 256 > m->all_cap) {
            cudaFree(m->d_all); m->d_all = nullptr; m->all_cap = 0;
            CU(cudaMalloc(&m->d_all, sum + sum / 8 + 4096));
            m->all_cap = sum + sum / 8 + 4096;
        }
        if (n_local) CU(cudaMemcpyAsync(m->d_all, d_local, n_local, cudaMemcpyDeviceToDevice, m->stream));
    }
    if (W > 1) {
        NC(n.GroupStart());
        if (m->rank == 0) {
            size_t off = (size_t) m->h_sizes[0];
            for (int r = 1; r < W; r++) {
                if (m->h_sizes[r]) NC(n.Recv(m->d_all + off, (size_t) m->h_sizes[r], ncclUint8, r, m->nccl, m->stream));
                off += (size_t) m->h_sizes[r];
            }
        } else if (n_local) {
            NC(n.Send(d_local, n_local, ncclUint8, 0, m->nccl, m->stream));
        }
        NC(n.GroupEnd());
    }
    return ZLB_OK;
}

int zlb_gather_packed(zlb_comm* m, const uint8_t* d_local, size_t n_local, uint8_t* out, size_t out_cap, size_t* out_len, uint64_t* sizes) {
    if (!m || (n_local && !d_local)) return fail(ZLB_E_ARG, "zlb_gather_packed: null argument");
    CU(cudaSetDevice(m->ctx->device));
    size_t total = 0;
    CU(cudaEventRecord(m->ev[2], m->stream));
    const int rc = gather_packed(m, d_local, n_local, &total);
    if (rc) return rc;
    if (m->rank == 0 && out) {
        if (total > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_gather_packed: output buffer too small");
        if (total) CU(cudaMemcpyAsync(out, m->d_all, total, cudaMemcpyDeviceToHost, m->stream));
    }
    CU(cudaEventRecord(m->ev[3], m->stream));
    CU(cudaStreamSynchronize(m->stream));
    { float t = 0; cudaEventElapsedTime(&t, m->ev[2], m->ev[3]); m->stats.ms_gather = t; }
    if (out_len) *out_len = total;
    if (sizes) for (int r = 0; r < m->world; r++) sizes[r] = m->h_sizes[r];
    return ZLB_OK;
}

int zlb_encode_blocks_gathered(zlb_encoder* e, zlb_comm* m, const uint8_t* in, size_t n, int in_on_device, uint8_t* out, size_t out_cap, size_t* out_len, uint64_t* sizes) {
    if (!e || !m || (n && !in) || e->ctx != m->ctx) return fail(ZLB_E_ARG, "zlb_encode_blocks_gathered: bad argument");
    zlb_ctx* c = e->ctx;
    int rc = check_encode_args(e, in, n, in, &n);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (out_len) *out_len = 0;
    m->stats = zlb_shard_stats{};
    size_t produced = 0;
    if (n) {
        int level_after = e->cur_level;
        CU(cudaEventRecord(c->ev[EV_START], c->stream));
        if (!in_on_device) {
            CU(cudaMemcpyAsync(c->d_in, in, n, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemsetAsync(c->d_in + n, 0, 64, c->stream));
        }
        CU(cudaEventRecord(c->ev[EV_H2D], c->stream));
        rc = encode_device(e, in_on_device ? in : c->d_in, n, c->d_out, c->out_cap, &produced, &level_after);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev[EV_END], c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        finish_stats(c, true);
        e->cur ^= 1; e->cur_level = level_after;
    }
    size_t total = 0;
    CU(cudaEventRecord(m->ev[2], m->stream));
    rc = gather_packed(m, c->d_out, produced, &total);
    if (rc) return rc;
    if (m->rank == 0 && out) {
        if (total > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_encode_blocks_gathered: output buffer too small");
        if (total) CU(cudaMemcpyAsync(out, m->d_all, total, cudaMemcpyDeviceToHost, m->stream));
    }
    CU(cudaEventRecord(m->ev[3], m->stream));
    CU(cudaStreamSynchronize(m->stream));
    { float t = 0; cudaEventElapsedTime(&t, m->ev[2], m->ev[3]); m->stats.ms_gather = t; }
    m->stats.local_bytes = produced; m->stats.total_bytes = total;
    if (out_len) *out_len = m->rank == 0 ? total : produced;
    if (sizes) for (int r = 0; r < m->world; r++) sizes[r] = m->h_sizes[r];
    return ZLB_OK;
}

int zlb_encode_stream_sharded(zlb_encoder* e, zlb_comm* m, const uint8_t* in, size_t n, int in_on_device, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!e || !m || (n && !in) || e->ctx != m->ctx) return fail(ZLB_E_ARG, "zlb_encode_stream_sharded: bad argument");
    zlb_ctx* c = e->ctx;
    if (c->pending) return fail(ZLB_E_ARG, "zlb_encode_stream_sharded: a submitted range is pending on this context");
    if (n > (size_t) c->max_blocks * kBlockBytes) return fail(ZLB_E_ARG, "zlb_encode_stream_sharded: more than max_blocks blocks in this rank's range");
    NcclApi& nc = nccl_api();
    CU(cudaSetDevice(c->device));
    if (out_len) *out_len = 0;
    m->stats = zlb_shard_stats{};
    const uint8_t* d_in = in_on_device ? in : c->d_in;
    size_t produced = 0;
    int level_after = e->cur_level;
    // 1. this rank's range: H2D + parse launch (returns at once; assumes the carried level is the requested one)
    CU(cudaEventRecord(c->ev[EV_START], c->stream));
    if (n) {
        if (!in_on_device) {
            CU(cudaMemcpyAsync(c->d_in, in, n, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemsetAsync(c->d_in + n, 0, 64, c->stream));
        }
        CU(cudaEventRecord(c->ev[EV_H2D], c->stream));
    }
    // 1a. the 64 KiB in front of every range travel to the rank that owns it (the parse predicts the level of the range's first
    // sub-block from them, like it does for every block behind the first of a call); an empty range forwards what it received
    uint8_t* tail_in = m->d_tail, *tail_out = m->d_tail + kTailBytes;
    if (m->world > 1) {
        CU(cudaEventRecord(m->ev[4], c->stream));
        CU(cudaStreamWaitEvent(m->stream, m->ev[4], 0));
        if (m->rank > 0) NC(nc.Recv(tail_in, kTailBytes, ncclUint8, m->rank - 1, m->nccl, m->stream));
        else CU(cudaMemsetAsync(tail_in, 0, kTailBytes, m->stream));
        if (m->rank + 1 < m->world) {
            if (n >= 65536) {
                CU(cudaMemcpyAsync(tail_out, d_in + n - 65536, 65536, cudaMemcpyDeviceToDevice, m->stream));
                CU(cudaMemsetAsync(tail_out + 65536, 1, 1, m->stream));
            } else if (n == 0) {
                CU(cudaMemcpyAsync(tail_out, tail_in, kTailBytes, cudaMemcpyDeviceToDevice, m->stream));
            } else {
                CU(cudaMemsetAsync(tail_out, 0, kTailBytes, m->stream));
            }
            NC(nc.Send(tail_out, kTailBytes, ncclUint8, m->rank + 1, m->nccl, m->stream));
        }
        CU(cudaEventRecord(m->ev[5], m->stream));
        CU(cudaStreamWaitEvent(c->stream, m->ev[5], 0));
    }
    const uint8_t* pre_tail = (m->world > 1 && m->rank > 0) ? tail_in : nullptr;
    if (n) {
        const int rc = encode_device(e, d_in, n, c->d_out, c->out_cap, &produced, &level_after, MODE_SUBMIT, pre_tail);
        if (rc) return rc;
    }
    // 1b. ranks behind the first: finish the level feedback of the range BEFORE the carried state is there.  The levels depend
    // on the Huffman sizes of the sub-blocks (libzling.cpp:261-266), those on the MTF ranks, those on the carried tables — but
    // only through the first few thousand literals of the range, so ranks computed from this encoder's current tables give
    // the same levels except on a knife's edge.  Mispredicted blocks are re-parsed now, concurrently on all ranks; the exact
    // pass below (step 3) verifies everything against the real carried state and repairs what is left.
    c->spec_reparsed = 0;
    if (n && m->rank > 0) {
        const int rc = encode_device(e, d_in, n, c->d_out, c->out_cap, &produced, &level_after, MODE_SPECULATE, pre_tail);
        if (rc) return rc;
    }
    m->stats.spec_reparsed_blocks = c->spec_reparsed;
    // 2. the carried state of the previous range, GPU -> GPU, while the parse runs
    CU(cudaEventRecord(m->ev[0], m->stream));
    if (m->rank > 0) {
        NC(nc.Recv(m->d_carry, ZLB_STATE_BYTES, ncclUint8, m->rank - 1, m->nccl, m->stream));
        CU(cudaMemcpyAsync(e->d_state[e->cur], m->d_carry, 65536, cudaMemcpyDeviceToDevice, m->stream));
        CU(cudaMemcpyAsync(m->h_level, m->d_carry + 65536, 4, cudaMemcpyDeviceToHost, m->stream));
        CU(cudaEventRecord(m->ev[1], m->stream));
        CU(cudaStreamSynchronize(m->stream));
        if (*m->h_level < 0 || *m->h_level > 4) return fail(ZLB_E_NCCL, "zlb_encode_stream_sharded: received a corrupt carried state");
        e->cur_level = *m->h_level;
        { float t = 0; cudaEventElapsedTime(&t, m->ev[0], m->ev[1]); m->stats.ms_wait_carry = t; }
    }
    // 3. MTF ranks + Huffman + framing of the range (re-parses its first block only if the carried level differs)
    if (n) {
        const int rc = encode_device(e, d_in, n, c->d_out, c->out_cap, &produced, &level_after, MODE_COMPLETE, pre_tail);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev[EV_END], c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        finish_stats(c, true);
        e->cur ^= 1; e->cur_level = level_after;
    }
    // 4. forward this range's final state
    if (m->rank + 1 < m->world) {
        *m->h_level = (int32_t) e->cur_level;
        CU(cudaMemcpyAsync(m->d_carry, e->d_state[e->cur], 65536, cudaMemcpyDeviceToDevice, m->stream));
        CU(cudaMemcpyAsync(m->d_carry + 65536, m->h_level, 4, cudaMemcpyHostToDevice, m->stream));
        NC(nc.Send(m->d_carry, ZLB_STATE_BYTES, ncclUint8, m->rank + 1, m->nccl, m->stream));
    }
    // 5. ONE gather of the framed ranges
    size_t total = 0;
    CU(cudaEventRecord(m->ev[2], m->stream));
    const int rc = gather_packed(m, c->d_out, produced, &total);
    if (rc) return rc;
    if (m->rank == 0 && out) {
        if (total > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_encode_stream_sharded: output buffer too small");
        if (total) CU(cudaMemcpyAsync(out, m->d_all, total, cudaMemcpyDeviceToHost, m->stream));
    }
    CU(cudaEventRecord(m->ev[3], m->stream));
    CU(cudaStreamSynchronize(m->stream));
    { float t = 0; cudaEventElapsedTime(&t, m->ev[2], m->ev[3]); m->stats.ms_gather = t; }
    m->stats.local_bytes = produced; m->stats.total_bytes = total;
    if (out_len) *out_len = m->rank == 0 ? total : produced;
    return ZLB_OK;
}

// ------------------------------------------------------------------------------------------ debug / tests
int zlb_debug_tokens(zlb_ctx* c, int blk, uint32_t* tok, size_t cap, size_t* ntok) {
    if (!c || !ntok || blk < 0 || blk >= c->last_nblocks) return fail(ZLB_E_ARG, "zlb_debug_tokens: bad argument");
    CU(cudaSetDevice(c->device));
    const size_t n = c->h_ntok[blk];
    *ntok = n;
    if (tok) CU(cudaMemcpy(tok, c->d_tok + (size_t) blk * kTokStride, (n < cap ? n : cap) * 4, cudaMemcpyDeviceToHost));
    return ZLB_OK;
}
int zlb_debug_subblocks(zlb_ctx* c, int blk, zlb_subblock* sub, size_t cap, size_t* nsub) {
    if (!c || !nsub || blk < 0 || blk >= c->last_nblocks) return fail(ZLB_E_ARG, "zlb_debug_subblocks: bad argument");
    static_assert(sizeof(zlb_subblock) == sizeof(SubBlock), "ABI mirror out of sync");
    const size_t n = c->h_nsub[blk];
    *nsub = n;
    if (sub) memcpy(sub, c->h_sub + (size_t) blk * kMaxSubPerBlock, (n < cap ? n : cap) * sizeof(SubBlock));
    return ZLB_OK;
}
int zlb_debug_huff_tables(zlb_ctx* c, const uint32_t* freq, int ntables, int nsym, int cap, uint8_t* len_out, uint16_t* code_out) {
    if (!c || !freq || !len_out || !code_out || ntables < 1 || nsym < 1 || nsym > kSyms1 || cap < 1 || cap > 15)
        return fail(ZLB_E_ARG, "zlb_debug_huff_tables: bad argument");
    CU(cudaSetDevice(c->device));
    uint32_t* d_f = nullptr; uint8_t* d_l = nullptr; uint16_t* d_c = nullptr;
    const size_t cnt = (size_t) ntables * nsym;
    struct Free { uint32_t*& f; uint8_t*& l; uint16_t*& c; ~Free() { cudaFree(f); cudaFree(l); cudaFree(c); } } guard{ d_f, d_l, d_c };
    CU(cudaMalloc(&d_f, cnt * 4)); CU(cudaMalloc(&d_l, cnt)); CU(cudaMalloc(&d_c, cnt * 2));
    CU(cudaMemcpy(d_f, freq, cnt * 4, cudaMemcpyHostToDevice));
    zl_huff_tables_only_kernel<<<ntables, 256, 0, c->stream>>>(d_f, nsym, cap, d_l, d_c);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaMemcpy(len_out, d_l, cnt, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(code_out, d_c, cnt * 2, cudaMemcpyDeviceToHost));
    return ZLB_OK;
}

// ------------------------------------------------------------------------------------------ decoder
zlb_decoder* zlb_decoder_begin(zlb_ctx* c) {
    if (!c) { fail(ZLB_E_ARG, "zlb_decoder_begin: null context"); return nullptr; }
    cudaSetDevice(c->device);
    zlb_decoder* d = new (std::nothrow) zlb_decoder();
    if (!d) { fail(ZLB_E_NOMEM, "zlb_decoder_begin: out of host memory"); return nullptr; }
    d->ctx = c; d->d_state = nullptr;
    uint8_t init[65536];
    for (int ctx = 0; ctx < 256; ctx++) memcpy(init + ctx * 256, kMtfInit, 256);          // lz.cpp:119-121
    if (cudaMalloc(&d->d_state, 65536) != cudaSuccess || cudaMemcpy(d->d_state, init, 65536, cudaMemcpyHostToDevice) != cudaSuccess) {
        fail(ZLB_E_CUDA, "zlb_decoder_begin: device allocation failed");
        if (d->d_state) cudaFree(d->d_state);
        delete d; return nullptr;
    }
    return d;
}
void zlb_decoder_end(zlb_decoder* d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaFree(d->d_state);
    delete d;
}

// walk the container of ONE stream on the host (flag / BE32 x3 / payload, libzling.cpp:313-332): appends a DecSub per framed
// sub-block (payload offsets relative to comp_base + the stream's own offset), stops after max_blocks complete blocks
static int walk_container(const uint8_t* in, size_t n, size_t comp_base, int block_base, int max_blocks, std::vector<DecSub>& subs, size_t* used, int* nblocks_out) {
    size_t at = 0;
    int nblocks = 0;
    while (at < n && nblocks < max_blocks) {
        uint32_t sym_off = 0;
        size_t p = at;
        bool closed = false;
        const size_t first_sub = subs.size();
        while (p < n) {
            const int flag = in[p++];
            if (flag != 0 && flag != 1) return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): invalid encflag.");   // :315-317
            if (flag == 0) { closed = true; break; }
            if (p + 12 > n) break;
            auto be32 = [&](size_t q) { return (uint32_t) in[q] << 24 | (uint32_t) in[q + 1] << 16 | (uint32_t) in[q + 2] << 8 | (uint32_t) in[q + 3]; };
            DecSub s;
            s.encpos = be32(p); s.rlen = be32(p + 4); s.olen = be32(p + 8); p += 12;
            if (s.rlen > (uint32_t) kSubSymbols || s.olen > (uint32_t) kSubBytesMax)
                return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): invalid block size.");                        // :326-328
            if (s.olen < (uint32_t) kTableBytes || s.encpos > (uint32_t) kBlockBytes)
                return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): invalid block size.");
            if (p + s.olen > n) { p = n + 1; break; }
            s.payload_off = comp_base + p; s.block = (uint32_t) (block_base + nblocks); s.sym_off = sym_off; s.pad = 0;
            if ((size_t) sym_off + s.rlen > 2 * kTokStride) return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): lzdecode failed.");
            sym_off += s.rlen;
            p += s.olen;
            subs.push_back(s);
        }
        if (!closed) { subs.resize(first_sub); break; }          // incomplete block: leave it for the next call
        at = p; nblocks++;
    }
    *used = at; *nblocks_out = nblocks;
    return ZLB_OK;
}

// decode the sub-blocks of `nr` streams (records in `subs`, ranges in c->h_drng) whose compressed bytes are already in c->d_comp
static int decode_ranges(zlb_ctx* c, const std::vector<DecSub>& subs, int nr, int nblocks) {
    cudaStream_t st = c->stream;
    const int ns = (int) subs.size();
    if ((size_t) ns > c->decsub_cap) return fail(ZLB_E_FORMAT, "zlb_decode_blocks: too many sub-blocks in one call");
    if (nr > c->decring_streams) {                                        // one 4 MiB offset ring per stream decoded concurrently
        cudaFree(c->d_decring); c->d_decring = nullptr; c->decring_streams = 0;
        CU(cudaMalloc(&c->d_decring, (size_t) nr * 256 * kRing * sizeof(uint32_t)));
        c->decring_streams = nr;
    }
    if (ns) CU(cudaMemcpyAsync(c->d_decsub, subs.data(), (size_t) ns * sizeof(DecSub), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_drng, c->h_drng, (size_t) nr * sizeof(DecRange), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->d_status, 0, ((size_t) ns + (size_t) nr + 4) * sizeof(int), st));
    CU(cudaMemsetAsync(c->d_nsub, 0, (size_t) nblocks * 4, st));
    CU(cudaEventRecord(c->ev[EV_H2D], st));
    uint16_t* d_sym = reinterpret_cast<uint16_t*>(c->d_tok);
    if (ns) {
        zl_huff_decode_kernel<<<ns, 256, 65536, st>>>(c->d_comp, c->d_decsub, d_sym, c->d_status);
        zl_rolz_decode_kernel<<<nr, 32, 65536, st>>>(c->d_decsub, c->d_drng, d_sym, c->d_in, c->d_decring, c->d_nsub, c->d_status + ns);
    }
    CU(cudaMemcpyAsync(c->h_status, c->d_status, ((size_t) ns + (size_t) nr + 4) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_nsub, c->d_nsub, (size_t) nblocks * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    for (int i = 0; i < ns; i++) {
        if (c->h_status[i] == 1) return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): invalid huffman stream. (bad code1)");   // :382
        if (c->h_status[i] == 2) return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): invalid huffman stream. (bad code2)");   // :392
        if (c->h_status[i] == 3) return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): invalid huffman stream. (bad ex-bits)"); // :399
    }
    for (int r = 0; r < nr; r++)
        if (c->h_status[ns + r] != 0) return fail(ZLB_E_FORMAT, "baidu::zling::Decode(): lzdecode failed.");                 // :407
    c->stats.launches = ns ? 2 : 0; c->stats.subblocks = (uint64_t) ns;
    return ZLB_OK;
}

int zlb_decode_blocks(zlb_decoder* d, const uint8_t* in, size_t n, size_t* consumed, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!d || !consumed || !out_len || (n && !in)) return fail(ZLB_E_ARG, "zlb_decode_blocks: null argument");
    zlb_ctx* c = d->ctx;
    if (c->pending) return fail(ZLB_E_ARG, "zlb_decode_blocks: a submitted encode range is pending on this context");
    CU(cudaSetDevice(c->device));
    *consumed = 0; *out_len = 0;
    if (n == 0) return ZLB_OK;
    std::vector<DecSub> subs;
    size_t at = 0;
    int nblocks = 0;
    int rc = walk_container(in, n, 0, 0, c->max_blocks, subs, &at, &nblocks);
    if (rc) return rc;
    if (nblocks == 0) return fail(ZLB_E_ARG, "zlb_decode_blocks: input holds no complete block");
    if (at > c->comp_cap) return fail(ZLB_E_ARG, "zlb_decode_blocks: compressed input larger than the context's buffer");
    cudaStream_t st = c->stream;
    CU(cudaEventRecord(c->ev[EV_START], st));
    CU(cudaMemcpyAsync(c->d_comp, in, at, cudaMemcpyHostToDevice, st));
    c->h_drng[0].s0 = 0; c->h_drng[0].s1 = (int) subs.size(); c->h_drng[0].state = d->d_state;
    rc = decode_ranges(c, subs, 1, nblocks);
    if (rc) return rc;
    size_t total = 0;
    for (int b = 0; b < nblocks; b++) total += c->h_nsub[b];
    if (total > out_cap) return fail(ZLB_E_OVERFLOW, "zlb_decode_blocks: output buffer too small");
    size_t w = 0;
    for (int b = 0; b < nblocks; b++) {
        if (c->h_nsub[b]) CU(cudaMemcpyAsync(out + w, c->d_in + (size_t) b * kBlockBytes, c->h_nsub[b], cudaMemcpyDeviceToHost, st));
        w += c->h_nsub[b];
    }
    CU(cudaEventRecord(c->ev[EV_END], st));
    CU(cudaStreamSynchronize(st));
    finish_stats(c, false);
    *consumed = at; *out_len = total;
    return ZLB_OK;
}

// batch: every `in` is a whole framed stream; one decode chain per stream runs concurrently (the way decode scales: a stream
// is one serial chain, src/libzling_lz.cpp:318-376 + the stream-lifetime MTF tables, src/libzling_lz.h:137)
int zlb_decode_batch(zlb_ctx* c, zlb_stream_io* sv, int nstreams) {
    if (!c || !sv || nstreams < 1) return fail(ZLB_E_ARG, "zlb_decode_batch: bad argument");
    if (c->pending) return fail(ZLB_E_ARG, "zlb_decode_batch: a submitted encode range is pending on this context");
    if (nstreams > c->max_blocks) return fail(ZLB_E_ARG, "zlb_decode_batch: more streams than max_blocks");
    CU(cudaSetDevice(c->device));
    std::vector<DecSub> subs;
    std::vector<int> b0(nstreams), nbv(nstreams);
    size_t comp_at = 0;
    int nblocks = 0, nr = 0;
    cudaStream_t st = c->stream;
    if (c->dstate_streams < nstreams) {
        cudaFree(c->d_dstate); c->d_dstate = nullptr; c->dstate_streams = 0;
        CU(cudaMalloc(&c->d_dstate, (size_t) nstreams * 65536));
        c->dstate_streams = nstreams;
    }
    uint8_t init[65536];
    for (int ctx = 0; ctx < 256; ctx++) memcpy(init + ctx * 256, kMtfInit, 256);          // lz.cpp:119-121
    CU(cudaEventRecord(c->ev[EV_START], st));
    for (int i = 0; i < nstreams; i++) {
        sv[i].out_len = 0; b0[i] = nblocks; nbv[i] = 0;
        if (sv[i].n == 0) continue;
        if (!sv[i].in || !sv[i].out) return fail(ZLB_E_ARG, "zlb_decode_batch: null stream buffer");
        const size_t s0 = subs.size();
        size_t used = 0; int nbi = 0;
        const int rc = walk_container(sv[i].in, sv[i].n, comp_at, nblocks, c->max_blocks - nblocks + 1, subs, &used, &nbi);
        if (rc) return rc;
        if (used != sv[i].n) return fail(nblocks + nbi > c->max_blocks ? ZLB_E_ARG : ZLB_E_FORMAT, "zlb_decode_batch: a stream is truncated, or the streams hold more than max_blocks blocks");
        if (comp_at + used > c->comp_cap) return fail(ZLB_E_ARG, "zlb_decode_batch: compressed input larger than the context's buffer");
        CU(cudaMemcpyAsync(c->d_comp + comp_at, sv[i].in, used, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->d_dstate + (size_t) nr * 65536, init, 65536, cudaMemcpyHostToDevice, st));
        c->h_drng[nr].s0 = (int) s0; c->h_drng[nr].s1 = (int) subs.size(); c->h_drng[nr].state = c->d_dstate + (size_t) nr * 65536;
        nr++;
        comp_at += used; nbv[i] = nbi; nblocks += nbi;
    }
    if (nblocks > c->max_blocks) return fail(ZLB_E_ARG, "zlb_decode_batch: the streams hold more than max_blocks blocks");
    if (nr == 0) return ZLB_OK;
    CU(cudaStreamSynchronize(st));                                       // `init` is on the stack: the copies above must be done before it goes away
    const int rc = decode_ranges(c, subs, nr, nblocks);
    if (rc) return rc;
    for (int i = 0; i < nstreams; i++) {
        size_t total = 0;
        for (int b = b0[i]; b < b0[i] + nbv[i]; b++) total += c->h_nsub[b];
        if (total > sv[i].out_cap) return fail(ZLB_E_OVERFLOW, "zlb_decode_batch: a stream's output buffer is too small");
        size_t w = 0;
        for (int b = b0[i]; b < b0[i] + nbv[i]; b++) {
            if (c->h_nsub[b]) CU(cudaMemcpyAsync(sv[i].out + w, c->d_in + (size_t) b * kBlockBytes, c->h_nsub[b], cudaMemcpyDeviceToHost, st));
            w += c->h_nsub[b];
        }
        sv[i].out_len = total;
    }
    CU(cudaEventRecord(c->ev[EV_END], st));
    CU(cudaStreamSynchronize(st));
    finish_stats(c, false);
    return ZLB_OK;
}

}  // extern "C"
