// Host-side mirror of the reference's public API (src/libzling.h:44-45, src/libzling_utils.{h,cpp}) on top of the
// C ABI in include/zlb.h.  This file is the "stream driver": it does what baidu::zling::Encode/Decode do AROUND
// the block pipeline (src/libzling.cpp:174-199,269-291 and :293-332,412-427) — pull bytes from the Inputter
// until a 16 MiB block is full or the input ends, hand whole blocks to the GPU, push the framed result to the
// Outputter, call the ActionHandler in the reference's order — and nothing of the codec itself.
//
// Ordering contract kept: every GetData/PutData/On* call is made from the calling thread; per
// block the Outputter sees the sub-block records, the stop flag, and only then OnProcess(block) is invoked.
// Batching: without an ActionHandler several blocks go to the GPU per call (unobservable); with one, encode still
// batches (the handler only ever sees blocks after their bytes were emitted, in order) but decode proceeds one
// block at a time because a handler may read from the Inputter inside OnProcess (demo/zling.cpp:124-132).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/libzling/libzling.h"
#include "../../include/zlb.h"

namespace baidu {
namespace zling {

// ---- src/libzling_utils.cpp:40-95 ---------------------------------------------------------------------------
int Inputter::GetChar() {
    unsigned char c = 0;
    GetData(&c, 1);
    return c;
}
uint32_t Inputter::GetUInt32() {
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v = (v << 8) | (uint32_t) GetChar();
    return v;
}
int Outputter::PutChar(int v) {
    unsigned char c = (unsigned char) v;
    PutData(&c, 1);
    return c;
}
uint32_t Outputter::PutUInt32(uint32_t v) {
    for (int shift = 24; shift >= 0; shift -= 8) PutChar((int) ((v >> shift) & 0xff));
    return v;
}

size_t FileInputter::GetData(unsigned char* buf, size_t len) {
    const size_t got = fread(buf, 1, len, m_fp);
    m_total_read += got;
    return got;
}
bool FileInputter::IsEnd() {
    const int c = fgetc(m_fp);
    return ungetc(c, m_fp) == EOF;      // peek, as the reference does (libzling_utils.cpp:72-74)
}
bool FileInputter::IsErr() { return ferror(m_fp) != 0; }
size_t FileInputter::GetInputSize() { return m_total_read; }

size_t FileOutputter::PutData(unsigned char* buf, size_t len) {
    const size_t put = fwrite(buf, 1, len, m_fp);
    m_total_write += put;
    return put;
}
bool FileOutputter::IsErr() { return ferror(m_fp) != 0; }
size_t FileOutputter::GetOutputSize() { return m_total_write; }

// ---- GPU contexts: a small pool, one context per call in flight -------------------------------------------------
// The reference's Encode/Decode are re-entrant (every call owns its EncodeResource/DecodeResource,
// src/libzling.cpp:108-163,180,299); so are these: a call takes a context (device buffers + page-locked staging)
// out of the pool for its whole duration and puts it back on every exit path; concurrent calls get different
// contexts.  Two sizes exist: a 1-block context (about 260 MB of device memory, 33 MB page-locked) serves inputs
// of up to 16 MiB, a ZLING_B200_BLOCKS-block one (default 8; about 2 GB / 260 MB) everything longer, so that a
// small input never pays for the large buffers.  ZLING_B200_DEVICE selects the GPU.
namespace {

// ZLING_B200_TRACE=1: wall-clock marks of the driver's phases on stderr (where does a CLI run spend its time?)
void trace(const char* what, size_t n = 0) {
    static const bool on = getenv("ZLING_B200_TRACE") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[zling-b200 %9.1f ms] %s %zu\n", ms, what, n);
}

int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    const long x = strtol(v, &end, 10);
    if (end == v || *end != 0 || x < lo || x > hi) return dflt;          // unparsable or out of range: keep the default
    return (int) x;
}

struct Gpu {
    zlb_ctx* ctx = nullptr;
    unsigned char* pin_in = nullptr;     // page-locked staging: max_blocks * 16 MiB
    unsigned char* pin_in2 = nullptr;    // large tier only: the second input buffer of the streaming pipeline (Encode)
    unsigned char* pin_out = nullptr;    // page-locked staging: zlb_encode_bound of the above
    size_t in_cap = 0, out_cap = 0;
    int max_blocks = 0;
    ~Gpu() {
        zlb_host_free(pin_in); zlb_host_free(pin_in2); zlb_host_free(pin_out);
        zlb_destroy(ctx);
    }
};

class Pool {
public:
    // a context able to hold `blocks` blocks (1 = the small tier, anything else = the large tier)
    Gpu* acquire(bool large) {
        const int want = large ? env_int("ZLING_B200_BLOCKS", 8, 1, 64) : 1;
        {
            std::lock_guard<std::mutex> lock(m_);
            for (size_t i = 0; i < free_.size(); i++) {
                if (free_[i]->max_blocks == want) { Gpu* g = free_[i]; free_.erase(free_.begin() + (long) i); return g; }
            }
        }
        // build everything in a local object and hand it out only when all of it exists
        std::unique_ptr<Gpu> g(new Gpu());
        g->max_blocks = want;
        g->ctx = zlb_create(env_int("ZLING_B200_DEVICE", 0, 0, 1023), want);
        if (!g->ctx) throw std::runtime_error(std::string("libzling (B200): ") + zlb_last_error());
        g->in_cap = (size_t) want * ZLB_BLOCK_BYTES;
        g->out_cap = zlb_encode_bound(g->in_cap);
        g->pin_in = (unsigned char*) zlb_host_alloc(g->in_cap + 64);
        g->pin_out = (unsigned char*) zlb_host_alloc(g->out_cap + 64);
        if (!g->pin_in || !g->pin_out) throw std::bad_alloc();      // (pin_in2 is allocated by the first Encode that needs a second batch)
        return g.release();
    }
    void release(Gpu* g) {
        if (!g) return;
        {
            std::lock_guard<std::mutex> lock(m_);
            if (free_.size() < 4) { free_.push_back(g); return; }        // keep a few warm, free the rest
        }
        delete g;
    }
    ~Pool() { for (Gpu* g : free_) delete g; }
private:
    std::mutex m_;
    std::vector<Gpu*> free_;
};
Pool& pool() { static Pool p; return p; }

struct GpuLease {                        // returns the context to the pool on every exit path, exceptions included
    Gpu* g;
    explicit GpuLease(bool large) : g(pool().acquire(large)) {}
    ~GpuLease() { pool().release(g); }
    void upgrade() { Gpu* big = pool().acquire(true); pool().release(g); g = big; }
    GpuLease(const GpuLease&) = delete;
    GpuLease& operator=(const GpuLease&) = delete;
};

struct EncoderGuard {
    zlb_encoder* e;
    explicit EncoderGuard(zlb_encoder* e_) : e(e_) {}
    ~EncoderGuard() { zlb_encoder_end(e); }
};
struct PendingGuard {                    // a submitted batch must not stay pending on a context that goes back to the pool
    zlb_encoder*& e; Gpu*& g; bool armed = false;
    PendingGuard(zlb_encoder*& e_, Gpu*& g_) : e(e_), g(g_) {}
    ~PendingGuard() { if (armed && e) { size_t n = 0; zlb_encode_complete(e, g->pin_out, g->out_cap, &n); } }
};
struct DecoderGuard {
    zlb_decoder* d;
    explicit DecoderGuard(zlb_decoder* d_) : d(d_) {}
    ~DecoderGuard() { zlb_decoder_end(d); }
};

[[noreturn]] void raise_zlb(int rc) {
    if (rc == ZLB_E_NOMEM) throw std::bad_alloc();
    throw std::runtime_error(zlb_last_error());
}

// push `len` bytes, retrying short writes (libzling.cpp:273-276); false on outputter error
bool put_all(Outputter* out, unsigned char* p, size_t len) {
    size_t off = 0;
    while (!out->IsErr() && off < len) off += out->PutData(p + off, len - off);
    return !out->IsErr();
}

}  // namespace

// ---- src/libzling.cpp:174-291 -------------------------------------------------------------------------------
int Encode(Inputter* inputter, Outputter* outputter, ActionHandler* action_handler, int level) {
    if (level < 0 || level > 4) return -1;       // the reference never terminates here; see libzling.h
    trace("Encode: enter");
    GpuLease lease(false);                       // acquired before OnInit: a call that cannot get a GPU throws without having started
    trace("Encode: context ready");
    if (action_handler) {
        action_handler->SetInputterOutputter(inputter, outputter, true);
        action_handler->OnInit();
    }
    EncoderGuard enc(zlb_encoder_begin(lease.g->ctx, level));
    if (!enc.e) raise_zlb(ZLB_E_CUDA);

    bool io_error = false;
    // pull bytes until the buffer is full or the input ends; every block except the stream's last is exactly 16 MiB (libzling.cpp:193-196)
    auto fill = [&](unsigned char* buf, size_t have, size_t cap) {
        while (have < cap && !inputter->IsEnd() && !inputter->IsErr()) {
            have += inputter->GetData(buf + have, cap - have);
            if (inputter->IsErr()) { io_error = true; break; }
        }
        return have;
    };
    // hand the frames of one batch to the outputter block by block so that OnProcess keeps its place in the order
    auto emit = [&](Gpu& g, unsigned char* in, size_t have) {
        size_t at = 0;
        for (size_t boff = 0; boff < have && !io_error; boff += ZLB_BLOCK_BYTES) {
            const size_t blen = have - boff < ZLB_BLOCK_BYTES ? have - boff : (size_t) ZLB_BLOCK_BYTES;
            size_t end = at;                      // find this block's stop flag by walking its sub-block headers
            while (g.pin_out[end] == 1) {
                const unsigned char* h = g.pin_out + end + 9;
                end += 13 + ((size_t) h[0] << 24 | (size_t) h[1] << 16 | (size_t) h[2] << 8 | (size_t) h[3]);
            }
            end += 1;
            if (!put_all(outputter, g.pin_out + at, end - at)) { io_error = true; break; }
            at = end;
            if (action_handler) action_handler->OnProcess(in + boff, blen);
        }
    };

    size_t have = fill(lease.g->pin_in, 0, lease.g->in_cap);
    trace("Encode: first block read", have);
    if (!io_error && have == lease.g->in_cap && !inputter->IsEnd()) {
        // longer than one block: move to the large context (the encoder holds no stream state yet) and keep reading
        std::vector<unsigned char> first(lease.g->pin_in, lease.g->pin_in + have);
        zlb_encoder_end(enc.e); enc.e = nullptr;
        lease.upgrade();
        enc.e = zlb_encoder_begin(lease.g->ctx, level);
        if (!enc.e) raise_zlb(ZLB_E_CUDA);
        memcpy(lease.g->pin_in, first.data(), have);
        trace("Encode: large context ready");
        have = fill(lease.g->pin_in, have, lease.g->in_cap);
        trace("Encode: first batch read", have);
    }
    // Streaming pipeline (two page-locked input buffers): the parse of a batch is launched without waiting for it
    // (zlb_encode_submit), the next batch is read from the Inputter while it runs, and the frames of a finished batch go to the
    // Outputter while the next one is already on the GPU.  All Inputter / Outputter / handler calls stay on this thread.
    PendingGuard pending(enc.e, lease.g);
    unsigned char* bufs[2] = { lease.g->pin_in, lease.g->pin_in2 };
    int cur = 0;
    if (!io_error && have > 0) {
        int rc = zlb_encode_submit(enc.e, bufs[cur], have);
        if (rc != ZLB_OK) raise_zlb(rc);
        pending.armed = true;
        while (true) {
            Gpu& g = *lease.g;
            size_t have_next = 0;
            if (g.max_blocks > 1 && !inputter->IsEnd() && !inputter->IsErr()) {
                if (!g.pin_in2) {                 // second staging buffer: only streams of more than one batch pay for it
                    g.pin_in2 = (unsigned char*) zlb_host_alloc(g.in_cap + 64);
                    if (!g.pin_in2) throw std::bad_alloc();
                    bufs[1] = g.pin_in2;
                }
                have_next = fill(bufs[cur ^ 1], 0, g.in_cap);
            }
            size_t produced = 0;
            pending.armed = false;
            trace("Encode: next batch read", have_next);
            rc = zlb_encode_complete(enc.e, g.pin_out, g.out_cap, &produced);
            if (rc != ZLB_OK) raise_zlb(rc);
            trace("Encode: batch complete", produced);
            if (!io_error && have_next > 0) {
                rc = zlb_encode_submit(enc.e, bufs[cur ^ 1], have_next);
                if (rc != ZLB_OK) raise_zlb(rc);
                pending.armed = true;
            }
            emit(g, bufs[cur], have);
            trace("Encode: batch written", have);
            if (io_error || have_next == 0) break;
            cur ^= 1;
            have = have_next;
        }
    }
    if (action_handler) action_handler->OnDone();
    return (inputter->IsErr() || outputter->IsErr()) ? -1 : 0;
}

// ---- src/libzling.cpp:293-427 -------------------------------------------------------------------------------
int Decode(Inputter* inputter, Outputter* outputter, ActionHandler* action_handler) {
    GpuLease lease(action_handler == NULL);      // with a handler decode goes block by block: the small context is enough
    if (action_handler) {
        action_handler->SetInputterOutputter(inputter, outputter, false);
        action_handler->OnInit();
    }
    Gpu& g = *lease.g;
    DecoderGuard dec(zlb_decoder_begin(g.ctx));
    if (!dec.d) raise_zlb(ZLB_E_CUDA);
    const int batch = action_handler ? 1 : g.max_blocks;
    unsigned char* comp = g.pin_out;             // compressed bytes are staged in the larger buffer
    bool io_error = false;

    while (!io_error && !inputter->IsEnd()) {
        // read exactly up to the stop flag of `batch` blocks: never consume input past a block end, because a
        // handler may read its own bytes (e.g. a checksum) from the inputter inside OnProcess
        size_t have = 0;
        int blocks = 0;
        while (blocks < batch && !inputter->IsEnd()) {
            bool closed = false;
            while (!inputter->IsEnd()) {
                const int flag = inputter->GetChar();
                if (flag != 0 && flag != 1) throw std::runtime_error("baidu::zling::Decode(): invalid encflag.");   // :315-317
                comp[have++] = (unsigned char) flag;
                if (flag == 0) { closed = true; break; }
                for (int i = 0; i < 12; i++) comp[have + i] = (unsigned char) inputter->GetChar();
                if (inputter->IsErr()) { io_error = true; break; }
                const unsigned char* h = comp + have;
                const uint32_t rlen = (uint32_t) h[4] << 24 | (uint32_t) h[5] << 16 | (uint32_t) h[6] << 8 | h[7];
                const uint32_t olen = (uint32_t) h[8] << 24 | (uint32_t) h[9] << 16 | (uint32_t) h[10] << 8 | h[11];
                have += 12;
                if (rlen > ZLB_SUBBLOCK_SYMBOLS || olen > ZLB_SUBBLOCK_BYTES)
                    throw std::runtime_error("baidu::zling::Decode(): invalid block size.");                        // :326-328
                if (have + olen + 16 > g.out_cap) throw std::runtime_error("baidu::zling::Decode(): invalid block size.");
                size_t off = 0;
                while (!inputter->IsEnd() && off < olen) {                                                          // :329-332
                    off += inputter->GetData(comp + have + off, olen - off);
                    if (inputter->IsErr()) { io_error = true; break; }
                }
                if (io_error) break;
                if (off < olen) throw std::runtime_error("baidu::zling::Decode(): invalid huffman stream. (truncated)");
                have += olen;
            }
            if (io_error) break;
            if (!closed) comp[have++] = 0;       // input ended without a stop flag: the reference still emits the block
            blocks++;
        }
        if (io_error || have == 0) break;
        size_t used = 0, produced = 0;
        const int rc = zlb_decode_blocks(dec.d, comp, have, &used, g.pin_in, g.in_cap, &produced);
        if (rc != ZLB_OK) raise_zlb(rc);
        // blocks come back concatenated; every block but the last of the stream decodes to 16 MiB, but lengths
        // are taken from the frames, not assumed: the last sub-block's encpos is the block length
        size_t cpos = 0, opos = 0;
        for (int b = 0; b < blocks && !io_error; b++) {
            size_t blen = 0;
            while (comp[cpos] == 1) {
                const unsigned char* h = comp + cpos + 1;
                blen = (size_t) h[0] << 24 | (size_t) h[1] << 16 | (size_t) h[2] << 8 | h[3];
                cpos += 13 + ((size_t) h[8] << 24 | (size_t) h[9] << 16 | (size_t) h[10] << 8 | (size_t) h[11]);
            }
            cpos += 1;
            if (!put_all(outputter, g.pin_in + opos, blen)) { io_error = true; break; }                             // :412-415
            if (action_handler) action_handler->OnProcess(g.pin_in + opos, blen);
            opos += blen;
        }
    }
    if (action_handler) action_handler->OnDone();
    return (inputter->IsErr() || outputter->IsErr()) ? -1 : 0;
}

}  // namespace zling
}  // namespace baidu
