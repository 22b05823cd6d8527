// Re-entrancy of the drop-in API (the reference's Encode/Decode own all their state per call, src/libzling.cpp:108-163,180,299):
// several threads call baidu::zling::Encode / Decode at the same time on in-memory Inputter/Outputter objects; every result
// must equal the one the same call gives when run alone.  Used by tests/test_gpu_cxx_api.py.   zl_threads <file> [threads]
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include "libzling/libzling.h"

struct MemIn : baidu::zling::Inputter {
    const std::vector<unsigned char>& d; size_t at = 0;
    explicit MemIn(const std::vector<unsigned char>& v) : d(v) {}
    size_t GetData(unsigned char* buf, size_t len) override {
        const size_t n = len < d.size() - at ? len : d.size() - at;
        memcpy(buf, d.data() + at, n); at += n; return n;
    }
    bool IsEnd() override { return at >= d.size(); }
    bool IsErr() override { return false; }
};
struct MemOut : baidu::zling::Outputter {
    std::vector<unsigned char> d;
    size_t PutData(unsigned char* buf, size_t len) override { d.insert(d.end(), buf, buf + len); return len; }
    bool IsErr() override { return false; }
};

static std::vector<unsigned char> enc(const std::vector<unsigned char>& in, int level) {
    MemIn i(in); MemOut o;
    if (baidu::zling::Encode(&i, &o, NULL, level) != 0) throw std::runtime_error("Encode failed");
    return o.d;
}
static std::vector<unsigned char> dec(const std::vector<unsigned char>& in) {
    MemIn i(in); MemOut o;
    if (baidu::zling::Decode(&i, &o, NULL) != 0) throw std::runtime_error("Decode failed");
    return o.d;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: zl_threads file [threads]\n"); return 2; }
    const int nthreads = argc > 2 ? atoi(argv[2]) : 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    std::vector<unsigned char> data;
    { unsigned char buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.insert(data.end(), buf, buf + n); }
    fclose(f);
    // every thread gets its own slice / level, so that mixed-up buffers cannot go unnoticed
    std::vector<std::vector<unsigned char>> in(nthreads), want(nthreads), got(nthreads), back(nthreads);
    for (int t = 0; t < nthreads; t++) {
        const size_t lo = data.size() / (nthreads + 1) * t / 2;
        in[t].assign(data.begin() + lo, data.end() - (size_t) t * 1000 % (data.size() / 2 + 1));
        want[t] = enc(in[t], t % 5);                                   // alone
    }
    int bad = 0;
    for (int round = 0; round < 3; round++) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() {
            try { got[t] = enc(in[t], t % 5); back[t] = dec(got[t]); } catch (const std::exception& e) { fprintf(stderr, "thread %d: %s\n", t, e.what()); got[t].clear(); }
        });
        for (auto& x : th) x.join();
        for (int t = 0; t < nthreads; t++) if (got[t] != want[t] || back[t] != in[t]) { fprintf(stderr, "round %d thread %d: result differs from the solo run\n", round, t); bad++; }
    }
    fprintf(stderr, "threads=%d bad=%d\n", nthreads, bad);
    return bad ? 1 : 0;
}
