// Minimal CLI written against the public API only (include/libzling/libzling.h), used by tests/test_gpu_cxx_api.py:
//   zl_cli e<level> src dst | zl_cli d src dst
// It installs an ActionHandler that checks the ordering contract: when OnProcess(block) fires, the
// outputter must already have received that block's frames.
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "libzling/libzling.h"

struct CountingHandler : baidu::zling::ActionHandler {
    int blocks = 0, inits = 0, dones = 0;
    size_t bytes = 0, last_out = 0;
    bool order_ok = true;
    void OnInit() override { inits++; }
    void OnDone() override { dones++; }
    void OnProcess(unsigned char* data, size_t size) override {
        (void) data;
        blocks++;
        bytes += size;
        if (IsEncode()) {
            baidu::zling::FileOutputter* o = dynamic_cast<baidu::zling::FileOutputter*>(GetOutputter());
            if (!o || o->GetOutputSize() <= last_out) order_ok = false;   // this block's frames were written before the callback
            if (o) last_out = o->GetOutputSize();
        }
    }
};

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: zl_cli e[0-4]|d src dst\n"); return 2; }
    FILE* fi = fopen(argv[2], "rb");
    FILE* fo = fopen(argv[3], "wb");
    if (!fi || !fo) { fprintf(stderr, "cannot open files\n"); return 2; }
    baidu::zling::FileInputter in(fi);
    baidu::zling::FileOutputter out(fo);
    CountingHandler h;
    int rc = -1;
    try {
        if (argv[1][0] == 'e') rc = baidu::zling::Encode(&in, &out, &h, argv[1][1] ? argv[1][1] - '0' : 0);
        else rc = baidu::zling::Decode(&in, &out, &h);
    } catch (const std::runtime_error& e) {
        fprintf(stderr, "runtime_error: %s\n", e.what());
        return 3;
    }
    fclose(fi); fclose(fo);
    fprintf(stderr, "rc=%d blocks=%d bytes=%zu inits=%d dones=%d order=%s in=%zu out=%zu\n", rc, h.blocks, h.bytes, h.inits, h.dones,
            h.order_ok ? "ok" : "BAD", in.GetInputSize(), out.GetOutputSize());
    return (rc == 0 && h.order_ok && h.inits == 1 && h.dones == 1) ? 0 : 1;
}
