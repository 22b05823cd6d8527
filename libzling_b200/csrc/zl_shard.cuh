// One stream over several GPUs, one process per GPU: the exchange steps of the path as C ABI entry points
// (include/zlb.h, "multi-GPU").  NCCL is loaded at run time (dlopen "libnccl.so.2": inside a torch process that is
// the library torch already uses), so libzling.so itself has no link-time dependency on it and single-GPU users
// never touch it.
//
// What couples the blocks of a stream in the reference is only what baidu::zling::Encode keeps outside its block
// loop: the MTF tables (m_mtf[256], src/libzling_lz.h:105 — never reset) and current_level
// (src/libzling.cpp:185,261-266).  The parse of a block range is independent of both (except the level of its first
// sub-block, which is verified afterwards), so every rank parses at once; the 65 540-byte carried state then moves
// GPU -> GPU in block order (ncclSend/ncclRecv, device buffers, no host bounce), each rank finishing MTF ranks +
// Huffman + framing of its range when its predecessor's state has arrived; ONE grouped variable-length gather
// (sizes by ncclAllGather of u64) brings the framed bytes to rank 0.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace zl {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* error = nullptr;
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[] = { "libnccl.so.2", "libnccl.so", nullptr };
    for (int i = 0; names[i] && !api.lib; i++) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) { api.error = "libnccl.so.2 not found (dlopen)"; return api; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) ok = false; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
    api.Send = (decltype(api.Send)) sym("ncclSend");
    api.Recv = (decltype(api.Recv)) sym("ncclRecv");
    api.AllGather = (decltype(api.AllGather)) sym("ncclAllGather");
    api.GroupStart = (decltype(api.GroupStart)) sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
    if (!ok) { api.error = "libnccl.so.2 lacks a required symbol"; dlclose(api.lib); api.lib = nullptr; }
    return api;
}

}  // namespace zl
