/*
 * xm_nccl.h -- NCCL for the sharded walk, bound at run time.
 *
 * libxenomapper_b200.so does not link against NCCL: a single-GPU user needs
 * none, and a launcher that has already loaded a libnccl.so.2 (PyTorch ships
 * one) must not end up with two.  The few entry points the sharded walk uses
 * are looked up with dlopen/dlsym on first use; their prototypes follow nccl.h
 * (2.x ABI: ncclUniqueId is 128 opaque bytes passed by value).
 */
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <string>

namespace xm {

struct NcclId { char internal[128]; };
typedef struct ncclComm *NcclComm;
enum { NCCL_SUM = 0, NCCL_MAX = 2 };                       /* ncclRedOp_t */
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 }; /* ncclDataType_t */

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;

    bool load()
    {
        if (lib) return true;
        const char *names[] = {getenv("XM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "dlopen failed") + "): the multi-GPU walk needs NCCL"; return false; }
        bool ok = true;
        auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p) { ok = false; err = std::string("libnccl lacks ") + n; } return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        Send = (decltype(Send))sym("ncclSend");
        Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(lib); lib = nullptr; }
        return ok;
    }
};

inline NcclApi &nccl_api()
{
    static NcclApi api;
    return api;
}

}  // namespace xm
