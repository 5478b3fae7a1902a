// tmb_nccl.h -- NCCL entry points resolved at run time (dlopen of libnccl.so.2), so that libtmap_b200.so carries no
// link-time dependency on NCCL: a single-GPU host never loads it, a multi-GPU job (one process per GPU) gets the
// copy its process already has (the one bundled with PyTorch shares the soname) or the system's.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdexcept>
#include <string>

namespace tmb
{
struct NcclApi
{
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    void *handle = nullptr;
    std::string error;

    static NcclApi &get()
    {
        static NcclApi api = load();
        return api;
    }
    bool ok() const { return handle != nullptr && error.empty(); }

  private:
    template <class F>
    static bool sym(void *h, const char *name, F &f, std::string &err)
    {
        f = reinterpret_cast<F>(dlsym(h, name));
        if (!f) err = std::string("libnccl.so.2 lacks ") + name;
        return f != nullptr;
    }
    static NcclApi load()
    {
        NcclApi a;
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle)
        {
            const char *e = dlerror();
            a.error = std::string("cannot load libnccl.so.2: ") + (e ? e : "unknown error");
            return a;
        }
        sym(a.handle, "ncclGetUniqueId", a.GetUniqueId, a.error) && sym(a.handle, "ncclCommInitRank", a.CommInitRank, a.error) &&
            sym(a.handle, "ncclCommDestroy", a.CommDestroy, a.error) && sym(a.handle, "ncclAllGather", a.AllGather, a.error) &&
            sym(a.handle, "ncclSend", a.Send, a.error) && sym(a.handle, "ncclRecv", a.Recv, a.error) &&
            sym(a.handle, "ncclGroupStart", a.GroupStart, a.error) && sym(a.handle, "ncclGroupEnd", a.GroupEnd, a.error) &&
            sym(a.handle, "ncclGetErrorString", a.GetErrorString, a.error);
        return a;
    }
};
}  // namespace tmb
