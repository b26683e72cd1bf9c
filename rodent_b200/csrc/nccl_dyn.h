// NCCL, bound at run time.  The library does not link libnccl: a process that also hosts PyTorch has PyTorch's own NCCL
// loaded under the same soname, and a C++ host (tools/rodent --gpus N) has the system's.  The first multi-device call
// dlopens "libnccl.so.2" -- which resolves to whatever the process already has, else to the system library -- and
// aborts with a message when there is none.  (The other order does not work: a process that loads the system library
// here and imports PyTorch afterwards has the wrong NCCL under PyTorch's soname; rodent_b200/render.py imports PyTorch
// first for that reason.)  Only the handful of calls the sharded paths use are bound: the film reduce
// (ncclReduce) and the hit-record gather (ncclSend / ncclRecv in one group).
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>

namespace rb200 {

struct Nccl {
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;

    static Nccl& get() {
        static Nccl n = load();
        return n;
    }

private:
    static Nccl load() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { std::fprintf(stderr, "rodent_b200: multi-device call without NCCL: %s\n", dlerror()); std::abort(); }
        Nccl n;
        auto bind = [h](auto& fn, const char* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, name));
            if (!fn) { std::fprintf(stderr, "rodent_b200: libnccl has no %s\n", name); std::abort(); }
        };
        bind(n.CommInitAll, "ncclCommInitAll"); bind(n.CommDestroy, "ncclCommDestroy"); bind(n.Reduce, "ncclReduce");
        bind(n.Send, "ncclSend"); bind(n.Recv, "ncclRecv"); bind(n.GroupStart, "ncclGroupStart"); bind(n.GroupEnd, "ncclGroupEnd");
        bind(n.GetErrorString, "ncclGetErrorString"); bind(n.GetVersion, "ncclGetVersion");
        return n;
    }
};

#define RB_NCCL_CHECK(expr)                                                                                  \
    do {                                                                                                     \
        ncclResult_t r__ = (expr);                                                                           \
        if (r__ != ncclSuccess) {                                                                            \
            std::fprintf(stderr, "rodent_b200: NCCL error '%s' at %s:%d (%s)\n",                              \
                         rb200::Nccl::get().GetErrorString(r__), __FILE__, __LINE__, #expr);                  \
            std::abort();                                                                                    \
        }                                                                                                    \
    } while (0)

}  // namespace rb200
