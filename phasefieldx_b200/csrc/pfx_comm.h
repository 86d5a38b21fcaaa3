// pfx_comm.h — NCCL bound at run time (dlopen), so libpfx_b200.so loads on a single GPU without
// NCCL and, inside a torch process, shares the libnccl.so.2 torch already loaded.
// Replaces the MPI collectives of the reference's hot path: PETSc ghostUpdate /
// scatter_forward -> grouped ncclSend/ncclRecv; comm.allreduce(SUM) -> ncclAllReduce.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <string>

namespace pfx {

typedef struct ncclComm* nccl_comm_t;
struct nccl_uid { char internal[128]; };
enum { kNcclFloat64 = 8, kNcclSum = 0 };

struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    std::string load() {
        if (h) return "";
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return std::string("cannot dlopen libnccl.so.2: ") + dlerror();
#define PFX_SYM(field, name)                                        \
    *(void**)(&field) = dlsym(h, name);                             \
    if (!field) return std::string("missing NCCL symbol ") + name;
        PFX_SYM(GetUniqueId, "ncclGetUniqueId")
        PFX_SYM(CommInitRank, "ncclCommInitRank")
        PFX_SYM(CommDestroy, "ncclCommDestroy")
        PFX_SYM(AllReduce, "ncclAllReduce")
        PFX_SYM(Send, "ncclSend")
        PFX_SYM(Recv, "ncclRecv")
        PFX_SYM(GroupStart, "ncclGroupStart")
        PFX_SYM(GroupEnd, "ncclGroupEnd")
        PFX_SYM(GetErrorString, "ncclGetErrorString")
#undef PFX_SYM
        return "";
    }
};

inline Nccl& nccl() {
    static Nccl n;
    return n;
}

}  // namespace pfx
