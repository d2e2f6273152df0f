// Collectives of the column-distributed factorisation: NCCL, bound at run time.
//
// The single-GPU library has no NCCL dependency: the few entry points used here are looked up with
// dlopen/dlsym the first time a handle joins a group (the process that calls feba_dist_init has
// normally loaded libnccl already through torch.distributed; FEBA_NCCL_LIB names another file).
// Types below restate the stable part of nccl.h (NCCL 2.x ABI): opaque communicator pointer,
// 128-byte unique id, result / datatype / reduction enums by value.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "feba_kernels.h"

namespace feba {

namespace {

typedef struct { char internal[128]; } nccl_unique_id;
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    char why[256] = {0};
};

NcclApi g_api;

bool load_api() {
    if (g_api.lib) return true;
    const char* names[] = {std::getenv("FEBA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        snprintf(g_api.why, sizeof(g_api.why), "libnccl.so.2 not found (%s); set FEBA_NCCL_LIB", dlerror());
        return false;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(lib, name);
        if (!p) {
            snprintf(g_api.why, sizeof(g_api.why), "%s missing in the NCCL library", name);
            ok = false;
        }
        return p;
    };
    g_api.GetUniqueId = reinterpret_cast<decltype(g_api.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_api.CommInitRank = reinterpret_cast<decltype(g_api.CommInitRank)>(sym("ncclCommInitRank"));
    g_api.CommDestroy = reinterpret_cast<decltype(g_api.CommDestroy)>(sym("ncclCommDestroy"));
    g_api.Broadcast = reinterpret_cast<decltype(g_api.Broadcast)>(sym("ncclBroadcast"));
    g_api.AllReduce = reinterpret_cast<decltype(g_api.AllReduce)>(sym("ncclAllReduce"));
    g_api.Reduce = reinterpret_cast<decltype(g_api.Reduce)>(sym("ncclReduce"));
    g_api.GroupStart = reinterpret_cast<decltype(g_api.GroupStart)>(sym("ncclGroupStart"));
    g_api.GroupEnd = reinterpret_cast<decltype(g_api.GroupEnd)>(sym("ncclGroupEnd"));
    g_api.GetErrorString = reinterpret_cast<decltype(g_api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
        dlclose(lib);
        return false;
    }
    g_api.lib = lib;
    return true;
}

int nccl_fail(DistCtx* D, const char* what, int rc) {
    snprintf(D ? D->err : g_api.why, 256, "%s failed: %s", what,
             g_api.GetErrorString ? g_api.GetErrorString(rc) : "?");
    return rc ? rc : -1;
}

}  // namespace

const char* dist_load_error() { return g_api.why; }

int dist_unique_id(void* id128) {
    if (!load_api()) return -1;
    nccl_unique_id id;
    const int rc = g_api.GetUniqueId(&id);
    if (rc) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
    std::memcpy(id128, id.internal, sizeof(id.internal));
    return 0;
}

int dist_comm_init(DistCtx* D, int rank, int world, const void* id128) {
    if (!load_api()) {
        snprintf(D->err, sizeof(D->err), "%s", g_api.why);
        return -1;
    }
    nccl_unique_id id;
    std::memcpy(id.internal, id128, sizeof(id.internal));
    void* comm = nullptr;
    const int rc = g_api.CommInitRank(&comm, world, id, rank);
    if (rc) return nccl_fail(D, "ncclCommInitRank", rc);
    D->comm = comm;
    D->rank = rank;
    D->world = world;
    return 0;
}

void dist_comm_destroy(DistCtx* D) {
    if (D->comm && g_api.CommDestroy) g_api.CommDestroy(D->comm);
    D->comm = nullptr;
}

int dist_bcast_f64(DistCtx* D, double* buf, size_t count, int root, cudaStream_t st) {
    const int rc = g_api.Broadcast(buf, buf, count, kNcclFloat64, root, D->comm, st);
    return rc ? nccl_fail(D, "ncclBroadcast", rc) : 0;
}

int dist_reduce_f64(DistCtx* D, double* buf, size_t count, int root, cudaStream_t st) {
    const int rc = g_api.Reduce(buf, buf, count, kNcclFloat64, kNcclSum, root, D->comm, st);
    return rc ? nccl_fail(D, "ncclReduce", rc) : 0;
}

int dist_allreduce_f64(DistCtx* D, double* buf, size_t count, cudaStream_t st) {
    const int rc = g_api.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, D->comm, st);
    return rc ? nccl_fail(D, "ncclAllReduce", rc) : 0;
}

int dist_allreduce_max_i32(DistCtx* D, int* buf, size_t count, cudaStream_t st) {
    const int rc = g_api.AllReduce(buf, buf, count, kNcclInt32, kNcclMax, D->comm, st);
    return rc ? nccl_fail(D, "ncclAllReduce", rc) : 0;
}

int dist_group_start() { return g_api.GroupStart(); }
int dist_group_end(DistCtx* D) {
    const int rc = g_api.GroupEnd();
    return rc ? nccl_fail(D, "ncclGroupEnd", rc) : 0;
}

}  // namespace feba
