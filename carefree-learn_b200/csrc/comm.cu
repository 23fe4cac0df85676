// Gradient exchange for the data-parallel step: a thin C-ABI over NCCL (SURVEY.md section 8b `b200_comm_*`).
//
// Replaces the DDP wrap the reference gets from `accelerator.prepare` (cflearn/trainer.py:226-229,266-273): the only
// collective on the path is the sum all-reduce of parameter gradients between backward and `optimizer.step`
// (cflearn/schema.py:980-984).  The library owns its communicator (ncclCommInitRank) instead of going through
// torch.distributed's ProcessGroup so that `ncclAllReduce` is an ordinary stream operation: it can be captured into the
// CUDA graph of the training step, forked onto a communication stream and overlapped with the remaining backward,
// bucket by bucket, with no watchdog thread or Work object in the way.
//
// libnccl.so.2 is resolved at run time (dlopen; torch already has it mapped in the processes that use this), so the
// shared object has no link-time dependency on NCCL and loads on the CPU-only build box.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "b200_internal.h"

namespace {

// ---- the handful of NCCL declarations used (ABI-stable since NCCL 2.x; values from nccl.h) ----------------------
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;
enum { kNcclSuccess = 0, kNcclInProgress = 7 };
enum { kNcclSum = 0, kNcclAvg = 4 };
enum { kNcclFloat32 = 7 };

typedef int (*PFN_GetUniqueId)(NcclUniqueId*);
typedef int (*PFN_CommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*PFN_CommInitRankConfig)(ncclComm_t*, int, NcclUniqueId, int, void*);
typedef int (*PFN_CommDestroy)(ncclComm_t);
typedef int (*PFN_CommAbort)(ncclComm_t);
typedef int (*PFN_CommGetAsyncError)(ncclComm_t, int*);
typedef int (*PFN_AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*PFN_GetErrorString)(int);
typedef int (*PFN_GetVersion)(int*);

struct NcclApi {
    void* handle = nullptr;
    PFN_GetUniqueId GetUniqueId = nullptr;
    PFN_CommInitRank CommInitRank = nullptr;
    PFN_CommInitRankConfig CommInitRankConfig = nullptr;
    PFN_CommDestroy CommDestroy = nullptr;
    PFN_CommAbort CommAbort = nullptr;
    PFN_CommGetAsyncError CommGetAsyncError = nullptr;
    PFN_AllReduce AllReduce = nullptr;
    PFN_GetErrorString GetErrorString = nullptr;
    PFN_GetVersion GetVersion = nullptr;
};

NcclApi g_api;

int load_api() {
    if (g_api.handle != nullptr) return 0;
    // prefer the copy that is already mapped (torch's bundled libnccl), then the loader path
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return b200::set_error(B200_ERR_DRIVER, "libnccl.so.2 not found (import torch first, or put NCCL on the loader path)");
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<PFN_GetUniqueId>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<PFN_CommInitRank>(dlsym(h, "ncclCommInitRank"));
    a.CommInitRankConfig = reinterpret_cast<PFN_CommInitRankConfig>(dlsym(h, "ncclCommInitRankConfig"));
    a.CommDestroy = reinterpret_cast<PFN_CommDestroy>(dlsym(h, "ncclCommDestroy"));
    a.CommAbort = reinterpret_cast<PFN_CommAbort>(dlsym(h, "ncclCommAbort"));
    a.CommGetAsyncError = reinterpret_cast<PFN_CommGetAsyncError>(dlsym(h, "ncclCommGetAsyncError"));
    a.AllReduce = reinterpret_cast<PFN_AllReduce>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<PFN_GetErrorString>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<PFN_GetVersion>(dlsym(h, "ncclGetVersion"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString)
        return b200::set_error(B200_ERR_DRIVER, "libnccl.so.2 lacks an expected symbol");
    g_api = a;
    return 0;
}

int nccl_fail(const char* what, int rc) {
    char msg[256];
    snprintf(msg, sizeof(msg), "%s: NCCL error %d (%s)", what, rc, g_api.GetErrorString ? g_api.GetErrorString(rc) : "?");
    return b200::set_error(B200_ERR_LAUNCH, msg);
}

// ncclConfig_t as of NCCL 2.17 (nccl.h: size / magic / version, then the user fields).  NCCL copies min(size, its own sizeof)
// bytes and fills every field the caller's `version` does not know with its defaults, so this prefix is valid for all later
// releases (2.28.9 here).  Only maxCTAs is set: the limit applies to THIS communicator, not to torch.distributed's.
struct NcclConfigV21700 {
    size_t size;
    unsigned int magic;
    unsigned int version;
    int blocking;
    int cgaClusterSize;
    int minCTAs;
    int maxCTAs;
    const char* netName;
};
constexpr int kNcclUndefInt = -2147483647 - 1;  // NCCL_CONFIG_UNDEF_INT (INT_MIN)

struct Comm {
    ncclComm_t nccl;
    int rank, world;
};

}  // namespace

extern "C" int b200_comm_unique_id(void* id_out_128) {
    if (id_out_128 == nullptr) return b200::set_error(B200_ERR_ARG, "comm_unique_id: null output");
    int rc = load_api();
    if (rc) return rc;
    NcclUniqueId id;
    rc = g_api.GetUniqueId(&id);
    if (rc != kNcclSuccess) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id_out_128, id.internal, sizeof(id.internal));
    return 0;
}

extern "C" int b200_comm_init(const void* id_128, int rank, int world, int max_ctas, void** comm_out) {
    if (id_128 == nullptr || comm_out == nullptr || world < 1 || rank < 0 || rank >= world)
        return b200::set_error(B200_ERR_ARG, "comm_init: bad arguments");
    int rc = load_api();
    if (rc) return rc;
    NcclUniqueId id;
    memcpy(id.internal, id_128, sizeof(id.internal));
    ncclComm_t c = nullptr;
    if (max_ctas > 0 && g_api.CommInitRankConfig != nullptr) {
        NcclConfigV21700 cfg{sizeof(NcclConfigV21700), 0xcafebeefu, 21700u, kNcclUndefInt, kNcclUndefInt, kNcclUndefInt, max_ctas, nullptr};
        if (cfg.maxCTAs > 32) cfg.maxCTAs = 32;
        rc = g_api.CommInitRankConfig(&c, world, id, rank, &cfg);  // collective over the ranks; uses the CURRENT device
        if (rc == 4 /* ncclInvalidArgument: the config is checked before any communication, identically on every rank */) {
            c = nullptr;
            rc = g_api.CommInitRank(&c, world, id, rank);
        }
        if (rc != kNcclSuccess) return nccl_fail("ncclCommInitRankConfig", rc);
    } else {
        rc = g_api.CommInitRank(&c, world, id, rank);
        if (rc != kNcclSuccess) return nccl_fail("ncclCommInitRank", rc);
    }
    Comm* out = new Comm{c, rank, world};
    *comm_out = out;
    return 0;
}

extern "C" int b200_comm_allreduce_bucket(void* comm, float* buf, long long n, int average, cudaStream_t stream) {
    if (comm == nullptr || buf == nullptr || n <= 0) return b200::set_error(B200_ERR_ARG, "comm_allreduce_bucket: bad arguments");
    Comm* c = static_cast<Comm*>(comm);
    const int rc = g_api.AllReduce(buf, buf, static_cast<size_t>(n), kNcclFloat32, average ? kNcclAvg : kNcclSum, c->nccl, stream);
    if (rc != kNcclSuccess) return nccl_fail("ncclAllReduce", rc);
    return 0;
}

// 0: healthy; negative: an asynchronous error was recorded on the communicator (a peer died, a transport failed): the
// caller aborts the job instead of waiting forever -- the NCCL-async-error -> abort path of SURVEY.md section 5.
extern "C" int b200_comm_async_error(void* comm) {
    if (comm == nullptr) return b200::set_error(B200_ERR_ARG, "comm_async_error: null communicator");
    if (g_api.CommGetAsyncError == nullptr) return 0;
    Comm* c = static_cast<Comm*>(comm);
    int state = kNcclSuccess;
    const int rc = g_api.CommGetAsyncError(c->nccl, &state);
    if (rc != kNcclSuccess) return nccl_fail("ncclCommGetAsyncError", rc);
    if (state != kNcclSuccess && state != kNcclInProgress) return nccl_fail("asynchronous communicator error", state);
    return 0;
}

extern "C" int b200_comm_finalize(void* comm, int abort) {
    if (comm == nullptr) return 0;
    Comm* c = static_cast<Comm*>(comm);
    int rc = kNcclSuccess;
    if (abort && g_api.CommAbort) rc = g_api.CommAbort(c->nccl);
    else rc = g_api.CommDestroy(c->nccl);
    delete c;
    if (rc != kNcclSuccess) return nccl_fail("ncclCommDestroy", rc);
    return 0;
}

extern "C" int b200_comm_nccl_version(void) {
    if (load_api() != 0 || g_api.GetVersion == nullptr) return 0;
    int v = 0;
    g_api.GetVersion(&v);
    return v;
}
