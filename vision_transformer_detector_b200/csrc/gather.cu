// The path's only inter-GPU exchange: all-gather of the fixed-size detection records over NCCL / NVLink.
// The reference has no distributed code; this is the C-ABI form of what parallel.py does through torch.distributed, so
// that a C caller (and the timed multi-GPU step) needs neither torch tensors nor eager packing kernels: head_tail_kernel
// writes the 13-float record itself (DecodeOut::packed) and ncclAllGather moves it.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded in the process — torch's bundled copy under
// torchrun — or VITDET_NCCL_LIB): the library keeps loading on machines without NCCL, where these entry points fail
// with VITDET_E_INVALID instead.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/vitdet_b200.h"
#include "kernels.h"

namespace {

// the few NCCL declarations used (nccl.h 2.x; ncclUniqueId is passed BY VALUE to ncclCommInitRank)
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclFloat = 7 };      // ncclFloat32
struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    char why[256] = "";
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[3] = {getenv("VITDET_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { snprintf(api.why, sizeof(api.why), "libnccl.so.2 could not be loaded (%s)", dlerror()); return api; }
    api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(dlsym(lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(dlsym(lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(lib, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(dlsym(lib, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(lib, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    if (!api.ok) snprintf(api.why, sizeof(api.why), "libnccl.so.2 lacks an expected symbol");
    return api;
}

#define NCCL_TRY(expr)                                                                                    \
    do {                                                                                                  \
        int r__ = (expr);                                                                                 \
        if (r__ != 0) return vitdet::fail(VITDET_E_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(r__)); \
    } while (0)

}  // namespace

extern "C" {

int vitdet_nccl_unique_id(char id_out[128]) {
    if (!id_out) return vitdet::fail(VITDET_E_INVALID, "nccl_unique_id: null argument");
    NcclApi& a = nccl();
    if (!a.ok) return vitdet::fail(VITDET_E_INVALID, "NCCL unavailable: %s", a.why);
    NcclUniqueId id;
    NCCL_TRY(a.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return 0;
}

int vitdet_nccl_comm_create(const char id[128], int rank, int world, void** nccl_comm_out) {
    if (!id || !nccl_comm_out || world <= 0 || rank < 0 || rank >= world) return vitdet::fail(VITDET_E_INVALID, "nccl_comm_create: bad arguments");
    NcclApi& a = nccl();
    if (!a.ok) return vitdet::fail(VITDET_E_INVALID, "NCCL unavailable: %s", a.why);
    NcclUniqueId uid;
    memcpy(uid.internal, id, 128);
    NcclComm c = nullptr;
    NCCL_TRY(a.CommInitRank(&c, world, uid, rank));
    *nccl_comm_out = c;
    return 0;
}

int vitdet_nccl_comm_destroy(void* nccl_comm) {
    if (!nccl_comm) return 0;
    NcclApi& a = nccl();
    if (!a.ok) return vitdet::fail(VITDET_E_INVALID, "NCCL unavailable: %s", a.why);
    NCCL_TRY(a.CommDestroy(nccl_comm));
    return 0;
}

int vitdet_gather_detections(void* nccl_comm, const float* packed_local_dev, int rows_local, float* packed_all_dev, void* stream) {
    if (!nccl_comm || !packed_local_dev || !packed_all_dev || rows_local < 0) return vitdet::fail(VITDET_E_INVALID, "gather_detections: bad arguments");
    if (rows_local == 0) return 0;
    NcclApi& a = nccl();
    if (!a.ok) return vitdet::fail(VITDET_E_INVALID, "NCCL unavailable: %s", a.why);
    NCCL_TRY(a.AllGather(packed_local_dev, packed_all_dev, static_cast<size_t>(rows_local) * VITDET_RECORD_FLOATS, kNcclFloat, nccl_comm,
                         static_cast<cudaStream_t>(stream)));
    return 0;
}

}  // extern "C"
