// slr_nccl.cu — the one collective of the path: an in-place NCCL all-gather of every rank's cloud over NVLink
// (the north star's "single NCCL all-gather of the output point cloud").
//
// Scans are independent, so nothing on the data path needs a collective; this entry point only ASSEMBLES the clouds
// where a caller wants all of them on every GPU.  libnccl is bound at run time (dlopen of the library already in the
// process — torch's — or the system one), so libslr_b200.so has no link-time NCCL dependency and single-GPU users
// need no NCCL at all.  Only the four NCCL calls used here are declared; their prototypes follow nccl.h 2.x.
#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

#include "slr_internal.h"

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccessV = 0 };
enum { ncclUint8V = 1, ncclFloat32V = 7 };   // ncclDataType_t values of nccl.h (ncclUint8 = 1, ncclFloat32 = 7)

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_api;

static void nccl_bind()
{
    NcclApi &api = g_api;
    const char *names[] = {getenv("SLR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllGather");
    api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
    api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GroupStart || !api.GroupEnd) {
        dlclose(api.lib);
        api.lib = nullptr;
    }
}

NcclApi *nccl_api()   // thread-safe one-time binding (slr_last_error is per thread, so callers may be too)
{
    static std::once_flag once;
    std::call_once(once, nccl_bind);
    return g_api.lib ? &g_api : nullptr;
}

#define SLR_CHECK_NCCL(api, expr)                                                                              \
    do {                                                                                                       \
        const int _r = (expr);                                                                                 \
        if (_r != ncclSuccessV) {                                                                              \
            slr_set_error("%s failed: %s", #expr, (api)->GetErrorString ? (api)->GetErrorString(_r) : "NCCL error"); \
            return SLR_ERR_CUDA;                                                                               \
        }                                                                                                      \
    } while (0)

}  // namespace

extern "C" slr_status slr_nccl_unique_id(void *id128)
{
    SLR_REQUIRE(id128 != nullptr, "slr_nccl_unique_id: NULL argument");
    NcclApi *a = nccl_api();
    SLR_REQUIRE(a != nullptr, "NCCL is not available (libnccl.so.2 not found; set SLR_NCCL_LIB)");
    SLR_CHECK_NCCL(a, a->GetUniqueId((ncclUniqueId *)id128));
    return SLR_OK;
}

extern "C" slr_status slr_nccl_comm_create(slr_engine *e, void **comm, int world, int rank, const void *id128)
{
    SLR_REQUIRE(e && comm && id128 && world > 0 && rank >= 0 && rank < world, "slr_nccl_comm_create: bad argument");
    NcclApi *a = nccl_api();
    SLR_REQUIRE(a != nullptr, "NCCL is not available (libnccl.so.2 not found; set SLR_NCCL_LIB)");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    ncclComm_t c = nullptr;
    SLR_CHECK_NCCL(a, a->CommInitRank(&c, world, *(const ncclUniqueId *)id128, rank));
    *comm = c;
    return SLR_OK;
}

extern "C" slr_status slr_nccl_comm_destroy(void *comm)
{
    NcclApi *a = nccl_api();
    if (a && comm) SLR_CHECK_NCCL(a, a->CommDestroy((ncclComm_t)comm));
    return SLR_OK;
}

extern "C" slr_status slr_allgather(slr_engine *e, void *nccl_comm, int world, int rank, int scans_per_rank,
                                    float *d_xyz_all, uint8_t *d_valid_all)
{
    SLR_REQUIRE(e && nccl_comm && d_xyz_all && d_valid_all && world > 0 && rank >= 0 && rank < world && scans_per_rank > 0,
                "slr_allgather: bad argument");
    NcclApi *a = nccl_api();
    SLR_REQUIRE(a != nullptr, "NCCL is not available (libnccl.so.2 not found; set SLR_NCCL_LIB)");
    SLR_CHECK_CUDA(cudaSetDevice(e->device));
    const size_t px = (size_t)scans_per_rank * e->W * e->H;
    // in place: this rank's block already sits at rank * count of the receive buffer.  Both tensors travel in ONE
    // NCCL group: a single communication call (one launch) per step
    SLR_CHECK_NCCL(a, a->GroupStart());
    const int r1 = a->AllGather(d_xyz_all + (size_t)rank * px * 3, d_xyz_all, px * 3, ncclFloat32V, (ncclComm_t)nccl_comm, e->stream);
    const int r2 = a->AllGather(d_valid_all + (size_t)rank * px, d_valid_all, px, ncclUint8V, (ncclComm_t)nccl_comm, e->stream);
    SLR_CHECK_NCCL(a, a->GroupEnd());
    SLR_CHECK_NCCL(a, r1);
    SLR_CHECK_NCCL(a, r2);
    return SLR_OK;
}
