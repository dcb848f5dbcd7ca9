// The one collective of the hot paths: the Monte-Carlo error accumulators of independent filter shards are
// summed over the GPUs of a box with one ncclAllReduce over NVLink / NVSwitch.  Filters and image sequences never exchange
// data, so nothing else communicates.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries NCCL — a PyTorch process does — keeps
// using that very copy, a plain C++ host takes the system library, and a single-GPU user needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "../../include/ekfvio_c.h"

namespace ekfvio {
int fail_msg(const std::string& msg);
}

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
            api.CommCount = (decltype(api.CommCount))dlsym(api.lib, "ncclCommCount");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.CommCount && api.GetErrorString;
        }
    }
    return api;
}
int nccl_fail(const char* what, ncclResult_t r) { return ekfvio::fail_msg(std::string(what) + ": " + nccl().GetErrorString(r)); }
}  // namespace

struct ekfvio_comm {
    ncclComm_t comm = nullptr;
    int nranks = 0, rank = 0;
};

extern "C" {

int ekfvio_comm_unique_id(unsigned char* id128) {
    if (!id128) return ekfvio::fail_msg("ekfvio_comm_unique_id: null buffer");
    if (!nccl().ok) return ekfvio::fail_msg("ekfvio_comm_unique_id: NCCL (libnccl.so.2) could not be loaded");
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
    static_assert(sizeof(id) == EKFVIO_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int ekfvio_comm_create(ekfvio_comm** out, int device, int nranks, int rank, const unsigned char* id128) {
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return ekfvio::fail_msg("ekfvio_comm_create: bad arguments");
    if (!nccl().ok) return ekfvio::fail_msg("ekfvio_comm_create: NCCL (libnccl.so.2) could not be loaded");
    if (cudaSetDevice(device) != cudaSuccess) return ekfvio::fail_msg("ekfvio_comm_create: cudaSetDevice failed");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ekfvio_comm* c = new ekfvio_comm();
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail("ncclCommInitRank", r); }
    c->nranks = nranks; c->rank = rank;
    *out = c;
    return 0;
}

int ekfvio_comm_destroy(ekfvio_comm* c) {
    if (!c) return 0;
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
    return 0;
}

int ekfvio_comm_size(const ekfvio_comm* c) {
    int n = 0;
    if (c && c->comm && nccl().CommCount(c->comm, &n) == ncclSuccess) return n;
    return 0;
}

int ekfvio_stats_allreduce(ekfvio_comm* c, double* d_buf, int n, void* stream) {
    if (!c || !c->comm || !d_buf || n < 0) return ekfvio::fail_msg("ekfvio_stats_allreduce: bad arguments");
    ncclResult_t r = nccl().AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream);
    if (r != ncclSuccess) return nccl_fail("ncclAllReduce", r);
    return 0;
}

}  // extern "C"
