// comm.cu -- multi-GPU exchange behind the C-ABI: an NCCL communicator owned by the ipf_ctx.
//
// The path shards by configuration (rows of different configurations are disjoint, src/lsq_db.jl:242-247); the only
// exchange step of a fit is the sum of the per-rank (n+1)^2 Gram blocks of a CholeskyQR pass, plus a handful of
// scalars (residual sums, error-table sums).  With a communicator attached, ipf_solve_gram all-reduces its block on
// the ctx stream (no host synchronisation, no torch), so a Julia caller gets the multi-GPU solve from the same calls.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy already loaded into the process by the host
// framework if there is one, else the system library); libipf_b200.so has no link-time dependency on it.
#include <dlfcn.h>

#include "ipf_common.cuh"

namespace {

struct NcclId { char internal[128]; };
typedef void* NcclComm;
// ncclDataType_t: ncclFloat64 = 8; ncclRedOp_t: ncclSum = 0 (stable since NCCL 2.0)
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return &api; }
    api.GetUniqueId = (int (*)(NcclId*))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))dlsym(api.handle, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString)
        api.err = "libnccl.so.2 lacks an expected symbol";
    return &api;
}

int nccl_fail(ipf_ctx* ctx, NcclApi* api, const char* what, int rc) {
    return ipf_set_error(ctx, IPF_ENCCL, "%s: %s", what, api->GetErrorString ? api->GetErrorString(rc) : "NCCL error");
}

}  // namespace

// in-place sum over the ranks of n doubles at device address buf, on the ctx stream; no-op without a communicator
int ipf_comm_allreduce_dev(ipf_ctx* ctx, double* buf, size_t n) {
    if (!ctx->nccl_comm || ctx->nranks <= 1 || n == 0 || ctx->comm_paused) return IPF_OK;
    NcclApi* api = nccl_api();
    ProfScope prof_(ctx, "allreduce");
    const int rc = api->AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, (NcclComm)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return nccl_fail(ctx, api, "ncclAllReduce", rc);
    return IPF_OK;
}

extern "C" {

int32_t ipf_comm_unique_id(ipf_ctx* ctx, uint8_t* id) {
    if (!ctx || !id) return ipf_set_error(ctx, IPF_EINVAL, "ipf_comm_unique_id: bad arguments");
    NcclApi* api = nccl_api();
    if (!api->err.empty()) return ipf_set_error(ctx, IPF_ENCCL, "%s", api->err.c_str());
    NcclId nid;
    const int rc = api->GetUniqueId(&nid);
    if (rc != 0) return nccl_fail(ctx, api, "ncclGetUniqueId", rc);
    memcpy(id, nid.internal, IPF_UNIQUE_ID_BYTES);
    return IPF_OK;
}

int32_t ipf_comm_init(ipf_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || !id)
        return ipf_set_error(ctx, IPF_EINVAL, "ipf_comm_init: bad arguments");
    if (ctx->nccl_comm) return ipf_set_error(ctx, IPF_ESTATE, "ipf_comm_init: a communicator is already attached");
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->nranks = nranks; ctx->rank = rank;
    if (nranks == 1) return IPF_OK;
    NcclApi* api = nccl_api();
    if (!api->err.empty()) return ipf_set_error(ctx, IPF_ENCCL, "%s", api->err.c_str());
    NcclId nid;
    memcpy(nid.internal, id, IPF_UNIQUE_ID_BYTES);
    NcclComm comm = nullptr;
    const int rc = api->CommInitRank(&comm, nranks, nid, rank);
    if (rc != 0) return nccl_fail(ctx, api, "ncclCommInitRank", rc);
    ctx->nccl_comm = comm;
    return IPF_OK;
}

int32_t ipf_comm_destroy(ipf_ctx* ctx) {
    if (!ctx) return IPF_EINVAL;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl_api()->CommDestroy((NcclComm)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1; ctx->rank = 0;
    return IPF_OK;
}

int32_t ipf_comm_pause(ipf_ctx* ctx, int32_t paused) {
    if (!ctx) return IPF_EINVAL;
    ctx->comm_paused = paused != 0;
    return IPF_OK;
}

int32_t ipf_comm_size(const ipf_ctx* ctx) { return ctx ? ctx->nranks : 0; }
int32_t ipf_comm_rank(const ipf_ctx* ctx) { return ctx ? ctx->rank : 0; }

// sum over the ranks of a small host vector (residual sums, error-table sums), staged through the device
int32_t ipf_comm_allreduce_f64(ipf_ctx* ctx, double* x, int64_t n) {
    if (!ctx || (n > 0 && !x) || n < 0) return ipf_set_error(ctx, IPF_EINVAL, "ipf_comm_allreduce_f64: bad arguments");
    if (!ctx->nccl_comm || ctx->nranks <= 1 || n == 0) return IPF_OK;
    IPF_CUDA(ctx, cudaSetDevice(ctx->device));
    void* d = nullptr;
    IPF_CHECK(ipf_scratch(ctx, 9, sizeof(double) * (size_t)n, &d));
    IPF_CUDA(ctx, cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    const bool paused = ctx->comm_paused;
    ctx->comm_paused = false;           // an explicit exchange is never paused
    const int rc_ar = ipf_comm_allreduce_dev(ctx, (double*)d, (size_t)n);
    ctx->comm_paused = paused;
    IPF_CHECK(rc_ar);
    IPF_CUDA(ctx, cudaMemcpyAsync(x, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    IPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IPF_OK;
}

}  // extern "C"
