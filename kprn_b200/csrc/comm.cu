// comm.cu -- multi-GPU inside the library: one process per GPU, an NCCL communicator owned by the ctx.
// The reference is single-device (SP/model/OneModel.lua:173-179); a LuaJIT host scales out by starting one process per GPU and
// calling kprn_comm_unique_id on rank 0, shipping the 128 bytes to the others (a file, a socket) and kprn_comm_init everywhere.
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy already mapped into the process when the host is a torch
// program, the system one otherwise), so single-GPU users need no NCCL at all and the library has no link-time dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

namespace kprn {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return KPRN_OK;
    const char* names[] = {getenv("KPRN_NCCL_LIB"), "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !n[0]) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { set_error("NCCL not found (libnccl.so.2): %s", dlerror()); return KPRN_ECUDA; }
#define KPRN_NCCL_SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                            \
    if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); dlclose(h); return KPRN_ECUDA; }
    KPRN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    KPRN_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    KPRN_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    KPRN_NCCL_SYM(AllReduce, "ncclAllReduce")
    KPRN_NCCL_SYM(AllGather, "ncclAllGather")
    KPRN_NCCL_SYM(Broadcast, "ncclBroadcast")
    KPRN_NCCL_SYM(GroupStart, "ncclGroupStart")
    KPRN_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    KPRN_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    KPRN_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef KPRN_NCCL_SYM
    g_nccl.handle = h;
    return KPRN_OK;
}

#define KPRN_NCCL(expr)                                                                              \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            set_error("NCCL error %d (%s) at %s:%d: %s", (int)_r, g_nccl.GetErrorString(_r), __FILE__, __LINE__, #expr); \
            return KPRN_ECUDA;                                                                       \
        }                                                                                            \
    } while (0)

int comm_allreduce_f32(kprn_ctx* ctx, float* buf, int64_t n, bool max_op) {
    if (ctx->world <= 1 || n <= 0) return KPRN_OK;
    KPRN_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, max_op ? ncclMax : ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    ctx->comm_bytes += 2 * (int64_t)sizeof(float) * n * (ctx->world - 1) / ctx->world;   // ring all-reduce: bytes sent per rank
    ctx->comm_calls++;
    return KPRN_OK;
}
int comm_allreduce_f32_2(kprn_ctx* ctx, float* a, int64_t na, float* b, int64_t nb) {   // two ranges, one group
    if (ctx->world <= 1) return KPRN_OK;
    KPRN_NCCL(g_nccl.GroupStart());
    if (na > 0) KPRN_NCCL(g_nccl.AllReduce(a, a, (size_t)na, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    if (nb > 0) KPRN_NCCL(g_nccl.AllReduce(b, b, (size_t)nb, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    KPRN_NCCL(g_nccl.GroupEnd());
    ctx->comm_bytes += 2 * (int64_t)sizeof(float) * (na + nb) * (ctx->world - 1) / ctx->world;
    ctx->comm_calls += 2;
    return KPRN_OK;
}
int comm_allreduce_i64(kprn_ctx* ctx, int64_t* buf, int64_t n, bool max_op) {
    if (ctx->world <= 1 || n <= 0) return KPRN_OK;
    KPRN_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclInt64, max_op ? ncclMax : ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    ctx->comm_bytes += 2 * 8 * n * (ctx->world - 1) / ctx->world;
    ctx->comm_calls++;
    return KPRN_OK;
}
// every rank contributes `bytes` (same on all ranks); recv holds world * bytes in rank order
int comm_allgather_bytes(kprn_ctx* ctx, const void* send, void* recv, int64_t bytes) {
    if (ctx->world <= 1) return KPRN_OK;
    KPRN_NCCL(g_nccl.AllGather(send, recv, (size_t)bytes, ncclInt8, (ncclComm_t)ctx->comm, ctx->stream));
    ctx->comm_bytes += bytes * (ctx->world - 1);     // bytes received per rank
    ctx->comm_calls++;
    return KPRN_OK;
}
int comm_broadcast_f32(kprn_ctx* ctx, float* buf, int64_t n, int root) {
    if (ctx->world <= 1) return KPRN_OK;
    KPRN_NCCL(g_nccl.Broadcast(buf, buf, (size_t)n, ncclFloat32, root, (ncclComm_t)ctx->comm, ctx->stream));
    ctx->comm_bytes += (int64_t)sizeof(float) * n;
    ctx->comm_calls++;
    return KPRN_OK;
}

}  // namespace kprn

using namespace kprn;

extern "C" {

int kprn_comm_unique_id(void* id128) {
    if (!id128) { set_error("null argument"); return KPRN_EINVAL; }
    KPRN_TRY(nccl_load());
    ncclUniqueId id;
    KPRN_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, sizeof(id));
    return KPRN_OK;
}

int kprn_comm_init(kprn_ctx* ctx, int32_t rank, int32_t world, const void* id128) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) { set_error("comm_init: bad arguments"); return KPRN_EINVAL; }
    if (ctx->comm) { set_error("comm_init: communicator already initialised"); return KPRN_ESTATE; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    ctx->rank = rank; ctx->world = world;
    if (world == 1) return KPRN_OK;
    KPRN_TRY(nccl_load());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    KPRN_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
    ctx->comm = c;
    return KPRN_OK;
}

int kprn_comm_destroy(kprn_ctx* ctx) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
    }
    ctx->world = 1; ctx->rank = 0;
    return KPRN_OK;
}

int kprn_comm_info(kprn_ctx* ctx, int32_t* rank, int32_t* world, int64_t* bytes, int64_t* calls, int32_t reset) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    if (bytes) *bytes = ctx->comm_bytes;
    if (calls) *calls = ctx->comm_calls;
    if (reset) { ctx->comm_bytes = 0; ctx->comm_calls = 0; }
    return KPRN_OK;
}

int kprn_broadcast_params(kprn_ctx* ctx, int32_t root) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    if (ctx->world <= 1) return KPRN_OK;
    KPRN_TRY(train_flush(ctx));
    // NCCL counts are size_t; one call moves the whole 2.9 GB vector
    KPRN_TRY(comm_broadcast_f32(ctx, ctx->d_params, ctx->lay.total, root));
    ctx->params_version++;
    return KPRN_OK;
}

int kprn_allreduce_grads(kprn_ctx* ctx, int32_t average) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    if (ctx->world <= 1) return KPRN_OK;
    KPRN_TRY(comm_allreduce_f32(ctx, ctx->d_grads, ctx->lay.total, false));
    if (average) KPRN_TRY(kprn_grad_scale(ctx, 1.f / (float)ctx->world));
    return KPRN_OK;
}

int kprn_allreduce_host_f64(kprn_ctx* ctx, double* vals, int64_t n, int32_t op_max) {
    if (!ctx || !vals || n < 0 || n > 64) { set_error("allreduce_host_f64: bad arguments (n <= 64)"); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    if (ctx->world <= 1 || n == 0) return KPRN_OK;
    // doubles travel as two floats each would lose bits: reduce them as int64 bit patterns is wrong for sums, so use a
    // float64-exact path: NCCL supports ncclFloat64
    KPRN_TRY(ensure(ctx->tmp_comm, 64 * sizeof(double)));
    KPRN_CUDA(cudaMemcpyAsync(ctx->tmp_comm.p, vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    KPRN_NCCL(g_nccl.AllReduce(ctx->tmp_comm.p, ctx->tmp_comm.p, (size_t)n, ncclFloat64, op_max ? ncclMax : ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    KPRN_CUDA(cudaMemcpyAsync(vals, ctx->tmp_comm.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->comm_calls++;
    return KPRN_OK;
}

}  // extern "C"
