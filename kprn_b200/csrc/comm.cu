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
// ---- path-sharded pooling: LogSumExp over the paths of a (user,item) pair that live on different ranks ----
// (SP/model/module/LogSumExp.lua:14-30 evaluated on partials: every rank holds (m_r, z_r) of its own paths; the pooled value is
//  (M + log sum_r z_r e^{m_r - M}) / ig with M = max_r m_r.  Two small all-reduces (max, sum) over S x ncols floats.)
__global__ void lse_part_max_kernel(const float* __restrict__ seg_logits, const int32_t* __restrict__ seg, int64_t S, int L, int col0, int ncols,
                                    float ig, float* __restrict__ m) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * ncols) return;
    const int64_t k = i / ncols; const int col = col0 + (int)(i % ncols);
    m[i] = seg[k + 1] > seg[k] ? seg_logits[k * L + col] * ig : -INFINITY;          // a rank without paths of this pair contributes nothing
}
__global__ void lse_part_sum_kernel(const float* __restrict__ seg_logits, const int32_t* __restrict__ seg, int64_t S, int L, int col0, int ncols,
                                    float ig, const float* __restrict__ M, float* __restrict__ z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * ncols) return;
    const int64_t k = i / ncols; const int col = col0 + (int)(i % ncols);
    z[i] = (seg[k + 1] > seg[k] && M[i] > -INFINITY) ? expf(seg_logits[k * L + col] * ig - M[i]) : 0.f;
}
__global__ void lse_part_finish_kernel(float* __restrict__ seg_logits, int64_t S, int L, int col0, int ncols, float ig, int cls,
                                       const float* __restrict__ M, const float* __restrict__ z, float* __restrict__ scores,
                                       float* __restrict__ out_scores, float* __restrict__ out_logits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * ncols) return;
    const int64_t k = i / ncols; const int col = col0 + (int)(i % ncols);
    const float v = M[i] > -INFINITY ? (M[i] + logf(z[i])) / ig : 0.f;             // no path anywhere: 0, as the local pool does
    seg_logits[k * L + col] = v;
    if (out_logits) out_logits[k * L + col] = v;
    if (col == cls) {
        const float pr = 1.f / (1.f + expf(-v));
        scores[k] = pr;
        if (out_scores) out_scores[k] = pr;
    }
}

int pool_allreduce_lse(kprn_ctx* ctx, const int32_t* d_seg, int64_t S, float* d_out_scores, float* d_out_logits) {
    const kprn_config& c = ctx->cfg;
    if (!ctx->comm || ctx->world <= 1) return KPRN_OK;
    if (c.pool != KPRN_POOL_LSE) { set_error("path_sharded: only the LogSumExp pool combines across ranks"); return KPRN_EINVAL; }
    const int col0 = ctx->cls_only ? c.class_id - 1 : 0, ncols = ctx->cls_only ? 1 : c.L;
    const int64_t n = S * ncols;
    KPRN_TRY(ensure(ctx->tmp_pool, sizeof(float) * 2 * (size_t)n));
    float* M = ctx->tmp_pool.as<float>(); float* z = M + n;
    const float ig = c.inv_gamma == 0.f ? 1.f : c.inv_gamma;
    const unsigned grid = (unsigned)cdiv(n, 256);
    lse_part_max_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->seg_logits.as<float>(), d_seg, S, c.L, col0, ncols, ig, M);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(comm_allreduce_f32(ctx, M, n, true));
    lse_part_sum_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->seg_logits.as<float>(), d_seg, S, c.L, col0, ncols, ig, M, z);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(comm_allreduce_f32(ctx, z, n, false));
    lse_part_finish_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->seg_logits.as<float>(), S, c.L, col0, ncols, ig, c.class_id - 1, M, z,
                                                          ctx->scores.as<float>(), d_out_scores, d_out_logits);
    KPRN_LAUNCH_CHECK(ctx);
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
