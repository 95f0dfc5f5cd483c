// api.cu -- the C ABI of include/kprn_b200.h: ctx lifetime, parameter I/O, forward/loss/backward/optimizer
// entry points.  Each entry point cites the reference interface it replaces in the header.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "adam.cuh"
#include "common.cuh"

namespace kprn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return KPRN_ECUDA;
}

int ensure(DBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return KPRN_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        b.p = nullptr;
        return KPRN_ENOMEM;
    }
    b.cap = want;
    return KPRN_OK;
}

int ensure_stage(kprn_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_stage_cap) return KPRN_OK;
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMallocHost(&ctx->h_stage, want);
    if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); return KPRN_ENOMEM; }
    ctx->h_stage_cap = want;
    return KPRN_OK;
}

int prof_begin(kprn_ctx* ctx, cudaEvent_t* e0) {
    *e0 = nullptr;
    if (!ctx->profiling) return KPRN_OK;
    KPRN_CUDA(cudaEventCreate(e0));
    KPRN_CUDA(cudaEventRecord(*e0, ctx->stream));
    return KPRN_OK;
}
int prof_end(kprn_ctx* ctx, cudaEvent_t e0, int slot) {
    if (!ctx->profiling || !e0) return KPRN_OK;
    cudaEvent_t e1;
    KPRN_CUDA(cudaEventCreate(&e1));
    KPRN_CUDA(cudaEventRecord(e1, ctx->stream));
    ctx->prof_pending.push_back({e0, e1, slot});
    return KPRN_OK;
}

static void make_layout(const kprn_config& c, Layout& l) {
    int64_t o = 0;
    l.D = c.d_t + c.d_e + c.d_r;
    l.G = (c.cell == KPRN_CELL_LSTM ? 4 : 1) * c.H;
    l.E_t = o; o += c.V_t * c.d_t;
    l.E_e = o; o += c.V_e * c.d_e;
    l.E_r = o; o += c.V_r * c.d_r;
    l.Wx = o; o += (int64_t)l.G * l.D;
    l.bx = o; o += l.G;
    l.Wh = o; o += (int64_t)l.G * c.H;
    l.bh = -1;
    if (c.cell == KPRN_CELL_RNN) { l.bh = o; o += c.H; }
    l.W1 = l.b1 = -1;
    if (c.head_layers == 2) { l.W1 = o; o += (int64_t)c.H * c.H; l.b1 = o; o += c.H; }
    l.Wo = o; o += (int64_t)c.L * c.H;
    l.bo = o; o += c.L;
    l.total = o;
}

// ---------------------------------------------------------------------------------------------
// small device kernels owned by the API layer

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
// reference tensor layout: DoubleTensor of 1-based ids (SP/util/Util.lua:121-201) -> int32
__global__ void f64ids_to_i32_kernel(const double* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)llrint(in[i]);
}
__global__ void dense_seg_kernel(int32_t* __restrict__ seg, int64_t S, int32_t P) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= S) seg[i] = (int32_t)(i * P);
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void init_uniform_kernel(float* __restrict__ p, int64_t n, float scale, uint64_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    float u = (float)((r >> 40) * (1.0 / 16777216.0));   // [0,1)
    p[i] = (2.f * u - 1.f) * scale;
}
__global__ void scale_kernel(float* __restrict__ x, int64_t n, float s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= s;
}
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n, float a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = fmaf(a, x[i], y[i]);
}
__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = x[i];
        acc += (double)v * (double)v;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
// optim.adam (OneModel.lua:340-355): dense over the flat vector, eps added after the sqrt.
__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float step, float b1, float b2, float eps) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float xi = x[i], mi = m[i], vi = v[i];
        adam_elem(xi, mi, vi, g[i], step, b1, b2, eps);
        x[i] = xi; m[i] = mi; v[i] = vi;
    }
}
// nn.BCECriterion (eps 1e-12, sizeAverage) + chain through nn.Sigmoid/nn.Select: writes dL/dlse_c.
__global__ void bce_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t S, double* __restrict__ loss,
                           float* __restrict__ dlse) {
    double acc = 0.0;
    const float eps = 1e-12f;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < S; k += (int64_t)gridDim.x * blockDim.x) {
        float pk = p[k], yk = y[k];
        acc -= (double)(yk * logf(pk + eps) + (1.f - yk) * logf(1.f - pk + eps));
        float dp = -(yk - pk) / ((1.f - pk + eps) * (pk + eps)) / (float)S;
        dlse[k] = dp * pk * (1.f - pk);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc / (double)S);
}
// nn.BPRLoss (criterion/BPRLoss.lua:18-32): out = sigmoid(pos-neg)+1e-4; err = -mean(log out);
// gradInput = {-(1-out), +(1-out)} (not divided by n).  Chained to dL/dlse_c.
__global__ void bpr_kernel(const float* __restrict__ scores, const float* __restrict__ seg_logits, int L, int cls,
                           const int32_t* __restrict__ pos, const int32_t* __restrict__ neg, int64_t n, int on_logits, int64_t S,
                           double* __restrict__ loss, float* __restrict__ dlse) {
    double acc = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int32_t a = pos[k], b = neg[k];
        if (a < 0 || a >= S || b < 0 || b >= S) continue;       // flagged by check_pairs_kernel
        float va, vb;
        if (on_logits) { va = seg_logits[(int64_t)a * L + cls]; vb = seg_logits[(int64_t)b * L + cls]; }
        else { va = scores[a]; vb = scores[b]; }
        float out = 1.f / (1.f + expf(-(va - vb))) + 1e-4f;
        acc -= (double)logf(out);
        float step = 1.f - out;
        float ja = on_logits ? 1.f : scores[a] * (1.f - scores[a]);
        float jb = on_logits ? 1.f : scores[b] * (1.f - scores[b]);
        atomicAdd(dlse + a, -step * ja);
        atomicAdd(dlse + b, step * jb);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc / (double)n);
}
__global__ void chain_scores_kernel(const float* __restrict__ scores, const float* __restrict__ dscore, int64_t S,
                                    float* __restrict__ dlse) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < S) dlse[k] = dscore[k] * scores[k] * (1.f - scores[k]);
}
__global__ void check_seg_kernel(const int32_t* __restrict__ seg, int64_t S, int64_t N, int32_t* err) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > S) return;
    if (k == 0 && seg[0] != 0) atomicExch(err, 2);
    if (k == S && seg[S] != N) atomicExch(err, 2);
    if (k < S && seg[k + 1] < seg[k]) atomicExch(err, 2);
}
__global__ void validate_ids_kernel2(const int32_t* __restrict__ ids, int64_t n_steps, int64_t V_t, int64_t V_e,
                                     int64_t V_r, int32_t* err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_steps) return;
    int32_t a = ids[i * 3], b = ids[i * 3 + 1], c = ids[i * 3 + 2];
    if (a < 1 || a > V_t || b < 1 || b > V_e || c < 1 || c > V_r) atomicExch(err, 1);
}

static int check_ctx(kprn_ctx* ctx) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
    return KPRN_OK;
}

static int check_err_flag(kprn_ctx* ctx) {
    int32_t flag = 0;
    KPRN_CUDA(cudaMemcpyAsync(&flag, ctx->d_err, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        KPRN_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream));
        set_error(flag == 1 ? "id out of range (ids are 1-based; nn.LookupTable would raise)"
                            : (flag == 3 ? "BPR pair: segment index out of range" : "bad segment offsets"));
        ctx->have_fwd = false;
        return KPRN_EINVAL;
    }
    return KPRN_OK;
}

static int params_changed(kprn_ctx* ctx) { ctx->params_version++; return KPRN_OK; }

int forward_device_impl(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S,
                               bool train, bool want_logits, float* d_out_scores, float* d_out_logits) {
    const kprn_config& c = ctx->cfg;
    if (N <= 0 || S <= 0) { set_error("empty batch: N=%lld S=%lld", (long long)N, (long long)S); return KPRN_EINVAL; }
    if (N * (int64_t)c.T * 3 >= (1ll << 40)) { set_error("batch too large"); return KPRN_EINVAL; }
    if (!ctx->in_train_step) KPRN_TRY(train_flush(ctx));     // rows a sparse training step left behind (train.cu) must be current
    // the fused tensor-core forward validates the ids inside its first kernel (the deduplication's hash insert)
    if (!(!train && c.precision != KPRN_PREC_FP32 && tc_supported(ctx) && tc_validates_ids(ctx, N))) {
        validate_ids_kernel2<<<(unsigned)cdiv(N * c.T, 256), 256, 0, ctx->stream>>>(d_ids, N * c.T, c.V_t, c.V_e, c.V_r, ctx->d_err);
        KPRN_LAUNCH_CHECK(ctx);
    }
    check_seg_kernel<<<(unsigned)cdiv(S + 1, 256), 256, 0, ctx->stream>>>(d_seg, S, N, ctx->d_err);
    KPRN_LAUNCH_CHECK(ctx);
    ctx->N = N; ctx->S = S; ctx->d_ids = d_ids; ctx->d_seg = d_seg;
    ctx->have_loss_grad = false;
    // scores-only fast path: nothing downstream of nn.Select(2,classId) needs the other L-1 columns
    ctx->cls_only = (!train || ctx->in_train_step) && !want_logits && c.head_layers == 1;
    ctx->head_part4 = nullptr;      // set by the fused tensor-core forward when it leaves the head to the pool kernel
    bool use_tc = c.precision != KPRN_PREC_FP32 && tc_supported(ctx);
    // shapes the fused kernels do not cover (H or d_e > 256, Elman cell, large type x relation product): the generic
    // engine runs, with its large products on tcgen05 (bf16x3) for scoring; TC_SPLIT training stays on fp32 FFMA
    ctx->gemm_on_tc = c.precision != KPRN_PREC_FP32 && !use_tc && !train;
    cudaEvent_t e0;
    KPRN_TRY(prof_begin(ctx, &e0));
    ctx->tc_train_fwd = false;
    if (use_tc) KPRN_TRY(tc_forward(ctx, d_ids, d_seg, N, S, train));
    else KPRN_TRY(simple_forward(ctx, d_ids, d_seg, N, S, train));
    KPRN_TRY(prof_end(ctx, e0, 0));
    ctx->gemm_on_tc = false;
    KPRN_TRY(pool_forward(ctx, d_seg, S, d_out_scores, d_out_logits));
    if (ctx->opt_path_sharded > 0 && ctx->world > 1) KPRN_TRY(pool_allreduce_lse(ctx, d_seg, S, d_out_scores, d_out_logits));
    ctx->have_fwd = true;
    ctx->have_train_cache = train;
    return KPRN_OK;
}

}  // namespace kprn

using namespace kprn;

extern "C" {

const char* kprn_last_error(void) { return g_err; }
const char* kprn_version(void) { return "kprn_b200 0.1 (sm_100a)"; }

int kprn_create(const kprn_config* cfg, kprn_ctx** out) {
    if (!cfg || !out) { set_error("null argument"); return KPRN_EINVAL; }
    const kprn_config& c = *cfg;
    if (c.T < 1 || c.T > 64 || c.d_t < 1 || c.d_e < 1 || c.d_r < 1 || c.H < 1 || c.L < 1 || c.V_t < 1 || c.V_e < 1 ||
        c.V_r < 1 || c.class_id < 1 || c.class_id > c.L || (c.cell != KPRN_CELL_LSTM && c.cell != KPRN_CELL_RNN) ||
        (c.head_layers != 1 && c.head_layers != 2) || c.pool < 0 || c.pool > 2 || (c.pool == KPRN_POOL_TOPK && c.topk < 1) ||
        c.precision < 0 || c.precision > 3) {
        set_error("invalid kprn_config");
        return KPRN_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no usable CUDA device (%s); kprn_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
        return KPRN_ECUDA;
    }
    if (c.device < 0 || c.device >= ndev) { set_error("device %d out of range (%d devices)", c.device, ndev); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(c.device));
    kprn_ctx* ctx = new (std::nothrow) kprn_ctx();
    if (!ctx) { set_error("out of host memory"); return KPRN_ENOMEM; }
    ctx->cfg = c;
    ctx->device = c.device;
    make_layout(c, ctx->lay);
    int rc = KPRN_OK;
    do {
        if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = KPRN_ECUDA; break; }
        ctx->stream = ctx->own_stream;
        size_t bytes = sizeof(float) * (size_t)ctx->lay.total;
        if (cudaMalloc(&ctx->d_params, bytes) != cudaSuccess || cudaMalloc(&ctx->d_grads, bytes) != cudaSuccess ||
            cudaMalloc(&ctx->d_err, sizeof(int32_t)) != cudaSuccess) { rc = KPRN_ENOMEM; break; }
        cudaMemsetAsync(ctx->d_params, 0, bytes, ctx->stream);
        cudaMemsetAsync(ctx->d_grads, 0, bytes, ctx->stream);
        cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = KPRN_ECUDA; break; }
    } while (0);
    if (rc != KPRN_OK) {
        set_error("kprn_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        kprn_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return KPRN_OK;
}

int kprn_destroy(kprn_ctx* ctx) {
    if (!ctx) return KPRN_OK;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    kprn_comm_destroy(ctx);
    train_destroy(ctx);
    tc_destroy(ctx);
    if (ctx->tmp_comm.p) cudaFree(ctx->tmp_comm.p);
    if (ctx->tmp_pool.p) cudaFree(ctx->tmp_pool.p);
    if (ctx->ws_part.p) cudaFree(ctx->ws_part.p);
    DBuf* bufs[] = {&ctx->ids_buf, &ctx->seg_buf, &ctx->path_logits, &ctx->seg_logits, &ctx->scores, &ctx->dscores,
                    &ctx->dseg_logits, &ctx->labels, &ctx->tmp_small, &ctx->act, &ctx->cst, &ctx->hst, &ctx->z1,
                    &ctx->ws_x, &ctx->ws_xw, &ctx->ws_h0, &ctx->ws_h1, &ctx->ws_c, &ctx->ws_dg, &ctx->ws_dx,
                    &ctx->ws_dh, &ctx->ws_dc, &ctx->ws_ds, &ctx->ws_dz, &ctx->pair_buf};
    for (DBuf* b : bufs) if (b->p) cudaFree(b->p);
    for (auto& pr : ctx->prof_pending) { cudaEventDestroy(pr.e0); cudaEventDestroy(pr.e1); }
    if (ctx->d_params) cudaFree(ctx->d_params);
    if (ctx->d_grads) cudaFree(ctx->d_grads);
    if (ctx->d_m) cudaFree(ctx->d_m);
    if (ctx->d_v) cudaFree(ctx->d_v);
    if (ctx->d_err) cudaFree(ctx->d_err);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return KPRN_OK;
}

int kprn_set_stream(kprn_ctx* ctx, void* cuda_stream) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return KPRN_OK;
}

int kprn_sync(kprn_ctx* ctx) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_err_flag(ctx);
}

int kprn_param_count(kprn_ctx* ctx, int64_t* n) {
    if (!ctx || !n) { set_error("null argument"); return KPRN_EINVAL; }
    *n = ctx->lay.total;
    return KPRN_OK;
}

int kprn_param_offsets(kprn_ctx* ctx, int64_t* o) {
    if (!ctx || !o) { set_error("null argument"); return KPRN_EINVAL; }
    const Layout& l = ctx->lay;
    o[0] = l.E_t; o[1] = l.E_e; o[2] = l.E_r; o[3] = l.Wx; o[4] = l.bx; o[5] = l.Wh; o[6] = l.bh; o[7] = l.W1;
    o[8] = l.b1; o[9] = l.Wo; o[10] = l.bo;
    return KPRN_OK;
}

int kprn_set_params(kprn_ctx* ctx, const float* host, int64_t n) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_flush(ctx));       // rows a sparse training step left behind are brought up to date first
    if (!host || n != ctx->lay.total) { set_error("set_params: expected %lld floats", (long long)ctx->lay.total); return KPRN_EINVAL; }
    KPRN_CUDA(cudaMemcpyAsync(ctx->d_params, host, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    return params_changed(ctx);
}

int kprn_set_params_f64(kprn_ctx* ctx, const double* host, int64_t n) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_flush(ctx));       // rows a sparse training step left behind are brought up to date first
    if (!host || n != ctx->lay.total) { set_error("set_params_f64: expected %lld doubles", (long long)ctx->lay.total); return KPRN_EINVAL; }
    const int64_t chunk = 1ll << 24;
    KPRN_TRY(ensure(ctx->ws_dx, sizeof(double) * (size_t)(n < chunk ? n : chunk)));
    for (int64_t o = 0; o < n; o += chunk) {
        int64_t m = (n - o) < chunk ? (n - o) : chunk;
        KPRN_CUDA(cudaMemcpyAsync(ctx->ws_dx.p, host + o, sizeof(double) * m, cudaMemcpyHostToDevice, ctx->stream));
        f64_to_f32_kernel<<<(unsigned)cdiv(m, 256), 256, 0, ctx->stream>>>(ctx->ws_dx.as<double>(), ctx->d_params + o, m);
        KPRN_LAUNCH_CHECK(ctx);
        KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return params_changed(ctx);
}

int kprn_get_params(kprn_ctx* ctx, float* host, int64_t n) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_flush(ctx));       // rows a sparse training step left behind are brought up to date first
    if (!host || n != ctx->lay.total) { set_error("get_params: expected %lld floats", (long long)ctx->lay.total); return KPRN_EINVAL; }
    KPRN_CUDA(cudaMemcpyAsync(host, ctx->d_params, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    return KPRN_OK;
}

int kprn_get_grads(kprn_ctx* ctx, float* host, int64_t n) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_materialize_grads(ctx));   // dense view of a sparse step's entity gradient
    if (!host || n != ctx->lay.total) { set_error("get_grads: expected %lld floats", (long long)ctx->lay.total); return KPRN_EINVAL; }
    KPRN_CUDA(cudaMemcpyAsync(host, ctx->d_grads, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    return KPRN_OK;
}

int kprn_init_uniform(kprn_ctx* ctx, float scale, uint64_t seed) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_flush(ctx));       // rows a sparse training step left behind are brought up to date first
    int64_t n = ctx->lay.total;
    init_uniform_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(ctx->d_params, n, scale, seed);
    KPRN_LAUNCH_CHECK(ctx);
    return params_changed(ctx);
}

int kprn_params_device(kprn_ctx* ctx, float** dptr, int64_t* n) {
    if (!ctx || !dptr || !n) { set_error("null argument"); return KPRN_EINVAL; }
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_flush(ctx));
    *dptr = ctx->d_params; *n = ctx->lay.total;
    ctx->params_version++;   // the caller may write through the pointer
    return KPRN_OK;
}

int kprn_params_dirty(kprn_ctx* ctx) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    return params_changed(ctx);
}

int kprn_grads_device(kprn_ctx* ctx, float** dptr, int64_t* n) {
    if (!ctx || !dptr || !n) { set_error("null argument"); return KPRN_EINVAL; }
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_materialize_grads(ctx));
    *dptr = ctx->d_grads; *n = ctx->lay.total;
    return KPRN_OK;
}

int kprn_zero_row(kprn_ctx* ctx, int table, int64_t row) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_flush(ctx));       // rows a sparse training step left behind are brought up to date first
    const kprn_config& c = ctx->cfg;
    int64_t V, off; int d;
    switch (table) {
        case KPRN_TABLE_TYPE: V = c.V_t; d = c.d_t; off = ctx->lay.E_t; break;
        case KPRN_TABLE_ENTITY: V = c.V_e; d = c.d_e; off = ctx->lay.E_e; break;
        case KPRN_TABLE_RELATION: V = c.V_r; d = c.d_r; off = ctx->lay.E_r; break;
        default: set_error("zero_row: bad table %d", table); return KPRN_EINVAL;
    }
    if (row < 1 || row > V) { set_error("zero_row: row %lld out of [1,%lld]", (long long)row, (long long)V); return KPRN_EINVAL; }
    KPRN_CUDA(cudaMemsetAsync(ctx->d_params + off + (row - 1) * d, 0, sizeof(float) * d, ctx->stream));
    return params_changed(ctx);
}

int kprn_forward_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, int32_t train,
                        float* d_out_scores, float* d_out_logits) {
    KPRN_TRY(check_ctx(ctx));
    if (!d_ids || !d_seg) { set_error("forward_device: null ids/seg"); return KPRN_EINVAL; }
    return forward_device_impl(ctx, d_ids, d_seg, N, S, train != 0, d_out_logits != nullptr, d_out_scores, d_out_logits);
}

}  // extern "C"
namespace kprn {
static bool is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

// host -> device copy on the ctx stream; pageable sources go through the ctx's pinned staging buffer at `stage_off`
int host_to_device(kprn_ctx* ctx, void* dst, const void* src, size_t bytes, size_t stage_off) {
    if (is_pinned(src)) {
        KPRN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return KPRN_OK;
    }
    KPRN_TRY(ensure_stage(ctx, stage_off + bytes));
    memcpy((char*)ctx->h_stage + stage_off, src, bytes);
    KPRN_CUDA(cudaMemcpyAsync(dst, (char*)ctx->h_stage + stage_off, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return KPRN_OK;
}
}  // namespace kprn
extern "C" {

static int forward_host_common(kprn_ctx* ctx, int64_t N, int64_t S, int32_t train, float* out_scores, float* out_logits) {
    // ids / seg are already on the device in ctx->ids_buf / ctx->seg_buf
    const kprn_config& c = ctx->cfg;
    // a flag left by an earlier kprn_forward_device that nobody collected (kprn_sync) must not fail this call
    KPRN_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream));
    KPRN_TRY(forward_device_impl(ctx, ctx->ids_buf.as<int32_t>(), ctx->seg_buf.as<int32_t>(), N, S, train != 0, out_logits != nullptr, nullptr, nullptr));
    size_t need = sizeof(float) * (size_t)(S + S * c.L);
    KPRN_TRY(ensure_stage(ctx, need));
    float* hs = (float*)ctx->h_stage;
    if (out_scores) KPRN_CUDA(cudaMemcpyAsync(hs, ctx->scores.p, sizeof(float) * S, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_logits) KPRN_CUDA(cudaMemcpyAsync(hs + S, ctx->seg_logits.p, sizeof(float) * S * c.L, cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_TRY(check_err_flag(ctx));   // synchronises
    if (out_scores) memcpy(out_scores, hs, sizeof(float) * S);
    if (out_logits) memcpy(out_logits, hs + S, sizeof(float) * S * c.L);
    return KPRN_OK;
}

int kprn_forward(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S, int32_t paths_per_seg,
                 int32_t train, float* out_scores, float* out_logits) {
    KPRN_TRY(check_ctx(ctx));
    const kprn_config& c = ctx->cfg;
    if (!ids) { set_error("forward: null ids"); return KPRN_EINVAL; }
    if (!seg_offsets) {
        if (paths_per_seg <= 0 || N % paths_per_seg != 0) { set_error("forward: dense bucket needs N %% paths_per_seg == 0"); return KPRN_EINVAL; }
        S = N / paths_per_seg;
    }
    if (N <= 0 || S <= 0) { set_error("empty batch: N=%lld S=%lld", (long long)N, (long long)S); return KPRN_EINVAL; }
    size_t ids_bytes = sizeof(int32_t) * (size_t)N * c.T * 3, seg_bytes = sizeof(int32_t) * (size_t)(S + 1);
    KPRN_TRY(ensure(ctx->ids_buf, ids_bytes));
    KPRN_TRY(ensure(ctx->seg_buf, seg_bytes));
    // pinned caller memory is DMA'd directly; pageable memory is staged through pinned memory first
    if (!is_pinned(ids) || (seg_offsets && !is_pinned(seg_offsets))) KPRN_TRY(ensure_stage(ctx, ids_bytes + seg_bytes));
    KPRN_TRY(host_to_device(ctx, ctx->ids_buf.p, ids, ids_bytes, 0));
    if (seg_offsets) {
        KPRN_TRY(host_to_device(ctx, ctx->seg_buf.p, seg_offsets, seg_bytes, ids_bytes));
    } else {
        dense_seg_kernel<<<(unsigned)cdiv(S + 1, 256), 256, 0, ctx->stream>>>(ctx->seg_buf.as<int32_t>(), S, paths_per_seg);
        KPRN_LAUNCH_CHECK(ctx);
    }
    return forward_host_common(ctx, N, S, train, out_scores, out_logits);
}

int kprn_forward_f64ids(kprn_ctx* ctx, const double* ids_f64, int64_t B, int64_t P, int32_t train, float* out_scores,
                        float* out_logits) {
    KPRN_TRY(check_ctx(ctx));
    const kprn_config& c = ctx->cfg;
    if (!ids_f64 || B <= 0 || P <= 0) { set_error("forward_f64ids: bad arguments"); return KPRN_EINVAL; }
    int64_t N = B * P, n = N * c.T * 3;
    KPRN_TRY(ensure(ctx->ids_buf, sizeof(int32_t) * n));
    KPRN_TRY(ensure(ctx->seg_buf, sizeof(int32_t) * (B + 1)));
    if (is_pinned(ids_f64)) {
        // page-locked DoubleTensor storage (kprn_pinned_alloc): DMA the doubles as they are, convert on the device
        KPRN_TRY(ensure(ctx->ws_dx, sizeof(double) * n));
        KPRN_CUDA(cudaMemcpyAsync(ctx->ws_dx.p, ids_f64, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        f64ids_to_i32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ctx->stream>>>(ctx->ws_dx.as<double>(), ctx->ids_buf.as<int32_t>(), n);
        KPRN_LAUNCH_CHECK(ctx);
    } else {
        // pageable storage has to be staged anyway: convert while staging (half the DMA bytes, no device-side pass)
        KPRN_TRY(ensure_stage(ctx, sizeof(int32_t) * n));
        int32_t* st = (int32_t*)ctx->h_stage;
        for (int64_t i = 0; i < n; ++i) st[i] = (int32_t)(ids_f64[i] + 0.5);      // ids are positive integers stored as doubles
        KPRN_CUDA(cudaMemcpyAsync(ctx->ids_buf.p, st, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    dense_seg_kernel<<<(unsigned)cdiv(B + 1, 256), 256, 0, ctx->stream>>>(ctx->seg_buf.as<int32_t>(), B, (int32_t)P);
    KPRN_LAUNCH_CHECK(ctx);
    return forward_host_common(ctx, N, B, train, out_scores, out_logits);
}

int kprn_pinned_alloc(int64_t bytes, void** host_ptr) {
    if (!host_ptr || bytes <= 0) { set_error("pinned_alloc: bad arguments"); return KPRN_EINVAL; }
    cudaError_t e = cudaMallocHost(host_ptr, (size_t)bytes);
    if (e != cudaSuccess) { set_error("cudaMallocHost(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e)); return KPRN_ENOMEM; }
    return KPRN_OK;
}

int kprn_pinned_free(void* host_ptr) {
    if (host_ptr) cudaFreeHost(host_ptr);
    return KPRN_OK;
}

int kprn_get_path_logits(kprn_ctx* ctx, float* host, int64_t n_floats) {
    KPRN_TRY(check_ctx(ctx));
    if (!ctx->have_fwd) { set_error("get_path_logits: no forward yet"); return KPRN_ESTATE; }
    if (ctx->cls_only) { set_error("get_path_logits: the last forward was scores-only (request logits or train=1 to materialise all L columns)"); return KPRN_ESTATE; }
    if (!host || n_floats != ctx->N * ctx->cfg.L) { set_error("get_path_logits: expected %lld floats", (long long)(ctx->N * ctx->cfg.L)); return KPRN_EINVAL; }
    KPRN_CUDA(cudaMemcpyAsync(host, ctx->path_logits.p, sizeof(float) * n_floats, cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    return KPRN_OK;
}

static int read_loss(kprn_ctx* ctx, double* d_loss, float* loss) {
    double h = 0.0;
    KPRN_CUDA(cudaMemcpyAsync(&h, d_loss, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (loss) *loss = (float)h;
    return KPRN_OK;
}

}  // extern "C"
namespace kprn {
int loss_bce_device(kprn_ctx* ctx, const float* d_labels, int64_t S, double** d_loss_out) {
    if (S != ctx->S) { set_error("loss_bce: expected %lld labels", (long long)ctx->S); return KPRN_EINVAL; }
    KPRN_TRY(ensure(ctx->dscores, sizeof(float) * S));
    KPRN_TRY(ensure(ctx->tmp_small, 64));
    double* d_loss = ctx->tmp_small.as<double>();
    KPRN_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(double), ctx->stream));
    bce_kernel<<<(unsigned)(cdiv(S, 256) < 592 ? cdiv(S, 256) : 592), 256, 0, ctx->stream>>>(
        ctx->scores.as<float>(), d_labels, S, d_loss, ctx->dscores.as<float>());
    KPRN_LAUNCH_CHECK(ctx);
    ctx->have_loss_grad = true;
    *d_loss_out = d_loss;
    return KPRN_OK;
}

__global__ void check_pairs_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n, int64_t S, int32_t* err) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && (a[k] < 0 || a[k] >= S || b[k] < 0 || b[k] >= S)) atomicExch(err, 3);
}

// nn.BPRLoss on device-resident pair lists (0-based segment indices); bad indices raise the device error flag
int loss_bpr_device(kprn_ctx* ctx, const int32_t* d_pos, const int32_t* d_neg, int64_t n_pairs, int on_logits, double** d_loss_out) {
    const int64_t S = ctx->S;
    KPRN_TRY(ensure(ctx->dscores, sizeof(float) * S));
    KPRN_TRY(ensure(ctx->tmp_small, 64));
    double* d_loss = ctx->tmp_small.as<double>();
    KPRN_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(double), ctx->stream));
    KPRN_CUDA(cudaMemsetAsync(ctx->dscores.p, 0, sizeof(float) * S, ctx->stream));
    check_pairs_kernel<<<(unsigned)cdiv(n_pairs, 256), 256, 0, ctx->stream>>>(d_pos, d_neg, n_pairs, S, ctx->d_err);
    KPRN_LAUNCH_CHECK(ctx);
    bpr_kernel<<<(unsigned)(cdiv(n_pairs, 256) < 592 ? cdiv(n_pairs, 256) : 592), 256, 0, ctx->stream>>>(
        ctx->scores.as<float>(), ctx->seg_logits.as<float>(), ctx->cfg.L, ctx->cfg.class_id - 1, d_pos, d_neg, n_pairs,
        on_logits, S, d_loss, ctx->dscores.as<float>());
    KPRN_LAUNCH_CHECK(ctx);
    ctx->have_loss_grad = true;
    *d_loss_out = d_loss;
    return KPRN_OK;
}

int backward_device(kprn_ctx* ctx, float scale) {
    if (!ctx->have_fwd || !ctx->have_train_cache) { set_error("backward: needs a forward with train=1 first"); return KPRN_ESTATE; }
    if (!ctx->have_loss_grad) { set_error("backward: no loss gradient (call kprn_loss_* first or pass grad_scores)"); return KPRN_ESTATE; }
    return ctx->tc_train_fwd ? tc_backward(ctx, scale) : simple_backward(ctx, scale);
}
}  // namespace kprn
extern "C" {

int kprn_loss_bce(kprn_ctx* ctx, const float* labels, int64_t S, float* loss) {
    KPRN_TRY(check_ctx(ctx));
    if (!ctx->have_fwd) { set_error("loss_bce: no forward yet"); return KPRN_ESTATE; }
    if (!labels || S != ctx->S) { set_error("loss_bce: expected %lld labels", (long long)ctx->S); return KPRN_EINVAL; }
    KPRN_TRY(ensure(ctx->labels, sizeof(float) * S));
    KPRN_CUDA(cudaMemcpyAsync(ctx->labels.p, labels, sizeof(float) * S, cudaMemcpyHostToDevice, ctx->stream));
    double* d_loss = nullptr;
    // tmp_small is also used by the RNN mask; the loss scalar lives in its first 8 bytes after forward is done
    KPRN_TRY(loss_bce_device(ctx, ctx->labels.as<float>(), S, &d_loss));
    return read_loss(ctx, d_loss, loss);
}

int kprn_loss_bpr(kprn_ctx* ctx, const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs, int32_t on_logits, float* loss) {
    KPRN_TRY(check_ctx(ctx));
    if (!ctx->have_fwd) { set_error("loss_bpr: no forward yet"); return KPRN_ESTATE; }
    if (!pos_seg || !neg_seg || n_pairs <= 0) { set_error("loss_bpr: bad arguments"); return KPRN_EINVAL; }
    for (int64_t i = 0; i < n_pairs; ++i)
        if (pos_seg[i] < 0 || pos_seg[i] >= ctx->S || neg_seg[i] < 0 || neg_seg[i] >= ctx->S) {
            set_error("loss_bpr: segment index out of range at pair %lld", (long long)i);
            return KPRN_EINVAL;
        }
    KPRN_TRY(ensure(ctx->pair_buf, sizeof(int32_t) * 2 * n_pairs));
    int32_t* dp = ctx->pair_buf.as<int32_t>();
    KPRN_CUDA(cudaMemcpyAsync(dp, pos_seg, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    KPRN_CUDA(cudaMemcpyAsync(dp + n_pairs, neg_seg, sizeof(int32_t) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    double* d_loss = nullptr;
    KPRN_TRY(loss_bpr_device(ctx, dp, dp + n_pairs, n_pairs, on_logits, &d_loss));
    return read_loss(ctx, d_loss, loss);
}

int kprn_backward(kprn_ctx* ctx, const float* grad_scores, float scale) {
    KPRN_TRY(check_ctx(ctx));
    if (!ctx->have_fwd || !ctx->have_train_cache) { set_error("backward: needs a forward with train=1 first"); return KPRN_ESTATE; }
    // kprn_forward_device only raises the device flag on a bad id / bad segment table: read it before the scatter uses the ids
    KPRN_TRY(check_err_flag(ctx));
    const int64_t S = ctx->S;
    if (grad_scores) {
        KPRN_TRY(ensure(ctx->dscores, sizeof(float) * S));
        KPRN_TRY(ensure(ctx->labels, sizeof(float) * S));
        KPRN_CUDA(cudaMemcpyAsync(ctx->labels.p, grad_scores, sizeof(float) * S, cudaMemcpyHostToDevice, ctx->stream));
        chain_scores_kernel<<<(unsigned)cdiv(S, 256), 256, 0, ctx->stream>>>(ctx->scores.as<float>(), ctx->labels.as<float>(), S,
                                                                            ctx->dscores.as<float>());
        KPRN_LAUNCH_CHECK(ctx);
        ctx->have_loss_grad = true;
    } else if (!ctx->have_loss_grad) {
        set_error("backward: no cached loss gradient (call kprn_loss_* first or pass grad_scores)");
        return KPRN_ESTATE;
    }
    // reference-style flow (zeroGradParameters / backward / getParameters): the gradient accumulates in the dense flat buffer
    KPRN_TRY(train_materialize_grads(ctx));
    return backward_device(ctx, scale);
}

int kprn_zero_grad(kprn_ctx* ctx) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_materialize_grads(ctx));   // leaves the sparse state; the memset below then clears everything
    KPRN_CUDA(cudaMemsetAsync(ctx->d_grads, 0, sizeof(float) * (size_t)ctx->lay.total, ctx->stream));
    return KPRN_OK;
}

static int grad_norm_device(kprn_ctx* ctx, float* norm) {
    KPRN_TRY(ensure(ctx->tmp_small, 64));
    double* d = ctx->tmp_small.as<double>() + 1;
    KPRN_CUDA(cudaMemsetAsync(d, 0, sizeof(double), ctx->stream));
    sumsq_kernel<<<1184, 256, 0, ctx->stream>>>(ctx->d_grads, ctx->lay.total, d);
    KPRN_LAUNCH_CHECK(ctx);
    double h = 0.0;
    KPRN_CUDA(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    *norm = (float)sqrt(h);
    return KPRN_OK;
}

int kprn_grad_norm(kprn_ctx* ctx, float* norm) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_materialize_grads(ctx));   // dense view of a sparse step's entity gradient
    if (!norm) { set_error("null argument"); return KPRN_EINVAL; }
    return grad_norm_device(ctx, norm);
}

int kprn_grad_scale(kprn_ctx* ctx, float s) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_materialize_grads(ctx));   // dense view of a sparse step's entity gradient
    scale_kernel<<<1184, 256, 0, ctx->stream>>>(ctx->d_grads, ctx->lay.total, s);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

int kprn_grad_add_params(kprn_ctx* ctx, float l2) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_TRY(train_materialize_grads(ctx));   // dense view of a sparse step's entity gradient
    axpy_kernel<<<1184, 256, 0, ctx->stream>>>(ctx->d_grads, ctx->d_params, ctx->lay.total, l2);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

int kprn_adam_step(kprn_ctx* ctx, float lr, float b1, float b2, float eps) {
    KPRN_TRY(check_ctx(ctx));
    const int64_t n = ctx->lay.total;
    KPRN_TRY(train_flush(ctx));               // the dense kernel assumes every row's (x, m, v) is current
    KPRN_TRY(train_materialize_grads(ctx));
    if (!ctx->d_m || !ctx->d_v) {
        if (ctx->d_m) { cudaFree(ctx->d_m); ctx->d_m = nullptr; }
        if (ctx->d_v) { cudaFree(ctx->d_v); ctx->d_v = nullptr; }
        if (cudaMalloc(&ctx->d_m, sizeof(float) * n) != cudaSuccess || cudaMalloc(&ctx->d_v, sizeof(float) * n) != cudaSuccess) {
            cudaGetLastError();
            if (ctx->d_m) { cudaFree(ctx->d_m); ctx->d_m = nullptr; }
            ctx->d_v = nullptr;
            set_error("adam: out of device memory for optimizer state");
            return KPRN_ENOMEM;
        }
        KPRN_CUDA(cudaMemsetAsync(ctx->d_m, 0, sizeof(float) * n, ctx->stream));
        KPRN_CUDA(cudaMemsetAsync(ctx->d_v, 0, sizeof(float) * n, ctx->stream));
    }
    ctx->adam_t++;
    double bc1 = 1.0 - pow((double)b1, (double)ctx->adam_t), bc2 = 1.0 - pow((double)b2, (double)ctx->adam_t);
    float step = (float)((double)lr * sqrt(bc2) / bc1);
    adam_kernel<<<1184, 256, 0, ctx->stream>>>(ctx->d_params, ctx->d_grads, ctx->d_m, ctx->d_v, n, step, b1, b2, eps);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(train_note_dense_step(ctx, step));
    return params_changed(ctx);
}

static void batch_opts(kprn_train_opts& o, int loss, int32_t regularize, int32_t use_clip, float clip_norm, float l2, float lr, float b1,
                       float b2, float eps) {
    memset(&o, 0, sizeof(o));
    o.loss = loss; o.regularize = regularize; o.use_clip = use_clip; o.clip_norm = clip_norm; o.l2 = l2;
    o.lr = lr; o.beta1 = b1; o.beta2 = b2; o.eps = eps; o.grad_scale = 1.f;
}

int kprn_train_batch_bce(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg, int64_t N, int64_t S, const float* labels,
                         const int64_t* pad3, int32_t regularize, int32_t use_clip, float clip_norm, float l2, float lr,
                         float b1, float b2, float eps, float* loss) {
    kprn_train_opts o;
    batch_opts(o, 0, regularize, use_clip, clip_norm, l2, lr, b1, b2, eps);
    return kprn_train_step(ctx, ids, seg, N, S, labels, nullptr, nullptr, 0, pad3, &o, loss);
}

int kprn_train_batch_bpr(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg, int64_t N, int64_t S, const int32_t* pos_seg,
                         const int32_t* neg_seg, int64_t n_pairs, int32_t apply_to_logits, const int64_t* pad3, int32_t regularize,
                         int32_t use_clip, float clip_norm, float l2, float lr, float b1, float b2, float eps, float* loss) {
    kprn_train_opts o;
    batch_opts(o, 1, regularize, use_clip, clip_norm, l2, lr, b1, b2, eps);
    o.bpr_on_logits = apply_to_logits;
    return kprn_train_step(ctx, ids, seg, N, S, nullptr, pos_seg, neg_seg, n_pairs, pad3, &o, loss);
}

int kprn_set_option(kprn_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) { set_error("null argument"); return KPRN_EINVAL; }
    if (!strcmp(key, "dedup")) { ctx->opt_dedup = value < 0 ? -1 : (value != 0); return KPRN_OK; }
    if (!strcmp(key, "prefix")) { ctx->opt_prefix = value < 0 ? -1 : (value != 0); return KPRN_OK; }
    if (!strcmp(key, "path_sharded")) { ctx->opt_path_sharded = value > 0; return KPRN_OK; }
    set_error("set_option: unknown key '%s'", key);
    return KPRN_EINVAL;
}

int kprn_launch_count(kprn_ctx* ctx, int64_t* n, int32_t reset) {
    if (!ctx || !n) { set_error("null argument"); return KPRN_EINVAL; }
    *n = ctx->launches;
    if (reset) ctx->launches = 0;
    return KPRN_OK;
}

int kprn_profile(kprn_ctx* ctx, int32_t enable) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    ctx->profiling = enable != 0;
    return KPRN_OK;
}

int kprn_profile_read(kprn_ctx* ctx, float* ms3, int64_t* n3, int32_t reset) {
    KPRN_TRY(check_ctx(ctx));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& pr : ctx->prof_pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pr.e0, pr.e1) == cudaSuccess) { ctx->prof_ms[pr.slot] += t; ctx->prof_launches[pr.slot]++; }
        cudaEventDestroy(pr.e0); cudaEventDestroy(pr.e1);
    }
    ctx->prof_pending.clear();
    for (int i = 0; i < 3; ++i) {
        if (ms3) ms3[i] = ctx->prof_ms[i];
        if (n3) n3[i] = ctx->prof_launches[i];
        if (reset) { ctx->prof_ms[i] = 0.f; ctx->prof_launches[i] = 0; }
    }
    return KPRN_OK;
}

}  // extern "C"
