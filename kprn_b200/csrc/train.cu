// train.cu -- one MyOptimizer:trainBatch (SP/model/optimizer/MyOptimizer.lua:177-221) on the device, with the entity table
// handled ROW-SPARSELY: a batch of 65 536 paths touches ~2e5 of the 2.85 M entity rows, yet the reference's dense flow
// (zeroGradParameters over 730 M floats, dense optim.adam over parameters + m + v: 20 GB of HBM traffic per step) pays for
// all of them.  Here
//   * the entity gradient lives in a COMPACT buffer gE[slot][d_e], one slot per distinct entity row of the batch
//     (dedup_build mode 1); only the small dense tensors (type/relation tables, Wx, bx, Wh, head) use the flat gradient buffer;
//   * Adam is evaluated LAZILY but stays the reference's DENSE optim.adam: a row that received no gradient at step k still
//     decays (m *= b1, v *= b2) and moves (x -= step_k m/(sqrt v + eps)) in the reference; that update depends on nothing but
//     the row's own state, so it is postponed and replayed, step by step with the very same arithmetic (adam.cuh), the next
//     time the row is read -- before the forward of the batch that touches it, or by train_flush() before anything else
//     looks at the parameters (scoring forward, kprn_get_params, checkpoint, broadcast).  Results are bit-identical to the
//     dense kernel given the same gradients;
//   * across GPUs (one process each, comm.cu) the small tensors are all-reduced (4 MB) and the row gradients travel as an
//     all-gather of (row ids, row gradients); every rank then reduces the contributions of each row IN RANK ORDER and applies
//     the same update, so the replicas stay bit-identical without ever moving the 2.9 GB dense gradient.
#include <math.h>
#include <string.h>

#include <vector>

#include "adam.cuh"
#include "common.cuh"

namespace kprn {

struct TrainState {
    DedupBufs es;                 // distinct entity rows of the batch: es.ids[slot] = 1-based row, es.count[0] = K
    DBuf gE;                      // [slot][d_e] fp32
    int32_t* last_t = nullptr;    // [V_e] optimizer step up to which the row's (x, m, v) are current; 0 = never touched sparsely
    int64_t base_t = 0;           // every row is current at least up to this step (dense step / flush)
    bool dense_ever = false;      // a dense optimizer step ran: rows with last_t == 0 may still carry momentum
    bool pending = false;         // some rows are behind ctx->adam_t
    bool sparse_grads = false;    // the last backward left the entity gradient in gE (kprn_get_grads materialises it)
    int64_t sparse_N = 0;         // paths of the batch gE belongs to
    float* steps_pinned = nullptr; // step size of optimizer step k at [k] (k = 1..adam_t), page-locked mirror
    DBuf steps;                   // device copy read by the replay
    int64_t steps_cap = 0;
    float b1 = 0.f, b2 = 0.f, eps = 0.f;
    // multi-GPU exchange
    DBuf send_ids, recv_ids, recv_g, contrib, norm2;
    DedupBufs us;                 // union of the ranks' row lists
    // kprn_train_prefetch_device: the ids of the NEXT batch, known early to a host with a look-ahead batcher.  The step that is
    // running when the hint is pending brings that batch's rows up to date at its own tail, on the side stream beside the
    // weight-gradient kernel, so the next step's catch-up (critical path: the projection reads those rows) finds nothing to do.
    const int32_t* hint_ids = nullptr; int64_t hint_N = 0;
    DedupBufs hs;                 // distinct entity rows of the hinted batch
    int64_t last_K = 0, last_U = 0, last_maxK = 0;
    int64_t pads[3] = {0, 0, 0};  // pad rows of the last training step: re-zeroed after a flush (they are zero after every step)
    bool have_pads = false;
};

static TrainState* tr_get(kprn_ctx* ctx) {
    if (!ctx->tr) ctx->tr = new TrainState();
    return ctx->tr;
}

// ---------------------------------------------------------------------------------------------
// kernels

// replay of the postponed zero-gradient steps (t0, t1] of one element
__device__ __forceinline__ void adam_replay(float& x, float& m, float& v, int64_t t0, int64_t t1, const float* __restrict__ steps,
                                            float b1, float b2, float eps) {
    if (m == 0.f && v == 0.f) return;                 // the update is exactly zero while m = v = 0
    for (int64_t k = t0 + 1; k <= t1; ++k) adam_elem(x, m, v, 0.f, steps[k], b1, b2, eps);
}
__device__ __forceinline__ void split_store(float v, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t idx) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[idx] = h;
    lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

struct RowsArgs {
    const int32_t* rows;        // 1-based entity row per slot
    const int32_t* count;       // number of slots (device)
    float* x; float* m; float* v;          // entity table slices of the flat vectors
    int32_t* last_t; int64_t base_t; int dense_ever;
    int64_t t_now;              // bring rows up to t_now - 1 by replay, then (apply) take step t_now with the real gradient
    int apply;
    const float* g; int64_t g_stride;      // compact gradient [slot][d_e]           (world == 1)
    const float* recv_g; const int32_t* contrib; int world; int64_t maxK;   // all-gathered gradients + [slot][world] index table
    const float* gscale;        // optional device scalar multiplying every gradient (global-norm clip)
    const float* steps; float b1, b2, eps;
    int d_e; int KE;
    __nv_bfloat16* E_hi; __nv_bfloat16* E_lo;       // packed bf16 copies of the entity table (tensor-core forward), or null
};

// one WARP per slot (grid-stride over slots): catch the row up, optionally apply this step's gradient, refresh the packed
// copies.  NV4 > 0: d_e = 128 NV4 -- a lane owns one float4 per 128 columns of x, m, v (and of the gradient): all loads of a
// row are issued before any arithmetic; NV4 = 0: any width, scalar.  Rows with nothing to replay are skipped by the catch-up
// pass before any load.
template <int NV4>
__global__ void __launch_bounds__(256, 4) adam_rows_kernel(RowsArgs a) {
    const int K = *a.count;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float gs = a.gscale ? *a.gscale : 1.f;
    for (int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < K; slot += warps) {
        const int64_t row = (int64_t)a.rows[slot] - 1;
        const int32_t lt = a.last_t[row];                  // every lane reads it before lane 0 rewrites it (the __syncwarp below)
        const int64_t t0 = lt > a.base_t ? lt : a.base_t;
        const bool may_have_state = lt != 0 || a.dense_ever;
        if (!a.apply && (t0 >= a.t_now - 1 || !may_have_state)) {       // nothing postponed for this row
            __syncwarp();
            continue;
        }
        if (NV4 > 0) {
            constexpr int NV = NV4 > 0 ? NV4 : 1;
            float4 x[NV], m[NV], v[NV], g[NV];
            float* px = a.x + row * a.d_e; float* pm = a.m + row * a.d_e; float* pv = a.v + row * a.d_e;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                x[q] = *reinterpret_cast<const float4*>(px + q * 128 + lane * 4);
                m[q] = *reinterpret_cast<const float4*>(pm + q * 128 + lane * 4);
                v[q] = *reinterpret_cast<const float4*>(pv + q * 128 + lane * 4);
            }
            if (a.apply) {
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    if (a.recv_g) {
                        g[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int r = 0; r < a.world; ++r) {       // fixed rank order: identical sums on every rank
                            const int32_t c = a.contrib[(int64_t)slot * a.world + r];
                            if (c >= 0) {
                                const float4 t = *reinterpret_cast<const float4*>(a.recv_g + ((int64_t)r * a.maxK + c) * a.d_e + q * 128 + lane * 4);
                                g[q].x += t.x; g[q].y += t.y; g[q].z += t.z; g[q].w += t.w;
                            }
                        }
                    } else {
                        g[q] = *reinterpret_cast<const float4*>(a.g + (int64_t)slot * a.g_stride + q * 128 + lane * 4);
                    }
                    if (a.gscale) { g[q].x *= gs; g[q].y *= gs; g[q].z *= gs; g[q].w *= gs; }
                }
            }
            const float step = a.apply ? a.steps[a.t_now] : 0.f;
            // replay: the postponed step outermost, the lane's 4 NV elements inside -- independent dependency chains for the
            // divide / square root of every step (adam_replay's early-out for m = v = 0 is only a shortcut: the zero-gradient
            // update of such an element is x - step * 0 / (0 + eps) = x, so the unconditional form gives the same bits)
            if (may_have_state)
                for (int64_t k = t0 + 1; k <= a.t_now - 1; ++k) {
                    const float sk = a.steps[k];
#pragma unroll
                    for (int q = 0; q < NV; ++q) {
                        float* xs = reinterpret_cast<float*>(&x[q]); float* ms = reinterpret_cast<float*>(&m[q]); float* vs = reinterpret_cast<float*>(&v[q]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) adam_elem(xs[e], ms[e], vs[e], 0.f, sk, a.b1, a.b2, a.eps);
                    }
                }
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                float* xs = reinterpret_cast<float*>(&x[q]); float* ms = reinterpret_cast<float*>(&m[q]); float* vs = reinterpret_cast<float*>(&v[q]);
                const float* gq = reinterpret_cast<const float*>(&g[q]);
                if (a.apply) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) adam_elem(xs[e], ms[e], vs[e], gq[e], step, a.b1, a.b2, a.eps);
                }
                *reinterpret_cast<float4*>(px + q * 128 + lane * 4) = x[q];
                *reinterpret_cast<float4*>(pm + q * 128 + lane * 4) = m[q];
                *reinterpret_cast<float4*>(pv + q * 128 + lane * 4) = v[q];
                if (a.E_hi) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_store(xs[e], a.E_hi, a.E_lo, row * a.KE + q * 128 + lane * 4 + e);
                }
            }
        } else {
            for (int k = lane; k < a.d_e; k += 32) {
                const int64_t i = row * a.d_e + k;
                float x = a.x[i], m = a.m[i], v = a.v[i];
                if (may_have_state) adam_replay(x, m, v, t0, a.t_now - 1, a.steps, a.b1, a.b2, a.eps);
                if (a.apply) {
                    float g;
                    if (a.recv_g) {
                        g = 0.f;
                        for (int r = 0; r < a.world; ++r) {
                            const int32_t c = a.contrib[(int64_t)slot * a.world + r];
                            if (c >= 0) g += a.recv_g[((int64_t)r * a.maxK + c) * a.d_e + k];
                        }
                    } else {
                        g = a.g[(int64_t)slot * a.g_stride + k];
                    }
                    if (a.gscale) g *= gs;
                    adam_elem(x, m, v, g, a.steps[a.t_now], a.b1, a.b2, a.eps);
                }
                a.x[i] = x; a.m[i] = m; a.v[i] = v;
                if (a.E_hi) split_store(x, a.E_hi, a.E_lo, row * a.KE + k);
            }
        }
        __syncwarp();
        if (lane == 0) a.last_t[row] = (int32_t)(a.apply ? a.t_now : a.t_now - 1);
    }
}

// flush: every row that is behind t_now is replayed up to t_now (a block per row, grid-stride over the table)
__global__ void adam_flush_kernel(float* __restrict__ x, float* __restrict__ m, float* __restrict__ v, int32_t* __restrict__ last_t,
                                  int64_t V, int d_e, int64_t base_t, int dense_ever, int64_t t_now, const float* __restrict__ steps,
                                  float b1, float b2, float eps, int KE, __nv_bfloat16* E_hi, __nv_bfloat16* E_lo) {
    for (int64_t row = blockIdx.x; row < V; row += gridDim.x) {
        const int32_t lt = last_t[row];
        const int64_t t0 = lt > base_t ? lt : base_t;
        if (t0 >= t_now || (lt == 0 && !dense_ever)) continue;
        for (int k = threadIdx.x; k < d_e; k += blockDim.x) {
            const int64_t i = row * d_e + k;
            float xi = x[i], mi = m[i], vi = v[i];
            adam_replay(xi, mi, vi, t0, t_now, steps, b1, b2, eps);
            x[i] = xi; m[i] = mi; v[i] = vi;
            if (E_hi) split_store(xi, E_hi, E_lo, row * KE + k);
        }
    }
}

__global__ void adam_dense_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                  int64_t n, float step, float b1, float b2, float eps, const float* __restrict__ gscale) {
    const float gs = gscale ? *gscale : 1.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float xi = x[i], mi = m[i], vi = v[i];
        float gi = g[i];
        if (gscale) gi *= gs;
        adam_elem(xi, mi, vi, gi, step, b1, b2, eps);
        x[i] = xi; m[i] = mi; v[i] = vi;
    }
}

__global__ void zero_rows_kernel(float* __restrict__ g, const int32_t* __restrict__ count, int d_e) {
    const int64_t n = (int64_t)(*count) * d_e;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] = 0.f;
}
// sum of squares of the compact rows (or of the rank-ordered reduced rows) -> out (double)
__global__ void rows_sumsq_kernel(const float* __restrict__ g, const int32_t* __restrict__ count, int d_e, const float* __restrict__ recv_g,
                                  const int32_t* __restrict__ contrib, int world, int64_t maxK, double* __restrict__ out) {
    const int64_t n = (int64_t)(*count) * d_e;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v;
        if (recv_g) {
            const int64_t slot = i / d_e; const int k = (int)(i % d_e);
            v = 0.f;
            for (int r = 0; r < world; ++r) { const int32_t c = contrib[slot * world + r]; if (c >= 0) v += recv_g[((int64_t)r * maxK + c) * d_e + k]; }
        } else v = g[i];
        acc += (double)v * (double)v;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
__global__ void range_sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += (double)x[i] * (double)x[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
// gscale = clip / ||g|| when the norm exceeds clip (MyOptimizer.lua:198-201), else 1
__global__ void clip_scale_kernel(const double* __restrict__ sumsq, float clip, float* __restrict__ gscale, float* __restrict__ norm_out) {
    const float nrm = (float)sqrt(*sumsq);
    *gscale = nrm > clip ? clip / nrm : 1.f;
    if (norm_out) *norm_out = nrm;
}
// send buffer of the row exchange: ids[0] = K, ids[1..K] = rows, the rest -1
__global__ void pack_ids_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ count, int32_t* __restrict__ out, int64_t n) {
    const int K = *count;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i == 0 ? K : (i - 1 < K ? rows[i - 1] : -1);
}
// keys of the union pass: entry (r, i) of the gathered lists, -1 beyond a rank's count
__global__ void union_keys_kernel(const int32_t* __restrict__ recv_ids, int world, int64_t maxK, int32_t* __restrict__ keys) {
    const int64_t n = (int64_t)world * maxK;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / maxK, i = e % maxK;
        const int32_t K = recv_ids[r * (maxK + 1)];
        keys[e] = i < K ? recv_ids[r * (maxK + 1) + 1 + i] : -1;
    }
}
__global__ void fill_i32_kernel(int32_t* __restrict__ p, int64_t n, int32_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// contrib[slot][r] = index i of rank r's entry for the slot's row (each rank lists a row at most once: no write conflicts)
__global__ void union_contrib_kernel(const int32_t* __restrict__ keys, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ slot_of,
                                     int world, int64_t maxK, int32_t* __restrict__ contrib) {
    const int64_t n = (int64_t)world * maxK;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        if (keys[e] < 0) continue;
        const uint32_t slot = slot_of[pos[e]];
        contrib[(int64_t)slot * world + e / maxK] = (int32_t)(e % maxK);
    }
}
// dense view of the compact gradient: gE_dense[row] = gE[slot]
__global__ void scatter_rows_kernel(const float* __restrict__ g, const int32_t* __restrict__ rows, const int32_t* __restrict__ count, int d_e,
                                    float* __restrict__ dense) {
    const int K = *count;
    for (int slot = blockIdx.x; slot < K; slot += gridDim.x) {
        const int64_t row = (int64_t)rows[slot] - 1;
        for (int k = threadIdx.x; k < d_e; k += blockDim.x) dense[row * d_e + k] = g[(int64_t)slot * d_e + k];
    }
}

// ---------------------------------------------------------------------------------------------
// host helpers

static int adam_state(kprn_ctx* ctx) {
    const int64_t n = ctx->lay.total;
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
    return KPRN_OK;
}

static float adam_step_size(float lr, float b1, float b2, int64_t t) {
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    return (float)((double)lr * sqrt(bc2) / bc1);
}

// records the step size of optimizer step ctx->adam_t (already incremented): steps[t] on the device, without a host sync
// (the value travels from a page-locked mirror that lives as long as the ctx)
static int steps_push(kprn_ctx* ctx, TrainState* tr, float step) {
    const int64_t t = ctx->adam_t;
    if (tr->steps_cap < t + 1) {
        const int64_t ncap = 2 * t + 1024;
        float* nh = nullptr;
        KPRN_CUDA(cudaMallocHost(&nh, sizeof(float) * (size_t)ncap));
        DBuf nb;
        if (ensure(nb, sizeof(float) * (size_t)ncap) != KPRN_OK) { cudaFreeHost(nh); return KPRN_ENOMEM; }
        KPRN_CUDA(cudaStreamSynchronize(ctx->stream));            // rare (capacity doubles): in-flight copies read the old mirror
        if (tr->steps_pinned) { memcpy(nh, tr->steps_pinned, sizeof(float) * (size_t)tr->steps_cap); cudaFreeHost(tr->steps_pinned); }
        else nh[0] = 0.f;
        if (tr->steps.p) {
            KPRN_CUDA(cudaMemcpyAsync(nb.p, tr->steps.p, sizeof(float) * (size_t)tr->steps_cap, cudaMemcpyDeviceToDevice, ctx->stream));
            KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(tr->steps.p);
        }
        tr->steps = nb; tr->steps_pinned = nh; tr->steps_cap = ncap;
    }
    tr->steps_pinned[t] = step;
    KPRN_CUDA(cudaMemcpyAsync(tr->steps.as<float>() + t, tr->steps_pinned + t, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return KPRN_OK;
}

// a dense optimizer step (kprn_adam_step) is about to run or just ran: see api.cu
int train_note_dense_step(kprn_ctx* ctx, float step) {
    TrainState* tr = tr_get(ctx);
    tr->dense_ever = true;
    tr->base_t = ctx->adam_t;             // every row is current after a dense step
    tr->pending = false;
    return steps_push(ctx, tr, step);
}

static void packed_entity(kprn_ctx* ctx, __nv_bfloat16** hi, __nv_bfloat16** lo, int* KE) { tc_packed_entity(ctx, hi, lo, KE); }
static int zero_pad_rows(kprn_ctx* ctx, const int64_t* pads);

int train_flush(kprn_ctx* ctx) {
    TrainState* tr = ctx->tr;
    if (!tr || !tr->pending) return KPRN_OK;
    const kprn_config& c = ctx->cfg;
    __nv_bfloat16 *hi = nullptr, *lo = nullptr; int KE = 0;
    packed_entity(ctx, &hi, &lo, &KE);
    const int threads = c.d_e >= 256 ? 256 : (c.d_e >= 128 ? 128 : 64);
    adam_flush_kernel<<<148 * 8, threads, 0, ctx->stream>>>(ctx->d_params + ctx->lay.E_e, ctx->d_m + ctx->lay.E_e, ctx->d_v + ctx->lay.E_e,
                                                           tr->last_t, c.V_e, c.d_e, tr->base_t, tr->dense_ever ? 1 : 0, ctx->adam_t,
                                                           tr->steps.as<float>(), tr->b1, tr->b2, tr->eps, KE, hi, lo);
    KPRN_LAUNCH_CHECK(ctx);
    tr->base_t = ctx->adam_t;
    tr->pending = false;
    if (tr->have_pads) KPRN_TRY(zero_pad_rows(ctx, tr->pads));     // the replay moved them; the reference holds them at zero
    return KPRN_OK;
}

void train_destroy(kprn_ctx* ctx) {
    TrainState* tr = ctx->tr;
    if (!tr) return;
    dedup_free(tr->es); dedup_free(tr->us); dedup_free(tr->hs);
    for (DBuf* b : {&tr->gE, &tr->steps, &tr->send_ids, &tr->recv_ids, &tr->recv_g, &tr->contrib, &tr->norm2})
        if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    if (tr->last_t) cudaFree(tr->last_t);
    if (tr->steps_pinned) cudaFreeHost(tr->steps_pinned);
    delete tr;
    ctx->tr = nullptr;
}

// dense flat view of the gradient for kprn_get_grads / kprn_grads_device after a sparse backward
int train_materialize_grads(kprn_ctx* ctx) {
    TrainState* tr = ctx->tr;
    if (!tr || !tr->sparse_grads) return KPRN_OK;
    const kprn_config& c = ctx->cfg;
    KPRN_CUDA(cudaMemsetAsync(ctx->d_grads + ctx->lay.E_e, 0, sizeof(float) * (size_t)c.V_e * c.d_e, ctx->stream));
    scatter_rows_kernel<<<148 * 8, 128, 0, ctx->stream>>>(tr->gE.as<float>(), tr->es.ids.as<int32_t>(), tr->es.count.as<int32_t>(), c.d_e,
                                                         ctx->d_grads + ctx->lay.E_e);
    KPRN_LAUNCH_CHECK(ctx);
    tr->sparse_grads = false;
    ctx->sparse_gE = nullptr;
    return KPRN_OK;
}

}  // namespace kprn

// ---------------------------------------------------------------------------------------------
// the step

namespace kprn {

// params[1][idx]:zero() of MyOptimizer:zeroPadTokens (MyOptimizer.lua:74-93) without invalidating the packed entity table
static int zero_pad_rows(kprn_ctx* ctx, const int64_t* pads) {
    if (!pads) return KPRN_OK;
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    bool small = false;
    if (pads[0] > 0) {
        if (pads[0] > c.V_t) { set_error("pad row (type) out of range"); return KPRN_EINVAL; }
        KPRN_CUDA(cudaMemsetAsync(ctx->d_params + l.E_t + (pads[0] - 1) * c.d_t, 0, sizeof(float) * c.d_t, ctx->stream)); small = true;
    }
    if (pads[2] > 0) {
        if (pads[2] > c.V_r) { set_error("pad row (relation) out of range"); return KPRN_EINVAL; }
        KPRN_CUDA(cudaMemsetAsync(ctx->d_params + l.E_r + (pads[2] - 1) * c.d_r, 0, sizeof(float) * c.d_r, ctx->stream)); small = true;
    }
    if (pads[1] > 0) {
        if (pads[1] > c.V_e) { set_error("pad row (entity) out of range"); return KPRN_EINVAL; }
        KPRN_CUDA(cudaMemsetAsync(ctx->d_params + l.E_e + (pads[1] - 1) * c.d_e, 0, sizeof(float) * c.d_e, ctx->stream));
        __nv_bfloat16 *hi = nullptr, *lo = nullptr; int KE = 0;
        packed_entity(ctx, &hi, &lo, &KE);
        if (hi) {
            KPRN_CUDA(cudaMemsetAsync(hi + (pads[1] - 1) * KE, 0, sizeof(__nv_bfloat16) * KE, ctx->stream));
            KPRN_CUDA(cudaMemsetAsync(lo + (pads[1] - 1) * KE, 0, sizeof(__nv_bfloat16) * KE, ctx->stream));
        }
    }
    if (small) ctx->small_version++;
    return KPRN_OK;
}

static int rows_launch(kprn_ctx* ctx, TrainState* tr, const int32_t* rows, const int32_t* count, int64_t t_now, int apply, const float* g,
                       const float* recv_g, const int32_t* contrib, int64_t maxK, const float* gscale, int64_t grid_hint,
                       cudaStream_t stream = nullptr) {
    if (!stream) stream = ctx->stream;
    const kprn_config& c = ctx->cfg;
    RowsArgs a;
    a.rows = rows; a.count = count;
    a.x = ctx->d_params + ctx->lay.E_e; a.m = ctx->d_m + ctx->lay.E_e; a.v = ctx->d_v + ctx->lay.E_e;
    a.last_t = tr->last_t; a.base_t = tr->base_t; a.dense_ever = tr->dense_ever ? 1 : 0;
    a.t_now = t_now; a.apply = apply; a.g = g; a.g_stride = c.d_e; a.recv_g = recv_g; a.contrib = contrib; a.world = ctx->world; a.maxK = maxK;
    a.gscale = gscale; a.steps = tr->steps.as<float>(); a.b1 = tr->b1; a.b2 = tr->b2; a.eps = tr->eps; a.d_e = c.d_e;
    a.E_hi = nullptr; a.E_lo = nullptr; a.KE = 0;
    packed_entity(ctx, &a.E_hi, &a.E_lo, &a.KE);
    int64_t grid = cdiv(grid_hint, 8) < 148 * 8 ? (grid_hint < 1 ? 1 : cdiv(grid_hint, 8)) : 148 * 8;      // 8 warps per block, a warp per row
    if (c.d_e == 256) adam_rows_kernel<2><<<(unsigned)grid, 256, 0, stream>>>(a);
    else if (c.d_e == 128) adam_rows_kernel<1><<<(unsigned)grid, 256, 0, stream>>>(a);
    else adam_rows_kernel<0><<<(unsigned)grid, 256, 0, stream>>>(a);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

// all-gather of (row ids, row gradients), union slots, per-slot contribution table.  Returns maxK.
static int exchange_rows(kprn_ctx* ctx, TrainState* tr, int64_t N, int64_t* maxK_out) {
    const kprn_config& c = ctx->cfg;
    const int world = ctx->world;
    // K of this rank and the maximum over ranks (one host round trip: the all-gather needs equal counts)
    KPRN_TRY(ensure(ctx->tmp_comm, 64 * sizeof(double)));
    int32_t K32 = 0;
    KPRN_CUDA(cudaMemcpyAsync(&K32, tr->es.count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t kk = K32;
    int64_t* dk = ctx->tmp_comm.as<int64_t>();
    KPRN_CUDA(cudaMemcpyAsync(dk, &kk, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    KPRN_TRY(comm_allreduce_i64(ctx, dk, 1, true));
    int64_t maxK = 0;
    KPRN_CUDA(cudaMemcpyAsync(&maxK, dk, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    maxK = (maxK + 63) / 64 * 64;
    tr->last_K = K32; tr->last_maxK = maxK;
    if (sizeof(float) * (size_t)maxK * c.d_e > tr->gE.cap) {
        // a rank with a smaller batch: grow the send buffer (rows beyond K are never read by anyone, but must be addressable)
        DBuf nb;
        KPRN_TRY(ensure(nb, sizeof(float) * (size_t)maxK * c.d_e));
        KPRN_CUDA(cudaMemcpyAsync(nb.p, tr->gE.p, sizeof(float) * (size_t)K32 * c.d_e, cudaMemcpyDeviceToDevice, ctx->stream));
        KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(tr->gE.p);
        tr->gE = nb;
        ctx->sparse_gE = tr->gE.as<float>();
    }
    KPRN_TRY(ensure(tr->send_ids, sizeof(int32_t) * (size_t)(maxK + 1)));
    KPRN_TRY(ensure(tr->recv_ids, sizeof(int32_t) * (size_t)(maxK + 1) * world));
    KPRN_TRY(ensure(tr->recv_g, sizeof(float) * (size_t)maxK * c.d_e * world));
    pack_ids_kernel<<<(unsigned)cdiv(maxK + 1, 256), 256, 0, ctx->stream>>>(tr->es.ids.as<int32_t>(), tr->es.count.as<int32_t>(),
                                                                          tr->send_ids.as<int32_t>(), maxK + 1);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(comm_allgather_bytes(ctx, tr->send_ids.p, tr->recv_ids.p, (int64_t)sizeof(int32_t) * (maxK + 1)));
    KPRN_TRY(comm_allgather_bytes(ctx, tr->gE.p, tr->recv_g.p, (int64_t)sizeof(float) * maxK * c.d_e));
    // union of the rows: slot per distinct row over all ranks, then who contributes what to each slot
    const int64_t n_keys = (int64_t)world * maxK;
    KPRN_TRY(ensure(tr->contrib, sizeof(int32_t) * (size_t)n_keys * (world + 1)));     // keys [n_keys] then contrib [n_keys][world]
    int32_t* keys = tr->contrib.as<int32_t>();
    int32_t* contrib = keys + n_keys;
    union_keys_kernel<<<(unsigned)(cdiv(n_keys, 256) < 2368 ? cdiv(n_keys, 256) : 2368), 256, 0, ctx->stream>>>(tr->recv_ids.as<int32_t>(), world, maxK, keys);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(dedup_build(ctx, tr->us, 2, keys, 0, 1, n_keys));
    fill_i32_kernel<<<2368, 256, 0, ctx->stream>>>(contrib, n_keys * world, -1);
    KPRN_LAUNCH_CHECK(ctx);
    union_contrib_kernel<<<(unsigned)(cdiv(n_keys, 256) < 2368 ? cdiv(n_keys, 256) : 2368), 256, 0, ctx->stream>>>(
        keys, tr->us.pos.as<uint32_t>(), tr->us.slot_of.as<uint32_t>(), world, maxK, contrib);
    KPRN_LAUNCH_CHECK(ctx);
    *maxK_out = maxK;
    (void)N;
    return KPRN_OK;
}

static int train_step_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, const float* d_labels,
                             const int32_t* d_pos, const int32_t* d_neg, int64_t n_pairs, const int64_t* pads, const kprn_train_opts& o) {
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    TrainState* tr = tr_get(ctx);
    const bool sparse = !o.dense && !(o.regularize && o.l2 != 0.f);     // g += l2*theta makes every row's gradient non-zero
    const float gsc = o.grad_scale == 0.f ? 1.f : o.grad_scale;
    if (o.loss == 0 && !d_labels) { set_error("train_step: BCE needs labels"); return KPRN_EINVAL; }
    if (o.loss == 1 && (!d_pos || !d_neg || n_pairs <= 0)) { set_error("train_step: BPR needs pos/neg segment pairs"); return KPRN_EINVAL; }
    if (o.loss != 0 && o.loss != 1) { set_error("train_step: loss must be 0 (BCE) or 1 (BPR)"); return KPRN_EINVAL; }
    KPRN_TRY(adam_state(ctx));
    struct Guard { kprn_ctx* c; ~Guard() { c->in_train_step = false; } } guard{ctx};
    ctx->in_train_step = true;
    if (pads) { tr->pads[0] = pads[0]; tr->pads[1] = pads[1]; tr->pads[2] = pads[2]; tr->have_pads = true; }
    double* d_loss = nullptr;

    if (!sparse) {
        // ---- the reference's dense flow (MyOptimizer.lua:177-221 literally): parity mode and the l2 != 0 case ----
        KPRN_TRY(train_flush(ctx));
        KPRN_TRY(zero_pad_rows(ctx, pads));
        ctx->params_version++;                       // rows changed outside the packed copies' bookkeeping
        KPRN_CUDA(cudaMemsetAsync(ctx->d_grads, 0, sizeof(float) * (size_t)l.total, ctx->stream));
        tr->sparse_grads = false; ctx->sparse_gE = nullptr;
        KPRN_TRY(forward_device_impl(ctx, d_ids, d_seg, N, S, true, false, nullptr, nullptr));
        if (o.loss == 0) KPRN_TRY(loss_bce_device(ctx, d_labels, S, &d_loss));
        else KPRN_TRY(loss_bpr_device(ctx, d_pos, d_neg, n_pairs, o.bpr_on_logits, &d_loss));
        KPRN_TRY(backward_device(ctx, gsc));
        if (ctx->world > 1) KPRN_TRY(comm_allreduce_f32(ctx, ctx->d_grads, l.total, false));
        if (o.regularize) {
            if (o.use_clip) {
                float nrm = 0.f;
                KPRN_TRY(kprn_grad_norm(ctx, &nrm));
                if (nrm > o.clip_norm) KPRN_TRY(kprn_grad_scale(ctx, o.clip_norm / nrm));
            }
            if (o.l2 != 0.f) KPRN_TRY(kprn_grad_add_params(ctx, o.l2));
        }
        KPRN_TRY(kprn_adam_step(ctx, o.lr, o.beta1, o.beta2, o.eps));
        KPRN_TRY(zero_pad_rows(ctx, pads));
        ctx->params_version++;
        ctx->last_loss = d_loss;
        return KPRN_OK;
    }

    // ---- sparse flow ----
    static const bool env_defer = !(getenv("KPRN_TRAIN_DEFER") && getenv("KPRN_TRAIN_DEFER")[0] == '0');      // diagnostics (A/B timing)
    if (tr->pending && (tr->b1 != o.beta1 || tr->b2 != o.beta2 || tr->eps != o.eps)) KPRN_TRY(train_flush(ctx));   // the replay uses latched constants
    tr->b1 = o.beta1; tr->b2 = o.beta2; tr->eps = o.eps;
    if (!tr->last_t) {
        KPRN_CUDA(cudaMalloc(&tr->last_t, sizeof(int32_t) * (size_t)c.V_e));
        KPRN_CUDA(cudaMemsetAsync(tr->last_t, 0, sizeof(int32_t) * (size_t)c.V_e, ctx->stream));
    }
    // 1. slot per distinct entity row of the batch
    KPRN_TRY(dedup_build(ctx, tr->es, 1, d_ids, N, c.T, 0));
    KPRN_TRY(ensure(tr->gE, sizeof(float) * (size_t)N * c.T * c.d_e));
    const int64_t t_now = ctx->adam_t + 1;
    // 2. the rows this batch reads must be current before the forward
    //    (measured: refreshing only the packed copies here and replaying x, m, v in the update kernel instead saves 3 KB of
    //    traffic per row but runs the replay arithmetic twice -- 0.27 ms slower per step at the benchmark size)
    if (tr->pending) KPRN_TRY(rows_launch(ctx, tr, tr->es.ids.as<int32_t>(), tr->es.count.as<int32_t>(), t_now, 0, nullptr, nullptr, nullptr, 0, nullptr, N * c.T));
    // 3. pad rows, 4. zero the gradients this step writes
    KPRN_TRY(zero_pad_rows(ctx, pads));
    KPRN_CUDA(cudaMemsetAsync(ctx->d_grads, 0, sizeof(float) * (size_t)l.E_e, ctx->stream));
    KPRN_CUDA(cudaMemsetAsync(ctx->d_grads + l.E_r, 0, sizeof(float) * (size_t)(l.total - l.E_r), ctx->stream));
    if (!tc_train_supported(ctx)) {     // the generic backward scatter-adds into the compact rows; the tensor-core one stores every slot once
        zero_rows_kernel<<<148 * 8, 256, 0, ctx->stream>>>(tr->gE.as<float>(), tr->es.count.as<int32_t>(), c.d_e);
        KPRN_LAUNCH_CHECK(ctx);
    }
    // 5. forward (training caches) -> criterion -> backward; the entity gradient lands in gE[slot]
    ctx->sparse_gE = tr->gE.as<float>(); ctx->sparse_pos = tr->es.pos.as<uint32_t>(); ctx->sparse_slot_of = tr->es.slot_of.as<uint32_t>();
    tr->sparse_grads = true; tr->sparse_N = N;
    // the step counter and its step size go first: the row optimizer may run on the backward's side stream, which forks from this
    // one inside the backward (rolled back if the step fails before the update)
    ctx->adam_t = t_now;
    struct Roll { kprn_ctx* c; bool done; ~Roll() { if (!done) { c->adam_t--; c->defer_eg_join = false; c->eg_stream = nullptr; } } } roll{ctx, false};
    const float step = adam_step_size(o.lr, o.beta1, o.beta2, t_now);
    KPRN_TRY(steps_push(ctx, tr, step));
    const bool clip = o.regularize && o.use_clip;
    ctx->defer_eg_join = !clip && env_defer;
    ctx->eg_stream = nullptr;
    KPRN_TRY(forward_device_impl(ctx, d_ids, d_seg, N, S, true, false, nullptr, nullptr));
    if (o.loss == 0) KPRN_TRY(loss_bce_device(ctx, d_labels, S, &d_loss));
    else KPRN_TRY(loss_bpr_device(ctx, d_pos, d_neg, n_pairs, o.bpr_on_logits, &d_loss));
    KPRN_TRY(backward_device(ctx, gsc));
    ctx->defer_eg_join = false;
    // 6. across GPUs: small tensors all-reduced, rows all-gathered
    const float* recv_g = nullptr; const int32_t* contrib = nullptr; int64_t maxK = 0;
    const int32_t* rows = tr->es.ids.as<int32_t>(); const int32_t* count = tr->es.count.as<int32_t>();
    int64_t rows_bound = N * c.T;
    if (ctx->world > 1) {
        if (ctx->eg_stream) {
            // the row exchange (all-gathers, union of the slots) runs on the backward's side stream, beside the weight-gradient
            // kernel of the main stream; it is issued BEFORE the all-reduce of the small tensors (which has to wait for that kernel):
            // NCCL serialises the operations of one communicator in issue order.  The helpers launch on ctx->stream: swapped here.
            struct Swap { kprn_ctx* c; cudaStream_t saved; ~Swap() { c->stream = saved; } } swap{ctx, ctx->stream};
            ctx->stream = ctx->eg_stream;
            KPRN_TRY(exchange_rows(ctx, tr, N, &maxK));
        }
        KPRN_TRY(comm_allreduce_f32_2(ctx, ctx->d_grads, l.E_e, ctx->d_grads + l.E_r, l.total - l.E_r));
        if (!ctx->eg_stream) KPRN_TRY(exchange_rows(ctx, tr, N, &maxK));
        recv_g = tr->recv_g.as<float>(); contrib = tr->contrib.as<int32_t>() + (int64_t)ctx->world * maxK;
        rows = tr->us.ids.as<int32_t>(); count = tr->us.count.as<int32_t>();
        rows_bound = (int64_t)ctx->world * maxK;
    }
    // 7. global-norm clip (MyOptimizer.lua:196-201) as a device scalar folded into the optimizer kernels
    const float* gscale = nullptr;
    if (clip) {
        KPRN_TRY(ensure(tr->norm2, 64));
        double* n2 = tr->norm2.as<double>();
        float* gs = reinterpret_cast<float*>(n2 + 1);
        KPRN_CUDA(cudaMemsetAsync(n2, 0, sizeof(double), ctx->stream));
        range_sumsq_kernel<<<148 * 2, 256, 0, ctx->stream>>>(ctx->d_grads, l.E_e, n2);
        KPRN_LAUNCH_CHECK(ctx);
        range_sumsq_kernel<<<148 * 4, 256, 0, ctx->stream>>>(ctx->d_grads + l.E_r, l.total - l.E_r, n2);
        KPRN_LAUNCH_CHECK(ctx);
        rows_sumsq_kernel<<<148 * 8, 256, 0, ctx->stream>>>(tr->gE.as<float>(), count, c.d_e, recv_g, contrib, ctx->world, maxK, n2);
        KPRN_LAUNCH_CHECK(ctx);
        clip_scale_kernel<<<1, 1, 0, ctx->stream>>>(n2, o.clip_norm, gs, gs + 1);
        KPRN_LAUNCH_CHECK(ctx);
        gscale = gs;
    }
    // 8. optim.adam: dense on the small tensors, lazily-dense on the rows that received a gradient
    adam_dense_kernel<<<148 * 2, 256, 0, ctx->stream>>>(ctx->d_params, ctx->d_grads, ctx->d_m, ctx->d_v, l.E_e, step, o.beta1, o.beta2, o.eps, gscale);
    KPRN_LAUNCH_CHECK(ctx);
    adam_dense_kernel<<<148 * 8, 256, 0, ctx->stream>>>(ctx->d_params + l.E_r, ctx->d_grads + l.E_r, ctx->d_m + l.E_r, ctx->d_v + l.E_r,
                                                       l.total - l.E_r, step, o.beta1, o.beta2, o.eps, gscale);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(rows_launch(ctx, tr, rows, count, t_now, 1, tr->gE.as<float>(), recv_g, contrib, maxK, gscale, rows_bound, ctx->eg_stream));
    if (tr->hint_ids) {
        // the hinted next batch: its rows are replayed up to this step now (same kernel, same arithmetic as the catch-up of the next
        // step, which then skips them), on whichever stream the row update ran on
        struct Swap { kprn_ctx* c; cudaStream_t saved; ~Swap() { c->stream = saved; } } swap{ctx, ctx->stream};
        if (ctx->eg_stream) ctx->stream = ctx->eg_stream;
        const int32_t* hid = tr->hint_ids; const int64_t hn = tr->hint_N;
        tr->hint_ids = nullptr;
        KPRN_TRY(dedup_build(ctx, tr->hs, 1, hid, hn, c.T, 0));
        KPRN_TRY(rows_launch(ctx, tr, tr->hs.ids.as<int32_t>(), tr->hs.count.as<int32_t>(), t_now + 1, 0, nullptr, nullptr, nullptr, 0, nullptr, hn * c.T,
                             ctx->stream));
    }
    if (ctx->eg_stream) {          // the rows were updated beside the weight-gradient kernel: join
        KPRN_CUDA(cudaEventRecord(ctx->eg_event, ctx->eg_stream));
        KPRN_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->eg_event, 0));
        ctx->eg_stream = nullptr;
    }
    roll.done = true;
    tr->pending = true;                 // rows outside this batch are now one more step behind
    ctx->small_version++;
    // 9. pad rows again
    KPRN_TRY(zero_pad_rows(ctx, pads));
    ctx->last_loss = d_loss;
    return KPRN_OK;
}

}  // namespace kprn

using namespace kprn;

extern "C" {

int kprn_train_step_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, const float* d_labels,
                           const int32_t* d_pos_seg, const int32_t* d_neg_seg, int64_t n_pairs, const int64_t* pad_rows3,
                           const kprn_train_opts* opts) {
    if (!ctx || !opts || !d_ids || !d_seg) { set_error("train_step_device: null argument"); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    return train_step_device(ctx, d_ids, d_seg, N, S, d_labels, d_pos_seg, d_neg_seg, n_pairs, pad_rows3, *opts);
}

int kprn_train_prefetch_device(kprn_ctx* ctx, const int32_t* d_ids_next, int64_t N_next) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    TrainState* tr = tr_get(ctx);
    if (!d_ids_next || N_next <= 0) { tr->hint_ids = nullptr; tr->hint_N = 0; return KPRN_OK; }
    if (N_next * (int64_t)ctx->cfg.T >= (1ll << 30)) { set_error("train_prefetch: batch too large"); return KPRN_EINVAL; }
    tr->hint_ids = d_ids_next; tr->hint_N = N_next;
    return KPRN_OK;
}

int kprn_last_loss(kprn_ctx* ctx, float* loss) {
    if (!ctx || !loss) { set_error("null argument"); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->last_loss) { set_error("last_loss: no training step yet"); return KPRN_ESTATE; }
    double h = 0.0;
    KPRN_CUDA(cudaMemcpyAsync(&h, ctx->last_loss, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
    *loss = (float)h;
    return kprn_sync(ctx);         // surfaces a bad-id flag raised by the step's forward
}

int kprn_train_step(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg, int64_t N, int64_t S, const float* labels,
                    const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs, const int64_t* pad_rows3,
                    const kprn_train_opts* opts, float* loss) {
    if (!ctx || !opts || !ids || !seg) { set_error("train_step: null argument"); return KPRN_EINVAL; }
    KPRN_CUDA(cudaSetDevice(ctx->device));
    if (N <= 0 || S <= 0) { set_error("empty batch: N=%lld S=%lld", (long long)N, (long long)S); return KPRN_EINVAL; }
    const kprn_config& c = ctx->cfg;
    const size_t ids_bytes = sizeof(int32_t) * (size_t)N * c.T * 3, seg_bytes = sizeof(int32_t) * (size_t)(S + 1);
    KPRN_TRY(ensure(ctx->ids_buf, ids_bytes));
    KPRN_TRY(ensure(ctx->seg_buf, seg_bytes));
    KPRN_TRY(host_to_device(ctx, ctx->ids_buf.p, ids, ids_bytes, 0));
    KPRN_TRY(host_to_device(ctx, ctx->seg_buf.p, seg, seg_bytes, ids_bytes));
    const float* d_labels = nullptr; const int32_t *d_pos = nullptr, *d_neg = nullptr;
    if (opts->loss == 0) {
        if (!labels) { set_error("train_step: BCE needs labels"); return KPRN_EINVAL; }
        KPRN_TRY(ensure(ctx->labels, sizeof(float) * (size_t)S));
        KPRN_TRY(host_to_device(ctx, ctx->labels.p, labels, sizeof(float) * (size_t)S, ids_bytes + seg_bytes));
        d_labels = ctx->labels.as<float>();
    } else {
        if (!pos_seg || !neg_seg || n_pairs <= 0) { set_error("train_step: BPR needs pos/neg segment pairs"); return KPRN_EINVAL; }
        KPRN_TRY(ensure(ctx->pair_buf, sizeof(int32_t) * 2 * (size_t)n_pairs));
        int32_t* dp = ctx->pair_buf.as<int32_t>();
        KPRN_TRY(host_to_device(ctx, dp, pos_seg, sizeof(int32_t) * (size_t)n_pairs, ids_bytes + seg_bytes));
        KPRN_TRY(host_to_device(ctx, dp + n_pairs, neg_seg, sizeof(int32_t) * (size_t)n_pairs, ids_bytes + seg_bytes + sizeof(int32_t) * (size_t)n_pairs));
        d_pos = dp; d_neg = dp + n_pairs;
    }
    KPRN_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream));
    KPRN_TRY(train_step_device(ctx, ctx->ids_buf.as<int32_t>(), ctx->seg_buf.as<int32_t>(), N, S, d_labels, d_pos, d_neg, n_pairs, pad_rows3, *opts));
    float lv = 0.f;
    KPRN_TRY(kprn_last_loss(ctx, &lv));
    if (loss) *loss = lv;
    return KPRN_OK;
}

int kprn_train_info(kprn_ctx* ctx, int64_t* rows_local, int64_t* rows_union, int64_t* rows_gathered) {
    if (!ctx) { set_error("null ctx"); return KPRN_EINVAL; }
    TrainState* tr = ctx->tr;
    if (rows_local) {
        *rows_local = 0;
        if (tr && tr->es.count.p) {
            int32_t k = 0;
            KPRN_CUDA(cudaMemcpyAsync(&k, tr->es.count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
            *rows_local = k;
        }
    }
    if (rows_union) {
        *rows_union = 0;
        if (tr && ctx->world > 1 && tr->us.count.p) {
            int32_t k = 0;
            KPRN_CUDA(cudaMemcpyAsync(&k, tr->us.count.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
            *rows_union = k;
        }
    }
    if (rows_gathered) *rows_gathered = tr ? tr->last_maxK * ctx->world : 0;
    return KPRN_OK;
}

}  // extern "C"
