// tc_prefix.cu -- leading-token reuse for the scoring recurrence (A4 of SURVEY.md 8a; reference: the batcher left-pads every
// path to T steps with the same pad token and the LSTM runs all T steps without masking, SP/model/batcher/Batcher.lua:43-54,
// SP/model/OneModel.lua:236-275).
//
// The state after k leading copies of one token X is the same for every path that starts with k copies of X -- in the shipped
// data 70 % of the paths start with two pad tokens.  So the batch is sorted by the number p of leading copies of X (counting
// sort, longest paths first), the state chain (h, c after 1..T-1 copies of X) is computed once on one row, and a tile of the
// recurrence whose paths all have p >= t0 starts at step t0 from that state: same results as stepping through the pads row by
// row, (T-1-t0) instead of (T-1) recurrent products.  X is the leading token of the first path that begins with a repeated
// token (any choice is correct; a token that nobody repeats just gives p = 0 everywhere and the sort is a no-op).
#include <stdio.h>
#include <stdlib.h>

#include "tc_state.cuh"

namespace kprn {

constexpr int PX_BINS = 32;           // T <= 32

// token[0..2] = the (type, entity, relation) ids of X; zeroes the histogram (counts [0,32), cursors [32,64))
__global__ void prefix_find_kernel(const int32_t* __restrict__ ids, int64_t N, int T, int32_t* __restrict__ token, int32_t* __restrict__ hist) {
    __shared__ int best;
    if (threadIdx.x == 0) best = 0x7fffffff;
    if (threadIdx.x < 2 * PX_BINS) hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t lim = N < 4096 ? N : 4096;
    if (T >= 2)
        for (int64_t n = threadIdx.x; n < lim; n += blockDim.x) {
            const int32_t* p = ids + n * T * 3;
            if (p[0] == p[3] && p[1] == p[4] && p[2] == p[5]) { atomicMin(&best, (int)n); break; }
        }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int64_t n = best == 0x7fffffff ? 0 : best;
        token[threadIdx.x] = ids[n * T * 3 + threadIdx.x];
    }
}

// p[n] = number of leading steps of path n equal to X (at most T-1: the last step always runs), histogram of p
__global__ void prefix_count_kernel(const int32_t* __restrict__ ids, int64_t N, int T, const int32_t* __restrict__ token, uint8_t* __restrict__ p,
                                    int32_t* __restrict__ hist) {
    __shared__ int h[PX_BINS];
    if (threadIdx.x < PX_BINS) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) {
        const int32_t a = token[0], b = token[1], c = token[2];
        const int32_t* q = ids + n * T * 3;
        int k = 0;
        while (k < T - 1 && q[3 * k] == a && q[3 * k + 1] == b && q[3 * k + 2] == c) ++k;
        p[n] = (uint8_t)k;
        atomicAdd(&h[k], 1);
    }
    __syncthreads();
    if (threadIdx.x < T && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// counting sort by p, ascending (longest real paths first): perm[sorted position] = path, ps = p in sorted order, ids copied in
// sorted order.  A block reserves its range of every bin with one atomic; the order inside a bin is whatever the block-local
// atomics give (results do not depend on it: rows are independent).
__global__ void prefix_scatter_kernel(const int32_t* __restrict__ ids, int64_t N, int T, const uint8_t* __restrict__ p, int32_t* __restrict__ hist,
                                      int32_t* __restrict__ perm, uint8_t* __restrict__ ps, int32_t* __restrict__ ids_out) {
    __shared__ int cnt[PX_BINS], base[PX_BINS];
    if (threadIdx.x < PX_BINS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0, local = 0;
    if (n < N) { k = p[n]; local = atomicAdd(&cnt[k], 1); }
    __syncthreads();
    if (threadIdx.x < T) {
        int off = 0;
        for (int b = 0; b < (int)threadIdx.x; ++b) off += hist[b];
        base[threadIdx.x] = cnt[threadIdx.x] ? off + atomicAdd(&hist[PX_BINS + threadIdx.x], cnt[threadIdx.x]) : 0;
    }
    __syncthreads();
    if (n < N) {
        const int64_t d = base[k] + local;
        perm[d] = (int32_t)n; ps[d] = (uint8_t)k;
        const int32_t* src = ids + n * T * 3;
        int32_t* dst = ids_out + d * T * 3;
        for (int i = 0; i < 3 * T; ++i) dst[i] = src[i];
    }
}

// pre[gp] = gate_scale * (P_tr[type, relation][gp] + Wx[row, entity columns] . E_entity[X])   -- the x-projection of X, fp32
__global__ void chain_pre_kernel(const int32_t* __restrict__ token, const float* __restrict__ E_e, const float* __restrict__ Wx, const float* __restrict__ Ptr,
                                 int64_t V_e, int V_t, int V_r, int d_t, int d_e, int D, int H, int G4, float* __restrict__ pre) {
    const int gp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (gp >= G4) return;
    int u;
    const int row = perm_to_row(gp, H, u);
    float acc = 0.f;
    if (u < H) {
        int64_t en = token[1] - 1; en = en < 0 ? 0 : (en >= V_e ? V_e - 1 : en);
        int ta = token[0] - 1, rb = token[2] - 1;
        ta = ta < 0 ? 0 : (ta >= V_t ? V_t - 1 : ta); rb = rb < 0 ? 0 : (rb >= V_r ? V_r - 1 : rb);
        const float* x = E_e + en * d_e;
        const float* w = Wx + (int64_t)row * D + d_t;
        for (int k = lane; k < d_e; k += 32) acc = fmaf(w[k], x[k], acc);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc = (acc + Ptr[(int64_t)(ta * V_r + rb) * G4 + gp]) * gate_scale(gp);
    }
    if (lane == 0) pre[gp] = acc;
}

// step k (1-based) of the chain: the cell on pre + chain_u[k-1] with c = chain_c[k-1] gives (h, c) after k copies of X;
// chain_c[k] = c, chain_u[k][gp] = gate_scale * Wh[row] . h  (what the recurrence's accumulator would hold for such a row)
__global__ void chain_step_kernel(int k, const float* __restrict__ pre, float* __restrict__ chain_u, float* __restrict__ chain_c,
                                  const float* __restrict__ Wh, int H, int Hp, int G4) {
    extern __shared__ float hs[];                      // h after k copies, [Hp]
    for (int u = threadIdx.x; u < Hp; u += blockDim.x) {
        const int j = u >> 5, uu = u & 31;
        float x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gp = j * 128 + q * 32 + uu;
            x[q] = pre[gp] + (k > 1 ? chain_u[(size_t)(k - 1) * G4 + gp] : 0.f);
        }
        const float cp = k > 1 ? chain_c[(size_t)(k - 1) * Hp + u] : 0.f;
        const float i_ = 1.f / (1.f + exp2f(x[0])), g_ = 1.f - 2.f / (1.f + exp2f(x[1]));
        const float f_ = 1.f / (1.f + exp2f(x[2])), o_ = 1.f / (1.f + exp2f(x[3]));
        const float c = fmaf(i_, g_, f_ * cp);
        const float h = u < H ? o_ * (1.f - 2.f / (1.f + exp2f(c * 2.8853900817779268f))) : 0.f;
        hs[u] = h;
        if (blockIdx.x == 0) chain_c[(size_t)k * Hp + u] = u < H ? c : 0.f;
    }
    __syncthreads();
    const int gp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (gp >= G4) return;
    int u;
    const int row = perm_to_row(gp, H, u);
    float acc = 0.f;
    if (u < H) {
        const float* w = Wh + (int64_t)row * H;
        for (int kk = lane; kk < H; kk += 32) acc = fmaf(w[kk], hs[kk], acc);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc *= gate_scale(gp);
    }
    if (lane == 0) chain_u[(size_t)k * G4 + gp] = acc;
}

bool prefix_enabled(const kprn_ctx* ctx, int64_t N) {
    bool forced = false;
    if (ctx->opt_prefix >= 0) { if (!ctx->opt_prefix) return false; forced = true; }
    else { const char* d = getenv("KPRN_TC_PREFIX"); if (d && d[0] == '0') return false; }
    const kprn_config& c = ctx->cfg;
    // small batches: the ten extra launches cost more than the skipped steps save (kprn_set_option(ctx, "prefix", 1) forces it)
    return c.T >= 2 && c.T <= PX_BINS && (forced || N >= 4096) && N < (1ll << 31) / (3 * c.T);
}

int prefix_build(kprn_ctx* ctx, const int32_t* d_ids, int64_t N, PrefixOut* out) {
    TcState* s = ctx->tc;
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    if (!s->px) s->px = new PrefixState();
    PrefixState* px = s->px;
    if (!px->side) {
        KPRN_CUDA(cudaStreamCreateWithFlags(&px->side, cudaStreamNonBlocking));
        KPRN_CUDA(cudaEventCreateWithFlags(&px->ev_tok, cudaEventDisableTiming));
        KPRN_CUDA(cudaEventCreateWithFlags(&px->ev_chain, cudaEventDisableTiming));
    }
    const int T = c.T;
    KPRN_TRY(ensure(px->token, 64));
    KPRN_TRY(ensure(px->hist, sizeof(int32_t) * 2 * PX_BINS));
    KPRN_TRY(ensure(px->p, (size_t)N));
    KPRN_TRY(ensure(px->ps, (size_t)N + TILE_M));
    KPRN_TRY(ensure(px->perm, sizeof(int32_t) * (size_t)N));
    KPRN_TRY(ensure(px->ids, sizeof(int32_t) * (size_t)N * T * 3));
    KPRN_TRY(ensure(px->pre, sizeof(float) * (size_t)s->G4));
    KPRN_TRY(ensure(px->chain_u, sizeof(float) * (size_t)T * s->G4));
    KPRN_TRY(ensure(px->chain_c, sizeof(float) * (size_t)T * s->Hp));
    int32_t* token = px->token.as<int32_t>();
    int32_t* hist = px->hist.as<int32_t>();
    prefix_find_kernel<<<1, 1024, 0, ctx->stream>>>(d_ids, N, T, token, hist);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_CUDA(cudaEventRecord(px->ev_tok, ctx->stream));
    // side stream: the state chain of X (needs only X and the parameters), beside the sort / dedup / projection
    KPRN_CUDA(cudaStreamWaitEvent(px->side, px->ev_tok, 0));
    const unsigned gb = (unsigned)cdiv(s->G4, 8);
    chain_pre_kernel<<<gb, 256, 0, px->side>>>(token, ctx->d_params + l.E_e, ctx->d_params + l.Wx, s->Ptr, c.V_e, (int)c.V_t, (int)c.V_r, c.d_t,
                                              c.d_e, l.D, c.H, s->G4, px->pre.as<float>());
    KPRN_LAUNCH_CHECK(ctx);
    for (int k = 1; k < T; ++k) {
        chain_step_kernel<<<gb, 256, sizeof(float) * s->Hp, px->side>>>(k, px->pre.as<float>(), px->chain_u.as<float>(), px->chain_c.as<float>(),
                                                                       ctx->d_params + l.Wh, c.H, s->Hp, s->G4);
        KPRN_LAUNCH_CHECK(ctx);
    }
    KPRN_CUDA(cudaEventRecord(px->ev_chain, px->side));
    // main stream: count, sort
    prefix_count_kernel<<<(unsigned)cdiv(N, 256), 256, 0, ctx->stream>>>(d_ids, N, T, token, px->p.as<uint8_t>(), hist);
    KPRN_LAUNCH_CHECK(ctx);
    prefix_scatter_kernel<<<(unsigned)cdiv(N, 256), 256, 0, ctx->stream>>>(d_ids, N, T, px->p.as<uint8_t>(), hist, px->perm.as<int32_t>(),
                                                                         px->ps.as<uint8_t>(), px->ids.as<int32_t>());
    KPRN_LAUNCH_CHECK(ctx);
    out->ids = px->ids.as<int32_t>(); out->perm = px->perm.as<int32_t>(); out->ps = px->ps.as<uint8_t>();
    out->chain_u = px->chain_u.as<float>(); out->chain_c = px->chain_c.as<float>();
    return KPRN_OK;
}

// the recurrence may start once the chain is there
int prefix_join(kprn_ctx* ctx) {
    PrefixState* px = ctx->tc->px;
    KPRN_CUDA(cudaStreamWaitEvent(ctx->stream, px->ev_chain, 0));
    return KPRN_OK;
}

void prefix_destroy(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    if (!s || !s->px) return;
    PrefixState* px = s->px;
    if (px->side) { cudaStreamSynchronize(px->side); cudaEventDestroy(px->ev_tok); cudaEventDestroy(px->ev_chain); cudaStreamDestroy(px->side); }
    for (DBuf* b : {&px->token, &px->hist, &px->p, &px->ps, &px->perm, &px->ids, &px->pre, &px->chain_u, &px->chain_c})
        if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    delete px;
    s->px = nullptr;
}

}  // namespace kprn
