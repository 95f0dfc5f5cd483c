// kernels_simple.cu -- fp32 CUDA-core implementation of the whole path (forward with training caches,
// pooling, losses' chain rule, BPTT, embedding scatter-add).  This is the KPRN_PREC_FP32 engine and
// the backward for every precision mode.  Math follows SURVEY.md App. A; reference call sites are
// cited per kernel (paths relative to /root/reference/release/songPathRnn/).
#include <math.h>

#include "common.cuh"
#include "sgemm.cuh"

namespace kprn {

int sgemm(kprn_ctx* ctx, int64_t M, int64_t N, int64_t K, const float* A, int64_t sAm, int64_t sAk,
          const float* B, int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, int splitk) {
    if (M <= 0 || N <= 0 || K <= 0) return KPRN_OK;
    {   // tensor-core precision modes: large products run on tcgen05 (bf16x3 split), the rest on the FFMA kernel below
        bool handled = false;
        KPRN_TRY(tc_gemm(ctx, M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, accumulate, splitk, &handled));
        if (handled) return KPRN_OK;
    }
    if (splitk < 1) splitk = 1;
    int64_t kchunk = cdiv(cdiv(K, splitk), SG_BK) * SG_BK;
    splitk = (int)cdiv(K, kchunk);
    dim3 grid((unsigned)cdiv(N, SG_BN), (unsigned)cdiv(M, SG_BM), (unsigned)splitk);
    int use_atomic = splitk > 1;
    if (use_atomic && !accumulate) {
        // split-K needs a zeroed destination
        if (ldc == N) KPRN_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * M * N, ctx->stream));
        else KPRN_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, ctx->stream));
    }
    sgemm_kernel<<<grid, SG_THREADS, 0, ctx->stream>>>(M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc,
                                                       accumulate ? 1 : 0, kchunk, use_atomic);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// ids validation (a bad id raises in the reference: nn.LookupTable index out of range)
__global__ void validate_ids_kernel(const int32_t* __restrict__ ids, int64_t n_steps, int64_t V_t, int64_t V_e,
                                    int64_t V_r, int32_t* err) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_steps) return;
    int32_t a = ids[i * 3], b = ids[i * 3 + 1], c = ids[i * 3 + 2];
    if (a < 1 || a > V_t || b < 1 || b > V_e || c < 1 || c > V_r) atomicExch(err, 1);
}

// A3: three nn.LookupTable gathers + JoinTable(3), order type . entity . relation
// (model/net/FeatureEmbedding.lua:112-121).  Rows laid out step-major inside a chunk: r = t*CH + p.
__global__ void gather_x_kernel(const int32_t* __restrict__ ids, const float* __restrict__ P, Layout l,
                                int d_t, int d_e, int d_r, int64_t V_t, int64_t V_e, int64_t V_r, int T,
                                int64_t n0, int64_t CH, float* __restrict__ X) {
    int64_t r = blockIdx.x;
    int t = (int)(r / CH);
    int64_t p = r % CH;
    const int32_t* id = ids + ((n0 + p) * T + t) * 3;
    int64_t it = min(max((int64_t)id[0] - 1, (int64_t)0), V_t - 1);
    int64_t ie = min(max((int64_t)id[1] - 1, (int64_t)0), V_e - 1);
    int64_t ir = min(max((int64_t)id[2] - 1, (int64_t)0), V_r - 1);
    const float* st = P + l.E_t + it * d_t;
    const float* se = P + l.E_e + ie * d_e;
    const float* sr = P + l.E_r + ir * d_r;
    float* x = X + r * l.D;
    for (int k = threadIdx.x; k < l.D; k += blockDim.x) {
        float v;
        if (k < d_t) v = st[k];
        else if (k < d_t + d_e) v = se[k - d_t];
        else v = sr[k - d_t - d_e];
        x[k] = v;
    }
}

// MaskZero(rm,1) test (OneModel.lua:251): a step is masked when its whole input row is zero.
__global__ void row_zero_kernel(const float* __restrict__ X, int D, int64_t rows, uint8_t* __restrict__ flag) {
    int64_t r = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (r >= rows) return;
    int lane = threadIdx.x & 31;
    int nz = 0;
    for (int k = lane; k < D; k += 32) nz |= (X[r * D + k] != 0.f);
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) flag[r] = nz ? 0 : 1;
}

// A4: FastLSTM cell (Element-Research rnn): gate blocks [i sigma, g tanh, f sigma, o sigma];
// c = i*g + f*c_prev; h = o*tanh(c)  (OneModel.lua:236).  G already holds x Wx^T + h_prev Wh^T.
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ G, const float* __restrict__ bx, int H, int64_t nb,
                                     const float* __restrict__ c_prev, int64_t ldcp, float* __restrict__ c_out,
                                     int64_t ldc, float* __restrict__ h_out, int64_t ldh,
                                     float* __restrict__ act_out, int64_t ldact) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * H) return;
    int64_t p = idx / H;
    int u = (int)(idx % H);
    const float* g = G + p * 4 * H;
    float i = sigmoidf_(g[u] + bx[u]);
    float gg = tanhf(g[H + u] + bx[H + u]);
    float f = sigmoidf_(g[2 * H + u] + bx[2 * H + u]);
    float o = sigmoidf_(g[3 * H + u] + bx[3 * H + u]);
    float cp = c_prev ? c_prev[p * ldcp + u] : 0.f;
    float c = fmaf(i, gg, f * cp);
    c_out[p * ldc + u] = c;
    h_out[p * ldh + u] = o * tanhf(c);
    if (act_out) {
        float* a = act_out + p * ldact;
        a[u] = i; a[H + u] = gg; a[2 * H + u] = f; a[3 * H + u] = o;
    }
}

// f4: ReLU Elman step wrapped in MaskZero (OneModel.lua:240-254): h = relu(i2h(x) + h2h(h_prev)), 0 if x row is zero
__global__ void rnn_cell_fwd_kernel(const float* __restrict__ G, const float* __restrict__ bx,
                                    const float* __restrict__ bh, int H, int64_t nb, const uint8_t* __restrict__ masked,
                                    float* __restrict__ h_out, int64_t ldh, float* __restrict__ act_out, int64_t ldact) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * H) return;
    int64_t p = idx / H;
    int u = (int)(idx % H);
    float v = fmaxf(G[p * H + u] + bx[u] + bh[u], 0.f);
    bool mk = masked[p] != 0;
    if (mk) v = 0.f;
    h_out[p * ldh + u] = v;
    if (act_out) act_out[p * ldact + u] = mk ? -1.f : v;
}

__global__ void bias_rows_kernel(float* __restrict__ out, int64_t ld, const float* __restrict__ b, int n, int64_t rows) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * n) return;
    out[(idx / n) * ld + (idx % n)] = b[idx % n];
}
__global__ void relu_kernel(float* __restrict__ x, int64_t n) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) x[idx] = fmaxf(x[idx], 0.f);
}

// A5: SelectTable(-1) + nn.Linear(H,L) (OneModel.lua:274-275); optional 2-layer head.
int head_forward(kprn_ctx* ctx, const float* d_hT, int64_t ldh, int64_t n0, int64_t nb, float* z1_opt) {
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    float* pl = ctx->path_logits.as<float>() + n0 * c.L;
    const float* hin = d_hT;
    int64_t ldin = ldh;
    if (c.head_layers == 2) {
        bias_rows_kernel<<<(unsigned)cdiv(nb * c.H, 256), 256, 0, ctx->stream>>>(z1_opt, c.H, ctx->d_params + l.b1, c.H, nb);
        KPRN_LAUNCH_CHECK(ctx);
        KPRN_TRY(sgemm(ctx, nb, c.H, c.H, d_hT, ldh, 1, ctx->d_params + l.W1, 1, c.H, z1_opt, c.H, true));
        relu_kernel<<<(unsigned)cdiv(nb * c.H, 256), 256, 0, ctx->stream>>>(z1_opt, nb * c.H);
        KPRN_LAUNCH_CHECK(ctx);
        hin = z1_opt; ldin = c.H;
    }
    bias_rows_kernel<<<(unsigned)cdiv(nb * c.L, 256), 256, 0, ctx->stream>>>(pl, c.L, ctx->d_params + l.bo, c.L, nb);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(sgemm(ctx, nb, c.L, c.H, hin, ldin, 1, ctx->d_params + l.Wo, 1, c.H, pl, c.L, true));
    return KPRN_OK;
}

// scores-only fast path: only column classId of nn.Linear(H,L) feeds nn.Select(2,classId) (MyOptimizer.lua:126,
// test_from_checkpoint.lua:82), so a scoring pass needs one dot product per path.  One warp per path.
__global__ void head_cls_kernel(const float* __restrict__ hT, int64_t ldh, const float* __restrict__ wrow, const float* __restrict__ b, int H,
                                int64_t nb, float* __restrict__ out, int L, int cls) {
    int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= nb) return;
    int lane = threadIdx.x & 31;
    const float* h = hT + p * ldh;
    float acc = 0.f;
    for (int k = lane; k < H; k += 32) acc = fmaf(h[k], wrow[k], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[p * L + cls] = acc + *b;
}

int head_forward_cls(kprn_ctx* ctx, const float* d_hT, int64_t ldh, int64_t n0, int64_t nb) {
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    const int cls = c.class_id - 1;
    float* pl = ctx->path_logits.as<float>() + n0 * c.L;
    head_cls_kernel<<<(unsigned)cdiv(nb, 8), 256, 0, ctx->stream>>>(d_hT, ldh, ctx->d_params + l.Wo + (int64_t)cls * c.H, ctx->d_params + l.bo + cls, c.H,
                                                                    nb, pl, c.L, cls);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

// A6+A7: MapReduce views (MapReduce.lua:32-47) + reducer + nn.Sigmoid + nn.Select(2,classId).
//   LSE: module/LogSumExp.lua:13-30; max: nn.Max(2); top-k mean: module/TopK.lua + nn.Mean(2).
// One thread per (segment, column); adjacent threads read adjacent columns (coalesced).
__global__ void pool_fwd_kernel(const float* __restrict__ s, const int32_t* __restrict__ seg, int64_t S, int L,
                                int pool, int topk, float ig, int cls, int col0, int ncols, float* __restrict__ seg_logits,
                                float* __restrict__ scores, float* __restrict__ out_scores, float* __restrict__ out_logits) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * ncols) return;
    int64_t k = idx / ncols;
    int j = col0 + (int)(idx % ncols);
    int64_t a = seg[k], b = seg[k + 1];
    float out = 0.f;
    if (b > a) {
        if (pool == KPRN_POOL_LSE) {
            float m = -INFINITY;
            for (int64_t p = a; p < b; ++p) m = fmaxf(m, s[p * L + j] * ig);
            float z = 0.f;
            for (int64_t p = a; p < b; ++p) z += expf(s[p * L + j] * ig - m);
            out = (m + logf(z)) / ig;
        } else if (pool == KPRN_POOL_MAX) {
            float m = -INFINITY;
            for (int64_t p = a; p < b; ++p) m = fmaxf(m, s[p * L + j]);
            out = m;
        } else {
            // mean of the k largest: repeated selection (k is small; reference uses k<=5)
            int64_t n = b - a;
            int kk = (int)(topk < n ? topk : n);
            float acc = 0.f, prev = INFINITY;
            int64_t prev_idx = -1;
            for (int r = 0; r < kk; ++r) {
                float best = -INFINITY; int64_t bi = -1;
                for (int64_t p = a; p < b; ++p) {
                    float v = s[p * L + j];
                    bool after = (v < prev) || (v == prev && p > prev_idx);   // strict order (value desc, index asc)
                    if (after && (v > best || bi < 0)) { best = v; bi = p; }
                }
                acc += best; prev = best; prev_idx = bi;
            }
            out = acc / (float)kk;
        }
    }
    seg_logits[k * L + j] = out;
    if (out_logits) out_logits[k * L + j] = out;
    if (j == cls) {
        float pr = sigmoidf_(out);
        scores[k] = pr;
        if (out_scores) out_scores[k] = pr;
    }
}

// Scores-only LSE pool: a warp-shuffle segmented reduction keyed by the (user,item) segment.  G lanes per segment (G = 32: one
// warp per segment; G = 4: eight segments per warp, for batches whose segments hold a few paths each): the lanes stride over the
// segment's paths keeping an online (max, sum-exp) pair, merged across the group with a shuffle butterfly.
// Same math as LogSumExp.lua:13-30 for column classId, then nn.Sigmoid (OneModel.lua:294).
// Path logit p = s[p*stride + col], or -- after the fused tensor-core forward -- bias + the four partial dot products
// part4[g*npad + p] that the recurrence's epilogue groups left (summed in a fixed order: deterministic).
template <int G>
__global__ void pool_lse_cls_warp_kernel(const float* __restrict__ s, int64_t stride, int col, const float* __restrict__ part4,
                                         int64_t npad, const float* __restrict__ bias, const int32_t* __restrict__ seg, int64_t S, int L,
                                         float ig, int cls, float* __restrict__ seg_logits, float* __restrict__ scores,
                                         float* __restrict__ out_scores) {
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    const bool live = k < S;                      // whole groups are live or not; dead groups still take part in the shuffles
    const int64_t a = live ? seg[k] : 0, b = live ? seg[k + 1] : 0;
    const float b0 = part4 ? *bias : 0.f;
    float m = -INFINITY, z = 0.f;
    for (int64_t p = a + lane; p < b; p += G) {
        const float raw = part4 ? ((part4[p] + part4[npad + p]) + (part4[2 * npad + p] + part4[3 * npad + p])) + b0 : s[p * stride + col];
        const float v = raw * ig;
        const float nm = fmaxf(m, v);
        z = z * expf(m - nm) + expf(v - nm);      // exp(-inf) = 0 on the first element
        m = nm;
    }
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o), oz = __shfl_xor_sync(0xffffffffu, z, o);
        const float nm = fmaxf(m, om);
        const float za = (m == -INFINITY) ? 0.f : z * expf(m - nm), zb = (om == -INFINITY) ? 0.f : oz * expf(om - nm);
        z = za + zb;
        m = nm;
    }
    if (live && lane == 0) {
        const float out = (b > a) ? (m + logf(z)) / ig : 0.f;
        seg_logits[k * L + cls] = out;
        const float pr = sigmoidf_(out);
        scores[k] = pr;
        if (out_scores) out_scores[k] = pr;
    }
}

int pool_forward(kprn_ctx* ctx, const int32_t* d_seg, int64_t S, float* d_out_scores, float* d_out_logits) {
    const kprn_config& c = ctx->cfg;
    KPRN_TRY(ensure(ctx->seg_logits, sizeof(float) * S * c.L));
    KPRN_TRY(ensure(ctx->scores, sizeof(float) * S));
    float ig = c.inv_gamma == 0.f ? 1.f : c.inv_gamma;
    if (ctx->cls_only && c.pool == KPRN_POOL_LSE) {
        const float* part4 = ctx->head_part4;
        const float* bias = ctx->d_params + ctx->lay.bo + (c.class_id - 1);
        const float* pl = ctx->path_logits.as<float>();
        if (ctx->N <= 6 * S) {       // a few paths per (user,item) pair: 4 lanes per segment
            pool_lse_cls_warp_kernel<4><<<(unsigned)cdiv(S * 4, 256), 256, 0, ctx->stream>>>(
                pl, c.L, c.class_id - 1, part4, ctx->head_npad, bias, d_seg, S, c.L, ig, c.class_id - 1, ctx->seg_logits.as<float>(),
                ctx->scores.as<float>(), d_out_scores);
        } else {
            pool_lse_cls_warp_kernel<32><<<(unsigned)cdiv(S * 32, 256), 256, 0, ctx->stream>>>(
                pl, c.L, c.class_id - 1, part4, ctx->head_npad, bias, d_seg, S, c.L, ig, c.class_id - 1, ctx->seg_logits.as<float>(),
                ctx->scores.as<float>(), d_out_scores);
        }
        KPRN_LAUNCH_CHECK(ctx);
        return KPRN_OK;
    }
    const int col0 = ctx->cls_only ? c.class_id - 1 : 0, ncols = ctx->cls_only ? 1 : c.L;
    pool_fwd_kernel<<<(unsigned)cdiv(S * ncols, 256), 256, 0, ctx->stream>>>(
        ctx->path_logits.as<float>(), d_seg, S, c.L, c.pool, c.topk, ig, c.class_id - 1, col0, ncols,
        ctx->seg_logits.as<float>(), ctx->scores.as<float>(), d_out_scores, d_out_logits);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

static int64_t simple_chunk(const kprn_ctx* ctx, int64_t N) {
    // bound the [CH*T, max(D,G)] workspaces to ~512 MB each
    int64_t per_path = (int64_t)ctx->cfg.T * (ctx->lay.G > ctx->lay.D ? ctx->lay.G : ctx->lay.D) * 4;
    int64_t ch = (512ll << 20) / per_path;
    if (ch < 128) ch = 128;
    ch = (ch / 128) * 128;
    return N < ch ? N : ch;
}

int simple_forward(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, bool train) {
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    const int H = c.H, T = c.T, G = l.G, D = l.D;
    const int64_t CH = simple_chunk(ctx, N);
    KPRN_TRY(ensure(ctx->path_logits, sizeof(float) * N * c.L));
    KPRN_TRY(ensure(ctx->ws_x, sizeof(float) * CH * T * D));
    KPRN_TRY(ensure(ctx->ws_xw, sizeof(float) * CH * T * G));
    if (train) {
        KPRN_TRY(ensure(ctx->act, sizeof(float) * N * T * G));
        if (c.cell == KPRN_CELL_LSTM) KPRN_TRY(ensure(ctx->cst, sizeof(float) * N * T * H));
        KPRN_TRY(ensure(ctx->hst, sizeof(float) * N * T * H));
        if (c.head_layers == 2) KPRN_TRY(ensure(ctx->z1, sizeof(float) * N * H));
    } else {
        KPRN_TRY(ensure(ctx->ws_h0, sizeof(float) * CH * H));
        KPRN_TRY(ensure(ctx->ws_c, sizeof(float) * CH * H));
        if (c.head_layers == 2) KPRN_TRY(ensure(ctx->z1, sizeof(float) * CH * H));
    }
    if (c.cell == KPRN_CELL_RNN) KPRN_TRY(ensure(ctx->tmp_small, (size_t)CH * T));
    const float* P = ctx->d_params;
    float* X = ctx->ws_x.as<float>();
    float* XW = ctx->ws_xw.as<float>();
    for (int64_t n0 = 0; n0 < N; n0 += CH) {
        const int64_t nb = (N - n0) < CH ? (N - n0) : CH;
        gather_x_kernel<<<(unsigned)(nb * T), 128, 0, ctx->stream>>>(d_ids, P, l, c.d_t, c.d_e, c.d_r, c.V_t, c.V_e,
                                                                     c.V_r, T, n0, nb, X);
        KPRN_LAUNCH_CHECK(ctx);
        if (c.cell == KPRN_CELL_RNN) {
            row_zero_kernel<<<(unsigned)cdiv(nb * T, 8), 256, 0, ctx->stream>>>(X, D, nb * T, ctx->tmp_small.as<uint8_t>());
            KPRN_LAUNCH_CHECK(ctx);
        }
        // x W_x^T for all T steps of the chunk in one product (no recurrence in this term)
        KPRN_TRY(sgemm(ctx, nb * T, G, D, X, D, 1, P + l.Wx, 1, D, XW, G, false));
        for (int t = 0; t < T; ++t) {
            float* Gt = XW + (int64_t)t * nb * G;
            const float* hprev; int64_t ldhp;
            float* hout; int64_t ldho;
            if (train) {
                hprev = t ? ctx->hst.as<float>() + (n0 * T + (t - 1)) * H : nullptr; ldhp = (int64_t)T * H;
                hout = ctx->hst.as<float>() + (n0 * T + t) * H; ldho = (int64_t)T * H;
            } else {
                hprev = t ? ctx->ws_h0.as<float>() : nullptr; ldhp = H;
                hout = ctx->ws_h0.as<float>(); ldho = H;
            }
            if (hprev) KPRN_TRY(sgemm(ctx, nb, G, H, hprev, ldhp, 1, P + l.Wh, 1, H, Gt, G, true));
            unsigned blocks = (unsigned)cdiv(nb * H, 256);
            if (c.cell == KPRN_CELL_LSTM) {
                const float* cprev; int64_t ldcp; float* cout; int64_t ldco;
                if (train) {
                    cprev = t ? ctx->cst.as<float>() + (n0 * T + (t - 1)) * H : nullptr; ldcp = (int64_t)T * H;
                    cout = ctx->cst.as<float>() + (n0 * T + t) * H; ldco = (int64_t)T * H;
                } else {
                    cprev = t ? ctx->ws_c.as<float>() : nullptr; ldcp = H; cout = ctx->ws_c.as<float>(); ldco = H;
                }
                lstm_cell_fwd_kernel<<<blocks, 256, 0, ctx->stream>>>(
                    Gt, P + l.bx, H, nb, cprev, ldcp, cout, ldco, hout, ldho,
                    train ? ctx->act.as<float>() + (n0 * T + t) * G : nullptr, (int64_t)T * G);
            } else {
                rnn_cell_fwd_kernel<<<blocks, 256, 0, ctx->stream>>>(
                    Gt, P + l.bx, P + l.bh, H, nb, ctx->tmp_small.as<uint8_t>() + (int64_t)t * nb, hout, ldho,
                    train ? ctx->act.as<float>() + (n0 * T + t) * G : nullptr, (int64_t)T * G);
            }
            KPRN_LAUNCH_CHECK(ctx);
        }
        const float* hT = train ? ctx->hst.as<float>() + (n0 * T + (T - 1)) * H : ctx->ws_h0.as<float>();
        float* z1 = c.head_layers == 2 ? (train ? ctx->z1.as<float>() + n0 * H : ctx->z1.as<float>()) : nullptr;
        if (ctx->cls_only) KPRN_TRY(head_forward_cls(ctx, hT, train ? (int64_t)T * H : H, n0, nb));
        else KPRN_TRY(head_forward(ctx, hT, train ? (int64_t)T * H : H, n0, nb, z1));
    }
    (void)S; (void)d_seg;
    return KPRN_OK;
}

// ---------------------------------------------------------------------------------------------
// backward

// LogSumExp.lua:32-37 (softmax * gradOutput) for the selected class column only -- every other column of
// the reference's gradOutput is zero after nn.Select(2,classId):backward.  MapReduce.lua:74-85.
__global__ void pool_bwd_kernel(const float* __restrict__ s, const float* __restrict__ seg_logits,
                                const int32_t* __restrict__ seg, int64_t S, int L, int pool, int topk, float ig,
                                int cls, const float* __restrict__ dlse, float scale, float* __restrict__ ds_c) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S) return;
    int64_t a = seg[k], b = seg[k + 1];
    if (b <= a) return;
    float go = dlse[k] * scale;
    if (pool == KPRN_POOL_LSE) {
        float lse = seg_logits[k * L + cls];
        for (int64_t p = a; p < b; ++p) ds_c[p] = expf((s[p * L + cls] - lse) * ig) * go;
    } else if (pool == KPRN_POOL_MAX) {
        int64_t best = a;
        for (int64_t p = a; p < b; ++p) { ds_c[p] = 0.f; if (s[p * L + cls] > s[best * L + cls]) best = p; }
        ds_c[best] = go;
    } else {
        int64_t n = b - a;
        int kk = (int)(topk < n ? topk : n);
        for (int64_t p = a; p < b; ++p) ds_c[p] = 0.f;
        float prev = INFINITY; int64_t prev_idx = -1;
        for (int r = 0; r < kk; ++r) {
            float best = -INFINITY; int64_t bi = -1;
            for (int64_t p = a; p < b; ++p) {
                float v = s[p * L + cls];
                bool after = (v < prev) || (v == prev && p > prev_idx);
                if (after && (v > best || bi < 0)) { best = v; bi = p; }
            }
            ds_c[bi] = go / (float)kk; prev = best; prev_idx = bi;
        }
    }
}

// nn.Linear backward for the one live output column: dh = ds_c * Wo[c,:]
__global__ void head_bwd_dh_kernel(const float* __restrict__ ds_c, const float* __restrict__ wrow, int H, int64_t nb,
                                   const float* __restrict__ relu_gate, float* __restrict__ dh) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * H) return;
    int64_t p = idx / H; int u = (int)(idx % H);
    float v = ds_c[p] * wrow[u];
    if (relu_gate && !(relu_gate[idx] > 0.f)) v = 0.f;
    dh[idx] = v;
}

// out[j] += sum_r w[r] * X[r*ld + j]   (w == nullptr -> plain column sum).  Used for gWo[c,:], gbo[c], gbx, gb1.
__global__ void wcolsum_kernel(const float* __restrict__ X, int64_t ld, const float* __restrict__ w, int64_t rows,
                               int cols, int64_t rows_per_block, float* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc += (w ? w[r] : 1.f) * X[r * ld + j];
    out[(int64_t)blockIdx.y * cols + j] = acc;          // partial of this row range; summed in range order by the kernel below
}
// out[j] += sum over the row ranges, in range order: no float atomics, the same bits on every run
// block (32 columns, 8 groups of row ranges): a thread sums every 8th range, the 8 group sums are combined in group order
__global__ void wcolsum_fin_kernel(const float* __restrict__ part, int nparts, int cols, float* __restrict__ out) {
    __shared__ float red[8][32];
    const int j = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (j < cols)
        for (int b = threadIdx.y; b < nparts; b += 8) acc += part[(int64_t)b * cols + j];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && j < cols) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += red[g][threadIdx.x];
        out[j] += t;
    }
}

static int wcolsum(kprn_ctx* ctx, const float* X, int64_t ld, const float* w, int64_t rows, int cols, float* out) {
    if (rows <= 0) return KPRN_OK;
    int64_t rpb = cdiv(rows, 1184);        // 8 row ranges per SM and column block: the loop is a dependent chain of strided loads
    if (rpb < 32) rpb = 32;
    dim3 grid((unsigned)cdiv(cols, 128), (unsigned)cdiv(rows, rpb));
    KPRN_TRY(ensure(ctx->ws_part, sizeof(float) * (size_t)grid.y * cols));
    wcolsum_kernel<<<grid, 128, 0, ctx->stream>>>(X, ld, w, rows, cols, rpb, ctx->ws_part.as<float>());
    KPRN_LAUNCH_CHECK(ctx);
    wcolsum_fin_kernel<<<(unsigned)cdiv(cols, 32), dim3(32, 8), 0, ctx->stream>>>(ctx->ws_part.as<float>(), (int)grid.y, cols, out);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

// out += sum(x): ONE block, fixed assignment of elements to threads and a fixed reduction tree (no float atomics)
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    __shared__ float wsum[32];
    float acc = 0.f;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;      // 128-bit loads over the aligned body
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        acc += (v.x + v.y) + (v.z + v.w);
    }
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) *out += v;
    }
}

// FastLSTM BPTT step.  dG is written step-major [t][p][4H].
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ act, int64_t ldact, const float* __restrict__ c_t,
                                     const float* __restrict__ c_prev, int64_t ldc, const float* __restrict__ dh,
                                     float* __restrict__ dc, int H, int64_t nb, float* __restrict__ dG) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * H) return;
    int64_t p = idx / H; int u = (int)(idx % H);
    const float* a = act + p * ldact;
    float i = a[u], g = a[H + u], f = a[2 * H + u], o = a[3 * H + u];
    float tc = tanhf(c_t[p * ldc + u]);
    float dhu = dh[idx];
    float d_o = dhu * tc;
    float dcu = dc[idx] + dhu * o * (1.f - tc * tc);
    float cp = c_prev ? c_prev[p * ldc + u] : 0.f;
    dc[idx] = dcu * f;
    float* d = dG + p * 4 * H;
    d[u] = dcu * g * i * (1.f - i);
    d[H + u] = dcu * i * (1.f - g * g);
    d[2 * H + u] = dcu * cp * f * (1.f - f);
    d[3 * H + u] = d_o * o * (1.f - o);
}

__global__ void rnn_cell_bwd_kernel(const float* __restrict__ act, int64_t ldact, const float* __restrict__ dh,
                                    int H, int64_t nb, float* __restrict__ dG) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * H) return;
    int64_t p = idx / H; int u = (int)(idx % H);
    dG[idx] = act[p * ldact + u] > 0.f ? dh[idx] : 0.f;   // ReLU'; masked steps carry act=-1 -> 0
}

// LookupTable:accGradParameters -- scatter-add of dX rows.  The two tiny tables (type, relation) are reduced in
// shared memory per block first; entity rows go straight to global float atomics (RED.ADD).
__global__ void scatter_embed_kernel(const float* __restrict__ dX, const int32_t* __restrict__ ids, Layout l,
                                     int d_t, int d_e, int d_r, int64_t V_t, int64_t V_e, int64_t V_r, int T, int64_t n0, int64_t CH,
                                     int64_t rows, int rows_per_block, float* __restrict__ grads, int use_smem,
                                     float* __restrict__ gE_compact, const uint32_t* __restrict__ sp_pos, const uint32_t* __restrict__ sp_slot_of) {
    extern __shared__ float sm[];   // [V_t*d_t + V_r*d_r] when use_smem
    const int n_small = (int)(V_t * d_t + V_r * d_r);
    if (use_smem) {
        for (int i = threadIdx.x; i < n_small; i += blockDim.x) sm[i] = 0.f;
        __syncthreads();
    }
    int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    for (int64_t r = r0; r < r1; ++r) {
        int t = (int)(r / CH);
        int64_t p = r % CH;
        const int32_t* id = ids + ((n0 + p) * T + t) * 3;
        // out-of-range ids are flagged by the forward (kprn_sync / kprn_backward report them); never scatter outside the tables
        int64_t it = id[0] - 1, ie = id[1] - 1, ir = id[2] - 1;
        if (it < 0 || it >= V_t || ie < 0 || ie >= V_e || ir < 0 || ir >= V_r) continue;
        const float* dx = dX + r * l.D;
        // sparse training step (train.cu): the entity row's gradient goes to its slot of the compact buffer; the slot map is
        // indexed in the recurrence's read order o = ((n/128)*T + t)*128 + n%128
        float* ge = grads + l.E_e + ie * d_e;
        if (gE_compact) {
            const int64_t n = n0 + p;
            ge = gE_compact + (int64_t)sp_slot_of[sp_pos[((n >> 7) * T + t) * 128 + (n & 127)]] * d_e;
        }
        for (int k = threadIdx.x; k < l.D; k += blockDim.x) {
            float v = dx[k];
            if (k < d_t) {
                if (use_smem) atomicAdd(&sm[it * d_t + k], v); else atomicAdd(grads + l.E_t + it * d_t + k, v);
            } else if (k < d_t + d_e) {
                atomicAdd(ge + (k - d_t), v);
            } else {
                int kk = k - d_t - d_e;
                if (use_smem) atomicAdd(&sm[V_t * d_t + ir * d_r + kk], v); else atomicAdd(grads + l.E_r + ir * d_r + kk, v);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_small; i += blockDim.x) {
            float v = sm[i];
            if (v != 0.f) {
                if (i < V_t * d_t) atomicAdd(grads + l.E_t + i, v);
                else atomicAdd(grads + l.E_r + (i - V_t * d_t), v);
            }
        }
    }
}

// pool backward + the head's gradients for the one live output column (head_layers == 1): ds_c[p] into ctx->ws_ds,
// gbo[c] += sum ds_c, gWo[c,:] += sum_p ds_c[p] h_T[p,:].  Shared by the fused tensor-core backward (tc_train.cu).
int head_backward_cls(kprn_ctx* ctx, float scale, const float* hT, int64_t ldh) {
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    const int cls = c.class_id - 1;
    const int64_t N = ctx->N, S = ctx->S;
    float* gr = ctx->d_grads;
    KPRN_TRY(ensure(ctx->ws_ds, sizeof(float) * N));
    float ig = c.inv_gamma == 0.f ? 1.f : c.inv_gamma;
    float* ds_c = ctx->ws_ds.as<float>();
    KPRN_CUDA(cudaMemsetAsync(ds_c, 0, sizeof(float) * N, ctx->stream));
    pool_bwd_kernel<<<(unsigned)cdiv(S, 128), 128, 0, ctx->stream>>>(
        ctx->path_logits.as<float>(), ctx->seg_logits.as<float>(), ctx->d_seg, S, c.L, c.pool, c.topk, ig, cls,
        ctx->dscores.as<float>(), scale, ds_c);
    KPRN_LAUNCH_CHECK(ctx);
    sum_kernel<<<1, 1024, 0, ctx->stream>>>(ds_c, N, gr + l.bo + cls);
    KPRN_LAUNCH_CHECK(ctx);
    KPRN_TRY(wcolsum(ctx, hT, ldh, ds_c, N, c.H, gr + l.Wo + (int64_t)cls * c.H));
    return KPRN_OK;
}

int simple_backward(kprn_ctx* ctx, float scale) {
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    const int H = c.H, T = c.T, G = l.G, D = l.D, L = c.L, cls = c.class_id - 1;
    const int64_t N = ctx->N, S = ctx->S;
    const float* P = ctx->d_params;
    float* gr = ctx->d_grads;
    const int64_t CH = simple_chunk(ctx, N);
    KPRN_TRY(ensure(ctx->ws_ds, sizeof(float) * N));
    KPRN_TRY(ensure(ctx->ws_x, sizeof(float) * CH * T * D));
    KPRN_TRY(ensure(ctx->ws_dg, sizeof(float) * CH * T * G));
    KPRN_TRY(ensure(ctx->ws_dx, sizeof(float) * CH * T * D));
    KPRN_TRY(ensure(ctx->ws_dh, sizeof(float) * CH * H));
    KPRN_TRY(ensure(ctx->ws_dc, sizeof(float) * CH * H));
    if (c.head_layers == 2) KPRN_TRY(ensure(ctx->ws_dz, sizeof(float) * CH * H));
    float ig = c.inv_gamma == 0.f ? 1.f : c.inv_gamma;
    float* ds_c = ctx->ws_ds.as<float>();
    KPRN_CUDA(cudaMemsetAsync(ds_c, 0, sizeof(float) * N, ctx->stream));
    pool_bwd_kernel<<<(unsigned)cdiv(S, 128), 128, 0, ctx->stream>>>(
        ctx->path_logits.as<float>(), ctx->seg_logits.as<float>(), ctx->d_seg, S, L, c.pool, c.topk, ig, cls,
        ctx->dscores.as<float>(), scale, ds_c);
    KPRN_LAUNCH_CHECK(ctx);
    // gbo[c] += sum ds_c
    sum_kernel<<<1, 1024, 0, ctx->stream>>>(ds_c, N, gr + l.bo + cls);
    KPRN_LAUNCH_CHECK(ctx);
    float* X = ctx->ws_x.as<float>();
    float* dG = ctx->ws_dg.as<float>();
    float* dX = ctx->ws_dx.as<float>();
    float* dh = ctx->ws_dh.as<float>();
    float* dc = ctx->ws_dc.as<float>();
    const int64_t ldT = (int64_t)T * H;
    for (int64_t n0 = 0; n0 < N; n0 += CH) {
        const int64_t nb = (N - n0) < CH ? (N - n0) : CH;
        const float* hT = ctx->hst.as<float>() + (n0 * T + (T - 1)) * H;
        unsigned blocks = (unsigned)cdiv(nb * H, 256);
        if (c.head_layers == 2) {
            const float* z1 = ctx->z1.as<float>() + n0 * H;
            float* dz = ctx->ws_dz.as<float>();
            // gWo[c,:] += sum_p ds_c[p] z1[p,:]
            KPRN_TRY(wcolsum(ctx, z1, H, ds_c + n0, nb, H, gr + l.Wo + (int64_t)cls * H));
            head_bwd_dh_kernel<<<blocks, 256, 0, ctx->stream>>>(ds_c + n0, P + l.Wo + (int64_t)cls * H, H, nb, z1, dz);
            KPRN_LAUNCH_CHECK(ctx);
            KPRN_TRY(wcolsum(ctx, dz, H, nullptr, nb, H, gr + l.b1));
            // gW1[u,k] += sum_p dz[p,u] hT[p,k]
            KPRN_TRY(sgemm(ctx, H, H, nb, dz, 1, H, hT, ldT, 1, gr + l.W1, H, true, 32));
            // dh = dz W1
            KPRN_TRY(sgemm(ctx, nb, H, H, dz, H, 1, P + l.W1, H, 1, dh, H, false));
        } else {
            KPRN_TRY(wcolsum(ctx, hT, ldT, ds_c + n0, nb, H, gr + l.Wo + (int64_t)cls * H));
            head_bwd_dh_kernel<<<blocks, 256, 0, ctx->stream>>>(ds_c + n0, P + l.Wo + (int64_t)cls * H, H, nb, nullptr, dh);
            KPRN_LAUNCH_CHECK(ctx);
        }
        KPRN_CUDA(cudaMemsetAsync(dc, 0, sizeof(float) * nb * H, ctx->stream));
        for (int t = T - 1; t >= 0; --t) {
            float* dGt = dG + (int64_t)t * nb * G;
            const float* act = ctx->act.as<float>() + (n0 * T + t) * G;
            if (c.cell == KPRN_CELL_LSTM) {
                const float* ct = ctx->cst.as<float>() + (n0 * T + t) * H;
                const float* cp = t ? ctx->cst.as<float>() + (n0 * T + t - 1) * H : nullptr;
                lstm_cell_bwd_kernel<<<blocks, 256, 0, ctx->stream>>>(act, (int64_t)T * G, ct, cp, ldT, dh, dc, H, nb, dGt);
            } else {
                rnn_cell_bwd_kernel<<<blocks, 256, 0, ctx->stream>>>(act, (int64_t)T * G, dh, H, nb, dGt);
            }
            KPRN_LAUNCH_CHECK(ctx);
            if (t > 0) {
                const float* hp = ctx->hst.as<float>() + (n0 * T + t - 1) * H;
                // gWh[g,k] += sum_p dG[p,g] h_{t-1}[p,k]
                KPRN_TRY(sgemm(ctx, G, H, nb, dGt, 1, G, hp, ldT, 1, gr + l.Wh, H, true, 32));
                // dh = dG Wh
                KPRN_TRY(sgemm(ctx, nb, H, G, dGt, G, 1, P + l.Wh, H, 1, dh, H, false));
            }
        }
        // bias grads: gbx += colsum(dG over all steps); Elman also has the h2h bias
        KPRN_TRY(wcolsum(ctx, dG, G, nullptr, nb * T, G, gr + l.bx));
        if (c.cell == KPRN_CELL_RNN) KPRN_TRY(wcolsum(ctx, dG, G, nullptr, nb * T, G, gr + l.bh));
        gather_x_kernel<<<(unsigned)(nb * T), 128, 0, ctx->stream>>>(ctx->d_ids, P, l, c.d_t, c.d_e, c.d_r, c.V_t, c.V_e,
                                                                     c.V_r, T, n0, nb, X);
        KPRN_LAUNCH_CHECK(ctx);
        // gWx[g,k] += sum_r dG[r,g] x[r,k]
        KPRN_TRY(sgemm(ctx, G, D, nb * T, dG, 1, G, X, D, 1, gr + l.Wx, D, true, 64));
        // dX = dG Wx
        KPRN_TRY(sgemm(ctx, nb * T, D, G, dG, G, 1, P + l.Wx, D, 1, dX, D, false));
        {
            int64_t rows = nb * T;
            int rpb = 32;
            size_t smem = sizeof(float) * (size_t)(c.V_t * c.d_t + c.V_r * c.d_r);
            int use_smem = smem <= 40 * 1024;
            scatter_embed_kernel<<<(unsigned)cdiv(rows, rpb), 256, use_smem ? smem : 0, ctx->stream>>>(
                dX, ctx->d_ids, l, c.d_t, c.d_e, c.d_r, c.V_t, c.V_e, c.V_r, T, n0, nb, rows, rpb, gr, use_smem,
                ctx->sparse_gE, ctx->sparse_pos, ctx->sparse_slot_of);
            KPRN_LAUNCH_CHECK(ctx);
        }
    }
    return KPRN_OK;
}

}  // namespace kprn
