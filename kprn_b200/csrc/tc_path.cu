// tc_path.cu -- tensor-core (tcgen05 / TMEM / TMA-bulk) forward for the FastLSTM path (A3 + A4 of SURVEY.md 8a).
//
// Math:  G_t = x_t Wx^T + bx + h_{t-1} Wh^T,  x_t = [E_type | E_entity | E_relation].
// Restructuring (algebraically exact, every flop of the reference is still accounted in the roofline):
//   * the type and relation lookups have tiny vocabularies (6 and 9 rows in the reference), so their
//     projection is hoisted into a table  P_tr[(type,rel)][4H] = E_t[type] Wx_t^T + E_r[rel] Wx_r^T + bx
//     recomputed whenever the parameters change; only the entity slice of x goes through the tensor core;
//   * kernel "proj":  XW[(path,t), 4H] = E_e[entity] Wx_e^T + P_tr[(type,rel)]   for all T steps (no recurrence);
//   * kernel "recur": per 128-path tile, T steps of  G = XW_t + h Wh^T  with the cell update fused in the
//     epilogue; c lives in TMEM, h is re-fed to the tensor core from shared memory, weights are streamed
//     from L2 with 1-D TMA bulk copies into a ring.
// Precision: every fp32 operand v is split as v = hi + lo (bf16 each) and the product is formed as
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (3 tcgen05.mma passes, "bf16x3"); KPRN_PREC_TC_FAST
// uses the hi*hi pass only.
//
// Operand layouts are the canonical K-major no-swizzle UMMA layout [k/8][row][8 elements] (ptx.cuh), which the
// packing kernels below produce in global memory so that a slab is one contiguous TMA bulk copy.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_state.cuh"

namespace kprn {


bool tc_supported(const kprn_ctx* ctx) {
    const kprn_config& c = ctx->cfg;
    if (c.cell != KPRN_CELL_LSTM) return false;
    int KE = round_up(c.d_e, 64);
    if (c.H > 512 || KE > 512) return false;      // 256 < d_e <= 512: projection in two K halves; 256 < H <= 512: per-step recurrence (tc_wide.cu)
    int noh = (int)((c.V_t * c.V_r + 63) / 64);
    const int KE0 = KE > 256 ? 256 : KE;
    if (noh > 2 || KE0 / 64 + noh > MAX_NS || c.V_t + c.V_r > 160) return false;   // (type,relation) combos ride the MMA as a one-hot k-slice
    return true;
}

static bool dedup_enabled(const kprn_ctx* ctx) {
    if (ctx->opt_dedup >= 0) return ctx->opt_dedup != 0;            // kprn_set_option(ctx, "dedup", v)
    const char* d = getenv("KPRN_TC_DEDUP");
    return !(d && d[0] == '0');
}
bool tc_validates_ids(const kprn_ctx* ctx, int64_t N) {
    const bool dd = dedup_enabled(ctx);
    return dd && (N + TILE_M - 1) / TILE_M * TILE_M * (int64_t)ctx->cfg.T < (1ll << 30);
}

// ---------------------------------------------------------------------------------------------
// parameter packing (runs when the parameters change)

__global__ void split_entity_kernel(const float* __restrict__ E, int64_t V, int d_e, int KE, bf16* __restrict__ hi,
                                    bf16* __restrict__ lo) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= V * KE) return;
    int64_t row = idx / KE;
    int k = (int)(idx % KE);
    float v = k < d_e ? E[row * d_e + k] : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    hi[idx] = h; lo[idx] = l;
}

// proj B operand, slabs [chunk nt][slice s][hi|lo][kc 8][n 128][8]:
//   s <  KSe : Wx[:, d_t : d_t+d_e]  (entity columns of the input projection)
//   s >= KSe : P_tr^T, the hoisted (type,relation) table, multiplied by a one-hot k-slice of the A tile
__global__ void pack_wxe_kernel(const float* __restrict__ Wx, const float* __restrict__ Ptr, int H, int D, int d_t, int d_e,
                                int NCH, int KSe, int NS, int ncombo, int G4, int k_off, bf16* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)NCH * NS * B_SLAB_ELEMS;
    if (idx >= total) return;
    int e = (int)(idx & 7);
    int n = (int)((idx >> 3) & 127);
    int kc = (int)((idx >> 10) & 7);
    int64_t slab = idx >> 13;
    int s = (int)(slab % NS), nt = (int)(slab / NS);
    int gp = nt * 128 + n, u;
    int row = perm_to_row(gp, H, u);
    float v;
    if (s < KSe) {
        int k = k_off + s * 64 + kc * 8 + e;
        v = (u < H && k < d_e) ? Wx[(int64_t)row * D + d_t + k] : 0.f;
    } else {
        int combo = (s - KSe) * 64 + kc * 8 + e;
        v = combo < ncombo ? Ptr[(int64_t)combo * G4 + gp] : 0.f;
    }
    v *= gate_scale(gp);
    bf16 h, l;
    split_bf16(v, h, l);
    int64_t base = slab * 2 * B_SLAB_ELEMS + ((kc * 128 + n) << 3) + e;
    out[base] = h;
    out[base + B_SLAB_ELEMS] = l;
}

// recur B operand: Wh -> slabs [chunk][kslice][hi|lo][kc 8][n 128][8]
// pair = 1: each slab is stored as two 64-column halves [half][kc 8][n 64][8] (one contiguous 8 KB piece per CTA of a pair)
__global__ void pack_wh_kernel(const float* __restrict__ Wh, int H, int NCH, int KSh, int pair, bf16* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)NCH * KSh * B_SLAB_ELEMS;
    if (idx >= total) return;
    int e = (int)(idx & 7);
    int n = (int)((idx >> 3) & 127);
    int kc = (int)((idx >> 10) & 7);
    int64_t slab = idx >> 13;
    int s = (int)(slab % KSh), j = (int)(slab / KSh);
    int gp = j * 128 + n, k = s * 64 + kc * 8 + e, u;
    int row = perm_to_row(gp, H, u);
    float v = (u < H && k < H) ? Wh[(int64_t)row * H + k] * gate_scale(gp) : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    int64_t within = pair ? (int64_t)(n >> 6) * (B_SLAB_ELEMS / 2) + ((kc * 64 + (n & 63)) << 3) + e : ((kc * 128 + n) << 3) + e;
    int64_t base = slab * 2 * B_SLAB_ELEMS + within;
    out[base] = h;
    out[base + B_SLAB_ELEMS] = l;
}

// hoisted small-vocabulary projection: P_tr[a*V_r+b][g'] = bx + E_t[a] Wx_t^T + E_r[b] Wx_r^T  (fp32 FFMA).
// One block per gate column: the V_t + V_r dot products of that column's weight row (each warp takes some of them, lanes
// stride over k, shuffle reduction), then the V_t*V_r sums -- 15 dot products per column instead of 54 x 2.
__global__ void ptr_table_kernel(const float* __restrict__ P, Layout l, int d_t, int d_e, int d_r, int H, int V_t, int V_r,
                                 int G4, float* __restrict__ out) {
    __shared__ float dots[160];
    const int gp = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int u;
    const int row = perm_to_row(gp, H, u);
    if (u < H) {
        const float* w = P + l.Wx + (int64_t)row * l.D;
        for (int q = warp; q < V_t + V_r; q += nw) {
            float acc = 0.f;
            if (q < V_t) { const float* e = P + l.E_t + (int64_t)q * d_t; for (int k = lane; k < d_t; k += 32) acc = fmaf(e[k], w[k], acc); }
            else { const float* e = P + l.E_r + (int64_t)(q - V_t) * d_r; for (int k = lane; k < d_r; k += 32) acc = fmaf(e[k], w[d_t + d_e + k], acc); }
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) dots[q] = acc;
        }
    }
    __syncthreads();
    const float b = u < H ? P[l.bx + row] : 0.f;
    for (int combo = threadIdx.x; combo < V_t * V_r; combo += blockDim.x)
        out[(int64_t)combo * G4 + gp] = u < H ? (b + dots[combo / V_r]) + dots[V_t + combo % V_r] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// kernel "proj": XW tile (128 path-steps x 4Hp) = [gather(E_e) | onehot(type,rel)] [Wx_e | P_tr]^T on tcgen05.
// XW layout: [tile][column group (4Hp/8)][row 128][8 floats], columns in permuted (chunk) order; a tile is 128 rows of the
// projected list (distinct step inputs, or (path tile, t) without deduplication) = one contiguous 128 x 4Hp x 4 B block; every
// thread access is one 32-byte sector (256-bit LDG / STG), a warp access 1 KB contiguous.
// Roles: warp 0 TMA producer (weight slabs), warp 1 MMA issuer, warps 2-3 gather (LDGSTS of pre-split entity rows +
// one-hot slice), warps 4-11 epilogue (TMEM -> global).  4 accumulators of 128 columns; the A tile is released to the
// gather warps k-slice by k-slice while the last column chunk of the tile is still being multiplied.

struct ProjArgs {
    const int32_t* ids; int64_t N; int T; int64_t V_e; int V_t; int V_r;
    const bf16* E_hi; const bf16* E_lo; const bf16* W; float* XW;
    int KE, KSe, NOH, NS, NCH, G4; int64_t num_tiles; int passes;   // passes: 3 = bf16x3, 1 = hi*hi only
    unsigned long long* dbg;   // optional wait-cycle counters (KPRN_TC_DEBUG=1): see tc_debug_dump
    const int32_t* n_dev;      // when set: the row count lives on the device (distinct step inputs, see dedup below); T is 1
    int k_base;                // first entity column of this launch (d_e > 256 runs as two launches over column halves)
    int accumulate;            // the epilogue adds to XW (second half) instead of storing
};

// the projection of a deduplicated step list learns its size on the device (no host round trip): N = *n_dev rows, T = 1
__device__ __forceinline__ void proj_resolve_size(ProjArgs& a) {
    if (a.n_dev) { a.N = *a.n_dev; a.num_tiles = (a.N + TILE_M - 1) / TILE_M; }
}

// ---------------------------------------------------------------------------------------------
// kernel "proj3": the projection on CTA PAIRS (tcgen05 cta_group::2, cluster of 2: one MMA multiplies two 128-row tiles by
// the same 256 gate columns, each CTA staging its own 128-column half of every weight slab) with the entity rows gathered by
// the TMA unit (cp.async.bulk.tensor.2d tile::gather4 over a tensor map of the pre-split bf16 table, 4 rows x 128 B per
// instruction, hardware 128-byte swizzle).  (The 1-CTA and LDGSTS-gather generations it replaced are in the git history.)
// The entity k-slices of the A tile therefore use the K-major SWIZZLE_128B UMMA layout (rows of 128 B, 8-row atoms of
// 1 KB); the one-hot slice keeps the no-swizzle layout.  Warp 2 issues the gathers (lane l owns rows 4l..4l+3), warp 3
// forwards "my A slice has landed" to the leader CTA.

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
proj3_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, ProjArgs a) {
    proj_resolve_size(a);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int R = PROJ_RING;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int ah_chunks = a.KE / 8 + 8 * a.NOH;
    uint8_t* sA_hi = smem;                                   // entity slices: [s][128 rows][128 B] swizzled; then one-hot slices
    uint8_t* sA_lo = smem + (size_t)ah_chunks * 2048;
    uint8_t* sRing = sA_lo + (size_t)(a.KE / 8) * 2048;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + R * B_SLAB_ELEMS * 2);
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    // [0,R) slab_full [R,2R) slab_empty [2R,3R) peer_full(leader) a_full[6](leader) a_free[6] acc_full[2] acc_empty[2](leader) a_land[6]
    constexpr int B_PEER = 2 * R, B_AFULL = 3 * R, B_AFREE = 3 * R + MAX_NS, B_ACCF = 3 * R + 2 * MAX_NS, B_ACCE = B_ACCF + 2,
                  B_ALAND = B_ACCE + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ALAND + MAX_NS);

    if (threadIdx.x == 0) {
        for (int i = 0; i < R; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(R + i), 1); ptx::mbar_init(BAR(B_PEER + i), 1); }
        for (int i = 0; i < MAX_NS; ++i) { ptx::mbar_init(BAR(B_AFULL + i), 2); ptx::mbar_init(BAR(B_AFREE + i), 1); ptx::mbar_init(BAR(B_ALAND + i), 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(B_ACCF + i), 1); ptx::mbar_init(BAR(B_ACCE + i), 8); }
        ptx::fence_mbar_init();
        ptx::prefetch_tensormap(&tm_hi);
        ptx::prefetch_tensormap(&tm_lo);
    }
    if (warp == 2) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    long long dbg_acc[16] = {0};
    const long long dbg_t0 = clock64();
    const int64_t npairs = gridDim.x / 2, pair = blockIdx.x / 2;
    const int64_t n_tp = (a.num_tiles + 1) / 2;
    const int64_t my_tiles = pair < n_tp ? (n_tp - pair + npairs - 1) / npairs : 0;
    auto tile_of = [&](int64_t it) { int64_t m = 2 * (pair + it * npairs) + rank; return m < a.num_tiles ? m : a.num_tiles - 1; };
    const int NP = a.NCH / 2;
    const uint32_t idesc = ptx::make_idesc_f16(1, 256, 256);
    const int parts = a.passes == 3 ? 2 : 1;

    if (warp == 0) {
        // ===== TMA producer: this CTA's 128-column half (chunk 2*np + rank) of every weight slab =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int np = 0; np < NP; ++np)
                    for (int s = 0; s < a.NS; ++s)
                        for (int part = 0; part < parts; ++part) {
                            TWAIT(0, BAR(R + stage), phase ^ 1);
                            ptx::mbar_arrive_expect_tx(BAR(stage), B_SLAB_ELEMS * 2);
                            const bf16* src = a.W + ((int64_t)((2 * np + rank) * a.NS + s) * 2 + part) * B_SLAB_ELEMS;
                            ptx::tma_bulk_g2s(ptx::smem_u32(sRing + stage * B_SLAB_ELEMS * 2), src, B_SLAB_ELEMS * 2, BAR(stage));
                            if (++stage == R) { stage = 0; phase ^= 1; }
                        }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer (leader CTA) =====
            const bool issue = ptx::elect_one();
            uint32_t stage = 0, phase = 0;
            int64_t au = 0;
            const uint32_t aHiSw = ptx::desc_lo_sw128(ptx::smem_u32(sA_hi)), aLoSw = ptx::desc_lo_sw128(ptx::smem_u32(sA_lo));
            const uint32_t aHi0 = ptx::desc_lo(ptx::smem_u32(sA_hi), 2048);
            const uint32_t ring0 = ptx::desc_lo(ptx::smem_u32(sRing), 2048);
            constexpr uint32_t DHI = ptx::desc_hi(128), DHI_SW = ptx::desc_hi_sw128();
            constexpr uint32_t SLAB_STEP = (B_SLAB_ELEMS * 2) >> 4, K16_STEP = (2 * 2048) >> 4, SLICE_STEP = (8 * 2048) >> 4;
            for (int64_t it = 0; it < my_tiles; ++it) {
                for (int np = 0; np < NP; ++np, ++au) {
                    const int buf = au & 1;
                    if (a.dbg) { long long t0 = clock64(); ptx::mbar_wait_cluster(BAR(B_ACCE + buf), ((au >> 1) & 1) ^ 1); dbg_acc[1] += clock64() - t0; }
                    else ptx::mbar_wait_cluster(BAR(B_ACCE + buf), ((au >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d = tmem + buf * 256;
                    for (int s = 0; s < a.NS; ++s) {
                        if (np == 0) {
                            if (a.dbg) { long long t0 = clock64(); ptx::mbar_wait_cluster(BAR(B_AFULL + s), it & 1); dbg_acc[2] += clock64() - t0; }
                            else ptx::mbar_wait_cluster(BAR(B_AFULL + s), it & 1);
                            ptx::tc_fence_after();
                        }
                        const bool onehot = s >= a.KSe;
                        // entity slices: SWIZZLE_128B descriptors (k16 step = 32 B); one-hot slice: no-swizzle descriptors
                        const uint32_t ah = onehot ? aHi0 + s * SLICE_STEP : aHiSw + s * SLICE_STEP;
                        const uint32_t al = aLoSw + s * SLICE_STEP;
                        const uint32_t a_hi_word = onehot ? DHI : DHI_SW;
                        const uint32_t a_k16 = onehot ? K16_STEP : 2u;
                        for (int part = 0; part < parts; ++part) {
                            if (a.dbg) { long long t0 = clock64(); ptx::mbar_wait(BAR(stage), phase); ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase); dbg_acc[3] += clock64() - t0; }
                            else { ptx::mbar_wait(BAR(stage), phase); ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase); }
                            ptx::tc_fence_after();
                            const uint32_t bl = ring0 + stage * SLAB_STEP;
                            if (issue) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma2_f16(d, ptx::desc_join(a_hi_word, ah + k * a_k16), ptx::desc_join(DHI, bl + k * K16_STEP), idesc,
                                                   (part | s | k) != 0);
                                if (part == 0 && a.passes == 3 && !onehot) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::umma2_f16(d, ptx::desc_join(DHI_SW, al + k * 2u), ptx::desc_join(DHI, bl + k * K16_STEP), idesc, 1);
                                }
                                ptx::umma2_commit(BAR(R + stage), 3);
                            }
                            if (++stage == R) { stage = 0; phase ^= 1; }
                        }
                        if (issue && np == NP - 1) ptx::umma2_commit(BAR(B_AFREE + s), 3);
                    }
                    if (issue) ptx::umma2_commit(BAR(B_ACCF + buf), 3);
                }
            }
        } else {
            // ===== peer forwarder: tells the leader when this CTA's half of a slab has landed =====
            if (lane == 0) {
                uint32_t stage = 0, phase = 0;
                const uint32_t remote0 = ptx::mapa(BAR(B_PEER), 0);
                for (int64_t it = 0; it < my_tiles; ++it)
                    for (int np = 0; np < NP; ++np)
                        for (int s = 0; s < a.NS; ++s)
                            for (int part = 0; part < parts; ++part) {
                                ptx::mbar_wait(BAR(stage), phase);
                                ptx::mbar_arrive_cluster(remote0 + 8u * stage);
                                if (++stage == R) { stage = 0; phase ^= 1; }
                            }
            }
        }
    } else if (warp == 2) {
        // ===== TMA gather warp: lane l owns rows 4l..4l+3 of this CTA's tile =====
        const uint32_t dHi = ptx::smem_u32(sA_hi), dLo = ptx::smem_u32(sA_lo);
        const uint32_t afull_leader = ptx::mapa(BAR(B_AFULL), 0);
        int32_t eid[4];
        int combo[4];
        auto load_ids = [&](int64_t it) {
            const int64_t m = tile_of(it);
            const int64_t pt = m / a.T;
            const int t = (int)(m % a.T);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int64_t n = pt * TILE_M + lane * 4 + b;
                if (n >= a.N) n = a.N - 1;
                const int32_t* idp = a.ids + (n * a.T + t) * 3;
                int64_t id = (int64_t)idp[1] - 1;
                eid[b] = (int32_t)(id < 0 ? 0 : (id >= a.V_e ? a.V_e - 1 : id));
                int ta = idp[0] - 1, rb = idp[2] - 1;
                ta = ta < 0 ? 0 : (ta >= a.V_t ? a.V_t - 1 : ta);
                rb = rb < 0 ? 0 : (rb >= a.V_r ? a.V_r - 1 : rb);
                combo[b] = ta * a.V_r + rb;
            }
        };
        if (my_tiles > 0) load_ids(0);
        for (int64_t it = 0; it < my_tiles; ++it) {
            for (int s = 0; s < a.NS; ++s) {
                TWAIT(4, BAR(B_AFREE + s), (it & 1) ^ 1);
                if (s < a.KSe) {
                    if (lane == 0) ptx::mbar_arrive_expect_tx(BAR(B_ALAND + s), 2 * TILE_M * 128);
                    __syncwarp();
                    const uint32_t off = s * 16384 + lane * 512;
                    ptx::tma_gather4(dHi + off, &tm_hi, a.k_base + s * 64, eid[0], eid[1], eid[2], eid[3], BAR(B_ALAND + s));
                    ptx::tma_gather4(dLo + off, &tm_lo, a.k_base + s * 64, eid[0], eid[1], eid[2], eid[3], BAR(B_ALAND + s));
                } else {
                    const int o = s - a.KSe;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int r = lane * 4 + b;
#pragma unroll
                        for (int kc = 0; kc < 8; ++kc) {
                            uint32_t w[4] = {0u, 0u, 0u, 0u};
                            if ((combo[b] >> 3) == o * 8 + kc) w[(combo[b] & 7) >> 1] = 0x3F80u << (16 * (combo[b] & 1));   // bf16 1.0
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dHi + (s * 8 + kc) * 2048 + r * 16),
                                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                        }
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(afull_leader + 8u * s);
                }
            }
            if (it + 1 < my_tiles) load_ids(it + 1);
        }
    } else if (warp == 3) {
        // ===== A forwarder: entity slice s of this CTA's tile has landed -> leader's a_full[s] =====
        if (lane == 0) {
            const uint32_t afull_leader = ptx::mapa(BAR(B_AFULL), 0);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int s = 0; s < a.KSe; ++s) {
                    ptx::mbar_wait(BAR(B_ALAND + s), it & 1);
                    ptx::mbar_arrive_cluster(afull_leader + 8u * s);
                }
        }
    } else {
        // ===== epilogue: each CTA drains its own 128 rows x 256 columns; the leader is told when both are done =====
        const int wg = (warp - 4) >> 2;
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t acce_leader = ptx::mapa(BAR(B_ACCE), 0);
        const int64_t total_uses = my_tiles * NP;
        for (int64_t au = wg; au < total_uses; au += 2) {
            const int64_t it = au / NP;
            const int np = (int)(au % NP);
            const int buf = au & 1;
            const int64_t m = tile_of(it);
            const bool dup = (2 * (pair + it * npairs) + rank) >= a.num_tiles;
            float* xw = a.XW + ((m * (a.G4 / 8) + np * 32) * TILE_M + r) * 8;
            TWAIT(5, BAR(B_ACCF + buf), (au >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + buf * 256;
#pragma unroll
            for (int blk = 0; blk < 8; ++blk) {
                float v[32];
                ptx::tmem_ld32(tbase + blk * 32, v);
                ptx::tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float* dstp = xw + (int64_t)(blk * 4 + i) * TILE_M * 8;
                    if (a.accumulate && dup) continue;      // an odd tail tile runs on both CTAs of the last pair: only one may add
                    if (a.accumulate) {
                        const float4 o0 = *reinterpret_cast<const float4*>(dstp), o1 = *reinterpret_cast<const float4*>(dstp + 4);
                        v[8 * i] += o0.x; v[8 * i + 1] += o0.y; v[8 * i + 2] += o0.z; v[8 * i + 3] += o0.w;
                        v[8 * i + 4] += o1.x; v[8 * i + 5] += o1.y; v[8 * i + 6] += o1.z; v[8 * i + 7] += o1.w;
                    }
                    ptx::stg256(dstp, v + 8 * i);     // one 32-byte sector per thread
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(acce_leader + 8u * buf);
        }
    }
    if (a.dbg && lane == 0 && leader) {
        if (warp < 4 || warp == 4 || warp == 8) {
            for (int i = 0; i < 15; ++i) if (dbg_acc[i]) atomicAdd(a.dbg + i, (unsigned long long)dbg_acc[i]);
            if (warp == 1) atomicAdd(a.dbg + 15, (unsigned long long)(clock64() - dbg_t0));
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// kernel "recur": per 128-path tile, T LSTM steps.  TMEM: columns [0,256) two 128-column accumulators
// (a chunk = 32 hidden units x 4 gates), columns [256,256+HP) the cell state c (fp32, lane = path).
// Roles: warp 0 TMA producer (Wh slabs), warp 1 MMA issuer, warp 2 TMEM allocator + h loader, warps 3-18 epilogue:
// four groups of four warps (one per TMEM lane quarter); group g owns sub-block g (8 hidden units x 4 gates) of every chunk.
// h_t round trip: the epilogue writes h_t (bf16 hi/lo, already in the UMMA operand layout) to a per-CTA scratch that
// stays in L2; once every MMA of step t has retired, the h loader pulls it into shared memory slice by slice with TMA
// bulk copies, so the next step's MMAs start on slice 0 while the last chunks are still in the epilogue.  Keeping h_t
// out of the register file leaves the cell math enough registers to overlap its MUFU chains.

struct RecurArgs {
    const float* XW; const bf16* W; float* hT; bf16* hscratch; int64_t N; int T; int H; int G4; int64_t num_tiles; int passes;
    unsigned long long* dbg;
    // scores-only path (A5 fused): row classId of the head weight and the [4][npad] partial dot products h_T . wo, one per epilogue
    // group; null = write h_T for the head kernels instead
    const float* wo; float* part4; int64_t npad;
    // XW row of (path tile, t, r) = slot_of[pos[(tile*T + t)*128 + r]] (a row of the projection of the DISTINCT step inputs,
    // see the dedup kernels); null = its own row (tile tile*T + t, row r)
    const uint32_t* pos; const uint32_t* slot_of;
    // training forward (TRAIN = true): the activations the backward needs, written by the epilogue as it computes them
    //   cache  [tile][t][chunk j][sub-block sb][i,g,f,o,c][half 2][row 128][4] fp32 (units 4 half .. +3 of the thread's 8): each
    //          st.global.v4 of a warp is 512 contiguous bytes
    //   hcache [tile][t][hi|lo][kc Hp/8][row 128][8] bf16: h_t in the UMMA operand layout (what the next step's MMAs read, and the
    //          B operand of the weight-gradient product dG^T h_{t-1}); replaces the per-CTA scratch
    float* cache; bf16* hcache;
    // leading-token reuse (tc_prefix.cu; scoring only): the batch arrives sorted by the number of leading copies of one token;
    // a pair of tiles whose first row has ps = t0 starts at step t0 from the chain state: chain_u[t0] = that state's recurrent
    // pre-activations (accumulator layout and scale), chain_c[t0] its cell state.  perm = sorted position -> path (outputs).
    const int32_t* perm; const uint8_t* ps; const float* chain_u; const float* chain_c;
};

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

__device__ __forceinline__ void pack_split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // packed cvt.rn.bf16x2.f32 for the hi pair, exact fp32 residuals, packed conversion for the lo pair
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const float r0 = v[2 * i] - __uint_as_float(hb << 16), r1 = v[2 * i + 1] - __uint_as_float(hb & 0xFFFF0000u);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        wh[i] = hb;
        wl[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    lo = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

// recur_kernel<HP, PAIR, TRAIN>.  PAIR = true: launched as clusters of two CTAs (cta_group::2).  Each CTA still owns one 128-path
// tile (its own h in shared memory, c and accumulators in its own TMEM, its own epilogue), but one MMA multiplies both
// tiles (M = 256) by the same 128 gate columns, and each CTA stages only ITS 64-column half of every weight slab: the
// L2 -> SM weight stream -- more than half of this kernel's L2 traffic, which runs at ~85 % of the chip's measured L2
// throughput in the 1-CTA form -- is halved.  The leader CTA's warp 1 issues the MMAs; the peer's warp 1 forwards
// "my half-slab landed" / "my h slice landed" to the leader; both epilogues hand accumulators back to the leader;
// MMA completion is multicast to both CTAs (tcgen05.commit ... multicast::cluster).
template <int HP, bool PAIR, bool TRAIN>
__global__ void __launch_bounds__(RECUR_THREADS, 1) recur_kernel(RecurArgs a) {
    constexpr int NCH = HP / 32;        // 128-column chunks per step
    constexpr int KS = HP / 64;         // 64-wide k slices of h
    constexpr int R = PAIR ? RECUR_RING_PAIR : RECUR_RING;
    constexpr uint32_t B_ROWS = PAIR ? 64 : 128;                    // gate columns of a slab staged by this CTA
    constexpr uint32_t STAGE_ELEMS = 8 * B_ROWS * 8;                // [kc 8][n B_ROWS][8]
    constexpr uint32_t H_BYTES = TILE_M * HP * 2;   // one of hi / lo
    constexpr uint32_t SLICE_BYTES = 8 * 2048;      // one k-slice (64 units) of hi or lo
    constexpr int EPI_WARPS = 16;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    uint8_t* sH_hi = smem;
    uint8_t* sH_lo = smem + H_BYTES;
    uint8_t* sRing = smem + 2 * H_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + R * STAGE_ELEMS * 2);
    // bars: [0,R) slab_full  [R,2R) slab_empty  acc_full[2]  acc_empty[2] (leader's: both CTAs' epilogue warps)  mma_step_done
    //       h_ready[KS] (TMA tx)  h_written[KS]  and on the leader: peer_slab_full[R]  peer_h_ready[KS]
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    constexpr int B_ACCF = 2 * R, B_ACCE = 2 * R + 2, B_STEP = 2 * R + 4, B_HREADY = 2 * R + 5, B_HWRIT = B_HREADY + KS,
                  B_PSLAB = B_HWRIT + KS, B_PH = B_PSLAB + R, NBARS = B_PH + KS;
    static_assert(NBARS * 8 + 4 <= RECUR_BAR_BYTES, "barrier block too small");
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);

    if (threadIdx.x == 0) {
        for (int i = 0; i < R; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(R + i), 1); ptx::mbar_init(BAR(B_PSLAB + i), 1); }
        ptx::mbar_init(BAR(B_ACCF), 1); ptx::mbar_init(BAR(B_ACCF + 1), 1);
        ptx::mbar_init(BAR(B_ACCE), (PAIR ? 2 : 1) * EPI_WARPS); ptx::mbar_init(BAR(B_ACCE + 1), (PAIR ? 2 : 1) * EPI_WARPS);
        ptx::mbar_init(BAR(B_STEP), 1);
        for (int s = 0; s < KS; ++s) {
            ptx::mbar_init(BAR(B_HREADY + s), 1);
            ptx::mbar_init(BAR(B_HWRIT + s), 2 * EPI_WARPS * 32);     // 2 chunks x 512 epilogue threads per k-slice
            ptx::mbar_init(BAR(B_PH + s), 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
        else { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    long long dbg_acc[16] = {0};
    const long long dbg_t0 = clock64();
    // tiles: 1-CTA form: blockIdx.x, +gridDim.x, ...; pair form: pair p takes tiles 2(p + i*npairs) + rank (an odd tail tile is
    // computed by both CTAs of the last pair: identical values, identical stores)
    const int64_t npairs = gridDim.x / 2, pair = blockIdx.x / 2, n_tp = (a.num_tiles + 1) / 2;
    const int64_t my_tiles = PAIR ? (pair < n_tp ? (n_tp - pair + npairs - 1) / npairs : 0)
                                  : (a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto tile_of = [&](int64_t it) -> int64_t {
        if (!PAIR) return blockIdx.x + it * gridDim.x;
        const int64_t m = 2 * (pair + it * npairs) + rank;
        return m < a.num_tiles ? m : a.num_tiles - 1;
    };
    const int T = a.T;
    // first step of the it-th tile (pair): the fewest leading copies of the shared token among its rows = those of its first row
    auto t0_of = [&](int64_t it) -> int {
        if (TRAIN || !a.ps) return 0;
        int64_t row0 = (PAIR ? 2 * (pair + it * npairs) : (int64_t)blockIdx.x + it * gridDim.x) * TILE_M;
        if (row0 >= a.N) row0 = a.N - 1;
        return a.ps[row0];
    };
    const uint32_t idesc = ptx::make_idesc_f16(1, PAIR ? 2 * TILE_M : TILE_M, 128);
    const int slabs_per_k = a.passes == 3 ? 2 : 1;
    uint8_t* const scratch0 = reinterpret_cast<uint8_t*>(a.hscratch) + (size_t)blockIdx.x * 2 * H_BYTES;   // [hi | lo] image of sH
    // where h_t of (tile, t) lives in global memory: the per-CTA L2-resident scratch, or (training) its slot of the h cache
    auto h_image = [&](int64_t tile, int t) -> uint8_t* {
        return TRAIN ? reinterpret_cast<uint8_t*>(a.hcache) + (size_t)(tile * T + t) * 2 * H_BYTES : scratch0;
    };

    if (warp == 0) {
        // ===== TMA producer: Wh slabs (this CTA's columns); the sequence is identical for every MMA step (t >= 1) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int64_t steps = 0;
            for (int64_t it = 0; it < my_tiles; ++it) steps += T - 1 - t0_of(it);
            for (int64_t st = 0; st < steps; ++st)
                for (int j = 0; j < NCH; ++j)
                    for (int s = 0; s < KS; ++s)
                        for (int part = 0; part < slabs_per_k; ++part) {
                            TWAIT(8, BAR(R + stage), phase ^ 1);
                            ptx::mbar_arrive_expect_tx(BAR(stage), STAGE_ELEMS * 2);
                            const bf16* src = a.W + ((int64_t)(j * KS + s) * 2 + part) * B_SLAB_ELEMS + (PAIR ? rank * STAGE_ELEMS : 0u);
                            ptx::tma_bulk_g2s(ptx::smem_u32(sRing + stage * STAGE_ELEMS * 2), src, STAGE_ELEMS * 2, BAR(stage));
                            if (++stage == R) { stage = 0; phase ^= 1; }
                        }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer: uniform control flow on the whole warp, one elected lane issues =====
            const bool issue = ptx::elect_one();
            uint32_t stage = 0, phase = 0;
            uint32_t ug0 = 0, ug1 = 0;
            const uint32_t hHi0 = ptx::desc_lo(ptx::smem_u32(sH_hi), 2048), hLo0 = ptx::desc_lo(ptx::smem_u32(sH_lo), 2048);
            const uint32_t ring0 = ptx::desc_lo(ptx::smem_u32(sRing), B_ROWS * 16);
            constexpr uint32_t DHI = ptx::desc_hi(128);
            constexpr uint32_t SLAB_STEP = (STAGE_ELEMS * 2) >> 4, K16_A = (2 * 2048) >> 4, K16_B = (2 * B_ROWS * 16) >> 4,
                               SLICE_STEP = (8 * 2048) >> 4;
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                if (PAIR) ptx::umma2_f16(d, da, db, idesc, acc); else ptx::umma_f16(d, da, db, idesc, acc);
            };
            auto commit = [&](uint32_t bar) { if (PAIR) ptx::umma2_commit(bar, 3); else ptx::umma_commit(bar); };
            auto wait_slab = [&]() {
                TWAIT(11, BAR(stage), phase);
                if (PAIR) TWAIT(11, BAR(B_PSLAB + stage), phase);
                ptx::tc_fence_after();
            };
            uint32_t cnt = 0;                                   // 1-based index of the MMA step
            for (int64_t it = 0; it < my_tiles; ++it) {
                for (int t = t0_of(it) + 1; t < T; ++t) {
                    ++cnt;
#pragma unroll 1
                    for (int j = 0; j < NCH; ++j) {
                        const int g = j & 1;
                        uint32_t& ug = g ? ug1 : ug0;
                        TWAIT(9, BAR(B_ACCE + g), (ug & 1) ^ 1);
                        ptx::tc_fence_after();
                        const uint32_t d = tmem + g * 128;
#pragma unroll 1
                        for (int s = 0; s < KS; ++s) {
                            if (j == 0) {          // h_{t-1} slice s has landed in shared memory (of both CTAs)
                                TWAIT(10, BAR(B_HREADY + s), (cnt - 1) & 1);
                                if (PAIR) TWAIT(10, BAR(B_PH + s), (cnt - 1) & 1);
                                ptx::tc_fence_after();
                            }
                            const uint32_t ah = hHi0 + s * SLICE_STEP, al = hLo0 + s * SLICE_STEP;
                            wait_slab();
                            uint32_t bl = ring0 + stage * SLAB_STEP;
                            if (issue) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma(d, ptx::desc_join(DHI, ah + k * K16_A), ptx::desc_join(DHI, bl + k * K16_B), (s | k) != 0);
                                if (a.passes == 3) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma(d, ptx::desc_join(DHI, al + k * K16_A), ptx::desc_join(DHI, bl + k * K16_B), 1);
                                }
                                commit(BAR(R + stage));
                            }
                            if (++stage == R) { stage = 0; phase ^= 1; }
                            if (a.passes == 3) {
                                wait_slab();
                                bl = ring0 + stage * SLAB_STEP;
                                if (issue) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma(d, ptx::desc_join(DHI, ah + k * K16_A), ptx::desc_join(DHI, bl + k * K16_B), 1);
                                    commit(BAR(R + stage));
                                }
                                if (++stage == R) { stage = 0; phase ^= 1; }
                            }
                        }
                        if (issue) commit(BAR(B_ACCF + g));
                        ++ug;
                    }
                    if (issue) commit(BAR(B_STEP));     // every MMA reading h_{t-1} has retired
                }
            }
        } else if (lane == 0) {
            // ===== peer forwarder: mirrors the issuer's wait sequence on this CTA's barriers and arrives on the leader's =====
            uint32_t stage = 0, phase = 0, cnt = 0;
            const uint32_t pslab = ptx::mapa(BAR(B_PSLAB), 0), ph = ptx::mapa(BAR(B_PH), 0);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int t = t0_of(it) + 1; t < T; ++t) {
                    ++cnt;
                    for (int j = 0; j < NCH; ++j)
                        for (int s = 0; s < KS; ++s) {
                            if (j == 0) {
                                ptx::mbar_wait(BAR(B_HREADY + s), (cnt - 1) & 1);
                                ptx::mbar_arrive_cluster(ph + 8u * s);
                            }
                            for (int part = 0; part < slabs_per_k; ++part) {
                                ptx::mbar_wait(BAR(stage), phase);
                                ptx::mbar_arrive_cluster(pslab + 8u * stage);
                                if (++stage == R) { stage = 0; phase ^= 1; }
                            }
                        }
                }
        }
    } else if (warp == 2) {
        // ===== h loader: scratch (L2) -> shared memory, one k-slice at a time =====
        if (lane == 0) {
            uint32_t idx = 0;                                   // 1-based index of the h vector being loaded
            for (int64_t it = 0; it < my_tiles; ++it) {
                for (int t = t0_of(it); t < T - 1; ++t) {
                    ++idx;
                    const uint8_t* scratch = h_image(tile_of(it), t);
                    // all MMAs that read the previous contents of sH must have retired: that is MMA step idx-1
                    if (idx >= 2) TWAIT(12, BAR(B_STEP), (idx - 2) & 1);
                    for (int s = 0; s < KS; ++s) {
                        TWAIT(13, BAR(B_HWRIT + s), (idx - 1) & 1);
                        ptx::mbar_arrive_expect_tx(BAR(B_HREADY + s), 2 * SLICE_BYTES);
                        ptx::tma_bulk_g2s(ptx::smem_u32(sH_hi) + s * SLICE_BYTES, scratch + s * SLICE_BYTES, SLICE_BYTES, BAR(B_HREADY + s));
                        ptx::tma_bulk_g2s(ptx::smem_u32(sH_lo) + s * SLICE_BYTES, scratch + H_BYTES + s * SLICE_BYTES, SLICE_BYTES, BAR(B_HREADY + s));
                    }
                }
            }
        }
    } else {
        // ===== epilogue: 4 groups of 4 warps; group g owns sub-block g (8 hidden units x 4 gates) of EVERY chunk, so all
        // 16 warps drain one accumulator together: the latency of a chunk's cell math (what the MMA issuer waits for when
        // it recycles an accumulator, and what delays h_t at the step boundary) is a quarter of a single group's =====
        const int sb = (warp - 3) >> 2;               // this group's sub-block (a group = 4 consecutive warps = the 4 TMEM lane quarters)
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const int64_t tile_xw = (int64_t)(a.G4 / 4) * TILE_M * 4;  // floats per tile-step
        const uint32_t acce_leader = PAIR ? ptx::mapa(BAR(B_ACCE), 0) : BAR(B_ACCE);
        const bool h_vec4 = (a.H & 3) == 0;
        const bool wo_vec4 = h_vec4 && (reinterpret_cast<uintptr_t>(a.wo) & 15) == 0;    // the flat parameter vector aligns nothing
        uint32_t ug0 = 0, ug1 = 0;
        // XW of (chunk j, this sub-block), gate q: the 8 columns j*128 + q*32 + sb*8 .. +7 = column group j*16 + q*4 + sb.
        // XW layout [tile][column group (4Hp/8)][row 128][8 floats]: the two 128-bit loads of a thread share one 32-byte sector
        // (a single 256-bit load measured 4 % slower here; the projection's 256-bit stores are fine).
#define KPRN_LOAD_XW(dst, base, j)                                                                                    \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                                  \
        const float4 x0 = *reinterpret_cast<const float4*>((base) + (int64_t)((j) * 16 + q * 4 + sb) * TILE_M * 8);     \
        const float4 x1 = *reinterpret_cast<const float4*>((base) + (int64_t)((j) * 16 + q * 4 + sb) * TILE_M * 8 + 4); \
        dst[q * 8 + 0] = x0.x; dst[q * 8 + 1] = x0.y; dst[q * 8 + 2] = x0.z; dst[q * 8 + 3] = x0.w;                   \
        dst[q * 8 + 4] = x1.x; dst[q * 8 + 5] = x1.y; dst[q * 8 + 6] = x1.z; dst[q * 8 + 7] = x1.w;                   \
    }
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t pt = tile_of(it);
            const int64_t n = pt * TILE_M + r;
            const int64_t n_out = (!TRAIN && a.perm && n < a.N) ? a.perm[n] : n;      // where this row's result goes (sorted batch: its path)
            // row index of this thread's XW row per step (rows past N in a ragged last tile: any valid row)
            auto row_of = [&](int t) -> uint32_t {
                const uint32_t p = a.pos[(pt * T + t) * TILE_M + r];
                return p == 0xFFFFFFFFu ? 0u : a.slot_of[p];
            };
            const int t0 = t0_of(it);
            const float* const cu0 = a.chain_u + (size_t)t0 * a.G4;        // used when t0 > 0
            const float* const cc0 = a.chain_c + (size_t)t0 * HP;
            uint32_t slot_next = a.pos ? row_of(t0) : 0u;
            for (int t = t0; t < T; ++t) {
                const uint32_t slot = slot_next;
                if (a.pos && t + 1 < T) slot_next = row_of(t + 1);       // next step's row index, in flight during this step
                // rows are numbered in the order the recurrence reads them (see the dedup kernels): the first occurrences inside a
                // warp are consecutive rows of one tile, repeated triples (pad steps, a pair's shared user / item) share a row
                const float* xw_t = a.pos ? a.XW + (int64_t)(slot >> 7) * tile_xw + (slot & 127) * 8
                                           : a.XW + (pt * T + t) * tile_xw + r * 8;
                const bool last = (t == T - 1);
                uint8_t* const scratch = h_image(pt, t);
                float* const cache_t = TRAIN ? a.cache + (size_t)(pt * T + t) * NCH * 20480 + r * 4 : nullptr;
                float xa[32], xb[32];                 // XW of the current / next chunk, ping-pong (no register copies)
                float dot = 0.f;                      // last step, scores-only: this thread's share of h_T . wo
                KPRN_LOAD_XW(xa, xw_t, 0)
                auto chunk = [&](const int j, float (&xc)[32], float (&xn)[32]) {
                    const int g = j & 1;
                    uint32_t& ug = g ? ug1 : ug0;
                    if (!last && j > 0) {
                        // publish chunk j-1's h writes: generic-proxy global writes -> async-proxy (TMA) reader
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        ptx::mbar_arrive(BAR(B_HWRIT + ((j - 1) >> 1)));
                    }
                    if (t > t0) {
                        TWAIT(14, BAR(B_ACCF + g), ug & 1);
                        ptx::tc_fence_after();
                    }
                    const uint32_t acc = tmem + lane_base + g * 128;
                    const uint32_t cst = tmem + lane_base + 256 + j * 32;
                    float cp[8];
                    if (t > t0) {
                        float ai[8], ag[8], af[8], ao[8];
                        ptx::tmem_ld8(acc + 0 * 32 + sb * 8, ai);
                        ptx::tmem_ld8(acc + 1 * 32 + sb * 8, ag);
                        ptx::tmem_ld8(acc + 2 * 32 + sb * 8, af);
                        ptx::tmem_ld8(acc + 3 * 32 + sb * 8, ao);
                        ptx::tmem_ld8(cst + sb * 8, cp);
                        ptx::tc_wait_ld();
                        // the accumulator has been read: hand it back to the MMA issuer before the cell math
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR) ptx::mbar_arrive_cluster(acce_leader + 8u * g); else ptx::mbar_arrive(BAR(B_ACCE + g));
                        }
                        ++ug;
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            ptx::unpack2(ptx::add2(ptx::pack2(xc[e], xc[e + 1]), ptx::pack2(ai[e], ai[e + 1])), xc[e], xc[e + 1]);
                            ptx::unpack2(ptx::add2(ptx::pack2(xc[8 + e], xc[9 + e]), ptx::pack2(ag[e], ag[e + 1])), xc[8 + e], xc[9 + e]);
                            ptx::unpack2(ptx::add2(ptx::pack2(xc[16 + e], xc[17 + e]), ptx::pack2(af[e], af[e + 1])), xc[16 + e], xc[17 + e]);
                            ptx::unpack2(ptx::add2(ptx::pack2(xc[24 + e], xc[25 + e]), ptx::pack2(ao[e], ao[e + 1])), xc[24 + e], xc[25 + e]);
                        }
                    } else if (t0 == 0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) cp[e] = 0.f;      // t = 0: h = c = 0 and no MMA ran
                    } else {
                        // first step after t0 shared leading tokens: the state every such row would have reached (tc_prefix.cu)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 u0 = *reinterpret_cast<const float4*>(cu0 + (j * 16 + q * 4 + sb) * 8);
                            const float4 u1 = *reinterpret_cast<const float4*>(cu0 + (j * 16 + q * 4 + sb) * 8 + 4);
                            xc[q * 8 + 0] += u0.x; xc[q * 8 + 1] += u0.y; xc[q * 8 + 2] += u0.z; xc[q * 8 + 3] += u0.w;
                            xc[q * 8 + 4] += u1.x; xc[q * 8 + 5] += u1.y; xc[q * 8 + 6] += u1.z; xc[q * 8 + 7] += u1.w;
                        }
                        const float4 c0 = *reinterpret_cast<const float4*>(cc0 + j * 32 + sb * 8), c1 = *reinterpret_cast<const float4*>(cc0 + j * 32 + sb * 8 + 4);
                        cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w; cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
                    }
                    if (j + 1 < NCH) { KPRN_LOAD_XW(xn, xw_t, j + 1) }     // next chunk's XW, in flight during the cell math
                    float cn[8], hv[8];
                    // FastLSTM cell, two cells per instruction on the FMA pipe (add/mul/fma .f32x2 = SASS FADD2 / FMUL2 / FFMA2).
                    // The pre-activations arrive pre-scaled (gate_scale): xc = -x log2e (i,f,o), +2x log2e (g), so each gate costs
                    // one ex2 and one rcp (10 MUFU per cell); no clamps needed: ex2 -> inf gives rcp -> 0.
                    {
                        const uint64_t one2 = ptx::pack2(1.f, 1.f), mtwo2 = ptx::pack2(-2.f, -2.f);
                        const uint64_t k2 = ptx::pack2(2.8853900817779268f, 2.8853900817779268f);
                        uint64_t ki = 0, kg = 0, kf = 0, ko = 0, kc2 = 0;      // TRAIN: the even pair, kept until its odd neighbour is ready
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int e = 2 * p;
                            uint64_t i_, f_, o_, g_, c2, th;
                            if constexpr (!TRAIN) {
                                // the XU pipe (16 MUFU lanes / clk / SM) is this kernel's bound: the four gate denominators of a cell
                                // share ONE reciprocal (batch inversion: 9 packed multiplies on the FMA pipe instead of 3 MUFU.RCP),
                                // the two tanh(c) denominators of a cell pair share one.  Exponents are clamped at 2^30 (a gate off by
                                // < 1e-9) and 2^60 so that the products stay finite.
                                const uint64_t pi = ptx::add2(ptx::pack2(ex2f(fminf(xc[e], 30.f)), ex2f(fminf(xc[e + 1], 30.f))), one2);
                                const uint64_t pg = ptx::add2(ptx::pack2(ex2f(fminf(xc[8 + e], 30.f)), ex2f(fminf(xc[9 + e], 30.f))), one2);
                                const uint64_t pf = ptx::add2(ptx::pack2(ex2f(fminf(xc[16 + e], 30.f)), ex2f(fminf(xc[17 + e], 30.f))), one2);
                                const uint64_t po = ptx::add2(ptx::pack2(ex2f(fminf(xc[24 + e], 30.f)), ex2f(fminf(xc[25 + e], 30.f))), one2);
                                const uint64_t P2 = ptx::mul2(pi, pg), P3 = ptx::mul2(P2, pf), P4 = ptx::mul2(P3, po);
                                const uint64_t R4 = ptx::pack2(rcpf(ptx::lo2(P4)), rcpf(ptx::hi2(P4)));
                                o_ = ptx::mul2(R4, P3); const uint64_t R3 = ptx::mul2(R4, po);
                                f_ = ptx::mul2(R3, P2); const uint64_t R2 = ptx::mul2(R3, pf);
                                i_ = ptx::mul2(R2, pg);
                                g_ = ptx::fma2(ptx::mul2(R2, pi), mtwo2, one2);
                                c2 = ptx::fma2(i_, g_, ptx::mul2(f_, ptx::pack2(cp[e], cp[e + 1])));
                                const uint64_t a2 = ptx::mul2(c2, k2);
                                const float t0_ = 1.f + ex2f(fminf(ptx::lo2(a2), 60.f)), t1_ = 1.f + ex2f(fminf(ptx::hi2(a2), 60.f));
                                const float rt = rcpf(t0_ * t1_);
                                th = ptx::fma2(ptx::pack2(rt * t1_, rt * t0_), mtwo2, one2);
                            } else {        // training forward: HBM-bound on its cache stores, and short of registers -- the plain form
                                const uint64_t pi = ptx::add2(ptx::pack2(ex2f(xc[e]), ex2f(xc[e + 1])), one2);
                                const uint64_t pg = ptx::add2(ptx::pack2(ex2f(xc[8 + e]), ex2f(xc[9 + e])), one2);
                                const uint64_t pf = ptx::add2(ptx::pack2(ex2f(xc[16 + e]), ex2f(xc[17 + e])), one2);
                                const uint64_t po = ptx::add2(ptx::pack2(ex2f(xc[24 + e]), ex2f(xc[25 + e])), one2);
                                i_ = ptx::pack2(rcpf(ptx::lo2(pi)), rcpf(ptx::hi2(pi)));
                                f_ = ptx::pack2(rcpf(ptx::lo2(pf)), rcpf(ptx::hi2(pf)));
                                o_ = ptx::pack2(rcpf(ptx::lo2(po)), rcpf(ptx::hi2(po)));
                                g_ = ptx::fma2(ptx::pack2(rcpf(ptx::lo2(pg)), rcpf(ptx::hi2(pg))), mtwo2, one2);
                                c2 = ptx::fma2(i_, g_, ptx::mul2(f_, ptx::pack2(cp[e], cp[e + 1])));
                                const uint64_t a2 = ptx::mul2(c2, k2);
                                const uint64_t pt2 = ptx::add2(ptx::pack2(ex2f(ptx::lo2(a2)), ex2f(ptx::hi2(a2))), one2);
                                th = ptx::fma2(ptx::pack2(rcpf(ptx::lo2(pt2)), rcpf(ptx::hi2(pt2))), mtwo2, one2);
                            }
                            const uint64_t h2 = ptx::mul2(o_, th);
                            cn[e] = ptx::lo2(c2); cn[e + 1] = ptx::hi2(c2);
                            hv[e] = ptx::lo2(h2); hv[e + 1] = ptx::hi2(h2);
                            if (TRAIN) {
                                // activation cache: two pairs per 128-bit store; a warp store is 512 contiguous bytes
                                if (p & 1) {
                                    float* cc = cache_t + (size_t)(j * 4 + sb) * 5120 + (p >> 1) * 512;
                                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cc), "f"(ptx::lo2(ki)), "f"(ptx::hi2(ki)), "f"(ptx::lo2(i_)), "f"(ptx::hi2(i_)) : "memory");
                                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cc + 1024), "f"(ptx::lo2(kg)), "f"(ptx::hi2(kg)), "f"(ptx::lo2(g_)), "f"(ptx::hi2(g_)) : "memory");
                                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cc + 2048), "f"(ptx::lo2(kf)), "f"(ptx::hi2(kf)), "f"(ptx::lo2(f_)), "f"(ptx::hi2(f_)) : "memory");
                                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cc + 3072), "f"(ptx::lo2(ko)), "f"(ptx::hi2(ko)), "f"(ptx::lo2(o_)), "f"(ptx::hi2(o_)) : "memory");
                                    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cc + 4096), "f"(ptx::lo2(kc2)), "f"(ptx::hi2(kc2)), "f"(ptx::lo2(c2)), "f"(ptx::hi2(c2)) : "memory");
                                } else { ki = i_; kg = g_; kf = f_; ko = o_; kc2 = c2; }
                            }
                        }
                    }
                    ptx::tmem_st8(cst + sb * 8, cn);
                    if (!last) {
                        uint4 hi, lo;
                        pack_split8(hv, hi, lo);
                        const int kc = 4 * j + sb;
                        *reinterpret_cast<uint4*>(scratch + kc * 2048 + r * 16) = hi;
                        *reinterpret_cast<uint4*>(scratch + H_BYTES + kc * 2048 + r * 16) = lo;
                        // publication of these writes (proxy fence + arrive) is deferred by one chunk: by then the
                        // stores have drained and the fence costs nothing (see the top of the lambda / after the loop)
                    } else {
                      if (a.part4) {
                        // A5 fused (nn.Linear row classId on SelectTable(-1)): partial dot product over this thread's 8 units;
                        // padded units (>= H) hold h = 0 exactly, their weights are not read
                        const int u0 = j * 32 + sb * 8;
                        if (wo_vec4 && u0 + 8 <= a.H) {
                            const float4 w0 = *reinterpret_cast<const float4*>(a.wo + u0), w1 = *reinterpret_cast<const float4*>(a.wo + u0 + 4);
                            dot = fmaf(hv[0], w0.x, dot); dot = fmaf(hv[1], w0.y, dot); dot = fmaf(hv[2], w0.z, dot); dot = fmaf(hv[3], w0.w, dot);
                            dot = fmaf(hv[4], w1.x, dot); dot = fmaf(hv[5], w1.y, dot); dot = fmaf(hv[6], w1.z, dot); dot = fmaf(hv[7], w1.w, dot);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (u0 + e < a.H) dot = fmaf(hv[e], a.wo[u0 + e], dot);
                        }
                      }
                      if ((TRAIN || !a.part4) && n < a.N) {
                        // SelectTable(-1): last hidden state, standard [N,H] layout for the head
                        float* dst = a.hT + n_out * a.H + j * 32 + sb * 8;
                        if (h_vec4 && j * 32 + sb * 8 + 8 <= a.H) {      // full-sector stores (scalar stores cost 8 partial sectors)
                            *reinterpret_cast<float4*>(dst) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                            *reinterpret_cast<float4*>(dst + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (j * 32 + sb * 8 + e < a.H) dst[e] = hv[e];
                        }
                      }
                    }
                };
#pragma unroll 1
                for (int j = 0; j < NCH; j += 2) {
                    chunk(j, xa, xb);
                    chunk(j + 1, xb, xa);
                }
                if (!last) {
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    ptx::mbar_arrive(BAR(B_HWRIT + ((NCH - 1) >> 1)));
                }
                if (last && a.part4 && n < a.N) a.part4[sb * a.npad + n_out] = dot;     // summed (fixed order) + bias in the pool kernel
                ptx::tc_wait_st();        // c_t is read back (tcgen05.ld) by this thread in the next step
            }
        }
#undef KPRN_LOAD_XW
    }
    if (a.dbg && lane == 0 && leader) {
        // one representative lane per role: warps 0..2 and the first warp of the first two epilogue groups
        if (warp <= 3 || warp == 7) {
            for (int i = 0; i < 15; ++i) if (dbg_acc[i]) atomicAdd(a.dbg + i, (unsigned long long)dbg_acc[i]);
            if (warp == 1) atomicAdd(a.dbg + 15, (unsigned long long)(clock64() - dbg_t0));
        }
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    if (warp == 2) { if (PAIR) ptx::tmem_dealloc2(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
}


// ---------------------------------------------------------------------------------------------
// Step-input deduplication.  The projection x_t Wx^T + bx depends only on the (type, entity, relation) triple of a step,
// and a batch repeats triples heavily: every left-pad step is the same triple, all paths of a (user,item) pair start at
// the same user and end at the same item, popular entities recur (55 % distinct in the synthetic bench batches, fewer in
// mined paths).  So the projection runs once per DISTINCT triple and the recurrence reads its XW rows through a
// row-index map.  Rows are numbered by FIRST OCCURRENCE IN THE ORDER THE RECURRENCE READS THEM -- o = ((tile*T + t)*128 + r)
// for path n = tile*128 + r, step t -- so the first occurrences inside one warp load are consecutive rows (coalesced as
// before) and repeats inside a warp hit the same sector.  Open-addressing hash table (64-bit keys, linear probing) with an
// atomicMin of o per key and a block-level prefix sum over the first-occurrence flags; two small kernels, no host round
// trip: the projection kernels read the distinct count from device memory, the recurrence resolves row = slot_of[pos[o]].
constexpr unsigned long long DEDUP_EMPTY = ~0ull;
constexpr int DEDUP_BLOCK = 1024;          // read-order indices per block of the flag scan

__device__ __forceinline__ uint32_t dedup_hash(unsigned long long k, uint32_t mask) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k & mask;
}
// read-order index o -> flat step index (n*T + t), or -1 for the rows of a ragged last tile
__device__ __forceinline__ int64_t dedup_step_of(int64_t o, int64_t N, int T) {
    const int64_t tile_t = o >> 7, tile = tile_t / T;
    const int64_t n = tile * TILE_M + (o & 127);
    return n < N ? n * T + (tile_t - tile * T) : -1;
}

// pos[o] = table position of the triple read at o; first[pos] = min o over its occurrences.  Equal triples inside a warp
// (a pad column of a tile is one triple) elect one lane; a plain load before the CAS keeps hot keys from serialising.
// mode 0: key = the (type, entity, relation) triple of ids[N,T,3], index o in READ ORDER (what the projection dedups);
// mode 1: key = the entity row alone, same order (distinct rows of a training batch: compact gradient / sparse optimizer);
// mode 2: key = keys[o] (a flat int32 list, negative = skip) -- the union of the row lists gathered from all ranks.
__global__ void dedup_insert_kernel(const int32_t* __restrict__ ids, int64_t n_order, int64_t N, int T, int64_t V_t, int64_t V_e,
                                    int64_t V_r, unsigned long long* table, uint32_t* first, uint32_t mask, uint32_t* __restrict__ pos,
                                    int32_t* err, int mode) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = o < n_order ? (mode == 2 ? (ids[o] >= 0 ? o : -1) : dedup_step_of(o, N, T)) : -1;
    const bool active = i >= 0;
    unsigned long long key = DEDUP_EMPTY;
    if (active && mode == 2) {
        key = (unsigned long long)ids[o];
    } else if (active) {
        int64_t ta = (int64_t)ids[i * 3] - 1, en = (int64_t)ids[i * 3 + 1] - 1, rb = (int64_t)ids[i * 3 + 2] - 1;
        // a bad id raises in the reference (nn.LookupTable index out of range): flagged here, reported by the forward call
        if (ta < 0 || ta >= V_t || en < 0 || en >= V_e || rb < 0 || rb >= V_r) atomicExch(err, 1);
        ta = ta < 0 ? 0 : (ta >= V_t ? V_t - 1 : ta);
        en = en < 0 ? 0 : (en >= V_e ? V_e - 1 : en);
        rb = rb < 0 ? 0 : (rb >= V_r ? V_r - 1 : rb);
        key = mode == 1 ? (unsigned long long)en
                        : ((unsigned long long)en * (unsigned long long)V_t + (unsigned long long)ta) * (unsigned long long)V_r + (unsigned long long)rb;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;           // lowest lane = lowest o of the group
    uint32_t p = 0;
    if ((int)(threadIdx.x & 31) == leader && active) {
        uint32_t h = dedup_hash(key, mask);
        for (;;) {
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(table + h);
            if (cur == DEDUP_EMPTY) cur = atomicCAS(table + h, DEDUP_EMPTY, key);
            if (cur == DEDUP_EMPTY || cur == key) break;
            h = (h + 1) & mask;
        }
        p = h;
        if (*reinterpret_cast<volatile uint32_t*>(first + p) > (uint32_t)o) atomicMin(first + p, (uint32_t)o);
    }
    p = __shfl_sync(0xffffffffu, p, leader);
    if (o < n_order) pos[o] = active ? p : 0xFFFFFFFFu;
}

// block-wide exclusive scan of one flag per thread x 4 (256 threads): returns this thread's exclusive offset, total in *tot
__device__ __forceinline__ int dedup_block_scan(int c, int* tot) {
    __shared__ int warp_tot[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    int off = 0, t = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { if (j < w) off += warp_tot[j]; t += warp_tot[j]; }
    __syncthreads();
    *tot = t;
    return off + incl - c;
}
__device__ __forceinline__ int dedup_flags(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ first, int64_t n_order,
                                           int64_t o0, bool (&f)[4]) {
    int c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t o = o0 + j;
        const uint32_t p = o < n_order ? pos[o] : 0xFFFFFFFFu;
        f[j] = p != 0xFFFFFFFFu && first[p] == (uint32_t)o;
        c += f[j];
    }
    return c;
}
// Row of every first occurrence: consecutive inside a block of 1024 read-order indices (what the recurrence's warp loads
// need), blocks placed by one atomicAdd each (their order varies from run to run; the values computed do not).  The
// first-occurrence thread records the row for its key and writes the triple to the list the projection reads.
// mode 1 (slots of the entity gradient): the block bases come from a count pass instead (block_counts != null), so the slots are
// numbered strictly by first occurrence -- the sorted segment reduce of the entity gradient then splits its runs at the same
// places on every run, i.e. training is bit-reproducible.
__global__ void dedup_count_kernel(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ first, int64_t n_order, int32_t* __restrict__ block_counts) {
    bool f[4];
    int tot;
    const int64_t o0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    dedup_block_scan(dedup_flags(pos, first, n_order, o0, f), &tot);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = tot;
}
__global__ void dedup_assign_kernel(const int32_t* __restrict__ ids, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ first,
                                    int64_t n_order, int64_t N, int T, int32_t* count, uint32_t* __restrict__ slot_of,
                                    int32_t* __restrict__ ids_d, int mode, const int32_t* __restrict__ block_counts) {
    __shared__ int base_s;
    __shared__ int red_s[8];
    bool f[4];
    int tot;
    const int64_t o0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const int c = dedup_flags(pos, first, n_order, o0, f);
    const int ex = dedup_block_scan(c, &tot);
    if (block_counts) {
        int part = 0;
        for (int i = threadIdx.x; i < (int)blockIdx.x; i += 256) part += block_counts[i];
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __syncthreads();                                   // dedup_block_scan's shared scratch is done with
        if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            int b = 0;
            for (int w = 0; w < 8; ++w) b += red_s[w];
            base_s = b;
            if (blockIdx.x == gridDim.x - 1) *count = b + tot;
        }
    } else if (threadIdx.x == 0) base_s = tot ? atomicAdd(count, tot) : 0;
    __syncthreads();
    int row = base_s + ex;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (f[j]) {
            slot_of[pos[o0 + j]] = (uint32_t)row;
            if (mode == 2) {
                ids_d[row] = ids[o0 + j];
            } else {
                const int64_t i = dedup_step_of(o0 + j, N, T);
                if (mode == 1) ids_d[row] = ids[i * 3 + 1];       // 1-based entity row of the slot
                else { ids_d[(int64_t)row * 3] = ids[i * 3]; ids_d[(int64_t)row * 3 + 1] = ids[i * 3 + 1]; ids_d[(int64_t)row * 3 + 2] = ids[i * 3 + 2]; }
            }
            ++row;
        }
}

// Builds the slot map of one batch (see DedupBufs in common.cuh).  No host round trip: the slot count stays on the device.
int dedup_build(kprn_ctx* ctx, DedupBufs& b, int mode, const int32_t* ids, int64_t N, int T, int64_t n_keys) {
    const kprn_config& c = ctx->cfg;
    const int64_t ptiles = cdiv(N, TILE_M);
    const int64_t n_order = mode == 2 ? n_keys : ptiles * T * TILE_M;
    const int64_t n_max = mode == 2 ? n_keys : N * T;
    if (n_order >= (1ll << 30)) { set_error("dedup: batch too large"); return KPRN_EINVAL; }
    uint32_t M = 1024;
    while ((int64_t)M < 2 * n_max) M <<= 1;
    KPRN_TRY(ensure(b.table, (sizeof(unsigned long long) + sizeof(uint32_t)) * (size_t)M));     // keys, then first[]
    KPRN_TRY(ensure(b.slot_of, sizeof(uint32_t) * (size_t)M));
    KPRN_TRY(ensure(b.pos, sizeof(uint32_t) * (size_t)n_order));
    KPRN_TRY(ensure(b.ids, sizeof(int32_t) * (mode == 0 ? 3 : 1) * (size_t)n_max));
    const unsigned nblk = (unsigned)cdiv(n_order, DEDUP_BLOCK);
    KPRN_TRY(ensure(b.count, 256 + sizeof(int32_t) * (size_t)nblk));        // [0] = number of slots; [64 ..) = per-block counts (mode 1)
    unsigned long long* table = b.table.as<unsigned long long>();
    uint32_t* first = reinterpret_cast<uint32_t*>(table + M);
    int32_t* count = b.count.as<int32_t>();
    KPRN_CUDA(cudaMemsetAsync(table, 0xFF, (sizeof(unsigned long long) + sizeof(uint32_t)) * (size_t)M, ctx->stream));
    KPRN_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), ctx->stream));
    dedup_insert_kernel<<<(unsigned)cdiv(n_order, 256), 256, 0, ctx->stream>>>(ids, n_order, N, T, c.V_t, c.V_e, c.V_r, table, first, M - 1,
                                                                              b.pos.as<uint32_t>(), ctx->d_err, mode);
    KPRN_LAUNCH_CHECK(ctx);
    int32_t* block_counts = nullptr;
    if (mode == 1) {
        block_counts = count + 64;
        dedup_count_kernel<<<nblk, 256, 0, ctx->stream>>>(b.pos.as<uint32_t>(), first, n_order, block_counts);
        KPRN_LAUNCH_CHECK(ctx);
    }
    dedup_assign_kernel<<<nblk, 256, 0, ctx->stream>>>(ids, b.pos.as<uint32_t>(), first, n_order, N, T, count,
                                                      b.slot_of.as<uint32_t>(), b.ids.as<int32_t>(), mode, block_counts);
    KPRN_LAUNCH_CHECK(ctx);
    b.n_order = n_order;
    return KPRN_OK;
}
void dedup_free(DedupBufs& b) {
    for (DBuf* d : {&b.table, &b.slot_of, &b.pos, &b.ids, &b.count})
        if (d->p) { cudaFree(d->p); d->p = nullptr; d->cap = 0; }
}

// ---------------------------------------------------------------------------------------------
// host side

// 2-D tensor maps over the pre-split entity tables [V_e rows][KE cols] bf16: box = 64 columns x 1 row (tile::gather4
// fetches 4 such rows per instruction), 128-byte swizzle.  The driver entry point is resolved through the runtime.
static bool make_entity_tensor_maps(TcState* s, int64_t V_e) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { cudaGetLastError(); return false; }
    EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
    const cuuint64_t gdim[2] = {(cuuint64_t)s->KE, (cuuint64_t)V_e};
    const cuuint64_t gstride[1] = {(cuuint64_t)s->KE * 2};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(i ? &s->tm_lo : &s->tm_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, i ? (void*)s->E_lo : (void*)s->E_hi, gdim, gstride,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

static int tc_pack(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    if (s->packed_version == ctx->params_version && s->packed_small == ctx->small_version) return KPRN_OK;
    const bool full = s->packed_version != ctx->params_version;
    if (!s->E_hi || !s->E_lo || !s->WxE || !s->Wh || !s->Ptr || !s->WxE2) {
        // all or nothing: a failed allocation must not leave a half-built state that the next call takes for complete
        void** ptrs[6] = {(void**)&s->E_hi, (void**)&s->E_lo, (void**)&s->WxE, (void**)&s->Wh, (void**)&s->Ptr, (void**)&s->WxE2};
        for (void** pp : ptrs) if (*pp) { cudaFree(*pp); *pp = nullptr; }
        const size_t eb = sizeof(bf16) * (size_t)c.V_e * s->KE;
        const size_t sizes[6] = {eb, eb, sizeof(bf16) * (size_t)s->NCH * s->NS * 2 * B_SLAB_ELEMS,
                                 sizeof(bf16) * (size_t)s->NCH * s->KSh * 2 * B_SLAB_ELEMS, sizeof(float) * (size_t)c.V_t * c.V_r * s->G4,
                                 sizeof(bf16) * (size_t)s->NCH * (s->KSe1 > 0 ? s->KSe1 : 1) * 2 * B_SLAB_ELEMS};
        for (int i = 0; i < 6; ++i) {
            cudaError_t e = cudaMalloc(ptrs[i], sizes[i]);
            if (e != cudaSuccess) {
                for (void** pp : ptrs) if (*pp) { cudaFree(*pp); *pp = nullptr; }
                return cuda_fail(e, "cudaMalloc (packed operands)", __FILE__, __LINE__);
            }
        }
        s->have_tmaps = make_entity_tensor_maps(s, c.V_e);
    }
    if (full) {
        int64_t ne = c.V_e * (int64_t)s->KE;
        split_entity_kernel<<<(unsigned)cdiv(ne, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.E_e, c.V_e, c.d_e, s->KE, s->E_hi, s->E_lo);
        KPRN_LAUNCH_CHECK(ctx);
    }
    if (s->Hp > 256) {      // per-step recurrence (tc_wide.cu): 256-column blocks x 64-wide k slices; same byte count as the slab form
        KPRN_TRY(pack_wh_wide(ctx, ctx->d_params + l.Wh, c.H, s->Hp, s->Wh));
    } else {
        int64_t nh = (int64_t)s->NCH * s->KSh * B_SLAB_ELEMS;
        pack_wh_kernel<<<(unsigned)cdiv(nh, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.Wh, c.H, s->NCH, s->KSh, s->recur_pairs ? 1 : 0, s->Wh);
        KPRN_LAUNCH_CHECK(ctx);
    }
    ptr_table_kernel<<<(unsigned)s->G4, 128, 0, ctx->stream>>>(ctx->d_params, l, c.d_t, c.d_e, c.d_r, c.H, (int)c.V_t, (int)c.V_r, s->G4, s->Ptr);
    KPRN_LAUNCH_CHECK(ctx);
    int64_t nw = (int64_t)s->NCH * s->NS * B_SLAB_ELEMS;
    pack_wxe_kernel<<<(unsigned)cdiv(nw, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.Wx, s->Ptr, c.H, l.D, c.d_t, c.d_e, s->NCH,
                                                                      s->KSe, s->NS, (int)(c.V_t * c.V_r), s->G4, 0, s->WxE);
    KPRN_LAUNCH_CHECK(ctx);
    if (s->KSe1 > 0) {      // entity columns 256.. : a second projection launch accumulates them into XW
        int64_t nw1 = (int64_t)s->NCH * s->KSe1 * B_SLAB_ELEMS;
        pack_wxe_kernel<<<(unsigned)cdiv(nw1, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.Wx, s->Ptr, c.H, l.D, c.d_t, c.d_e, s->NCH,
                                                                           s->KSe1, s->KSe1, 0, s->G4, 256, s->WxE2);
        KPRN_LAUNCH_CHECK(ctx);
    }
    s->packed_version = ctx->params_version;
    s->packed_small = ctx->small_version;
    return KPRN_OK;
}

// the packed bf16 copies of the entity table, when they mirror the fp32 table (the sparse optimizer keeps them in step row by row)
void tc_packed_entity(kprn_ctx* ctx, bf16** hi, bf16** lo, int* KE) {
    *hi = nullptr; *lo = nullptr; *KE = 0;
    TcState* s = ctx->tc;
    if (!s || !s->E_hi || !s->E_lo || s->packed_version != ctx->params_version) return;
    *hi = s->E_hi; *lo = s->E_lo; *KE = s->KE;
}

static size_t proj_smem(int KE, int NOH) { return (size_t)(KE / 8 + 8 * NOH) * 2048 + (size_t)(KE / 8) * 2048 + PROJ_RING * B_SLAB_ELEMS * 2 + 512; }
static size_t recur_smem(int Hp, bool pair) {
    return (size_t)2 * TILE_M * Hp * 2 + (pair ? (size_t)RECUR_RING_PAIR * B_SLAB_ELEMS : (size_t)RECUR_RING * B_SLAB_ELEMS * 2) + RECUR_BAR_BYTES;
}
template <int HP, bool PAIR, bool TRAIN>
static cudaError_t launch_recur(const RecurArgs& a, int64_t grid, cudaStream_t stream) {
    {   // the opt-in is per device (a process may hold ctxs on several): set it on every launch, it costs nothing
        cudaError_t e = cudaFuncSetAttribute(recur_kernel<HP, PAIR, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)recur_smem(HP, PAIR));
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(RECUR_THREADS); cfg.dynamicSmemBytes = recur_smem(HP, PAIR); cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, recur_kernel<HP, PAIR, TRAIN>, a);
}

int tc_forward(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, bool train) {
    const kprn_config& c = ctx->cfg;
    // training: the fused kernels with activation caches (KPRN_PREC_TC_ALL, shapes of tc_train_supported), else the generic engine
    const bool tc_train = train && tc_train_supported(ctx);
    ctx->tc_train_fwd = false;
    if (train && !tc_train) return simple_forward(ctx, d_ids, d_seg, N, S, true);
    if (!ctx->tc) {
        ctx->tc = new TcState();
        TcState* s = ctx->tc;
        s->Hp = pad_hidden(c.H); s->KE = round_up(c.d_e, 64);
        s->NCH = s->Hp / 32; s->KSh = s->Hp / 64; s->G4 = 4 * s->Hp;
        s->KE0 = s->KE > 256 ? 256 : s->KE; s->KSe = s->KE0 / 64; s->KSe1 = (s->KE - s->KE0) / 64;      // entity columns in two K halves when d_e > 256
        s->NOH = (int)((c.V_t * c.V_r + 63) / 64); s->NS = s->KSe + s->NOH;
        cudaDeviceProp prop;
        KPRN_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
        s->num_sms = prop.multiProcessorCount;
        { const char* d = getenv("KPRN_TC_DEBUG"); s->debug = d && d[0] == '1'; }
        { const char* d = getenv("KPRN_TC_RECUR_PAIRS"); s->recur_pairs = !(d && d[0] == '0'); }
        if (prop.major != 10) {
            delete ctx->tc; ctx->tc = nullptr;      // a retry must meet this check again
            set_error("tensor-core path needs sm_100 (found sm_%d%d)", prop.major, prop.minor);
            return KPRN_ECUDA;
        }
    }
    TcState* s = ctx->tc;
    s->dedup = dedup_enabled(ctx);
    if (!s->attrs_set) {
        KPRN_CUDA(cudaFuncSetAttribute(proj3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)proj_smem(s->KE0, s->NOH)));
        s->attrs_set = true;
    }
    KPRN_TRY(tc_pack(ctx));
    const int64_t ptiles = cdiv(N, TILE_M);
    KPRN_TRY(ensure(s->xw, sizeof(float) * (size_t)ptiles * c.T * s->G4 * TILE_M));
    KPRN_TRY(ensure(s->hT, sizeof(float) * (size_t)N * c.H));
    KPRN_TRY(ensure(ctx->path_logits, sizeof(float) * N * c.L));
    const int passes = c.precision == KPRN_PREC_TC_FAST ? 1 : 3;   // TC_SPLIT and TC_ALL: bf16x3
    unsigned long long* dbg = nullptr;
    if (s->debug) {
        KPRN_TRY(ensure(s->dbg, 2 * 16 * sizeof(unsigned long long)));
        dbg = s->dbg.as<unsigned long long>();
        KPRN_CUDA(cudaMemsetAsync(dbg, 0, 2 * 16 * sizeof(unsigned long long), ctx->stream));
    }
    // leading-token reuse (scoring on the paired recurrence): everything below sees the batch sorted by leading-token count
    PrefixOut pfx{nullptr, nullptr, nullptr, nullptr, nullptr};
    const bool use_prefix = !train && s->Hp <= 256 && s->recur_pairs && prefix_enabled(ctx, N);
    if (use_prefix) { KPRN_TRY(prefix_build(ctx, d_ids, N, &pfx)); d_ids = pfx.ids; }
    {
        ProjArgs a;
        a.ids = d_ids; a.N = N; a.T = c.T; a.V_e = c.V_e; a.V_t = (int)c.V_t; a.V_r = (int)c.V_r;
        a.E_hi = s->E_hi; a.E_lo = s->E_lo; a.W = s->WxE; a.XW = s->xw.as<float>();
        a.KE = s->KE0; a.KSe = s->KSe; a.NOH = s->NOH; a.NS = s->NS; a.NCH = s->NCH; a.G4 = s->G4; a.num_tiles = ptiles * c.T; a.passes = passes; a.dbg = dbg;
        a.n_dev = nullptr; a.k_base = 0; a.accumulate = 0;
        cudaEvent_t e0;
        KPRN_TRY(prof_begin(ctx, &e0));
        const int64_t n_steps = N * c.T, n_order = ptiles * c.T * TILE_M;       // n_order: read-order indices incl. the ragged tail
        if (s->dedup && n_order < (1ll << 30)) {
            KPRN_TRY(dedup_build(ctx, s->dd, 0, d_ids, N, c.T, 0));
            // the projection sees a list of (at most n_steps) single-step "paths"; its real length is read on the device
            a.ids = s->dd.ids.as<int32_t>(); a.N = n_steps; a.T = 1; a.num_tiles = cdiv(n_steps, TILE_M); a.n_dev = s->dd.count.as<int32_t>();
        }
        if (!s->have_tmaps) { set_error("tensor-core path: cuTensorMapEncodeTiled unavailable (driver too old for sm_100a TMA)"); return KPRN_ECUDA; }
        {   // a single (or odd last) tile is computed by both CTAs of a pair: identical values, identical stores
            int64_t pairs = (a.num_tiles + 1) / 2;
            int64_t g2 = 2 * (pairs < s->num_sms / 2 ? pairs : s->num_sms / 2);
            proj3_kernel<<<(unsigned)g2, NUM_THREADS, proj_smem(s->KE0, s->NOH), ctx->stream>>>(s->tm_hi, s->tm_lo, a);
            KPRN_LAUNCH_CHECK(ctx);
            if (s->KSe1 > 0) {      // entity columns 256..: same kernel, no one-hot slice, accumulating epilogue
                ProjArgs b = a;
                b.KE = s->KE - s->KE0; b.KSe = s->KSe1; b.NOH = 0; b.NS = s->KSe1; b.W = s->WxE2; b.k_base = 256; b.accumulate = 1;
                proj3_kernel<<<(unsigned)g2, NUM_THREADS, proj_smem(s->KE0, s->NOH), ctx->stream>>>(s->tm_hi, s->tm_lo, b);
                KPRN_LAUNCH_CHECK(ctx);
            }
        }
        KPRN_TRY(prof_end(ctx, e0, 1));
    }
    {
        RecurArgs a;
        int64_t grid_r = ptiles < s->num_sms ? ptiles : s->num_sms;
        int sms_r = s->num_sms;
        { const char* d = getenv("KPRN_TC_RECUR_SMS"); if (d && atoi(d) >= 2 && atoi(d) <= s->num_sms) sms_r = atoi(d); }   // experiment knob
        if (grid_r > sms_r) grid_r = sms_r;
        if (s->recur_pairs) { const int64_t tp = (ptiles + 1) / 2; grid_r = 2 * (tp < sms_r / 2 ? tp : sms_r / 2); }
        KPRN_TRY(ensure(s->hscratch, (size_t)grid_r * 2 * TILE_M * s->Hp * 2));
        a.XW = s->xw.as<float>(); a.W = s->Wh; a.hT = s->hT.as<float>(); a.hscratch = s->hscratch.as<bf16>(); a.N = N; a.T = c.T; a.H = c.H; a.G4 = s->G4;
        a.num_tiles = ptiles; a.passes = passes; a.dbg = dbg ? dbg + 16 : nullptr;
        a.wo = nullptr; a.part4 = nullptr; a.npad = ptiles * TILE_M;
        const bool dd = s->dedup && ptiles * c.T * TILE_M < (1ll << 30);
        a.pos = dd ? s->dd.pos.as<uint32_t>() : nullptr; a.slot_of = dd ? s->dd.slot_of.as<uint32_t>() : nullptr;
        ctx->head_part4 = nullptr;
        a.cache = nullptr; a.hcache = nullptr;
        a.perm = pfx.perm; a.ps = pfx.ps; a.chain_u = pfx.chain_u; a.chain_c = pfx.chain_c;
        if (use_prefix) KPRN_TRY(prefix_join(ctx));
        if (tc_train) KPRN_TRY(tc_train_buffers(ctx, N, &a.cache, &a.hcache));
        if (ctx->cls_only && c.pool == KPRN_POOL_LSE) {      // the pool kernel finishes the head: sum of 4 partials + bias
            KPRN_TRY(ensure(s->part4, sizeof(float) * 4 * (size_t)a.npad));
            a.part4 = s->part4.as<float>();
            a.wo = ctx->d_params + ctx->lay.Wo + (int64_t)(c.class_id - 1) * c.H;
            ctx->head_part4 = a.part4; ctx->head_npad = a.npad;
        }

        cudaEvent_t e0;
        KPRN_TRY(prof_begin(ctx, &e0));
        if (s->Hp > 256) {        // 256 < H <= 512: one launch per step, state through HBM (tc_wide.cu)
            const size_t hbytes = (size_t)ptiles * 2 * TILE_M * s->Hp * 2;
            KPRN_TRY(ensure(s->wide_h0, hbytes));
            KPRN_TRY(ensure(s->wide_h1, hbytes));
            KPRN_TRY(ensure(s->wide_c, sizeof(float) * (size_t)ptiles * s->Hp * TILE_M));
            RecurWide w;
            w.XW = a.XW; w.W = s->Wh; w.h0 = s->wide_h0.as<bf16>(); w.h1 = s->wide_h1.as<bf16>(); w.c = s->wide_c.as<float>();
            w.N = N; w.T = c.T; w.H = c.H; w.Hp = s->Hp; w.num_sms = s->num_sms; w.pos = a.pos; w.slot_of = a.slot_of;
            w.wo = a.wo; w.part4 = a.part4; w.npad = a.npad; w.hT = a.part4 ? nullptr : a.hT;
            KPRN_TRY(wide_recurrence(ctx, w));
        } else if (tc_train) {           // CTA pairs, epilogue also writes the activation caches of the backward
            grid_r = 2 * ((ptiles + 1) / 2 < s->num_sms / 2 ? (ptiles + 1) / 2 : s->num_sms / 2);
            if (s->Hp == 64) KPRN_CUDA((launch_recur<64, true, true>(a, grid_r, ctx->stream)));
            else if (s->Hp == 128) KPRN_CUDA((launch_recur<128, true, true>(a, grid_r, ctx->stream)));
            else KPRN_CUDA((launch_recur<256, true, true>(a, grid_r, ctx->stream)));
        } else if (s->recur_pairs) {     // a single tile is computed by both CTAs of the one pair
            if (s->Hp == 64) KPRN_CUDA((launch_recur<64, true, false>(a, grid_r, ctx->stream)));
            else if (s->Hp == 128) KPRN_CUDA((launch_recur<128, true, false>(a, grid_r, ctx->stream)));
            else KPRN_CUDA((launch_recur<256, true, false>(a, grid_r, ctx->stream)));
        } else {
            if (s->Hp == 64) KPRN_CUDA((launch_recur<64, false, false>(a, grid_r, ctx->stream)));
            else if (s->Hp == 128) KPRN_CUDA((launch_recur<128, false, false>(a, grid_r, ctx->stream)));
            else KPRN_CUDA((launch_recur<256, false, false>(a, grid_r, ctx->stream)));
        }
        KPRN_LAUNCH_CHECK(ctx);
        KPRN_TRY(prof_end(ctx, e0, 2));
    }
    if (s->debug) {
        unsigned long long h[32];
        KPRN_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
        const double np_ = (double)(ptiles * c.T < s->num_sms ? ptiles * c.T : s->num_sms), nr_ = (double)(ptiles < s->num_sms ? ptiles : s->num_sms);
        fprintf(stderr, "[kprn tc debug] proj: kernel cycles/CTA %.0f | waits/CTA: producer.slab_empty %.0f, mma.acc_empty %.0f, mma.a_full %.0f, "
                        "mma.slab_full %.0f, gather.a_free %.0f, epilogue(wg0+wg1).acc_full %.0f\n",
                h[15] / np_, h[0] / np_, h[1] / np_, h[2] / np_, h[3] / np_, h[4] / np_ / 2, h[5] / np_);
        fprintf(stderr, "[kprn tc debug] recur: kernel cycles/CTA %.0f | waits/CTA: producer.slab_empty %.0f, mma.acc_empty %.0f, mma.h_ready %.0f, "
                        "mma.slab_full %.0f, hloader.mma_done %.0f, hloader.h_written %.0f, epilogue(wg0+wg1).acc_full %.0f\n",
                h[31] / nr_, h[24] / nr_, h[25] / nr_, h[26] / nr_, h[27] / nr_, h[28] / nr_, h[29] / nr_, h[30] / nr_);
    }
    // A5: head on the last hidden state (fp32 CUDA cores; 0.2 % of the flops)
    if (ctx->head_part4) {
        // nothing to launch: see pool_forward
    } else if (ctx->cls_only) {
        KPRN_TRY(head_forward_cls(ctx, s->hT.as<float>(), c.H, 0, N));
    } else {
        float* z1 = nullptr;
        if (c.head_layers == 2) { KPRN_TRY(ensure(ctx->z1, sizeof(float) * N * c.H)); z1 = ctx->z1.as<float>(); }
        KPRN_TRY(head_forward(ctx, s->hT.as<float>(), c.H, 0, N, z1));
    }
    (void)S; (void)d_seg;
    ctx->tc_train_fwd = tc_train;
    return KPRN_OK;
}

void tc_destroy(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    if (!s) return;
    tc_train_destroy(ctx);
    prefix_destroy(ctx);
    if (s->E_hi) cudaFree(s->E_hi);
    if (s->E_lo) cudaFree(s->E_lo);
    if (s->WxE) cudaFree(s->WxE);
    if (s->Wh) cudaFree(s->Wh);
    if (s->Ptr) cudaFree(s->Ptr);
    if (s->WxE2) cudaFree(s->WxE2);
    for (DBuf* b : {&s->xw, &s->hT, &s->hscratch, &s->dbg, &s->part4, &s->wide_h0, &s->wide_h1, &s->wide_c})
        if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    dedup_free(s->dd);
    delete s;
    ctx->tc = nullptr;
}

}  // namespace kprn
