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
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace kprn {

using bf16 = __nv_bfloat16;
#define KDESC(addr, lbo, sbo) (a.swap_ls ? ptx::make_desc_kmajor((addr), (sbo), (lbo)) : ptx::make_desc_kmajor((addr), (lbo), (sbo)))

constexpr int TILE_M = 128;           // paths per CTA tile (UMMA M)
constexpr int A_SLAB_ELEMS = 8 * 256 * 8;   // proj B slab: 64 k x 256 cols  (32 KB)
constexpr int B_SLAB_ELEMS = 8 * 128 * 8;   // recur B slab: 64 k x 128 cols (16 KB)
constexpr int PROJ_RING = 3;
constexpr int RECUR_RING = 6;
constexpr int NUM_THREADS = 384;      // warps: 0 TMA producer, 1 MMA issuer, 2-3 gather/alloc, 4-11 epilogue

struct TcState {
    uint64_t packed_version = 0;
    int Hp = 0, KE = 0, NCH = 0, KSh = 0, KSe = 0, G4 = 0, NT = 0;
    bf16* E_hi = nullptr; bf16* E_lo = nullptr;
    bf16* WxE = nullptr; bf16* Wh = nullptr;
    float* Ptr = nullptr;
    DBuf xw, hT;
    int num_sms = 148;
    bool attrs_set = false;
    int swap_ls = 0;
};

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

bool tc_supported(const kprn_ctx* ctx) {
    const kprn_config& c = ctx->cfg;
    if (c.cell != KPRN_CELL_LSTM) return false;
    int Hp = round_up(c.H, 64), KE = round_up(c.d_e, 64);
    if (Hp > 256 || KE > 256) return false;
    if (c.V_t * c.V_r > 4096) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// parameter packing (runs when the parameters change)

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

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

// permuted gate column g' = chunk*128 + gate*32 + uu  <->  reference row gate*H + chunk*32 + uu
__device__ __forceinline__ int perm_to_row(int gp, int H, int& u) {
    int j = gp >> 7, q = (gp >> 5) & 3, uu = gp & 31;
    u = j * 32 + uu;
    return q * H + u;
}

// proj B operand: Wx[:, d_t : d_t+d_e] -> slabs [ntile][kslice][hi|lo][kc 8][n 256][8]
__global__ void pack_wxe_kernel(const float* __restrict__ Wx, int H, int D, int d_t, int d_e, int NT, int KSe,
                                bf16* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)NT * KSe * A_SLAB_ELEMS;
    if (idx >= total) return;
    int e = (int)(idx & 7);
    int n = (int)((idx >> 3) & 255);
    int kc = (int)((idx >> 11) & 7);
    int64_t slab = idx >> 14;
    int s = (int)(slab % KSe), nt = (int)(slab / KSe);
    int gp = nt * 256 + n, k = s * 64 + kc * 8 + e, u;
    int row = perm_to_row(gp, H, u);
    float v = (u < H && k < d_e) ? Wx[(int64_t)row * D + d_t + k] : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    int64_t base = slab * 2 * A_SLAB_ELEMS + ((kc * 256 + n) << 3) + e;
    out[base] = h;
    out[base + A_SLAB_ELEMS] = l;
}

// recur B operand: Wh -> slabs [chunk][kslice][hi|lo][kc 8][n 128][8]
__global__ void pack_wh_kernel(const float* __restrict__ Wh, int H, int NCH, int KSh, bf16* __restrict__ out) {
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
    float v = (u < H && k < H) ? Wh[(int64_t)row * H + k] : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    int64_t base = slab * 2 * B_SLAB_ELEMS + ((kc * 128 + n) << 3) + e;
    out[base] = h;
    out[base + B_SLAB_ELEMS] = l;
}

// hoisted small-vocabulary projection: P_tr[a*V_r+b][g'] = bx + E_t[a] Wx_t^T + E_r[b] Wx_r^T  (fp32 FFMA)
__global__ void ptr_table_kernel(const float* __restrict__ P, Layout l, int d_t, int d_e, int d_r, int H, int V_r,
                                 int G4, float* __restrict__ out) {
    int combo = blockIdx.x;
    int a = combo / V_r, b = combo % V_r;
    const float* et = P + l.E_t + (int64_t)a * d_t;
    const float* er = P + l.E_r + (int64_t)b * d_r;
    for (int gp = threadIdx.x; gp < G4; gp += blockDim.x) {
        int u;
        int row = perm_to_row(gp, H, u);
        float acc = 0.f;
        if (u < H) {
            const float* w = P + l.Wx + (int64_t)row * l.D;
            acc = P[l.bx + row];
            for (int k = 0; k < d_t; ++k) acc = fmaf(et[k], w[k], acc);
            for (int k = 0; k < d_r; ++k) acc = fmaf(er[k], w[d_t + d_e + k], acc);
        }
        out[(int64_t)combo * G4 + gp] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// kernel "proj": XW tile (128 rows x 4Hp) = gather(E_e) Wx_e^T + P_tr, tcgen05 with bf16x3 split.
// XW layout: [tile m = pathtile*T + t][column block (4Hp/8)][row 128][8 floats], columns in permuted order.

struct ProjArgs {
    const int32_t* ids; int64_t N; int T; int64_t V_e; int V_t; int V_r;
    const bf16* E_hi; const bf16* E_lo; const bf16* W; const float* Ptr; float* XW;
    int KE, KSe, NT, G4; int64_t num_tiles; int passes;   // passes: 3 = bf16x3, 1 = hi*hi only
    int swap_ls;   // debug: swap the LBO/SBO descriptor fields (KPRN_TC_SWAP_LS=1)
};

__global__ void __launch_bounds__(NUM_THREADS, 1) proj_kernel(ProjArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = (uint32_t)TILE_M * a.KE * 2;          // one of hi / lo
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + a_bytes;
    uint8_t* sRing = smem + 2 * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + PROJ_RING * A_SLAB_ELEMS * 2);
    // bars: [0,R) slab_full, [R,2R) slab_empty, 2R a_full, 2R+1 a_free, 2R+2.. acc_full[2], acc_empty[2]
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    constexpr int R = PROJ_RING;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 6);

    if (threadIdx.x == 0) {
        for (int i = 0; i < R; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(R + i), 1); }
        ptx::mbar_init(BAR(2 * R), 64);       // a_full: 64 gather threads
        ptx::mbar_init(BAR(2 * R + 1), 1);    // a_free: tcgen05.commit
        ptx::mbar_init(BAR(2 * R + 2), 1); ptx::mbar_init(BAR(2 * R + 3), 1);        // acc_full
        ptx::mbar_init(BAR(2 * R + 4), 128); ptx::mbar_init(BAR(2 * R + 5), 128);    // acc_empty
        ptx::fence_mbar_init();
    }
    if (warp == 2) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t my_tiles = (a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles blockIdx.x + i*gridDim.x
    const uint32_t idesc = ptx::make_idesc_f16(1, TILE_M, 256);
    const int slabs_per_k = a.passes == 3 ? 2 : 1;

    if (warp == 0) {
        // ===== TMA producer: weight slabs, the same sequence for every tile =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int nt = 0; nt < a.NT; ++nt)
                    for (int s = 0; s < a.KSe; ++s)
                        for (int part = 0; part < slabs_per_k; ++part) {
                            ptx::mbar_wait(BAR(R + stage), phase ^ 1);
                            ptx::mbar_arrive_expect_tx(BAR(stage), A_SLAB_ELEMS * 2);
                            const bf16* src = a.W + ((int64_t)(nt * a.KSe + s) * 2 + part) * A_SLAB_ELEMS;
                            ptx::tma_bulk_g2s(ptx::smem_u32(sRing + stage * A_SLAB_ELEMS * 2), src, A_SLAB_ELEMS * 2, BAR(stage));
                            if (++stage == R) { stage = 0; phase ^= 1; }
                        }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int64_t au = 0;
            const uint32_t aHi = ptx::smem_u32(sA_hi), aLo = ptx::smem_u32(sA_lo);
            for (int64_t it = 0; it < my_tiles; ++it) {
                ptx::mbar_wait(BAR(2 * R), it & 1);
                ptx::tc_fence_after();
                for (int nt = 0; nt < a.NT; ++nt, ++au) {
                    const int buf = au & 1;
                    ptx::mbar_wait(BAR(2 * R + 4 + buf), ((au >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d = tmem + buf * 256;
                    for (int s = 0; s < a.KSe; ++s) {
                        ptx::mbar_wait(BAR(stage), phase);
                        ptx::tc_fence_after();
                        uint32_t bAddr = ptx::smem_u32(sRing + stage * A_SLAB_ELEMS * 2);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint64_t db = KDESC(bAddr + k * 2 * 4096, 4096, 128);
                            uint64_t da = KDESC(aHi + (s * 8 + k * 2) * 2048, 2048, 128);
                            ptx::umma_f16(d, da, db, idesc, (s | k) != 0);
                        }
                        if (a.passes == 3) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                uint64_t db = KDESC(bAddr + k * 2 * 4096, 4096, 128);
                                uint64_t da = KDESC(aLo + (s * 8 + k * 2) * 2048, 2048, 128);
                                ptx::umma_f16(d, da, db, idesc, 1);
                            }
                        }
                        ptx::umma_commit(BAR(R + stage));
                        if (++stage == R) { stage = 0; phase ^= 1; }
                        if (a.passes == 3) {
                            ptx::mbar_wait(BAR(stage), phase);
                            ptx::tc_fence_after();
                            bAddr = ptx::smem_u32(sRing + stage * A_SLAB_ELEMS * 2);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                uint64_t db = KDESC(bAddr + k * 2 * 4096, 4096, 128);
                                uint64_t da = KDESC(aHi + (s * 8 + k * 2) * 2048, 2048, 128);
                                ptx::umma_f16(d, da, db, idesc, 1);
                            }
                            ptx::umma_commit(BAR(R + stage));
                            if (++stage == R) { stage = 0; phase ^= 1; }
                        }
                    }
                    ptx::umma_commit(BAR(2 * R + 2 + buf));
                }
                ptx::umma_commit(BAR(2 * R + 1));   // A tile may be overwritten
            }
        }
    } else if (warp < 4) {
        // ===== gather warps: pre-split entity rows -> canonical A layout via 16-byte LDGSTS =====
        const int gw = warp - 2;
        const int chunks = a.KE / 8;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t m = blockIdx.x + it * gridDim.x;
            const int64_t pt = m / a.T;
            const int t = (int)(m % a.T);
            ptx::mbar_wait(BAR(2 * R + 1), (it & 1) ^ 1);
            for (int blk = gw; blk < TILE_M / 16; blk += 2) {
                const int r = blk * 16 + (lane >> 1);
                int64_t n = pt * TILE_M + r;
                if (n >= a.N) n = a.N - 1;
                int64_t id = (int64_t)a.ids[(n * a.T + t) * 3 + 1] - 1;
                id = id < 0 ? 0 : (id >= a.V_e ? a.V_e - 1 : id);
                const bf16* sh = a.E_hi + id * a.KE;
                const bf16* sl = a.E_lo + id * a.KE;
                const uint32_t dh = ptx::smem_u32(sA_hi) + r * 16, dl = ptx::smem_u32(sA_lo) + r * 16;
                for (int kc = (lane & 1); kc < chunks; kc += 2) {
                    ptx::cp_async16(dh + kc * 2048, sh + kc * 8);
                    ptx::cp_async16(dl + kc * 2048, sl + kc * 8);
                }
            }
            ptx::cp_async_wait_all();
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(BAR(2 * R));
        }
    } else {
        // ===== epilogue: TMEM -> registers, + P_tr row, -> XW =====
        const int wg = (warp - 4) >> 2;               // accumulator buffer / parity of the use index
        const int q4 = warp & 3;                      // TMEM lane quarter of this warp
        const int r = q4 * 32 + lane;
        const int64_t total_uses = my_tiles * a.NT;
        for (int64_t au = wg; au < total_uses; au += 2) {
            const int64_t it = au / a.NT;
            const int nt = (int)(au % a.NT);
            const int64_t m = blockIdx.x + it * gridDim.x;
            const int64_t pt = m / a.T;
            const int t = (int)(m % a.T);
            int64_t n = pt * TILE_M + r;
            if (n >= a.N) n = a.N - 1;
            const int32_t* idp = a.ids + (n * a.T + t) * 3;
            int ta = idp[0] - 1, rb = idp[2] - 1;
            ta = ta < 0 ? 0 : (ta >= a.V_t ? a.V_t - 1 : ta);
            rb = rb < 0 ? 0 : (rb >= a.V_r ? a.V_r - 1 : rb);
            const float* prow = a.Ptr + (int64_t)(ta * a.V_r + rb) * a.G4 + nt * 256;
            float* xw = a.XW + ((m * (a.G4 / 8) + nt * 32) * TILE_M + r) * 8;
            ptx::mbar_wait(BAR(2 * R + 2 + wg), (au >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + wg * 256;
#pragma unroll 4
            for (int cb = 0; cb < 32; ++cb) {
                float v[8];
                ptx::tmem_ld8(tbase + cb * 8, v);
                const float4 p0 = __ldg(reinterpret_cast<const float4*>(prow + cb * 8));
                const float4 p1 = __ldg(reinterpret_cast<const float4*>(prow + cb * 8 + 4));
                ptx::tc_wait_ld();
                float4 o0 = make_float4(v[0] + p0.x, v[1] + p0.y, v[2] + p0.z, v[3] + p0.w);
                float4 o1 = make_float4(v[4] + p1.x, v[5] + p1.y, v[6] + p1.z, v[7] + p1.w);
                float4* dst = reinterpret_cast<float4*>(xw + (int64_t)cb * TILE_M * 8);
                dst[0] = o0; dst[1] = o1;
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(BAR(2 * R + 4 + wg));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// kernel "recur": per 128-path tile, T LSTM steps.  TMEM: columns [0,256) two 128-column accumulators
// (a chunk = 32 hidden units x 4 gates), columns [256,256+HP) the cell state c (fp32, lane = path).

struct RecurArgs {
    const float* XW; const bf16* W; float* hT; int64_t N; int T; int H; int G4; int64_t num_tiles; int passes; int swap_ls;
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

__device__ __forceinline__ uint4 pack_hi8(const float* v, float* rem) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bf16 a = __float2bfloat16_rn(v[2 * i]), b = __float2bfloat16_rn(v[2 * i + 1]);
        rem[2 * i] = v[2 * i] - __bfloat162float(a);
        rem[2 * i + 1] = v[2 * i + 1] - __bfloat162float(b);
        w[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 pack_lo8(const float* rem) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bf16 a = __float2bfloat16_rn(rem[2 * i]), b = __float2bfloat16_rn(rem[2 * i + 1]);
        w[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int HP>
__global__ void __launch_bounds__(NUM_THREADS, 1) recur_kernel(RecurArgs a) {
    constexpr int NCH = HP / 32;        // 128-column chunks per step
    constexpr int KS = HP / 64;         // 64-wide k slices of h
    constexpr int R = RECUR_RING;
    constexpr uint32_t H_BYTES = TILE_M * HP * 2;   // one of hi / lo
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sH_hi = smem;
    uint8_t* sH_lo = smem + H_BYTES;
    uint8_t* sRing = smem + 2 * H_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + R * B_SLAB_ELEMS * 2);
    // bars: [0,R) slab_full, [R,2R) slab_empty, 2R..2R+1 acc_full, 2R+2..2R+3 acc_empty, 2R+4 mma_step_done,
    //       2R+5.. h_ready[KS]
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 5 + KS);

    if (threadIdx.x == 0) {
        for (int i = 0; i < R; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(R + i), 1); }
        ptx::mbar_init(BAR(2 * R), 1); ptx::mbar_init(BAR(2 * R + 1), 1);
        ptx::mbar_init(BAR(2 * R + 2), 128); ptx::mbar_init(BAR(2 * R + 3), 128);
        ptx::mbar_init(BAR(2 * R + 4), 1);
        for (int s = 0; s < KS; ++s) ptx::mbar_init(BAR(2 * R + 5 + s), 256);
        ptx::fence_mbar_init();
    }
    if (warp == 2) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t my_tiles = (a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int T = a.T;
    const uint32_t idesc = ptx::make_idesc_f16(1, TILE_M, 128);
    const int slabs_per_k = a.passes == 3 ? 2 : 1;

    if (warp == 0) {
        // ===== TMA producer: Wh slabs; the sequence is identical for every MMA step (t >= 1) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const int64_t steps = my_tiles * (T - 1);
            for (int64_t st = 0; st < steps; ++st)
                for (int j = 0; j < NCH; ++j)
                    for (int s = 0; s < KS; ++s)
                        for (int part = 0; part < slabs_per_k; ++part) {
                            ptx::mbar_wait(BAR(R + stage), phase ^ 1);
                            ptx::mbar_arrive_expect_tx(BAR(stage), B_SLAB_ELEMS * 2);
                            const bf16* src = a.W + ((int64_t)(j * KS + s) * 2 + part) * B_SLAB_ELEMS;
                            ptx::tma_bulk_g2s(ptx::smem_u32(sRing + stage * B_SLAB_ELEMS * 2), src, B_SLAB_ELEMS * 2, BAR(stage));
                            if (++stage == R) { stage = 0; phase ^= 1; }
                        }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int64_t ug0 = 0, ug1 = 0;
            const uint32_t hHi = ptx::smem_u32(sH_hi), hLo = ptx::smem_u32(sH_lo);
            for (int64_t it = 0; it < my_tiles; ++it) {
                for (int t = 1; t < T; ++t) {
                    const int64_t cnt = it * (T - 1) + t;          // 1-based index of this MMA step
                    for (int j = 0; j < NCH; ++j) {
                        const int g = j & 1;
                        int64_t& ug = g ? ug1 : ug0;
                        ptx::mbar_wait(BAR(2 * R + 2 + g), (ug & 1) ^ 1);
                        ptx::tc_fence_after();
                        const uint32_t d = tmem + g * 128;
                        for (int s = 0; s < KS; ++s) {
                            if (j == 0) {          // h_{t-1} slice s written by the epilogue warps
                                ptx::mbar_wait(BAR(2 * R + 5 + s), (cnt - 1) & 1);
                                ptx::tc_fence_after();
                            }
                            ptx::mbar_wait(BAR(stage), phase);
                            ptx::tc_fence_after();
                            uint32_t bAddr = ptx::smem_u32(sRing + stage * B_SLAB_ELEMS * 2);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                uint64_t db = KDESC(bAddr + k * 2 * 2048, 2048, 128);
                                uint64_t da = KDESC(hHi + (s * 8 + k * 2) * 2048, 2048, 128);
                                ptx::umma_f16(d, da, db, idesc, (s | k) != 0);
                            }
                            if (a.passes == 3) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    uint64_t db = KDESC(bAddr + k * 2 * 2048, 2048, 128);
                                    uint64_t da = KDESC(hLo + (s * 8 + k * 2) * 2048, 2048, 128);
                                    ptx::umma_f16(d, da, db, idesc, 1);
                                }
                            }
                            ptx::umma_commit(BAR(R + stage));
                            if (++stage == R) { stage = 0; phase ^= 1; }
                            if (a.passes == 3) {
                                ptx::mbar_wait(BAR(stage), phase);
                                ptx::tc_fence_after();
                                bAddr = ptx::smem_u32(sRing + stage * B_SLAB_ELEMS * 2);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    uint64_t db = KDESC(bAddr + k * 2 * 2048, 2048, 128);
                                    uint64_t da = KDESC(hHi + (s * 8 + k * 2) * 2048, 2048, 128);
                                    ptx::umma_f16(d, da, db, idesc, 1);
                                }
                                ptx::umma_commit(BAR(R + stage));
                                if (++stage == R) { stage = 0; phase ^= 1; }
                            }
                        }
                        ptx::umma_commit(BAR(2 * R + g));
                        ++ug;
                    }
                    ptx::umma_commit(BAR(2 * R + 4));     // every MMA reading h_{t-1} has completed
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue warpgroups: wg 0 owns even chunks / accumulator 0, wg 1 odd chunks / accumulator 1 =====
        const int wg = (warp - 4) >> 2;
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const uint32_t hHi = ptx::smem_u32(sH_hi), hLo = ptx::smem_u32(sH_lo);
        float hnew[HP / 2];
        int64_t ug = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t pt = blockIdx.x + it * gridDim.x;
            const int64_t n = pt * TILE_M + r;
            for (int t = 0; t < T; ++t) {
                const float* xw_t = a.XW + (((pt * T + t) * (a.G4 / 8)) * TILE_M + r) * 8;
#pragma unroll
                for (int lc = 0; lc < NCH / 2; ++lc) {
                    const int j = 2 * lc + wg;
                    if (t > 0) {
                        ptx::mbar_wait(BAR(2 * R + wg), ug & 1);
                        ptx::tc_fence_after();
                    }
                    const uint32_t acc = tmem + lane_base + wg * 128;
                    const uint32_t cst = tmem + lane_base + 256 + j * 32;
#pragma unroll
                    for (int sb = 0; sb < 4; ++sb) {
                        float gi[8], gg[8], gf[8], go[8], cp[8];
                        const float* xp = xw_t + (int64_t)(j * 16 + sb) * TILE_M * 8;
                        float4 x0, x1;
#define KPRN_LOAD_GATE(dst, q)                                                                        \
    x0 = *reinterpret_cast<const float4*>(xp + (int64_t)(q) * 4 * TILE_M * 8);                        \
    x1 = *reinterpret_cast<const float4*>(xp + (int64_t)(q) * 4 * TILE_M * 8 + 4);                    \
    dst[0] = x0.x; dst[1] = x0.y; dst[2] = x0.z; dst[3] = x0.w; dst[4] = x1.x; dst[5] = x1.y; dst[6] = x1.z; dst[7] = x1.w;
                        KPRN_LOAD_GATE(gi, 0)
                        KPRN_LOAD_GATE(gg, 1)
                        KPRN_LOAD_GATE(gf, 2)
                        KPRN_LOAD_GATE(go, 3)
#undef KPRN_LOAD_GATE
                        if (t > 0) {
                            float ai[8], ag[8], af[8], ao[8];
                            ptx::tmem_ld8(acc + 0 * 32 + sb * 8, ai);
                            ptx::tmem_ld8(acc + 1 * 32 + sb * 8, ag);
                            ptx::tmem_ld8(acc + 2 * 32 + sb * 8, af);
                            ptx::tmem_ld8(acc + 3 * 32 + sb * 8, ao);
                            ptx::tmem_ld8(cst + sb * 8, cp);
                            ptx::tc_wait_ld();
#pragma unroll
                            for (int e = 0; e < 8; ++e) { gi[e] += ai[e]; gg[e] += ag[e]; gf[e] += af[e]; go[e] += ao[e]; }
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) cp[e] = 0.f;
                        }
                        float cn[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float i_ = fast_sigmoid(gi[e]), g_ = fast_tanh(gg[e]), f_ = fast_sigmoid(gf[e]), o_ = fast_sigmoid(go[e]);
                            cn[e] = fmaf(i_, g_, f_ * cp[e]);
                            hnew[lc * 32 + sb * 8 + e] = o_ * fast_tanh(cn[e]);
                        }
                        ptx::tmem_st8(cst + sb * 8, cn);
                    }
                    ptx::tc_wait_st();
                    if (t > 0) {
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(BAR(2 * R + 2 + wg));
                        ++ug;
                    }
                }
                if (t < T - 1) {
                    // all MMAs that read the h buffer must be complete before it is overwritten
                    const int64_t need = it * (T - 1) + t;
                    if (need >= 1) ptx::mbar_wait(BAR(2 * R + 4), (need - 1) & 1);
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float rem[8];
                            const int kc = 8 * s + 4 * wg + i;
                            uint4 hi = pack_hi8(&hnew[s * 32 + i * 8], rem);
                            uint4 lo = pack_lo8(rem);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hHi + kc * 2048 + r * 16), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hLo + kc * 2048 + r * 16), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
                        }
                        ptx::fence_proxy_async_smem();
                        ptx::mbar_arrive(BAR(2 * R + 5 + s));
                    }
                } else if (n < a.N) {
                    // SelectTable(-1): last hidden state, standard [N,H] layout for the head
                    float* dst = a.hT + n * a.H;
#pragma unroll
                    for (int lc = 0; lc < NCH / 2; ++lc) {
                        const int u0 = (2 * lc + wg) * 32;
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (u0 + e < a.H) dst[u0 + e] = hnew[lc * 32 + e];
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side

static int tc_pack(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    if (s->packed_version == ctx->params_version) return KPRN_OK;
    if (!s->E_hi) {
        size_t eb = sizeof(bf16) * (size_t)c.V_e * s->KE;
        KPRN_CUDA(cudaMalloc(&s->E_hi, eb));
        KPRN_CUDA(cudaMalloc(&s->E_lo, eb));
        KPRN_CUDA(cudaMalloc(&s->WxE, sizeof(bf16) * (size_t)s->NT * s->KSe * 2 * A_SLAB_ELEMS));
        KPRN_CUDA(cudaMalloc(&s->Wh, sizeof(bf16) * (size_t)s->NCH * s->KSh * 2 * B_SLAB_ELEMS));
        KPRN_CUDA(cudaMalloc(&s->Ptr, sizeof(float) * (size_t)c.V_t * c.V_r * s->G4));
    }
    int64_t ne = c.V_e * (int64_t)s->KE;
    split_entity_kernel<<<(unsigned)cdiv(ne, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.E_e, c.V_e, c.d_e, s->KE, s->E_hi, s->E_lo);
    KPRN_LAUNCH_CHECK(ctx);
    int64_t nw = (int64_t)s->NT * s->KSe * A_SLAB_ELEMS;
    pack_wxe_kernel<<<(unsigned)cdiv(nw, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.Wx, c.H, l.D, c.d_t, c.d_e, s->NT, s->KSe, s->WxE);
    KPRN_LAUNCH_CHECK(ctx);
    int64_t nh = (int64_t)s->NCH * s->KSh * B_SLAB_ELEMS;
    pack_wh_kernel<<<(unsigned)cdiv(nh, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.Wh, c.H, s->NCH, s->KSh, s->Wh);
    KPRN_LAUNCH_CHECK(ctx);
    ptr_table_kernel<<<(unsigned)(c.V_t * c.V_r), 256, 0, ctx->stream>>>(ctx->d_params, l, c.d_t, c.d_e, c.d_r, c.H, (int)c.V_r, s->G4, s->Ptr);
    KPRN_LAUNCH_CHECK(ctx);
    s->packed_version = ctx->params_version;
    return KPRN_OK;
}

static size_t proj_smem(int KE) { return (size_t)2 * TILE_M * KE * 2 + PROJ_RING * A_SLAB_ELEMS * 2 + 256; }
static size_t recur_smem(int Hp) { return (size_t)2 * TILE_M * Hp * 2 + RECUR_RING * B_SLAB_ELEMS * 2 + 256; }

int tc_forward(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, bool train) {
    const kprn_config& c = ctx->cfg;
    if (train) return simple_forward(ctx, d_ids, d_seg, N, S, true);   // training caches come from the fp32 kernels
    if (!ctx->tc) {
        ctx->tc = new TcState();
        TcState* s = ctx->tc;
        s->Hp = round_up(c.H, 64); s->KE = round_up(c.d_e, 64);
        s->NCH = s->Hp / 32; s->KSh = s->Hp / 64; s->KSe = s->KE / 64; s->G4 = 4 * s->Hp; s->NT = s->G4 / 256;
        cudaDeviceProp prop;
        KPRN_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
        s->num_sms = prop.multiProcessorCount;
        const char* sw = getenv("KPRN_TC_SWAP_LS");
        s->swap_ls = (sw && sw[0] == '1') ? 1 : 0;
        if (prop.major != 10) { set_error("tensor-core path needs sm_100 (found sm_%d%d)", prop.major, prop.minor); return KPRN_ECUDA; }
    }
    TcState* s = ctx->tc;
    if (!s->attrs_set) {
        KPRN_CUDA(cudaFuncSetAttribute(proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)proj_smem(s->KE)));
        KPRN_CUDA(cudaFuncSetAttribute(recur_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)recur_smem(64)));
        KPRN_CUDA(cudaFuncSetAttribute(recur_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)recur_smem(128)));
        KPRN_CUDA(cudaFuncSetAttribute(recur_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)recur_smem(256)));
        s->attrs_set = true;
    }
    KPRN_TRY(tc_pack(ctx));
    const int64_t ptiles = cdiv(N, TILE_M);
    KPRN_TRY(ensure(s->xw, sizeof(float) * (size_t)ptiles * c.T * s->G4 * TILE_M));
    KPRN_TRY(ensure(s->hT, sizeof(float) * (size_t)N * c.H));
    KPRN_TRY(ensure(ctx->path_logits, sizeof(float) * N * c.L));
    const int passes = c.precision == KPRN_PREC_TC_FAST ? 1 : 3;
    {
        ProjArgs a;
        a.ids = d_ids; a.N = N; a.T = c.T; a.V_e = c.V_e; a.V_t = (int)c.V_t; a.V_r = (int)c.V_r;
        a.E_hi = s->E_hi; a.E_lo = s->E_lo; a.W = s->WxE; a.Ptr = s->Ptr; a.XW = s->xw.as<float>();
        a.KE = s->KE; a.KSe = s->KSe; a.NT = s->NT; a.G4 = s->G4; a.num_tiles = ptiles * c.T; a.passes = passes; a.swap_ls = s->swap_ls;
        int64_t grid = a.num_tiles < s->num_sms ? a.num_tiles : s->num_sms;
        proj_kernel<<<(unsigned)grid, NUM_THREADS, proj_smem(s->KE), ctx->stream>>>(a);
        KPRN_LAUNCH_CHECK(ctx);
    }
    {
        RecurArgs a;
        a.XW = s->xw.as<float>(); a.W = s->Wh; a.hT = s->hT.as<float>(); a.N = N; a.T = c.T; a.H = c.H; a.G4 = s->G4;
        a.num_tiles = ptiles; a.passes = passes; a.swap_ls = s->swap_ls;
        int64_t grid = ptiles < s->num_sms ? ptiles : s->num_sms;
        if (s->Hp == 64) recur_kernel<64><<<(unsigned)grid, NUM_THREADS, recur_smem(64), ctx->stream>>>(a);
        else if (s->Hp == 128) recur_kernel<128><<<(unsigned)grid, NUM_THREADS, recur_smem(128), ctx->stream>>>(a);
        else recur_kernel<256><<<(unsigned)grid, NUM_THREADS, recur_smem(256), ctx->stream>>>(a);
        KPRN_LAUNCH_CHECK(ctx);
    }
    // A5: head on the last hidden state (fp32 CUDA cores; 0.2 % of the flops)
    float* z1 = nullptr;
    if (c.head_layers == 2) { KPRN_TRY(ensure(ctx->z1, sizeof(float) * N * c.H)); z1 = ctx->z1.as<float>(); }
    KPRN_TRY(head_forward(ctx, s->hT.as<float>(), c.H, 0, N, z1));
    (void)S; (void)d_seg;
    return KPRN_OK;
}

void tc_destroy(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    if (!s) return;
    if (s->E_hi) cudaFree(s->E_hi);
    if (s->E_lo) cudaFree(s->E_lo);
    if (s->WxE) cudaFree(s->WxE);
    if (s->Wh) cudaFree(s->Wh);
    if (s->Ptr) cudaFree(s->Ptr);
    if (s->xw.p) cudaFree(s->xw.p);
    if (s->hT.p) cudaFree(s->hT.p);
    delete s;
    ctx->tc = nullptr;
}

}  // namespace kprn
