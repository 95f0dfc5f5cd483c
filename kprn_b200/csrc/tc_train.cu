// tc_train.cu -- the training products of the FastLSTM path on tcgen05 (A10 of SURVEY.md 8a; math: SURVEY App. A.7;
// reference call site SP/model/optimizer/MyOptimizer.lua:184-217, modules SP/model/module/MapReduce.lua:74-85 + FastLSTM BPTT).
//
//   forward (tc_path.cu, recur_kernel<.., TRAIN>) leaves   cache  = i, g, f, o, c of every (path, step, unit)   (fp32)
//                                                          hcache = h_t as bf16 hi/lo in the UMMA operand layout
//   bptt_kernel         the whole time loop t = T-1 .. 0 in one persistent launch: cell warps turn (cache_t, c_{t-1}, dh_t, dc) into
//                       dG_t, written as bf16 hi/lo straight into the shared-memory operand stage (and to dgop for wgrad_kernel);
//                       [dh_{t-1} | dX_t] = dG_t [Wh | Wx_e]   (M = paths, N = Hp + KE, K = 4Hp; cta_group::2, bf16x3);
//                       dh -> tile-private fp32 buffer for the next step; dX rows -> staging, one row per path-step
//                       (tc_bptt_steps.cuh holds the same loop as a pair of launches per step: the cross-check, KPRN_TC_FUSED=0)
//   entity gradient     radix sort of (slot, path-step) + eg_reduce_kernel + eg_fixup_kernel: the staged dX rows summed per distinct
//                       entity row in a fixed order (no atomics), into the compact per-slot buffer of a sparse training step or
//                       the dense table
//   wgrad_kernel        gWh += dG^T h_{t-1}, gWx_e += dG^T x_e, gP += dG^T onehot(type,relation)   (M = 4Hp gate columns,
//                       K = all path-steps: the SAME bytes of dgop read as an MN-major operand; split-K over CTA pairs)
//   wgrad_reduce_kernel fixed-order sum of the K-split partials into the flat gradient; ptr_sums / ptr_embed / ptr_wx turn gP (the
//                       gradient of the hoisted (type,relation) table) into gE_type, gE_relation, gWx[:, type|relation], gbx.
// The head / pool / loss gradients stay on the fp32 kernels of kernels_simple.cu (0.2 % of the flops).  Nothing on this path uses
// floating-point atomics: a training step is bit-reproducible.
#include <cub/device/device_radix_sort.cuh>

#include <stdio.h>
#include <stdlib.h>

#include "tc_state.cuh"

namespace kprn {

constexpr int BG_THREADS = 384;       // bgemm: warp 0 TMA producer, 1 MMA issuer / peer forwarder, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int BG_STAGES = 2;
constexpr int WG_THREADS = 256;       // wgrad: warp 0 TMA producer, 1 issuer / forwarder, 2 gather + one-hot writer, 3 forwarder, 4-7 epilogue
constexpr int WG_STAGES = 3;
constexpr int WG_ROWS = 64;           // path-step rows (K) per pipeline stage

struct TcTrain {
    uint64_t packed_small = 0;
    bf16* WT = nullptr;               // bgemm B operand: [part 2][ks G4/64][hi|lo][half 2][kc 8][n Nh][8]
    DBuf cache, hcache, dgop, dh, dc, wpart, gP, cls;
    DBuf dX, eg_keys, eg_vals, eg_tmp, eg_part, eg_flags;     // entity-gradient staging + sorted segment reduce
    DBuf xop, hop, ohop, dbg;         // gathered entity rows / h / one-hot slice in the weight-gradient kernel's stage layout
    // a second stream for the memory-bound side work of the backward (operand staging, entity-gradient segment reduce): it runs
    // beside the tensor-bound kernels of the main stream; fork / join with events, so the caller's stream semantics are unchanged
    cudaStream_t side = nullptr;
    bool use_side = true;
    cudaEvent_t ev_fork = nullptr, ev_loop = nullptr, ev_prep = nullptr, ev_eg = nullptr;
    // the BPTT loop walks the batch as two tile groups on two streams: cell_bwd (HBM-bound, 256-thread blocks) of one group runs
    // on the SMs beside the resident bgemm CTAs (tensor-bound, registers capped) of the other
    cudaStream_t side2 = nullptr;
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr;
    bool use_groups = true;
    DedupBufs es;                     // distinct entity rows of the batch when the caller (a sparse training step) did not build them
};

// ---------------------------------------------------------------------------------------------
// packing of the backward's weight operands (unscaled: the gradients are w.r.t. the true pre-activations)

// part 0: Wh^T as B of dh = dG Wh     -> element (n = hidden unit k, K index gp) = Wh[row(gp)][n]
// part 1: Wx_e^T as B of dX = dG Wx_e -> element (n = entity column,  K index gp) = Wx[row(gp)][d_t + n]
__global__ void pack_wt_kernel(const float* __restrict__ Wh, const float* __restrict__ Wx, int H, int D, int d_t, int d_e, int G4,
                               int N0, int N1, bf16* __restrict__ out) {
    const int KSg = G4 / 64;
    const int64_t per0 = (int64_t)KSg * 2 * 8 * N0 * 8, per1 = (int64_t)KSg * 2 * 8 * N1 * 8;     // elements of a part (hi + lo)
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (per0 + per1) / 2) return;                                                          // one thread per (hi, lo) pair
    const int part = idx < per0 / 2 ? 0 : 1;
    const int64_t li = part ? idx - per0 / 2 : idx;
    const int N = part ? N1 : N0, Nh = N / 2;
    // li indexes [ks][half][kc 8][n Nh][8]
    const int e = (int)(li & 7);
    int64_t r = li >> 3;
    const int n = (int)(r % Nh); r /= Nh;
    const int kc = (int)(r & 7); r >>= 3;
    const int half = (int)(r & 1); r >>= 1;
    const int ks = (int)r;
    const int gp = ks * 64 + kc * 8 + e, nn = half * Nh + n;
    int u;
    const int row = perm_to_row(gp, H, u);
    float v = 0.f;
    if (u < H) {
        if (part == 0) { if (nn < H) v = Wh[(int64_t)row * H + nn]; }
        else if (nn < d_e) v = Wx[(int64_t)row * D + d_t + nn];
    }
    bf16 h, l;
    split_bf16(v, h, l);
    bf16* base = out + (part ? per0 : 0) + (((int64_t)ks * 2 + 0) * 2 + half) * (8 * Nh * 8) + ((int64_t)kc * Nh + n) * 8 + e;
    base[0] = h;
    base[(int64_t)2 * 8 * Nh * 8] = l;          // the lo block follows the two halves of hi
}

#include "tc_bptt_steps.cuh"      // cell_bwd_kernel + bgemm_kernel: the per-step form (KPRN_TC_FUSED=0), cross-check of bptt_kernel

// ---------------------------------------------------------------------------------------------
// bptt_kernel: the whole time loop of the backward in ONE persistent launch -- cell_bwd_kernel's arithmetic as the producer of
// bgemm_kernel's A operand.  A CTA pair keeps a pair of path tiles for all T steps:
//   cell warps (16; thread = path row x 8 hidden units, the mapping of the recurrence epilogue)
//       read cache_t / c_{t-1} / dh_t / dc, write dG_t as bf16 hi/lo straight into the shared-memory stage the MMAs read (and to
//       dgop for the weight-gradient kernel), a 64-column K slice at a time: slice 2j = gates i,g of units [32j, 32j+32), slice
//       2j+1 = gates f,o of the same units
//   warp 0 streams the matching [Wh^T | Wx_e^T] slices (TMA bulk), warp 1 of the leader issues the cta_group::2 MMAs
//   after the 16 slices of a step the cell warps drain TMEM: dh_{t-1} -> a tile-private fp32 buffer (written and re-read by the
//       same thread 20-30 us later: it lives in L2), dX_t -> the staged entity-gradient rows
// Against the cell_bwd + bgemm pair of launches per step this removes bgemm's read of dgop and the DRAM round trips of dh and dc
// (10 KB instead of 18.7 KB of DRAM traffic per path-step) and every launch boundary of the loop.
constexpr int BP_THREADS = 640;       // warp 0 TMA producer, 1 MMA issuer / peer forwarder, 2 TMEM alloc, 3 idle, 4-19 cell + drain

struct BpArgs {
    const float* cache; const bf16* WT; bf16* dgop; float* dh; float* dc; float* dX;
    const float* ds_c; const float* wo;
    int T, NCH, H, G4, N0, N1; int64_t num_tiles, N;
};

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const float r0 = v[2 * i] - __uint_as_float(hb << 16), r1 = v[2 * i + 1] - __uint_as_float(hb & 0xFFFF0000u);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        wh[i] = hb; wl[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    lo = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}
__device__ __forceinline__ void ld8(const float* p, float* v) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// one cache array of a thread's 8 units: [half 2][row 128][4]; streaming (read once) / c (read by two consecutive steps)
__device__ __forceinline__ void ld_cache8c(const float* p, float* v) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p + 512));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld_cache8(const float* p, float* v) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(p)), b = __ldcs(reinterpret_cast<const float4*>(p + 512));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BP_THREADS, 1) bptt_kernel(BpArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    constexpr int S = BG_STAGES;
    const int Nh0 = a.N0 / 2, Nh1 = a.N1 / 2;
    const uint32_t A_BYTES = 16384;                                   // one of hi / lo of a 64-wide k-slice: [kc 8][row 128][16 B]
    const uint32_t B0_BYTES = 8u * Nh0 * 16, B1_BYTES = 8u * Nh1 * 16;
    const uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B0_BYTES + 2 * B1_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    // [0,S) full (B bytes + 16 cell warps)  [S,2S) empty  [2S,3S) peer_full (leader)  3S acc_full  3S+1 acc_empty (leader)
    constexpr int B_PEER = 2 * S, B_ACCF = 3 * S, B_ACCE = 3 * S + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACCE + 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { ptx::mbar_init(BAR(i), 17); ptx::mbar_init(BAR(S + i), 1); ptx::mbar_init(BAR(B_PEER + i), 1); }
        ptx::mbar_init(BAR(B_ACCF), 1); ptx::mbar_init(BAR(B_ACCE), 32);
        ptx::fence_mbar_init();
    }
    if (warp == 2) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t npairs = gridDim.x / 2, pair = blockIdx.x / 2, n_tp = (a.num_tiles + 1) / 2;
    const int64_t my_tiles = pair < n_tp ? (n_tp - pair + npairs - 1) / npairs : 0;
    const int KSg = a.G4 / 64, T = a.T;
    const size_t hl = (size_t)(a.G4 / 8) * TILE_M * 8;               // elements of one of hi / lo of a dgop tile
    const int64_t per0 = (int64_t)KSg * 2 * 8 * a.N0 * 8;            // elements of part 0 of WT

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int t = T - 1; t >= 0; --t)
                    for (int ks = 0; ks < KSg; ++ks) {
                        ptx::mbar_wait(BAR(S + stage), phase ^ 1);
                        ptx::mbar_arrive_expect_tx(BAR(stage), (t > 0 ? 2 * B0_BYTES : 0) + 2 * B1_BYTES);
                        const uint32_t sb0 = ptx::smem_u32(smem + stage * STAGE_BYTES) + 2 * A_BYTES;
                        if (t > 0) {
                            const bf16* W0 = a.WT + (((int64_t)ks * 2 + 0) * 2 + rank) * (8 * Nh0 * 8);
                            ptx::tma_bulk_g2s(sb0, W0, B0_BYTES, BAR(stage));
                            ptx::tma_bulk_g2s(sb0 + B0_BYTES, W0 + (int64_t)2 * 8 * Nh0 * 8, B0_BYTES, BAR(stage));
                        }
                        const bf16* W1 = a.WT + per0 + (((int64_t)ks * 2 + 0) * 2 + rank) * (8 * Nh1 * 8);
                        ptx::tma_bulk_g2s(sb0 + 2 * B0_BYTES, W1, B1_BYTES, BAR(stage));
                        ptx::tma_bulk_g2s(sb0 + 2 * B0_BYTES + B1_BYTES, W1 + (int64_t)2 * 8 * Nh1 * 8, B1_BYTES, BAR(stage));
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        if (leader) {
            const bool issue = ptx::elect_one();
            uint32_t stage = 0, phase = 0, e = 0;
            constexpr uint32_t DHI = ptx::desc_hi(128);
            const uint32_t idesc0 = ptx::make_idesc_f16(1, 256, a.N0), idesc1 = ptx::make_idesc_f16(1, 256, a.N1);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int t = T - 1; t >= 0; --t, ++e) {
                    ptx::mbar_wait_cluster(BAR(B_ACCE), (e & 1) ^ 1);
                    ptx::tc_fence_after();
                    for (int ks = 0; ks < KSg; ++ks) {
                        ptx::mbar_wait(BAR(stage), phase);
                        ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase);
                        ptx::tc_fence_after();
                        const uint32_t sb0 = ptx::smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t ah = ptx::desc_lo(sb0, 2048), al = ptx::desc_lo(sb0 + A_BYTES, 2048);
                        if (issue) {
                            for (int part = (t > 0 ? 0 : 1); part < 2; ++part) {
                                const uint32_t Nh = part ? Nh1 : Nh0;
                                const uint32_t boff = 2 * A_BYTES + (part ? 2 * B0_BYTES : 0), bbytes = part ? B1_BYTES : B0_BYTES;
                                const uint32_t bh = ptx::desc_lo(sb0 + boff, Nh * 16), bl = ptx::desc_lo(sb0 + boff + bbytes, Nh * 16);
                                const uint32_t k16b = (2 * Nh * 16) >> 4;
                                const uint32_t d = tmem + part * 256;
                                const uint32_t idesc = part ? idesc1 : idesc0;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma2_f16(d, ptx::desc_join(DHI, ah + k * 256u), ptx::desc_join(DHI, bh + k * k16b), idesc, (ks | k) != 0);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma2_f16(d, ptx::desc_join(DHI, al + k * 256u), ptx::desc_join(DHI, bh + k * k16b), idesc, 1);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    ptx::umma2_f16(d, ptx::desc_join(DHI, ah + k * 256u), ptx::desc_join(DHI, bl + k * k16b), idesc, 1);
                            }
                            ptx::umma2_commit(BAR(S + stage), 3);
                        }
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                    if (issue) ptx::umma2_commit(BAR(B_ACCF), 3);
                }
        } else if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t remote = ptx::mapa(BAR(B_PEER), 0);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int64_t q = 0; q < (int64_t)T * KSg; ++q) {
                    ptx::mbar_wait(BAR(stage), phase);
                    ptx::mbar_arrive_cluster(remote + 8u * stage);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp >= 4) {
        // ===== cell warps: warp (q4, sb) owns path rows q4*32 + lane (its TMEM lane quarter) and the sb-th 8-unit group of a round =====
        const int q4 = warp & 3, sb = (warp - 4) >> 2, r = q4 * 32 + lane;
        const uint32_t acce_leader = ptx::mapa(BAR(B_ACCE), 0);
        const size_t step_floats = (size_t)a.NCH * 20480;
        const uint32_t tb = tmem + ((uint32_t)(q4 * 32) << 16);
        const int xcols = a.N1 / 4;                                   // dX columns this warp drains
        uint32_t stage = 0, phase = 0, e = 0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t mraw = 2 * (pair + it * npairs) + rank;
            const bool dup = mraw >= a.num_tiles;                     // odd tile count: the pair's second CTA recomputes the last tile, stores nothing
            const int64_t m = dup ? a.num_tiles - 1 : mraw;
            const int64_t n = m * TILE_M + r;
            const bool live = n < a.N;
            const float s_c = live ? a.ds_c[n] : 0.f;
            for (int t = T - 1; t >= 0; --t, ++e) {
                const float* cstep = a.cache + (size_t)(m * T + t) * step_floats + r * 4;
                bf16* gbase = a.dgop + (size_t)(m * T + t) * 2 * hl + (size_t)r * 8;
                for (int j = 0; j < a.NCH; ++j) {
                    const int grp = j * 4 + sb;
                    const float* ct = cstep + (size_t)grp * 5120;
                    const size_t gi8 = ((size_t)(m * (a.G4 / 32) + grp) * TILE_M + r) * 8;
                    float dcu[8], dO[8], di[8], dg[8];
                    {
                        float gi[8], gg[8], go[8], cc[8], dh[8], dc[8];
                        ld_cache8(ct, gi); ld_cache8(ct + 1024, gg); ld_cache8(ct + 3072, go); ld_cache8c(ct + 4096, cc);
                        if (t < T - 1) { ld8(a.dh + gi8, dh); ld8(a.dc + gi8, dc); }
                        else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) { const int u = grp * 8 + q; dh[q] = u < a.H ? s_c * a.wo[u] : 0.f; dc[q] = 0.f; }
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float tc = tanhf(cc[q]);
                            const float d_o = dh[q] * tc;
                            dcu[q] = dc[q] + dh[q] * go[q] * (1.f - tc * tc);
                            di[q] = live ? dcu[q] * gg[q] * gi[q] * (1.f - gi[q]) : 0.f;
                            dg[q] = live ? dcu[q] * gi[q] * (1.f - gg[q] * gg[q]) : 0.f;
                            dO[q] = live ? d_o * go[q] * (1.f - go[q]) : 0.f;
                        }
                    }
                    float gf[8], cp[8];
                    ld_cache8(ct + 2048, gf);
                    if (t > 0) ld_cache8c(ct - step_floats + 4096, cp);
                    else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) cp[q] = 0.f;
                    }
                    // ---- slice 2j: gates i, g ----
                    uint4 h0, l0, h1, l1;
                    split8(di, h0, l0); split8(dg, h1, l1);
                    ptx::mbar_wait(BAR(S + stage), phase ^ 1);
                    {
                        uint8_t* sA = smem + stage * STAGE_BYTES + ((size_t)sb * TILE_M + r) * 16;
                        *reinterpret_cast<uint4*>(sA) = h0; *reinterpret_cast<uint4*>(sA + A_BYTES) = l0;
                        *reinterpret_cast<uint4*>(sA + 4 * TILE_M * 16) = h1; *reinterpret_cast<uint4*>(sA + 4 * TILE_M * 16 + A_BYTES) = l1;
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(BAR(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                    if (!dup) {
                        bf16* g0 = gbase + (size_t)(j * 16 + sb) * TILE_M * 8;
                        *reinterpret_cast<uint4*>(g0) = h0; *reinterpret_cast<uint4*>(g0 + hl) = l0;
                        *reinterpret_cast<uint4*>(g0 + 4 * TILE_M * 8) = h1; *reinterpret_cast<uint4*>(g0 + 4 * TILE_M * 8 + hl) = l1;
                    }
                    // ---- slice 2j+1: gates f, o ----
                    float df[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        df[q] = live ? dcu[q] * cp[q] * gf[q] * (1.f - gf[q]) : 0.f;
                        dcu[q] *= gf[q];                                                       // dc for step t-1
                    }
                    split8(df, h0, l0); split8(dO, h1, l1);
                    ptx::mbar_wait(BAR(S + stage), phase ^ 1);
                    {
                        uint8_t* sA = smem + stage * STAGE_BYTES + ((size_t)sb * TILE_M + r) * 16;
                        *reinterpret_cast<uint4*>(sA) = h0; *reinterpret_cast<uint4*>(sA + A_BYTES) = l0;
                        *reinterpret_cast<uint4*>(sA + 4 * TILE_M * 16) = h1; *reinterpret_cast<uint4*>(sA + 4 * TILE_M * 16 + A_BYTES) = l1;
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(BAR(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                    if (!dup) {
                        bf16* g0 = gbase + (size_t)(j * 16 + 8 + sb) * TILE_M * 8;
                        *reinterpret_cast<uint4*>(g0) = h0; *reinterpret_cast<uint4*>(g0 + hl) = l0;
                        *reinterpret_cast<uint4*>(g0 + 4 * TILE_M * 8) = h1; *reinterpret_cast<uint4*>(g0 + 4 * TILE_M * 8 + hl) = l1;
                        if (t > 0) ptx::stg256(a.dc + gi8, dcu);
                    }
                }
                // ---- drain: dh_{t-1} (this thread's own groups) and this warp's quarter of the dX_t columns ----
                ptx::mbar_wait(BAR(B_ACCF), e & 1);
                ptx::tc_fence_after();
                if (t > 0) {
                    for (int j = 0; j < a.NCH; ++j) {
                        float v[8];
                        ptx::tmem_ld8(tb + (uint32_t)(j * 4 + sb) * 8, v);
                        ptx::tc_wait_ld();
                        if (!dup) ptx::stg256(a.dh + ((size_t)(m * (a.G4 / 32) + j * 4 + sb) * TILE_M + r) * 8, v);
                    }
                }
                {
                    float* dst = a.dX + ((size_t)(m * T + t) * TILE_M + r) * a.N1 + sb * xcols;
                    for (int blk = 0; blk < xcols / 32; ++blk) {
                        float v[32];
                        ptx::tmem_ld32(tb + 256 + sb * xcols + blk * 32, v);
                        ptx::tc_wait_ld();
                        if (!dup) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) ptx::stg256(dst + blk * 32 + 8 * i, v + 8 * i);
                        }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(acce_leader);
            }
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem, 512);
}


// Operands of the weight-gradient product, staged once per backward in the layout its stages consume.  A stage of
// wgrad_kernel is 64 path rows; the TMA unit sustains roughly one contiguous segment per ~55 cycles per SM whatever its size
// (measured: 64 one-kilobyte copies per stage made the kernel issue-bound), so everything that is free to choose its layout
// is stored [(tile*T + t)][hi|lo][row half 2][kc][row 64][8 bf16]: the [kc][64 rows] tile of a stage is ONE contiguous copy.
//   xop : the gathered entity rows x_e of every path-step (from the pre-split table), kc = KE/8
//   hop : h_t re-laid out from the forward's hcache ([kc][row 128][8], which its own MMAs need), kc = Hp/8, t < T-1
//   ohop: onehot(type*V_r + relation), 64 columns, exact in bf16 (hi only): [(tile*T + t)][row half][kc 8][row 64][8]
// One block per (tile, t), one thread per path row.
__global__ void __launch_bounds__(128) prep_operands_kernel(const int32_t* __restrict__ ids, int64_t N, int T, int64_t V_e, int V_t, int V_r,
                                                            const bf16* __restrict__ E_hi, const bf16* __restrict__ E_lo, int KE,
                                                            const bf16* __restrict__ hcache, int Hp,
                                                            bf16* __restrict__ xop, bf16* __restrict__ hop, bf16* __restrict__ ohop) {
    const int64_t b = blockIdx.x;                 // tile*T + t
    const int r = threadIdx.x, rh = r >> 6, rr = r & 63;
    const int64_t m = b / T;
    const int t = (int)(b - m * T);
    const int64_t n = m * TILE_M + r;
    const bool live = n < N;
    int64_t e = 0; int combo = -1;
    if (live) {
        const int32_t* idp = ids + (n * T + t) * 3;
        e = (int64_t)idp[1] - 1;
        e = e < 0 ? 0 : (e >= V_e ? V_e - 1 : e);
        int ta = idp[0] - 1, rb = idp[2] - 1;
        ta = ta < 0 ? 0 : (ta >= V_t ? V_t - 1 : ta);
        rb = rb < 0 ? 0 : (rb >= V_r ? V_r - 1 : rb);
        combo = ta * V_r + rb;
    }
    const int KC = KE / 8, KH = Hp / 8;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4* sh = reinterpret_cast<const uint4*>(E_hi + e * KE);
    const uint4* sl = reinterpret_cast<const uint4*>(E_lo + e * KE);
    // 16-byte units: [(b*2 + hl)*2 + rh][kc][rr]
    uint4* dxh = reinterpret_cast<uint4*>(xop) + ((size_t)((b * 2 + 0) * 2 + rh) * KC) * 64 + rr;
    uint4* dxl = reinterpret_cast<uint4*>(xop) + ((size_t)((b * 2 + 1) * 2 + rh) * KC) * 64 + rr;
    for (int kc = 0; kc < KC; ++kc) {             // a warp's store is 512 contiguous bytes
        dxh[(size_t)kc * 64] = live ? sh[kc] : z;
        dxl[(size_t)kc * 64] = live ? sl[kc] : z;
    }
    if (t < T - 1) {
        const uint4* hs = reinterpret_cast<const uint4*>(hcache) + (size_t)(b * 2) * KH * TILE_M + r;
        uint4* dhh = reinterpret_cast<uint4*>(hop) + ((size_t)((b * 2 + 0) * 2 + rh) * KH) * 64 + rr;
        uint4* dhl = reinterpret_cast<uint4*>(hop) + ((size_t)((b * 2 + 1) * 2 + rh) * KH) * 64 + rr;
        for (int kc = 0; kc < KH; ++kc) {
            dhh[(size_t)kc * 64] = hs[(size_t)kc * TILE_M];
            dhl[(size_t)kc * 64] = hs[(size_t)(KH + kc) * TILE_M];
        }
    }
    uint4* doh = reinterpret_cast<uint4*>(ohop) + ((size_t)(b * 2 + rh) * 8) * 64 + rr;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (combo >= 0 && (combo >> 3) == kc) w[(combo & 7) >> 1] = 0x3F80u << (16 * (combo & 1));      // bf16 1.0
        doh[(size_t)kc * 64] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// wgrad: G[gp, n] = sum over path-steps of dG[step, gp] * B[step, n]  for one 256-row block of gate columns (128 per CTA of a
// pair) and one operand family:
//   part 0: B = [h_{t-1} (Hp columns) | onehot(type,relation) (64 columns)]   -> gWh and gP (two accumulators)
//   part 1: B = x_e (KE columns)                                              -> gWx[:, entity columns]
// A = dgop read as an MN-major operand (M = gate column, K = path row): the bytes [kc][row][8] that bgemm reads K-major;
// B likewise (hcache / xop / ohop).  A stage is 64 path rows: every operand half is sixteen (or fewer) 1 KB pieces, one
// cp.async.bulk per lane of the producer warp.  K is split over gridDim.y CTA pairs; each writes its partial tile,
// wgrad_reduce_kernel sums them in a fixed order.

struct WgArgs {
    const bf16* dgop; const bf16* hop; const bf16* xop; const bf16* ohop;
    unsigned long long* dbg;                      // optional wait-cycle counters (KPRN_TC_DEBUG=1)
    int T; int64_t num_tiles;
    int G4, N0, KE;
    int64_t blocks_per_split;       // (tile, t) blocks per K-split
    float* part;                    // [split][part 2][gp G4][n 320] fp32
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1) wgrad_kernel(WgArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    constexpr int S = WG_STAGES;
    // work item: blockIdx.x / 2 = (mp, part) ; blockIdx.y = K split
    const int item = blockIdx.x / 2, part = item & 1, mp = item >> 1;
    const int split = blockIdx.y;
    const int Nmain = part ? a.KE : a.N0, Nh = Nmain / 2;
    constexpr uint32_t A_BYTES = 16 * WG_ROWS * 16;                  // [kc 16][row 64][16 B] = 16 KB per hi / lo
    const uint32_t B_BYTES = (uint32_t)Nh * WG_ROWS * 2;             // per hi / lo: [kc Nh/8][row 64][16 B]
    constexpr uint32_t OH_BYTES = 4 * WG_ROWS * 16;                  // one-hot: this CTA's 32 combos x 64 rows (hi only)
    constexpr uint32_t B_OFF = 2 * A_BYTES, B_LO_OFF = B_OFF + 16384, OH_OFF = B_LO_OFF + 16384, STAGE_BYTES = OH_OFF + OH_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    // [0,S) full (TMA tx)  [S,2S) empty  [2S,3S) peer_full (leader)  3S acc_full
    constexpr int B_PEER = 2 * S, B_ACCF = 3 * S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACCF + 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(S + i), 1); ptx::mbar_init(BAR(B_PEER + i), 1); }
        ptx::mbar_init(BAR(B_ACCF), 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t nblocks = a.num_tiles * a.T;
    const int64_t b0 = split * a.blocks_per_split, b1 = (b0 + a.blocks_per_split < nblocks) ? b0 + a.blocks_per_split : nblocks;
    const int kc_rows = a.G4 / 8;                                    // k-chunk rows of one of hi / lo of a dgop step
    const int bk_rows = (part ? a.KE : a.N0) / 8;                    // ... of an hcache / xop step
    const int kcA = (mp * 2 + (int)rank) * 16;                       // first k-chunk of this CTA's 128 gate columns
    const long long dbg_t0 = clock64();

    if (warp == 0) {
        // ===== TMA producer.  A (dG, the layout bgemm needs): lane = (hi|lo, k-chunk), one 1 KB bulk copy each; B (hop / xop)
        //       and the one-hot slice come from the wgrad-layout copies: one contiguous copy per operand half =====
        uint32_t stage = 0, phase = 0;
        const uint8_t* dg8 = reinterpret_cast<const uint8_t*>(a.dgop);
        const uint8_t* b8 = reinterpret_cast<const uint8_t*>(part ? a.xop : a.hop);
        const uint8_t* oh8 = reinterpret_cast<const uint8_t*>(a.ohop);
        const int hlx = lane >> 4, kc = lane & 15;
        for (int64_t b = b0; b < b1; ++b) {
            const int t = (int)(b % a.T);
            const int64_t bsrc = part ? b : b - 1;                   // h_{t-1} of the same tile (t > 0), x_e of the step itself
            const bool has_b = part == 1 || t > 0;
            for (int half = 0; half < 2; ++half) {
                if (lane == 0) {
                    if (a.dbg) { long long t0 = clock64(); ptx::mbar_wait(BAR(S + stage), phase ^ 1); if (leader) atomicAdd(a.dbg + 3, (unsigned long long)(clock64() - t0)); }
                    else ptx::mbar_wait(BAR(S + stage), phase ^ 1);
                    ptx::mbar_arrive_expect_tx(BAR(stage), 2 * A_BYTES + (has_b ? 2 * B_BYTES : 0) + (part == 0 ? OH_BYTES : 0));
                }
                __syncwarp();
                const uint32_t sb0 = ptx::smem_u32(smem + stage * STAGE_BYTES);
                const size_t rowA = (size_t)((b * 2 + hlx) * kc_rows + kcA + kc);
                ptx::tma_bulk_g2s(sb0 + hlx * A_BYTES + kc * 1024, dg8 + rowA * 2048 + half * 1024, 1024, BAR(stage));
                if (has_b && lane < 2) {          // lane = hi | lo: [kc Nh/8][row 64][16 B] of this CTA's column half, contiguous
                    const size_t off = ((size_t)((bsrc * 2 + lane) * 2 + half) * bk_rows + (size_t)rank * (Nh / 8)) * 1024;
                    ptx::tma_bulk_g2s(sb0 + (lane ? B_LO_OFF : B_OFF), b8 + off, B_BYTES, BAR(stage));
                }
                if (part == 0 && lane == 2) {
                    const size_t off = ((size_t)(b * 2 + half) * 8 + (size_t)rank * 4) * 1024;
                    ptx::tma_bulk_g2s(sb0 + OH_OFF, oh8 + off, OH_BYTES, BAR(stage));
                }
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===== MMA issuer =====
            const bool issue = ptx::elect_one();
            uint32_t stage = 0, phase = 0;
            // MN-major no-swizzle operand [k-chunk (8 mn)][row (k)][16 B]: LBO = 128 B between 8-row k groups, SBO = 1024 B
            // between 8-element mn groups (tools/umma_probe.cu); a K=16 step = 16 rows = 256 B
            constexpr uint32_t DHI_MN = ptx::desc_hi(1024);
            const uint32_t idesc_main = ptx::make_idesc_f16(1, 256, Nmain) | (1u << 15) | (1u << 16);
            const uint32_t idesc_oh = ptx::make_idesc_f16(1, 256, 64) | (1u << 15) | (1u << 16);
            bool first_main = true, first_oh = true;
            for (int64_t b = b0; b < b1; ++b) {
                const int t = (int)(b % a.T);
                const bool has_main = part == 1 || t > 0;
                for (int half = 0; half < 2; ++half) {
                    if (a.dbg) {
                        long long t0 = clock64(); ptx::mbar_wait(BAR(stage), phase); long long t1 = clock64();
                        ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase);
                        if (issue) { atomicAdd(a.dbg + 0, (unsigned long long)(t1 - t0)); atomicAdd(a.dbg + 1, (unsigned long long)(clock64() - t1)); }
                    } else {
                        ptx::mbar_wait(BAR(stage), phase);
                        ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase);
                    }
                    ptx::tc_fence_after();
                    const uint32_t sb0 = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t ah = ptx::desc_lo(sb0, 128), al = ptx::desc_lo(sb0 + A_BYTES, 128);
                    if (issue) {
                        if (has_main) {
                            const uint32_t bh = ptx::desc_lo(sb0 + B_OFF, 128), bl = ptx::desc_lo(sb0 + B_LO_OFF, 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma2_f16(tmem, ptx::desc_join(DHI_MN, ah + k * 16u), ptx::desc_join(DHI_MN, bh + k * 16u), idesc_main, !(first_main && k == 0));
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma2_f16(tmem, ptx::desc_join(DHI_MN, al + k * 16u), ptx::desc_join(DHI_MN, bh + k * 16u), idesc_main, 1);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma2_f16(tmem, ptx::desc_join(DHI_MN, ah + k * 16u), ptx::desc_join(DHI_MN, bl + k * 16u), idesc_main, 1);
                            first_main = false;
                        }
                        if (part == 0) {
                            const uint32_t oh = ptx::desc_lo(sb0 + OH_OFF, 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma2_f16(tmem + 256, ptx::desc_join(DHI_MN, ah + k * 16u), ptx::desc_join(DHI_MN, oh + k * 16u), idesc_oh, !(first_oh && k == 0));
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::umma2_f16(tmem + 256, ptx::desc_join(DHI_MN, al + k * 16u), ptx::desc_join(DHI_MN, oh + k * 16u), idesc_oh, 1);
                            first_oh = false;
                        }
                        ptx::umma2_commit(BAR(S + stage), 3);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
            if (issue) ptx::umma2_commit(BAR(B_ACCF), 3);
            if (issue && a.dbg) atomicAdd(a.dbg + 4, (unsigned long long)(clock64() - dbg_t0));
        } else if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t remote = ptx::mapa(BAR(B_PEER), 0);
            for (int64_t b = b0; b < b1; ++b)
                for (int half = 0; half < 2; ++half) {
                    ptx::mbar_wait(BAR(stage), phase);
                    ptx::mbar_arrive_cluster(remote + 8u * stage);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp >= 4) {
        // ===== epilogue: this CTA's 128 gate columns x (Nmain [+ 64]) -> partial buffer [split][part][gp][320] =====
        const int q4 = warp & 3, r = q4 * 32 + lane;
        ptx::mbar_wait(BAR(B_ACCF), 0);
        ptx::tc_fence_after();
        const int gp = mp * 256 + (int)rank * 128 + r;
        float* dst = a.part + (((size_t)split * 2 + part) * a.G4 + gp) * 320;
        const uint32_t tb = tmem + ((uint32_t)(q4 * 32) << 16);
        const bool any_main = b1 > b0 && (part == 1 || (a.T > 1 && (b1 - b0 > 1 || (b0 % a.T) != 0)));
        for (int blk = 0; blk < Nmain / 32; ++blk) {
            float v[32];
            if (any_main) { ptx::tmem_ld32(tb + blk * 32, v); ptx::tc_wait_ld(); }
            else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dst + blk * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (part == 0) {
            for (int blk = 0; blk < 2; ++blk) {
                float v[32];
                if (b1 > b0) { ptx::tmem_ld32(tb + 256 + blk * 32, v); ptx::tc_wait_ld(); }
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(dst + 256 + blk * 32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        }
        ptx::tc_fence_before();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem, 512);
}

// fixed-order sum of the K-split partials into the flat gradient (accumulating, as accGradParameters does):
//   part 0, n < N0        -> gWh[row(gp)][n]          part 0, n in [256, 320) -> gP[n - 256][gp] (temporary, overwritten)
//   part 1, n < KE        -> gWx[row(gp)][d_t + n]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int G4, int H, int D, int d_t, int d_e, int N0, int KE,
                                    float* __restrict__ gWh, float* __restrict__ gWx, float* __restrict__ gP) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)2 * G4 * 320) return;
    const int n = (int)(idx % 320);
    const int gp = (int)((idx / 320) % G4);
    const int pt = (int)(idx / ((int64_t)320 * G4));
    if (pt == 0 && n >= N0 && n < 256) return;
    if (pt == 1 && n >= KE) return;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += part[(((size_t)sp * 2 + pt) * G4 + gp) * 320 + n];
    int u;
    const int row = perm_to_row(gp, H, u);
    if (pt == 0 && n >= 256) { gP[(size_t)(n - 256) * G4 + gp] = acc; return; }
    if (u >= H) return;
    if (pt == 0) { if (n < H) gWh[(size_t)row * H + n] += acc; }
    else if (n < d_e) gWx[(size_t)row * D + d_t + n] += acc;
}

// gradient of the hoisted (type,relation) table -> bias, per-type / per-relation sums
__global__ void ptr_sums_kernel(const float* __restrict__ gP, int G4, int H, int V_t, int V_r, float* __restrict__ Tsum, float* __restrict__ Rsum,
                                float* __restrict__ gbx) {
    const int gp = blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= G4) return;
    int u;
    const int row = perm_to_row(gp, H, u);
    float all = 0.f;
    for (int a = 0; a < V_t; ++a) {
        float s = 0.f;
        for (int b = 0; b < V_r; ++b) s += gP[(size_t)(a * V_r + b) * G4 + gp];
        Tsum[(size_t)a * G4 + gp] = s;
        all += s;
    }
    for (int b = 0; b < V_r; ++b) {
        float s = 0.f;
        for (int a = 0; a < V_t; ++a) s += gP[(size_t)(a * V_r + b) * G4 + gp];
        Rsum[(size_t)b * G4 + gp] = s;
    }
    if (u < H) gbx[row] += all;
}
// gE_type[a][k] += sum_gp Tsum[a][gp] Wx[row(gp)][k];  gE_rel[b][k] += sum_gp Rsum[b][gp] Wx[row(gp)][d_t + d_e + k]
// grid (V_t + V_r, column chunks of 64), block (64 columns, 8 gate-column ranges): coalesced weight reads, the 8 range partials of
// an output element are combined through shared memory in range order (no float atomics: the same bits on every run)
__global__ void ptr_embed_kernel(const float* __restrict__ Tsum, const float* __restrict__ Rsum, const float* __restrict__ Wx, int G4, int H,
                                 int D, int d_t, int d_e, int d_r, int V_t, float* __restrict__ gEt, float* __restrict__ gEr) {
    const int which = blockIdx.x;                // [0, V_t): type rows; then relation rows
    const bool is_t = which < V_t;
    const int rowi = is_t ? which : which - V_t;
    const int width = is_t ? d_t : d_r, coff = is_t ? 0 : d_t + d_e;
    const float* S = (is_t ? Tsum : Rsum) + (size_t)rowi * G4;
    __shared__ float part[8][64];
    const int k = blockIdx.y * 64 + threadIdx.x;
    const int g0 = threadIdx.y * (G4 / 8), g1 = g0 + G4 / 8;
    float acc = 0.f;
    if (k < width)
        for (int gp = g0; gp < g1; ++gp) {
            int u;
            const int row = perm_to_row(gp, H, u);
            if (u < H) acc = fmaf(S[gp], Wx[(size_t)row * D + coff + k], acc);
        }
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && k < width) {
        float t = 0.f;
#pragma unroll
        for (int z = 0; z < 8; ++z) t += part[z][threadIdx.x];
        ((is_t ? gEt : gEr) + (size_t)rowi * width)[k] += t;
    }
}
// gWx[row(gp)][k] += sum_a Tsum[a][gp] E_t[a][k]  (type columns)  and the same with Rsum / E_r for the relation columns
__global__ void ptr_wx_kernel(const float* __restrict__ Tsum, const float* __restrict__ Rsum, const float* __restrict__ Et, const float* __restrict__ Er,
                              int G4, int H, int D, int d_t, int d_e, int d_r, int V_t, int V_r, float* __restrict__ gWx) {
    const int gp = blockIdx.x;
    int u;
    const int row = perm_to_row(gp, H, u);
    if (u >= H) return;
    for (int k = threadIdx.x; k < d_t + d_r; k += blockDim.x) {
        float acc = 0.f;
        if (k < d_t) { for (int a = 0; a < V_t; ++a) acc = fmaf(Tsum[(size_t)a * G4 + gp], Et[(size_t)a * d_t + k], acc); }
        else { const int kk = k - d_t; for (int b = 0; b < V_r; ++b) acc = fmaf(Rsum[(size_t)b * G4 + gp], Er[(size_t)b * d_r + kk], acc); }
        gWx[(size_t)row * D + (k < d_t ? k : d_e + k)] += acc;
    }
}
// class logit of every path from the fused head partials (what pool_bwd_kernel reads): s[p*L + cls] = sum of 4 partials + bias
__global__ void cls_logit_kernel(const float* __restrict__ part4, int64_t npad, const float* __restrict__ bias, int64_t N, int L, int cls,
                                 float* __restrict__ path_logits) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) path_logits[p * L + cls] = ((part4[p] + part4[npad + p]) + (part4[2 * npad + p] + part4[3 * npad + p])) + *bias;
}

// ---------------------------------------------------------------------------------------------
// LookupTable:accGradParameters for the entity table WITHOUT atomics: dX[o] (one row per path-step, staged by bgemm) is summed
// per distinct entity row in a fixed order.  (slot, o) pairs are radix-sorted by slot (stable: occurrences stay in index
// order); a warp per chunk of EG_CHUNK sorted entries sums runs of equal slot; a run that lies inside one chunk is added to
// its gradient row directly (no other writer exists), runs crossing chunk boundaries leave partial rows that one warp per run
// combines in chunk order.  Deterministic, and immune to the contention float atomics suffer on hot rows (the pad row alone
// is ~20 % of all steps of a batch).

constexpr int EG_CHUNK = 64;

__global__ void eg_keys_kernel(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ slot_of, int64_t n, uint32_t* __restrict__ keys,
                               uint32_t* __restrict__ vals) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const uint32_t p = pos[o];
    keys[o] = p == 0xFFFFFFFFu ? (uint32_t)n : slot_of[p];      // rows of a ragged last tile sort to the end
    vals[o] = (uint32_t)o;
}

struct EgArgs {
    const uint32_t* keys; const uint32_t* vals; int64_t n; int64_t n_chunks;
    const float* dX; int ldx; int d_e;
    float* target; const int32_t* rows;       // target row of slot s: rows ? rows[s] - 1 : s   (dense table / compact buffer)
    int assign;                               // compact buffer of a sparse step: every slot has exactly one writer, which stores (no zero-fill pass, no read)
    float* part;                              // [chunk][2][d_e]: 0 = run touching the chunk's begin, 1 = run touching its end
    int32_t* flags;                           // [chunk][4]: head key (-1 none), head continues into the next chunk, tail key (-1 none), unused
};

// NV4 = d_e / 128: a lane owns the float4 at columns [4 lane, +4) of every 128-column block of a row (coalesced 512-byte warp loads)
template <int NV4>
__global__ void __launch_bounds__(256) eg_reduce_kernel(EgArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (chunk >= a.n_chunks) return;
    const int64_t e0 = chunk * EG_CHUNK, e1 = (e0 + EG_CHUNK < a.n) ? e0 + EG_CHUNK : a.n;
    const int cnt = (int)(e1 - e0);
    const uint32_t INVALID = (uint32_t)a.n;
    const uint32_t prev_key = e0 > 0 ? a.keys[e0 - 1] : INVALID;
    const uint32_t next_key = e1 < a.n ? a.keys[e1] : INVALID;
    // the chunk's keys and row indices, two per lane
    const uint32_t k0 = lane < cnt ? a.keys[e0 + lane] : INVALID, k1 = lane + 32 < cnt ? a.keys[e0 + 32 + lane] : INVALID;
    const uint32_t v0 = lane < cnt ? a.vals[e0 + lane] : 0u, v1 = lane + 32 < cnt ? a.vals[e0 + 32 + lane] : 0u;
    int32_t head_key = -1, head_cont = 0, tail_key = -1;
    float4 acc[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t cur = INVALID;
    bool cur_starts = false;
    auto flush = [&](bool ends_here) {
        if (cur == INVALID) return;
        float* dst;
        if (cur_starts && ends_here) {
            dst = a.target + (size_t)(a.rows ? a.rows[cur] - 1 : cur) * a.d_e;
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                float4 o = a.assign ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(dst + i * 128 + lane * 4);
                o.x += acc[i].x; o.y += acc[i].y; o.z += acc[i].z; o.w += acc[i].w;
                *reinterpret_cast<float4*>(dst + i * 128 + lane * 4) = o;
            }
        } else {
            const int which = cur_starts ? 1 : 0;
            dst = a.part + ((size_t)chunk * 2 + which) * a.d_e;
#pragma unroll
            for (int i = 0; i < NV4; ++i) *reinterpret_cast<float4*>(dst + i * 128 + lane * 4) = acc[i];
            if (which) tail_key = (int32_t)cur;
            else { head_key = (int32_t)cur; head_cont = ends_here ? 0 : 1; }
        }
#pragma unroll
        for (int i = 0; i < NV4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    for (int base = 0; base < cnt; base += 4) {
        uint32_t key[4];
        float4 row[4][NV4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                   // four rows in flight
            const int i = base + u;
            key[u] = __shfl_sync(0xffffffffu, i < 32 ? k0 : k1, i & 31);
            const uint32_t val = __shfl_sync(0xffffffffu, i < 32 ? v0 : v1, i & 31);
            if (i >= cnt) key[u] = INVALID;
            if (key[u] != INVALID) {
                const float* src = a.dX + (size_t)val * a.ldx;
#pragma unroll
                for (int q = 0; q < NV4; ++q) row[u][q] = *reinterpret_cast<const float4*>(src + q * 128 + lane * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (key[u] == INVALID) continue;            // sorted: padding comes last
            if (key[u] != cur) {
                flush(true);
                cur = key[u];
                cur_starts = !(base + u == 0 && prev_key == cur);
            }
#pragma unroll
            for (int q = 0; q < NV4; ++q) { acc[q].x += row[u][q].x; acc[q].y += row[u][q].y; acc[q].z += row[u][q].z; acc[q].w += row[u][q].w; }
        }
    }
    // the last run ends here unless the next chunk starts with the same key (a run ending exactly at the array's valid end ends here)
    {
        const uint32_t last_key = __shfl_sync(0xffffffffu, (cnt - 1) < 32 ? k0 : k1, (cnt - 1) & 31);
        flush(!(cur == last_key && next_key == cur));
    }
    if (lane == 0) { a.flags[chunk * 4] = head_key; a.flags[chunk * 4 + 1] = head_cont; a.flags[chunk * 4 + 2] = tail_key; }
}

// one BLOCK per chunk that owns a boundary-crossing run (the chunk where the run starts): its tail partial plus the head
// partials of the following chunks (contiguous, the keys are sorted).  Warp w takes chunks start+1+w, +16, ... ; the warps'
// partial sums are combined in warp order: deterministic.
template <int NV4>
__global__ void __launch_bounds__(512) eg_fixup_kernel(EgArgs a) {
    __shared__ float4 red[16][NV4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t chunk = blockIdx.x;
    const int32_t key = a.flags[chunk * 4 + 2];
    if (key < 0) return;
    float4 acc[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t j = chunk + 1 + w; j < a.n_chunks; j += 16) {
        if (a.flags[j * 4] != key) break;
        const float* s2 = a.part + ((size_t)j * 2) * a.d_e;
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(s2 + q * 128 + lane * 4);
            acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
        }
    }
#pragma unroll
    for (int q = 0; q < NV4; ++q) red[w][q][lane] = acc[q];
    __syncthreads();
    if (w == 0) {
        const float* t0 = a.part + ((size_t)chunk * 2 + 1) * a.d_e;
        float* dst = a.target + (size_t)(a.rows ? a.rows[key] - 1 : key) * a.d_e;
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            float4 t = *reinterpret_cast<const float4*>(t0 + q * 128 + lane * 4);
            for (int ww = 0; ww < 16; ++ww) { const float4 v = red[ww][q][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
            float4 o = a.assign ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(dst + q * 128 + lane * 4);
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            *reinterpret_cast<float4*>(dst + q * 128 + lane * 4) = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side

static TcTrain* tt_get(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    if (!s->tt) s->tt = new TcTrain();
    return s->tt;
}

bool tc_train_supported(const kprn_ctx* ctx) {
    const kprn_config& c = ctx->cfg;
    if (!tc_supported(ctx) || c.precision != KPRN_PREC_TC_ALL || c.head_layers != 1 || c.H > 256) return false;
    const int KE = round_up(c.d_e, 64);
    if (c.d_e != 128 && c.d_e != 256) return false;      // wgrad stages the entity operand in 64-column swizzled panels per CTA; the
                                                         // segment reduce moves rows as float4 per lane per 128 columns
    if (c.V_t * c.V_r > 64) return false;                // one one-hot slice
    { const char* d = getenv("KPRN_TC_TRAIN"); if (d && d[0] == '0') return false; }
    return true;
}

int tc_train_pack(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    TcTrain* tt = tt_get(ctx);
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    const int KSg = s->G4 / 64;
    const int64_t elems = (int64_t)KSg * 2 * 8 * (s->Hp + s->KE) * 8;
    if (!tt->WT) KPRN_CUDA(cudaMalloc(&tt->WT, sizeof(bf16) * (size_t)elems));
    if (tt->packed_small == ctx->small_version + ctx->params_version * 1000003ull) return KPRN_OK;
    pack_wt_kernel<<<(unsigned)cdiv(elems / 2, 256), 256, 0, ctx->stream>>>(ctx->d_params + l.Wh, ctx->d_params + l.Wx, c.H, l.D, c.d_t, c.d_e,
                                                                           s->G4, s->Hp, s->KE, tt->WT);
    KPRN_LAUNCH_CHECK(ctx);
    tt->packed_small = ctx->small_version + ctx->params_version * 1000003ull;
    return KPRN_OK;
}

// buffers the TRAIN forward writes (called by tc_forward before the recurrence)
int tc_train_buffers(kprn_ctx* ctx, int64_t N, float** cache, bf16** hcache) {
    TcState* s = ctx->tc;
    TcTrain* tt = tt_get(ctx);
    const int T = ctx->cfg.T;
    const int64_t ptiles = cdiv(N, TILE_M);
    KPRN_TRY(ensure(tt->cache, sizeof(float) * (size_t)ptiles * T * s->NCH * 20480));
    KPRN_TRY(ensure(tt->hcache, (size_t)ptiles * T * 2 * TILE_M * s->Hp * 2));
    *cache = tt->cache.as<float>(); *hcache = tt->hcache.as<bf16>();
    return KPRN_OK;
}

int tc_backward(kprn_ctx* ctx, float scale) {
    TcState* s = ctx->tc;
    TcTrain* tt = tt_get(ctx);
    const kprn_config& c = ctx->cfg;
    const Layout& l = ctx->lay;
    const int T = c.T, G4 = s->G4, Hp = s->Hp, KE = s->KE;
    const int64_t N = ctx->N, ptiles = cdiv(N, TILE_M);
    float* gr = ctx->d_grads;
    const float* P = ctx->d_params;
    KPRN_TRY(tc_train_pack(ctx));
    if (!tt->side) {
        KPRN_CUDA(cudaStreamCreateWithFlags(&tt->side, cudaStreamNonBlocking));
        KPRN_CUDA(cudaStreamCreateWithFlags(&tt->side2, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&tt->ev_fork, &tt->ev_loop, &tt->ev_prep, &tt->ev_eg, &tt->ev_g0, &tt->ev_g1}) KPRN_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    { const char* d = getenv("KPRN_TC_GROUPS"); tt->use_groups = !(d && d[0] == '0'); }
    { const char* d = getenv("KPRN_TC_SIDE"); tt->use_side = !(d && d[0] == '0'); }
    cudaStream_t side = tt->use_side ? tt->side : ctx->stream;
    // ---- pool / head gradients on the fp32 kernels (kernels_simple.cu): ds_c[p], gbo[c], gWo[c,:] ----
    if (ctx->head_part4) {
        cls_logit_kernel<<<(unsigned)cdiv(N, 256), 256, 0, ctx->stream>>>(ctx->head_part4, ctx->head_npad, P + l.bo + (c.class_id - 1), N, c.L,
                                                                        c.class_id - 1, ctx->path_logits.as<float>());
        KPRN_LAUNCH_CHECK(ctx);
    }
    KPRN_TRY(head_backward_cls(ctx, scale, s->hT.as<float>(), c.H));
    const float* ds_c = ctx->ws_ds.as<float>();
    // ---- buffers ----
    const size_t dg_bytes = (size_t)ptiles * T * 2 * (G4 / 8) * 2048;
    KPRN_TRY(ensure(tt->dgop, dg_bytes));
    KPRN_TRY(ensure(tt->dh, sizeof(float) * (size_t)ptiles * Hp * TILE_M));
    KPRN_TRY(ensure(tt->dc, sizeof(float) * (size_t)ptiles * Hp * TILE_M));
    const int items = (G4 / 256) * 2;
    int splits = (s->num_sms / 2) / items;
    if (splits < 1) splits = 1;
    const int64_t nblocks = ptiles * T;
    if (splits > nblocks) splits = (int)nblocks;
    const int64_t bps = cdiv(nblocks, splits);
    splits = (int)cdiv(nblocks, bps);
    KPRN_TRY(ensure(tt->wpart, sizeof(float) * (size_t)splits * 2 * G4 * 320));
    KPRN_TRY(ensure(tt->gP, sizeof(float) * (size_t)(64 + c.V_t + c.V_r) * G4));
    // entity-row gradients: bgemm stages dX per path-step; the sorted segment reduce below adds them per distinct row into the
    // compact buffer of a sparse training step (slot-indexed) or into the dense table
    const int64_t n_order = ptiles * T * TILE_M;
    KPRN_TRY(ensure(tt->dX, sizeof(float) * (size_t)n_order * KE));
    const uint32_t* eg_pos; const uint32_t* eg_slot_of; const int32_t* eg_rows = nullptr; float* eg_target;
    if (ctx->sparse_gE) { eg_pos = ctx->sparse_pos; eg_slot_of = ctx->sparse_slot_of; eg_target = ctx->sparse_gE; }
    else {
        KPRN_TRY(dedup_build(ctx, tt->es, 1, ctx->d_ids, N, T, 0));
        eg_pos = tt->es.pos.as<uint32_t>(); eg_slot_of = tt->es.slot_of.as<uint32_t>(); eg_rows = tt->es.ids.as<int32_t>(); eg_target = gr + l.E_e;
    }
    // ---- side stream: the weight-gradient kernel's operands (needs only the forward's outputs) ----
    {
        const size_t xop_bytes = (size_t)nblocks * 2 * (KE / 8) * 2048, hop_bytes = (size_t)nblocks * 2 * (Hp / 8) * 2048, oh_bytes = (size_t)nblocks * 8 * 2048;
        KPRN_TRY(ensure(tt->xop, xop_bytes));
        KPRN_TRY(ensure(tt->hop, hop_bytes));
        KPRN_TRY(ensure(tt->ohop, oh_bytes));
        if (tt->use_side) { KPRN_CUDA(cudaEventRecord(tt->ev_fork, ctx->stream)); KPRN_CUDA(cudaStreamWaitEvent(side, tt->ev_fork, 0)); }
        prep_operands_kernel<<<(unsigned)nblocks, 128, 0, side>>>(ctx->d_ids, N, T, c.V_e, (int)c.V_t, (int)c.V_r, s->E_hi, s->E_lo, KE,
                                                                  tt->hcache.as<bf16>(), Hp, tt->xop.as<bf16>(), tt->hop.as<bf16>(),
                                                                  tt->ohop.as<bf16>());
        KPRN_LAUNCH_CHECK(ctx);
        if (tt->use_side) KPRN_CUDA(cudaEventRecord(tt->ev_prep, side));
    }
    const size_t bg_smem = BG_STAGES * (2 * 16384 + 2 * 8 * (Hp / 2) * 16 + 2 * 8 * (KE / 2) * 16) + 256;
    KPRN_CUDA(cudaFuncSetAttribute(bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bg_smem));
    bool fused = true;
    { const char* d = getenv("KPRN_TC_FUSED"); if (d && d[0] == '0') fused = false; }
    if (fused) {
        // the whole time loop in one persistent launch (bptt_kernel)
        BpArgs pa;
        pa.cache = tt->cache.as<float>(); pa.WT = tt->WT; pa.dgop = tt->dgop.as<bf16>(); pa.dh = tt->dh.as<float>(); pa.dc = tt->dc.as<float>();
        pa.dX = tt->dX.as<float>(); pa.ds_c = ds_c; pa.wo = P + l.Wo + (int64_t)(c.class_id - 1) * c.H;
        pa.T = T; pa.NCH = s->NCH; pa.H = c.H; pa.G4 = G4; pa.N0 = Hp; pa.N1 = KE; pa.num_tiles = ptiles; pa.N = N;
        KPRN_CUDA(cudaFuncSetAttribute(bptt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bg_smem));
        const int64_t n_tp = (ptiles + 1) / 2;
        const unsigned bp_grid = (unsigned)(2 * (n_tp < s->num_sms / 2 ? n_tp : s->num_sms / 2));
        bptt_kernel<<<bp_grid, BP_THREADS, bg_smem, ctx->stream>>>(pa);
        KPRN_LAUNCH_CHECK(ctx);
    }
    const int groups = fused ? 0 : (tt->use_groups && ptiles >= 8) ? 2 : 1;
    const int64_t half = groups == 2 ? (ptiles / 2 + 1) / 2 * 2 : ptiles;          // even: no duplicated tail tile inside a group
    cudaStream_t gstream[2] = {ctx->stream, tt->side2};
    if (groups == 2) { KPRN_CUDA(cudaEventRecord(tt->ev_g0, ctx->stream)); KPRN_CUDA(cudaStreamWaitEvent(tt->side2, tt->ev_g0, 0)); }
    for (int t = T - 1; t >= 0; --t) {
        for (int g = 0; g < groups; ++g) {
            const int64_t mb = g ? half : 0, cnt = groups == 2 ? (g ? ptiles - half : half) : ptiles;
            CellBwdArgs ca;
            ca.cache = tt->cache.as<float>(); ca.T = T; ca.t = t; ca.NCH = s->NCH; ca.H = c.H; ca.N = N;
            ca.dh = t == T - 1 ? nullptr : tt->dh.as<float>(); ca.ds_c = ds_c; ca.wo = P + l.Wo + (int64_t)(c.class_id - 1) * c.H;
            ca.dc = tt->dc.as<float>(); ca.dgop = tt->dgop.as<bf16>(); ca.G4 = G4; ca.m_begin = mb;
            cell_bwd_kernel<<<dim3((unsigned)cnt, (unsigned)s->NCH, 2), dim3(TILE_M, 2), 0, gstream[g]>>>(ca);
            KPRN_LAUNCH_CHECK(ctx);
            BgArgs ba;
            ba.dgop = tt->dgop.as<bf16>(); ba.WT = tt->WT; ba.T = T; ba.t = t; ba.G4 = G4; ba.N0 = Hp; ba.N1 = KE; ba.num_tiles = cnt; ba.N = N;
            ba.m_begin = mb;
            ba.part_begin = t == 0 ? 1 : 0; ba.dh = tt->dh.as<float>();
            ba.dX = tt->dX.as<float>();
            const int64_t n_tp = (cnt + 1) / 2;
            const unsigned bg_grid = (unsigned)(2 * (n_tp < s->num_sms / 2 ? n_tp : s->num_sms / 2));
            bgemm_kernel<<<bg_grid, BG_THREADS, bg_smem, gstream[g]>>>(ba);
            KPRN_LAUNCH_CHECK(ctx);
        }
    }
    if (groups == 2) { KPRN_CUDA(cudaEventRecord(tt->ev_g1, tt->side2)); KPRN_CUDA(cudaStreamWaitEvent(ctx->stream, tt->ev_g1, 0)); }
    // ---- entity-table gradient: sorted segment reduce of the staged dX rows (side stream, beside the weight-gradient kernel) ----
    {
        if (tt->use_side) { KPRN_CUDA(cudaEventRecord(tt->ev_loop, ctx->stream)); KPRN_CUDA(cudaStreamWaitEvent(side, tt->ev_loop, 0)); }
        int end_bit = 1;
        while ((1ll << end_bit) <= n_order) ++end_bit;
        KPRN_TRY(ensure(tt->eg_keys, sizeof(uint32_t) * 2 * (size_t)n_order));
        KPRN_TRY(ensure(tt->eg_vals, sizeof(uint32_t) * 2 * (size_t)n_order));
        uint32_t* k_in = tt->eg_keys.as<uint32_t>(); uint32_t* k_out = k_in + n_order;
        uint32_t* v_in = tt->eg_vals.as<uint32_t>(); uint32_t* v_out = v_in + n_order;
        eg_keys_kernel<<<(unsigned)cdiv(n_order, 256), 256, 0, side>>>(eg_pos, eg_slot_of, n_order, k_in, v_in);
        KPRN_LAUNCH_CHECK(ctx);
        size_t tmp_bytes = 0;
        KPRN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int)n_order, 0, end_bit, side));
        KPRN_TRY(ensure(tt->eg_tmp, tmp_bytes));
        KPRN_CUDA(cub::DeviceRadixSort::SortPairs(tt->eg_tmp.p, tmp_bytes, k_in, k_out, v_in, v_out, (int)n_order, 0, end_bit, side));
        ctx->launches += 4;
        EgArgs ea;
        ea.keys = k_out; ea.vals = v_out; ea.n = n_order; ea.n_chunks = cdiv(n_order, EG_CHUNK);
        ea.dX = tt->dX.as<float>(); ea.ldx = KE; ea.d_e = c.d_e; ea.target = eg_target; ea.rows = eg_rows; ea.assign = ctx->sparse_gE ? 1 : 0;
        KPRN_TRY(ensure(tt->eg_part, sizeof(float) * (size_t)ea.n_chunks * 2 * c.d_e));
        KPRN_TRY(ensure(tt->eg_flags, sizeof(int32_t) * (size_t)ea.n_chunks * 4));
        ea.part = tt->eg_part.as<float>(); ea.flags = tt->eg_flags.as<int32_t>();
        if (c.d_e == 256) {
            eg_reduce_kernel<2><<<(unsigned)cdiv(ea.n_chunks, 8), 256, 0, side>>>(ea);
            KPRN_LAUNCH_CHECK(ctx);
            eg_fixup_kernel<2><<<(unsigned)ea.n_chunks, 512, 0, side>>>(ea);
        } else {        // d_e = 128 (tc_train_supported admits KE of 128 or 256 with d_e a multiple of 128)
            eg_reduce_kernel<1><<<(unsigned)cdiv(ea.n_chunks, 8), 256, 0, side>>>(ea);
            KPRN_LAUNCH_CHECK(ctx);
            eg_fixup_kernel<1><<<(unsigned)ea.n_chunks, 512, 0, side>>>(ea);
        }
        KPRN_LAUNCH_CHECK(ctx);
        if (tt->use_side) KPRN_CUDA(cudaEventRecord(tt->ev_eg, side));
    }
    // ---- weight gradients ----
    {
        if (tt->use_side) KPRN_CUDA(cudaStreamWaitEvent(ctx->stream, tt->ev_prep, 0));
        WgArgs wa;
        wa.dgop = tt->dgop.as<bf16>(); wa.hop = tt->hop.as<bf16>(); wa.xop = tt->xop.as<bf16>(); wa.ohop = tt->ohop.as<bf16>();
        wa.dbg = nullptr;
        if (s->debug) {
            KPRN_TRY(ensure(tt->dbg, 8 * sizeof(unsigned long long)));
            KPRN_CUDA(cudaMemsetAsync(tt->dbg.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
            wa.dbg = tt->dbg.as<unsigned long long>();
        }
        wa.T = T; wa.num_tiles = ptiles; wa.G4 = G4; wa.N0 = Hp; wa.KE = KE; wa.blocks_per_split = bps; wa.part = tt->wpart.as<float>();
        const size_t wg_smem = (size_t)WG_STAGES * (2 * 16384 + 2 * 16384 + 4096) + 256;
        KPRN_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem));
        wgrad_kernel<<<dim3((unsigned)(items * 2), (unsigned)splits), WG_THREADS, wg_smem, ctx->stream>>>(wa);
        KPRN_LAUNCH_CHECK(ctx);
        if (wa.dbg) {
            unsigned long long h[8];
            KPRN_CUDA(cudaMemcpyAsync(h, wa.dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
            KPRN_CUDA(cudaStreamSynchronize(ctx->stream));
            const double np_ = (double)items * splits;
            fprintf(stderr, "[kprn tc debug] wgrad: issuer cycles/pair %.0f | waits/pair: issuer.full %.0f, issuer.peer %.0f, producer.empty %.0f\n",
                    h[4] / np_, h[0] / np_, h[1] / np_, h[3] / np_);
        }
    }
    float* gP = tt->gP.as<float>();
    float* Tsum = gP + (size_t)64 * G4;
    float* Rsum = Tsum + (size_t)c.V_t * G4;
    wgrad_reduce_kernel<<<(unsigned)cdiv((int64_t)2 * G4 * 320, 256), 256, 0, ctx->stream>>>(tt->wpart.as<float>(), splits, G4, c.H, l.D, c.d_t, c.d_e,
                                                                                          Hp, KE, gr + l.Wh, gr + l.Wx, gP);
    KPRN_LAUNCH_CHECK(ctx);
    ptr_sums_kernel<<<(unsigned)cdiv(G4, 128), 128, 0, ctx->stream>>>(gP, G4, c.H, (int)c.V_t, (int)c.V_r, Tsum, Rsum, gr + l.bx);
    KPRN_LAUNCH_CHECK(ctx);
    {
        const int wmax = c.d_t > c.d_r ? c.d_t : c.d_r;
        ptr_embed_kernel<<<dim3((unsigned)(c.V_t + c.V_r), (unsigned)cdiv(wmax, 64)), dim3(64, 8), 0, ctx->stream>>>(
            Tsum, Rsum, P + l.Wx, G4, c.H, l.D, c.d_t, c.d_e, c.d_r, (int)c.V_t, gr + l.E_t, gr + l.E_r);
    }
    KPRN_LAUNCH_CHECK(ctx);
    ptr_wx_kernel<<<(unsigned)G4, 256, 0, ctx->stream>>>(Tsum, Rsum, P + l.E_t, P + l.E_r, G4, c.H, l.D, c.d_t, c.d_e, c.d_r, (int)c.V_t, (int)c.V_r,
                                                        gr + l.Wx);
    KPRN_LAUNCH_CHECK(ctx);
    if (tt->use_side) {
        if (ctx->defer_eg_join) { ctx->eg_stream = side; ctx->eg_event = tt->ev_eg; }      // the caller continues on the side stream, then joins
        else KPRN_CUDA(cudaStreamWaitEvent(ctx->stream, tt->ev_eg, 0));                    // join: everything after sees the entity gradient
    }
    return KPRN_OK;
}

void tc_train_destroy(kprn_ctx* ctx) {
    TcState* s = ctx->tc;
    if (!s || !s->tt) return;
    TcTrain* tt = s->tt;
    if (tt->WT) cudaFree(tt->WT);
    if (tt->side) {
        cudaStreamSynchronize(tt->side);
        if (tt->side2) { cudaStreamSynchronize(tt->side2); cudaStreamDestroy(tt->side2); }
        for (cudaEvent_t e : {tt->ev_fork, tt->ev_loop, tt->ev_prep, tt->ev_eg, tt->ev_g0, tt->ev_g1}) if (e) cudaEventDestroy(e);
        cudaStreamDestroy(tt->side);
    }
    for (DBuf* b : {&tt->cache, &tt->hcache, &tt->dgop, &tt->dh, &tt->dc, &tt->wpart, &tt->gP, &tt->cls, &tt->dX, &tt->eg_keys, &tt->eg_vals,
                    &tt->eg_tmp, &tt->eg_part, &tt->eg_flags, &tt->xop, &tt->hop, &tt->ohop, &tt->dbg})
        if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    dedup_free(tt->es);
    delete tt;
    s->tt = nullptr;
}

}  // namespace kprn
