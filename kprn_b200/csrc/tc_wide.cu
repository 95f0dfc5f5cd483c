// tc_wide.cu -- the recurrence for hidden sizes the persistent recurrence cannot hold on one SM (256 < H <= 512: the cell
// state of a 128-path tile alone would fill TMEM, h_t alone the shared memory).  BASELINE configs[4], d = 512.
// One launch per time step: state (h_t as bf16 hi/lo in the UMMA operand layout, c_t fp32) makes one trip through HBM per
// step (16 B per unit and path: 0.5 GB per step at N = 65 536, a tenth of the step's MMA time), and the kernel is a CTA-pair
// GEMM  G[256 paths, 4Hp] = h_{t-1} Wh^T  (K = Hp, bf16x3, cta_group::2) walked over 256-column blocks (two chunks of 32
// hidden units x 4 gates) with two TMEM accumulators, so the FastLSTM cell epilogue of block b (the same packed-f32x2 math as
// recur_kernel) overlaps the MMAs of block b+1.  The input projection XW comes from proj3_kernel as usual.
// Math: SURVEY.md App. A.3 / SP/model/OneModel.lua:236,268-273.
#include "tc_state.cuh"

namespace kprn {

constexpr int WS_THREADS = 384;     // warp 0 TMA producer, 1 MMA issuer / peer forwarder, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int WS_STAGES = 3;
constexpr uint32_t WS_A = 16384, WS_B = 16384, WS_STAGE = 2 * WS_A + 2 * WS_B;

struct WideArgs {
    const float* XW; const bf16* W;            // W: [blk G4/256][ks Hp/64][hi|lo][half 2][kc 8][n 128][8], gate-scaled
    const bf16* h_in; bf16* h_out;             // [tile][hi|lo][kc Hp/8][row 128][8]
    float* c;                                  // [tile][Hp/8][row 128][8], in place
    int64_t N; int T; int t; int H; int Hp; int G4; int64_t num_tiles;
    const uint32_t* pos; const uint32_t* slot_of;
    const float* wo; float* part4; int64_t npad; float* hT;      // last step: fused head partials (scores-only) and / or h_T
};

__device__ __forceinline__ float wex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float wrcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WS_THREADS, 1) wide_step_kernel(WideArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    constexpr int S = WS_STAGES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * WS_STAGE);
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    // [0,S) full  [S,2S) empty  [2S,3S) peer_full (leader)  acc_full[2]  acc_empty[2] (leader)
    constexpr int B_PEER = 2 * S, B_ACCF = 3 * S, B_ACCE = 3 * S + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACCE + 2);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(S + i), 1); ptx::mbar_init(BAR(B_PEER + i), 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(B_ACCF + i), 1); ptx::mbar_init(BAR(B_ACCE + i), 16); }
        ptx::fence_mbar_init();
    }
    if (warp == 2) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t npairs = gridDim.x / 2, pair = blockIdx.x / 2, n_tp = (a.num_tiles + 1) / 2;
    const int64_t my_tiles = pair < n_tp ? (n_tp - pair + npairs - 1) / npairs : 0;
    auto tile_of = [&](int64_t it) -> int64_t { const int64_t m = 2 * (pair + it * npairs) + rank; return m < a.num_tiles ? m : a.num_tiles - 1; };
    const int NB = a.G4 / 256, KS = a.Hp / 64;
    const bool has_mma = a.t > 0;
    const size_t hl = (size_t)(a.Hp / 8) * TILE_M * 8;                   // elements of one of hi / lo of an h tile

    if (warp == 0) {
        if (lane == 0 && has_mma) {
            uint32_t stage = 0, phase = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const bf16* A = a.h_in + (size_t)tile_of(it) * 2 * hl;
                for (int blk = 0; blk < NB; ++blk)
                    for (int ks = 0; ks < KS; ++ks) {
                        ptx::mbar_wait(BAR(S + stage), phase ^ 1);
                        ptx::mbar_arrive_expect_tx(BAR(stage), WS_STAGE);
                        const uint32_t sb0 = ptx::smem_u32(smem + stage * WS_STAGE);
                        ptx::tma_bulk_g2s(sb0, A + (size_t)ks * 8 * TILE_M * 8, WS_A, BAR(stage));
                        ptx::tma_bulk_g2s(sb0 + WS_A, A + hl + (size_t)ks * 8 * TILE_M * 8, WS_A, BAR(stage));
                        const bf16* Wb = a.W + ((((size_t)blk * KS + ks) * 2 + 0) * 2 + rank) * (8 * 128 * 8);
                        ptx::tma_bulk_g2s(sb0 + 2 * WS_A, Wb, WS_B, BAR(stage));
                        ptx::tma_bulk_g2s(sb0 + 2 * WS_A + WS_B, Wb + (size_t)2 * 8 * 128 * 8, WS_B, BAR(stage));
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (!has_mma) {
            // nothing to issue at t = 0 (h_{-1} = 0)
        } else if (leader) {
            const bool issue = ptx::elect_one();
            uint32_t stage = 0, phase = 0, ug[2] = {0, 0};
            constexpr uint32_t DHI = ptx::desc_hi(128);
            const uint32_t idesc = ptx::make_idesc_f16(1, 256, 256);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int blk = 0; blk < NB; ++blk) {
                    const int g = blk & 1;
                    ptx::mbar_wait_cluster(BAR(B_ACCE + g), (ug[g] & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d = tmem + g * 256;
                    for (int ks = 0; ks < KS; ++ks) {
                        ptx::mbar_wait(BAR(stage), phase);
                        ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase);
                        ptx::tc_fence_after();
                        const uint32_t sb0 = ptx::smem_u32(smem + stage * WS_STAGE);
                        const uint32_t ah = ptx::desc_lo(sb0, 2048), al = ptx::desc_lo(sb0 + WS_A, 2048);
                        const uint32_t bh = ptx::desc_lo(sb0 + 2 * WS_A, 2048), bl = ptx::desc_lo(sb0 + 2 * WS_A + WS_B, 2048);
                        if (issue) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) ptx::umma2_f16(d, ptx::desc_join(DHI, ah + k * 256u), ptx::desc_join(DHI, bh + k * 256u), idesc, (ks | k) != 0);
#pragma unroll
                            for (int k = 0; k < 4; ++k) ptx::umma2_f16(d, ptx::desc_join(DHI, al + k * 256u), ptx::desc_join(DHI, bh + k * 256u), idesc, 1);
#pragma unroll
                            for (int k = 0; k < 4; ++k) ptx::umma2_f16(d, ptx::desc_join(DHI, ah + k * 256u), ptx::desc_join(DHI, bl + k * 256u), idesc, 1);
                            ptx::umma2_commit(BAR(S + stage), 3);
                        }
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                    if (issue) ptx::umma2_commit(BAR(B_ACCF + g), 3);
                    ++ug[g];
                }
        } else if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t remote = ptx::mapa(BAR(B_PEER), 0);
            for (int64_t it = 0; it < my_tiles; ++it)
                for (int i = 0; i < NB * KS; ++i) {
                    ptx::mbar_wait(BAR(stage), phase);
                    ptx::mbar_arrive_cluster(remote + 8u * stage);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp >= 4) {
        // ===== epilogue: lane quarter q4, sub-blocks {2 hf, 2 hf + 1} of both chunks of every 256-column block =====
        const int q4 = warp & 3, hf = (warp - 4) >> 2, r = q4 * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const uint32_t acce_leader = ptx::mapa(BAR(B_ACCE), 0);
        const int64_t tile_xw = (int64_t)(a.G4 / 8) * TILE_M * 8;
        const bool last = a.t == a.T - 1;
        const bool h_vec4 = (a.H & 3) == 0;
        uint32_t ug[2] = {0, 0};
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t m = tile_of(it);
            // an odd tail tile runs on both CTAs of the last pair (the MMA needs both): the duplicate takes part in the MMAs and the
            // barriers but writes nothing -- c is updated in place, two writers would race
            const bool dup = (2 * (pair + it * npairs) + rank) >= a.num_tiles;
            const int64_t n = dup ? a.N : m * TILE_M + r;
            uint32_t slot = 0;
            if (a.pos) { const uint32_t p = a.pos[(m * a.T + a.t) * TILE_M + r]; slot = p == 0xFFFFFFFFu ? 0u : a.slot_of[p]; }
            const float* xw_t = a.pos ? a.XW + (int64_t)(slot >> 7) * tile_xw + (slot & 127) * 8 : a.XW + (m * a.T + a.t) * tile_xw + r * 8;
            float dot[2] = {0.f, 0.f};
            for (int blk = 0; blk < NB; ++blk) {
                const int g = blk & 1;
                if (has_mma) { ptx::mbar_wait(BAR(B_ACCF + g), ug[g] & 1); ptx::tc_fence_after(); }
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = blk * 2 + jj;
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const int sb = hf * 2 + s2;
                        float xc[32], cp[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float* px = xw_t + (int64_t)(j * 16 + q * 4 + sb) * TILE_M * 8;
                            const float4 x0 = *reinterpret_cast<const float4*>(px), x1 = *reinterpret_cast<const float4*>(px + 4);
                            xc[q * 8 + 0] = x0.x; xc[q * 8 + 1] = x0.y; xc[q * 8 + 2] = x0.z; xc[q * 8 + 3] = x0.w;
                            xc[q * 8 + 4] = x1.x; xc[q * 8 + 5] = x1.y; xc[q * 8 + 6] = x1.z; xc[q * 8 + 7] = x1.w;
                        }
                        float* cptr = a.c + ((size_t)(m * (a.Hp / 8) + j * 4 + sb) * TILE_M + r) * 8;
                        if (has_mma) {
                            float ai[8], ag[8], af[8], ao[8];
                            const uint32_t acc = tmem + lane_base + g * 256 + jj * 128 + sb * 8;
                            ptx::tmem_ld8(acc, ai); ptx::tmem_ld8(acc + 32, ag); ptx::tmem_ld8(acc + 64, af); ptx::tmem_ld8(acc + 96, ao);
                            const float4 c0 = *reinterpret_cast<const float4*>(cptr), c1 = *reinterpret_cast<const float4*>(cptr + 4);
                            cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w; cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
                            ptx::tc_wait_ld();
#pragma unroll
                            for (int e = 0; e < 8; ++e) { xc[e] += ai[e]; xc[8 + e] += ag[e]; xc[16 + e] += af[e]; xc[24 + e] += ao[e]; }
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) cp[e] = 0.f;
                        }
                        float cn[8], hv[8];
                        {
                            const uint64_t one2 = ptx::pack2(1.f, 1.f), mtwo2 = ptx::pack2(-2.f, -2.f);
                            const uint64_t k2 = ptx::pack2(2.8853900817779268f, 2.8853900817779268f);
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const int e = 2 * p;
                                const uint64_t pi = ptx::add2(ptx::pack2(wex2f(xc[e]), wex2f(xc[e + 1])), one2);
                                const uint64_t pg = ptx::add2(ptx::pack2(wex2f(xc[8 + e]), wex2f(xc[9 + e])), one2);
                                const uint64_t pf = ptx::add2(ptx::pack2(wex2f(xc[16 + e]), wex2f(xc[17 + e])), one2);
                                const uint64_t po = ptx::add2(ptx::pack2(wex2f(xc[24 + e]), wex2f(xc[25 + e])), one2);
                                const uint64_t i_ = ptx::pack2(wrcpf(ptx::lo2(pi)), wrcpf(ptx::hi2(pi)));
                                const uint64_t f_ = ptx::pack2(wrcpf(ptx::lo2(pf)), wrcpf(ptx::hi2(pf)));
                                const uint64_t o_ = ptx::pack2(wrcpf(ptx::lo2(po)), wrcpf(ptx::hi2(po)));
                                const uint64_t g_ = ptx::fma2(ptx::pack2(wrcpf(ptx::lo2(pg)), wrcpf(ptx::hi2(pg))), mtwo2, one2);
                                const uint64_t c2 = ptx::fma2(i_, g_, ptx::mul2(f_, ptx::pack2(cp[e], cp[e + 1])));
                                const uint64_t a2 = ptx::mul2(c2, k2);
                                const uint64_t pt2 = ptx::add2(ptx::pack2(wex2f(ptx::lo2(a2)), wex2f(ptx::hi2(a2))), one2);
                                const uint64_t th = ptx::fma2(ptx::pack2(wrcpf(ptx::lo2(pt2)), wrcpf(ptx::hi2(pt2))), mtwo2, one2);
                                const uint64_t h2 = ptx::mul2(o_, th);
                                cn[e] = ptx::lo2(c2); cn[e + 1] = ptx::hi2(c2);
                                hv[e] = ptx::lo2(h2); hv[e + 1] = ptx::hi2(h2);
                            }
                        }
                        if (dup) {
                            // nothing to store
                        } else if (!last) {
                            *reinterpret_cast<float4*>(cptr) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                            *reinterpret_cast<float4*>(cptr + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                            uint32_t wh[4], wl[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const __nv_bfloat162 h2b = __floats2bfloat162_rn(hv[2 * i], hv[2 * i + 1]);
                                const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2b);
                                const float r0 = hv[2 * i] - __uint_as_float(hb << 16), r1 = hv[2 * i + 1] - __uint_as_float(hb & 0xFFFF0000u);
                                const __nv_bfloat162 l2b = __floats2bfloat162_rn(r0, r1);
                                wh[i] = hb; wl[i] = *reinterpret_cast<const uint32_t*>(&l2b);
                            }
                            bf16* hd = a.h_out + (size_t)m * 2 * hl + ((size_t)(j * 4 + sb) * TILE_M + r) * 8;
                            *reinterpret_cast<uint4*>(hd) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                            *reinterpret_cast<uint4*>(hd + hl) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                        } else {
                            const int u0 = j * 32 + sb * 8;
                            if (a.part4) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) if (u0 + e < a.H) dot[s2] = fmaf(hv[e], a.wo[u0 + e], dot[s2]);
                            }
                            if (a.hT && n < a.N) {
                                float* dst = a.hT + n * a.H + u0;
                                if (h_vec4 && u0 + 8 <= a.H) {
                                    *reinterpret_cast<float4*>(dst) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                                    *reinterpret_cast<float4*>(dst + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) if (u0 + e < a.H) dst[e] = hv[e];
                                }
                            }
                        }
                    }
                }
                if (has_mma) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(acce_leader + 8u * g);
                    ++ug[g];
                }
            }
            if (last && a.part4 && n < a.N) { a.part4[(size_t)(hf * 2) * a.npad + n] = dot[0]; a.part4[(size_t)(hf * 2 + 1) * a.npad + n] = dot[1]; }
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem, 512);
}

// recurrent weights for wide_step_kernel: [blk G4/256][ks Hp/64][hi|lo][half 2][kc 8][n 128][8], gate-scaled like pack_wh_kernel
__global__ void pack_wh_wide_kernel(const float* __restrict__ Wh, int H, int Hp, int G4, bf16* __restrict__ out) {
    const int KS = Hp / 64;
    const int64_t total = (int64_t)(G4 / 256) * KS * 2 * 8 * 128 * 8;       // (hi, lo) pairs
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int e = (int)(idx & 7);
    const int n = (int)((idx >> 3) & 127);
    const int kc = (int)((idx >> 10) & 7);
    const int half = (int)((idx >> 13) & 1);
    const int64_t r = idx >> 14;
    const int ks = (int)(r % KS), blk = (int)(r / KS);
    const int gp = blk * 256 + half * 128 + n, k = ks * 64 + kc * 8 + e;
    int u;
    const int row = perm_to_row(gp, H, u);
    const int q = (gp >> 5) & 3;
    const float sc = q == 1 ? 2.8853900817779268f : -1.4426950408889634f;
    const float v = (u < H && k < H) ? Wh[(int64_t)row * H + k] * sc : 0.f;
    bf16 h, l;
    split_bf16(v, h, l);
    bf16* base = out + ((((int64_t)blk * KS + ks) * 2 + 0) * 2 + half) * (8 * 128 * 8) + ((int64_t)kc * 128 + n) * 8 + e;
    base[0] = h;
    base[(int64_t)2 * 8 * 128 * 8] = l;
}

// host: T launches of wide_step_kernel over ping-pong h buffers
int wide_recurrence(kprn_ctx* ctx, const RecurWide& w) {
    WideArgs a;
    a.XW = w.XW; a.W = w.W; a.c = w.c; a.N = w.N; a.T = w.T; a.H = w.H; a.Hp = w.Hp; a.G4 = 4 * w.Hp; a.num_tiles = cdiv(w.N, TILE_M);
    a.pos = w.pos; a.slot_of = w.slot_of; a.wo = w.wo; a.part4 = w.part4; a.npad = w.npad; a.hT = w.hT;
    const size_t smem = (size_t)WS_STAGES * WS_STAGE + 256;
    KPRN_CUDA(cudaFuncSetAttribute(wide_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_tp = (a.num_tiles + 1) / 2;
    const unsigned grid = (unsigned)(2 * (n_tp < w.num_sms / 2 ? n_tp : w.num_sms / 2));
    for (int t = 0; t < w.T; ++t) {
        a.t = t;
        a.h_in = (t & 1) ? w.h0 : w.h1;          // step t reads what step t-1 wrote
        a.h_out = (t & 1) ? w.h1 : w.h0;
        wide_step_kernel<<<grid, WS_THREADS, smem, ctx->stream>>>(a);
        KPRN_LAUNCH_CHECK(ctx);
    }
    return KPRN_OK;
}

int pack_wh_wide(kprn_ctx* ctx, const float* Wh, int H, int Hp, bf16* out) {
    const int64_t total = (int64_t)(4 * Hp / 256) * (Hp / 64) * 2 * 8 * 128 * 8;
    pack_wh_wide_kernel<<<(unsigned)cdiv(total, 256), 256, 0, ctx->stream>>>(Wh, H, Hp, 4 * Hp, out);
    KPRN_LAUNCH_CHECK(ctx);
    return KPRN_OK;
}

}  // namespace kprn
