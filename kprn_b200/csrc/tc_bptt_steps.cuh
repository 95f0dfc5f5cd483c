// tc_bptt_steps.cuh -- the per-step form of the backward's time loop: cell_bwd_kernel + bgemm_kernel, one pair of launches per
// time step (two tile groups on two streams).  Superseded on the hot path by bptt_kernel (tc_train.cu), which runs the same
// arithmetic in one persistent launch; kept, under KPRN_TC_FUSED=0, as the cross-check that tests/test_gpu_tc.py holds
// bptt_kernel to.  Included by tc_train.cu only (uses its constants and ptx.cuh wrappers).
#pragma once

// ---------------------------------------------------------------------------------------------
// cell backward: one thread = one path row x 8 hidden units (the recurrence epilogue's mapping).
// grid (tiles, NCH), block (128 rows, 4 sub-blocks)

struct CellBwdArgs {
    const float* cache; int T; int t; int NCH; int H; int64_t N;
    const float* dh;          // [tile][Hp/8][row 128][8] from bgemm (t < T-1) or null (t = T-1: dh = ds_c[n] * wo[u])
    const float* ds_c; const float* wo;
    float* dc;                // [tile][Hp/8][row 128][8] in/out (first step: treated as 0)
    bf16* dgop;               // [tile][t][hi|lo][kc G4/8][row 128][8]
    int G4;
    int64_t m_begin;          // first tile of this launch (the backward walks the batch in tile groups on two streams)
};

__device__ __forceinline__ void split8_store(const float* v, bf16* hi, bf16* lo) {
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const float r0 = v[2 * i] - __uint_as_float(hb << 16), r1 = v[2 * i + 1] - __uint_as_float(hb & 0xFFFF0000u);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        wh[i] = hb; wl[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    *reinterpret_cast<uint4*>(lo) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

// grid (tiles of the group, NCH, 2), block (128 rows, 2 sub-blocks): 256-thread blocks of ~21 k registers, so that they fit on an SM
// beside a resident bgemm CTA of the other tile group (see tc_backward)
__global__ void __launch_bounds__(256) cell_bwd_kernel(CellBwdArgs a) {
    const int64_t m = a.m_begin + blockIdx.x;
    const int j = blockIdx.y, r = threadIdx.x, sb = blockIdx.z * 2 + threadIdx.y;
    const int64_t n = m * TILE_M + r;
    const bool live = n < a.N;
    const size_t step_floats = (size_t)a.NCH * 20480;
    const float* ct = a.cache + (size_t)(m * a.T + a.t) * step_floats + (size_t)(j * 4 + sb) * 5120 + r * 4;
    const float* cprev = a.t > 0 ? ct - step_floats + 4096 : nullptr;
    float gi[8], gg[8], gf[8], go[8], cc[8], cp[8];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {            // cache layout [i,g,f,o,c][half 2][row 128][4]: 128-bit loads, 512 B per warp
        const float4 vi = *reinterpret_cast<const float4*>(ct + hf * 512), vg = *reinterpret_cast<const float4*>(ct + 1024 + hf * 512);
        const float4 vf = *reinterpret_cast<const float4*>(ct + 2048 + hf * 512), vo = *reinterpret_cast<const float4*>(ct + 3072 + hf * 512);
        const float4 vc = *reinterpret_cast<const float4*>(ct + 4096 + hf * 512);
        float4 vp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cprev) vp = *reinterpret_cast<const float4*>(cprev + hf * 512);
        const int b = 4 * hf;
        gi[b] = vi.x; gi[b + 1] = vi.y; gi[b + 2] = vi.z; gi[b + 3] = vi.w; gg[b] = vg.x; gg[b + 1] = vg.y; gg[b + 2] = vg.z; gg[b + 3] = vg.w;
        gf[b] = vf.x; gf[b + 1] = vf.y; gf[b + 2] = vf.z; gf[b + 3] = vf.w; go[b] = vo.x; go[b + 1] = vo.y; go[b + 2] = vo.z; go[b + 3] = vo.w;
        cc[b] = vc.x; cc[b + 1] = vc.y; cc[b + 2] = vc.z; cc[b + 3] = vc.w; cp[b] = vp.x; cp[b + 1] = vp.y; cp[b + 2] = vp.z; cp[b + 3] = vp.w;
    }
    const int grp = j * 4 + sb;                                 // 8-unit group of this thread
    const size_t gi8 = ((size_t)(m * (a.G4 / 32) + grp) * TILE_M + r) * 8;      // G4/32 = Hp/8 groups per tile
    float dh[8], dc[8];
    if (a.dh) {
        const float4 x0 = *reinterpret_cast<const float4*>(a.dh + gi8), x1 = *reinterpret_cast<const float4*>(a.dh + gi8 + 4);
        dh[0] = x0.x; dh[1] = x0.y; dh[2] = x0.z; dh[3] = x0.w; dh[4] = x1.x; dh[5] = x1.y; dh[6] = x1.z; dh[7] = x1.w;
        const float4 y0 = *reinterpret_cast<const float4*>(a.dc + gi8), y1 = *reinterpret_cast<const float4*>(a.dc + gi8 + 4);
        dc[0] = y0.x; dc[1] = y0.y; dc[2] = y0.z; dc[3] = y0.w; dc[4] = y1.x; dc[5] = y1.y; dc[6] = y1.z; dc[7] = y1.w;
    } else {
        // t = T-1: dh = ds_c * Wo[classId, :]   (nn.Linear backward for the one live output column), dc = 0
        const float s = live ? a.ds_c[n] : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { const int u = grp * 8 + e; dh[e] = u < a.H ? s * a.wo[u] : 0.f; dc[e] = 0.f; }
    }
    float di[8], dg[8], df[8], dO[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float tc = tanhf(cc[e]);
        const float d_o = dh[e] * tc;
        const float dcu = dc[e] + dh[e] * go[e] * (1.f - tc * tc);
        dc[e] = dcu * gf[e];
        di[e] = live ? dcu * gg[e] * gi[e] * (1.f - gi[e]) : 0.f;
        dg[e] = live ? dcu * gi[e] * (1.f - gg[e] * gg[e]) : 0.f;
        df[e] = live ? dcu * cp[e] * gf[e] * (1.f - gf[e]) : 0.f;
        dO[e] = live ? d_o * go[e] * (1.f - go[e]) : 0.f;
    }
    *reinterpret_cast<float4*>(a.dc + gi8) = make_float4(dc[0], dc[1], dc[2], dc[3]);
    *reinterpret_cast<float4*>(a.dc + gi8 + 4) = make_float4(dc[4], dc[5], dc[6], dc[7]);
    // operand layout: k-chunk kc = j*16 + gate*4 + sb (8 permuted gate columns), row r
    const size_t hl = (size_t)(a.G4 / 8) * TILE_M * 8;                       // elements of one of hi / lo
    bf16* base = a.dgop + (size_t)(m * a.T + a.t) * 2 * hl + ((size_t)(j * 16 + sb) * TILE_M + r) * 8;
    const size_t gstep = (size_t)4 * TILE_M * 8;                              // next gate = 4 k-chunks further
    split8_store(di, base, base + hl);
    split8_store(dg, base + gstep, base + gstep + hl);
    split8_store(df, base + 2 * gstep, base + 2 * gstep + hl);
    split8_store(dO, base + 3 * gstep, base + 3 * gstep + hl);
}

// ---------------------------------------------------------------------------------------------
// bgemm: [dh | dX] (128 paths per CTA, 256 per pair) = dG_t (K = G4) x [Wh^T | Wx_e^T], bf16x3, cta_group::2.

struct BgArgs {
    const bf16* dgop; const bf16* WT; int T; int t; int G4; int N0, N1; int64_t num_tiles; int64_t N;
    int64_t m_begin;          // tiles [m_begin, m_begin + num_tiles) of the batch
    int part_begin;           // 1 at t = 0 (no dh needed)
    float* dh;                // [tile][N0/8][row 128][8]
    float* dX;                // [(tile*T + t)*128 + row][N1] fp32: the entity-slice input gradient of every path-step (row index =
                              // the read-order index o of the slot maps); summed per entity row by the sorted segment reduce below
};

// __maxnreg__(112): 384 x 112 registers leave room on the SM for a 256-thread cell_bwd block of the other tile group
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(112) bgemm_kernel(BgArgs a) {
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
    // [0,S) full  [S,2S) empty  [2S,3S) peer_full (leader)  3S acc_full  3S+1 acc_empty (leader)
    constexpr int B_PEER = 2 * S, B_ACCF = 3 * S, B_ACCE = 3 * S + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACCE + 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(S + i), 1); ptx::mbar_init(BAR(B_PEER + i), 1); }
        ptx::mbar_init(BAR(B_ACCF), 1); ptx::mbar_init(BAR(B_ACCE), 16);
        ptx::fence_mbar_init();
    }
    if (warp == 2) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t npairs = gridDim.x / 2, pair = blockIdx.x / 2, n_tp = (a.num_tiles + 1) / 2;
    const int64_t my_tiles = pair < n_tp ? (n_tp - pair + npairs - 1) / npairs : 0;
    auto tile_of = [&](int64_t it) -> int64_t { const int64_t m = 2 * (pair + it * npairs) + rank; return a.m_begin + (m < a.num_tiles ? m : a.num_tiles - 1); };
    const int KSg = a.G4 / 64;
    const size_t hl = (size_t)(a.G4 / 8) * TILE_M * 8;               // elements of one of hi / lo of a dgop tile
    const int64_t per0 = (int64_t)KSg * 2 * 8 * a.N0 * 8;            // elements of part 0 of WT

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t it = 0; it < my_tiles; ++it) {
                const bf16* A = a.dgop + (size_t)(tile_of(it) * a.T + a.t) * 2 * hl;
                for (int ks = 0; ks < KSg; ++ks) {
                    ptx::mbar_wait(BAR(S + stage), phase ^ 1);
                    const uint32_t tx = 2 * A_BYTES + (a.part_begin == 0 ? 2 * B0_BYTES : 0) + 2 * B1_BYTES;
                    ptx::mbar_arrive_expect_tx(BAR(stage), tx);
                    const uint32_t sb0 = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    ptx::tma_bulk_g2s(sb0, A + (size_t)ks * 8 * TILE_M * 8, A_BYTES, BAR(stage));
                    ptx::tma_bulk_g2s(sb0 + A_BYTES, A + hl + (size_t)ks * 8 * TILE_M * 8, A_BYTES, BAR(stage));
                    if (a.part_begin == 0) {
                        const bf16* W0 = a.WT + (((int64_t)ks * 2 + 0) * 2 + rank) * (8 * Nh0 * 8);
                        ptx::tma_bulk_g2s(sb0 + 2 * A_BYTES, W0, B0_BYTES, BAR(stage));
                        ptx::tma_bulk_g2s(sb0 + 2 * A_BYTES + B0_BYTES, W0 + (int64_t)2 * 8 * Nh0 * 8, B0_BYTES, BAR(stage));
                    }
                    const bf16* W1 = a.WT + per0 + (((int64_t)ks * 2 + 0) * 2 + rank) * (8 * Nh1 * 8);
                    ptx::tma_bulk_g2s(sb0 + 2 * A_BYTES + 2 * B0_BYTES, W1, B1_BYTES, BAR(stage));
                    ptx::tma_bulk_g2s(sb0 + 2 * A_BYTES + 2 * B0_BYTES + B1_BYTES, W1 + (int64_t)2 * 8 * Nh1 * 8, B1_BYTES, BAR(stage));
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            const bool issue = ptx::elect_one();
            uint32_t stage = 0, phase = 0;
            constexpr uint32_t DHI = ptx::desc_hi(128);
            const uint32_t idesc0 = ptx::make_idesc_f16(1, 256, a.N0), idesc1 = ptx::make_idesc_f16(1, 256, a.N1);
            for (int64_t it = 0; it < my_tiles; ++it) {
                ptx::mbar_wait_cluster(BAR(B_ACCE), (it & 1) ^ 1);
                ptx::tc_fence_after();
                for (int ks = 0; ks < KSg; ++ks) {
                    ptx::mbar_wait(BAR(stage), phase);
                    ptx::mbar_wait_cluster(BAR(B_PEER + stage), phase);
                    ptx::tc_fence_after();
                    const uint32_t sb0 = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t ah = ptx::desc_lo(sb0, 2048), al = ptx::desc_lo(sb0 + A_BYTES, 2048);
                    if (issue) {
                        for (int part = a.part_begin; part < 2; ++part) {
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
                for (int ks = 0; ks < KSg; ++ks) {
                    ptx::mbar_wait(BAR(stage), phase);
                    ptx::mbar_arrive_cluster(remote + 8u * stage);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp >= 4) {
        // ===== epilogue: warps 4-7 drain part 0 (dh), warps 8-11 part 1 (dX); a warp owns the 32 rows of its TMEM lane quarter =====
        const int part = (warp - 4) >> 2, q4 = warp & 3, r = q4 * 32 + lane;
        const uint32_t acce_leader = ptx::mapa(BAR(B_ACCE), 0);
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t m = tile_of(it);
            ptx::mbar_wait(BAR(B_ACCF), it & 1);
            ptx::tc_fence_after();
            const uint32_t tb = tmem + ((uint32_t)(q4 * 32) << 16) + part * 256;
            if (part == 0) {
                if (a.part_begin == 0) {
                    float* dst = a.dh + ((size_t)m * (a.N0 / 8) * TILE_M + r) * 8;
                    for (int blk = 0; blk < a.N0 / 32; ++blk) {
                        float v[32];
                        ptx::tmem_ld32(tb + blk * 32, v);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 4; ++i) ptx::stg256(dst + (size_t)(blk * 4 + i) * TILE_M * 8, v + 8 * i);
                    }
                }
            } else {
                float* dst = a.dX + ((size_t)(m * a.T + a.t) * TILE_M + r) * a.N1;
                for (int blk = 0; blk < a.N1 / 32; ++blk) {
                    float v[32];
                    ptx::tmem_ld32(tb + blk * 32, v);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 4; ++i) ptx::stg256(dst + blk * 32 + 8 * i, v + 8 * i);     // one full 32-byte sector per store
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(acce_leader);
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2(tmem, 512);
}

