// tc_gemm.cu -- generic-stride GEMM on tcgen05 tensor cores with the bf16x3 split, used by the training path
// (forward with caches and every backward product) when the engine runs in a tensor-core precision mode.
//
//   C[m,n] (=|+=) sum_k A(m,k) * B(k,n),   A(m,k) = A[m*sAm + k*sAk],  B(k,n) = B[k*sBk + n*sBn]   (fp32 in HBM)
//
// Same contract as the FFMA kernel in sgemm.cuh (y = x W^T, dx = dy W, split-K gW += dy^T x).  The operands are
// converted on the fly: 8 producer warps read fp32 tiles with whatever stride is contiguous (K-contiguous or
// M/N-contiguous sources), split every value into bf16 hi + lo and store both in the canonical K-major UMMA layout
// [k/8][row][8]; one warp issues hi*hi + lo*hi + hi*lo tcgen05.mma passes (M=128, N=256, K=16) into a TMEM
// accumulator; the producer warps then drain TMEM to C (plain store, read-modify-write, or float atomics for split-K).
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace kprn {

using bf16 = __nv_bfloat16;

constexpr int GM = 128, GN = 256, GK = 64, G_STAGES = 2;
constexpr int G_PRODUCERS = 256;                 // warps 1..8
constexpr int G_THREADS = 32 + G_PRODUCERS;
constexpr uint32_t G_A_BYTES = GM * GK * 2;      // one of hi / lo : 16 KB
constexpr uint32_t G_B_BYTES = GN * GK * 2;      // 32 KB
constexpr uint32_t G_STAGE_BYTES = 2 * G_A_BYTES + 2 * G_B_BYTES;   // 96 KB

struct GemmArgs {
    int64_t M, N, K;
    const float* A; int64_t sAm, sAk;
    const float* B; int64_t sBk, sBn;
    float* C; int64_t ldc;
    int accumulate, use_atomic;
    int64_t kchunk;
    int passes;
};

__device__ __forceinline__ void split_store8(const float* v, uint32_t dst_hi, uint32_t dst_lo) {
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // packed conversions: cvt.rn.bf16x2.f32 for the hi pair, exact fp32 residuals, cvt.rn.bf16x2.f32 for the lo pair
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const float r0 = v[2 * i] - __uint_as_float(hb << 16), r1 = v[2 * i + 1] - __uint_as_float(hb & 0xFFFF0000u);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
        wh[i] = hb;
        wl[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst_hi), "r"(wh[0]), "r"(wh[1]), "r"(wh[2]), "r"(wh[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst_lo), "r"(wl[0]), "r"(wl[1]), "r"(wl[2]), "r"(wl[3]) : "memory");
}

// Fill one operand tile (ROWS x 64 k) of a stage.  elem(row, k) = P[row*sR + k*sK]; rows >= R or k >= kend read as 0.
// Every thread first issues ALL of its global loads for the tile (independent, so their latencies overlap), then
// converts and stores: the tile fill is then bounded by issue slots, not by one HBM round trip per task.
template <int ROWS>
__device__ __forceinline__ void fill_operand(const float* __restrict__ P, int64_t sR, int64_t sK, int64_t row0, int64_t R,
                                             int64_t k0, int64_t kend, uint32_t s_hi, uint32_t s_lo, int ptid) {
    constexpr uint32_t KC_STRIDE = ROWS * 16;     // bytes between k-chunks
    if (sK == 1) {
        // K-contiguous source: a thread owns (row, kc) = 8 consecutive k of one row (32 B); lanes take consecutive rows
        constexpr int TASKS = ROWS * 8 / G_PRODUCERS;
        const bool vec = ((sR & 3) == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0) && ((k0 & 3) == 0);
        float v[TASKS][8];
#pragma unroll
        for (int i = 0; i < TASKS; ++i) {
            const int task = ptid + i * G_PRODUCERS;
            const int row = task % ROWS, kc = task / ROWS;
            const int64_t gr = row0 + row, gk = k0 + kc * 8;
            if (gr < R && gk + 8 <= kend && vec) {
                const float4 x0 = *reinterpret_cast<const float4*>(P + gr * sR + gk);
                const float4 x1 = *reinterpret_cast<const float4*>(P + gr * sR + gk + 4);
                v[i][0] = x0.x; v[i][1] = x0.y; v[i][2] = x0.z; v[i][3] = x0.w; v[i][4] = x1.x; v[i][5] = x1.y; v[i][6] = x1.z; v[i][7] = x1.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[i][e] = (gr < R && gk + e < kend) ? P[gr * sR + gk + e] : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < TASKS; ++i) {
            const int task = ptid + i * G_PRODUCERS;
            const int row = task % ROWS, kc = task / ROWS;
            split_store8(v[i], s_hi + kc * KC_STRIDE + row * 16, s_lo + kc * KC_STRIDE + row * 16);
        }
    } else {
        // row-contiguous source (sR == 1 in practice): a thread owns (4 rows, kc): 8 float4 loads, one per k;
        // lanes take consecutive row quads (coalesced 512 B per k), shared stores are 64 B per lane, conflict-free
        constexpr int TASKS = (ROWS / 4) * 8 / G_PRODUCERS;
        const bool vec = (sR == 1) && ((sK & 3) == 0) && ((reinterpret_cast<uintptr_t>(P) & 15) == 0) && ((row0 & 3) == 0);
        float v[TASKS][4][8];
#pragma unroll
        for (int t = 0; t < TASKS; ++t) {
            const int task = ptid + t * G_PRODUCERS;
            const int rq = task % (ROWS / 4), kc = task / (ROWS / 4);
            const int64_t gr = row0 + rq * 4, gk = k0 + kc * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (vec && gr + 4 <= R && gk + e < kend) {
                    const float4 x = *reinterpret_cast<const float4*>(P + (gk + e) * sK + gr);
                    v[t][0][e] = x.x; v[t][1][e] = x.y; v[t][2][e] = x.z; v[t][3][e] = x.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[t][i][e] = (gr + i < R && gk + e < kend) ? P[(gr + i) * sR + (gk + e) * sK] : 0.f;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < TASKS; ++t) {
            const int task = ptid + t * G_PRODUCERS;
            const int rq = task % (ROWS / 4), kc = task / (ROWS / 4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                split_store8(v[t][i], s_hi + kc * KC_STRIDE + (rq * 4 + i) * 16, s_lo + kc * KC_STRIDE + (rq * 4 + i) * 16);
        }
    }
}

__global__ void __launch_bounds__(G_THREADS, 1) tc_gemm_kernel(GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES);
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };   // [0,S) full  [S,2S) empty  2S acc_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G_STAGES + 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < G_STAGES; ++i) { ptx::mbar_init(BAR(i), G_PRODUCERS); ptx::mbar_init(BAR(G_STAGES + i), 1); }
        ptx::mbar_init(BAR(2 * G_STAGES), 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 256); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int64_t m0 = (int64_t)blockIdx.y * GM, n0 = (int64_t)blockIdx.x * GN;
    const int64_t kbeg = (int64_t)blockIdx.z * a.kchunk;
    const int64_t kend = kbeg + a.kchunk < a.K ? kbeg + a.kchunk : a.K;
    const int nk = (int)((kend - kbeg + GK - 1) / GK);
    const uint32_t sbase = ptx::smem_u32(smem);

    if (warp == 0) {
        // ===== MMA issuer =====
        const bool issue = ptx::elect_one();
        const uint32_t idesc = ptx::make_idesc_f16(1, GM, GN);
        constexpr uint32_t DHI = ptx::desc_hi(128);
        for (int it = 0; it < nk; ++it) {
            const int st = it % G_STAGES;
            ptx::mbar_wait(BAR(st), (it / G_STAGES) & 1);
            ptx::tc_fence_after();
            const uint32_t sa_hi = sbase + st * G_STAGE_BYTES, sa_lo = sa_hi + G_A_BYTES;
            const uint32_t sb_hi = sa_lo + G_A_BYTES, sb_lo = sb_hi + G_B_BYTES;
            const uint32_t ah = ptx::desc_lo(sa_hi, GM * 16), al = ptx::desc_lo(sa_lo, GM * 16);
            const uint32_t bh = ptx::desc_lo(sb_hi, GN * 16), bl = ptx::desc_lo(sb_lo, GN * 16);
            constexpr uint32_t A_K16 = (2 * GM * 16) >> 4, B_K16 = (2 * GN * 16) >> 4;
            if (issue) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::umma_f16(tmem, ptx::desc_join(DHI, ah + k * A_K16), ptx::desc_join(DHI, bh + k * B_K16), idesc, (it | k) != 0);
                if (a.passes == 3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(tmem, ptx::desc_join(DHI, al + k * A_K16), ptx::desc_join(DHI, bh + k * B_K16), idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(tmem, ptx::desc_join(DHI, ah + k * A_K16), ptx::desc_join(DHI, bl + k * B_K16), idesc, 1);
                }
                ptx::umma_commit(BAR(G_STAGES + st));
            }
        }
        if (issue) ptx::umma_commit(BAR(2 * G_STAGES));
    } else {
        // ===== producers: fp32 global -> bf16 hi/lo operand layout =====
        const int ptid = threadIdx.x - 32;
        for (int it = 0; it < nk; ++it) {
            const int st = it % G_STAGES;
            ptx::mbar_wait(BAR(G_STAGES + st), ((it / G_STAGES) & 1) ^ 1);
            const uint32_t sa_hi = sbase + st * G_STAGE_BYTES, sa_lo = sa_hi + G_A_BYTES;
            const uint32_t sb_hi = sa_lo + G_A_BYTES, sb_lo = sb_hi + G_B_BYTES;
            const int64_t k0 = kbeg + (int64_t)it * GK;
            fill_operand<GM>(a.A, a.sAm, a.sAk, m0, a.M, k0, kend, sa_hi, sa_lo, ptid);
            fill_operand<GN>(a.B, a.sBn, a.sBk, n0, a.N, k0, kend, sb_hi, sb_lo, ptid);
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(BAR(st));
        }
        // ===== epilogue: TMEM -> shared (per-warp transpose tile, reusing the operand stages) -> C, coalesced:
        // a warp owns 32 rows x 128 columns; lane = row when reading TMEM, lane = column when writing C, so every
        // global access is 128 contiguous bytes instead of 32 scattered sectors =====
        ptx::mbar_wait(BAR(2 * G_STAGES), 0);        // all MMAs retired: the stage buffers are free
        ptx::tc_fence_after();
        const int q4 = warp & 3;                      // TMEM lane quarter
        const int ch = (warp - 1) >> 2;               // column half
        constexpr int TS = 129;                       // padded row stride (floats): conflict-free both ways
        float* tile = reinterpret_cast<float*>(smem) + (size_t)(warp - 1) * 32 * TS;
        const uint32_t tbase = tmem + ((uint32_t)(q4 * 32) << 16) + ch * 128;
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
            float v[32];
            if (nk > 0) { ptx::tmem_ld32(tbase + blk * 32, v); ptx::tc_wait_ld(); }
            else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) tile[lane * TS + blk * 32 + e] = v[e];
        }
        __syncwarp();
        const int64_t gn0 = n0 + ch * 128;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {          // 8 rows x 4 accesses in flight per lane
            float val[8][4], old[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t gm = m0 + q4 * 32 + r0 + i;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int col = c4 * 32 + lane;
                    val[i][c4] = tile[(r0 + i) * TS + col];
                    old[i][c4] = 0.f;
                    if (a.accumulate && !a.use_atomic && gm < a.M && gn0 + col < a.N) old[i][c4] = a.C[gm * a.ldc + gn0 + col];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t gm = m0 + q4 * 32 + r0 + i;
                if (gm >= a.M) continue;
                float* crow = a.C + gm * a.ldc + gn0;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int col = c4 * 32 + lane;
                    if (gn0 + col < a.N) {
                        if (a.use_atomic) atomicAdd(crow + col, val[i][c4]);
                        else crow[col] = val[i][c4] + old[i][c4];
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

// Returns KPRN_OK and sets *handled when the product ran on tensor cores.
int tc_gemm(kprn_ctx* ctx, int64_t M, int64_t N, int64_t K, const float* A, int64_t sAm, int64_t sAk, const float* B,
            int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, int splitk, bool* handled) {
    *handled = false;
    if (ctx->cfg.precision != KPRN_PREC_TC_ALL && ctx->cfg.precision != KPRN_PREC_TC_FAST && !ctx->gemm_on_tc) return KPRN_OK;
    if (M * N * K < (1ll << 24)) return KPRN_OK;                      // small products stay on the FFMA kernel
    if (!(sAk == 1 || sAm == 1) || !(sBk == 1 || sBn == 1)) return KPRN_OK;
    const size_t smem = G_STAGES * G_STAGE_BYTES + 256;
    // per-device opt-in: set on every launch (a process may hold ctxs on several devices)
    KPRN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (splitk < 1) splitk = 1;
    // split K until the grid covers the machine (the weight-gradient products have tiny M x N and a huge K)
    const int64_t tiles = cdiv(M, GM) * cdiv(N, GN);
    if (tiles * splitk < 148 && K > 8 * GK) splitk = (int)cdiv(148, tiles);
    int64_t kchunk = cdiv(cdiv(K, splitk), GK) * GK;
    splitk = (int)cdiv(K, kchunk);
    const int use_atomic = splitk > 1;
    if (use_atomic && !accumulate) {
        if (ldc == N) KPRN_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * M * N, ctx->stream));
        else KPRN_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, ctx->stream));
    }
    GemmArgs a;
    a.M = M; a.N = N; a.K = K; a.A = A; a.sAm = sAm; a.sAk = sAk; a.B = B; a.sBk = sBk; a.sBn = sBn; a.C = C; a.ldc = ldc;
    a.accumulate = accumulate ? 1 : 0; a.use_atomic = use_atomic; a.kchunk = kchunk;
    a.passes = ctx->cfg.precision == KPRN_PREC_TC_FAST ? 1 : 3;
    dim3 grid((unsigned)cdiv(N, GN), (unsigned)cdiv(M, GM), (unsigned)splitk);
    tc_gemm_kernel<<<grid, G_THREADS, smem, ctx->stream>>>(a);
    KPRN_LAUNCH_CHECK(ctx);
    *handled = true;
    return KPRN_OK;
}

}  // namespace kprn
