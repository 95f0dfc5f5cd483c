// sgemm.cuh -- generic-stride fp32 GEMM on CUDA cores (FFMA), used by the KPRN_PREC_FP32 path and
// by every backward product.  C[m,n] (=|+=) sum_k A(m,k) * B(k,n) with arbitrary element strides, so the
// same kernel serves y = x W^T (nn.Linear forward), dx = dy W (updateGradInput) and
// gW += dy^T x (accGradParameters, split-K with float atomics).
#pragma once
#include "common.cuh"

namespace kprn {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_TM = 8, SG_TN = 8, SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS)
sgemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t sAm, int64_t sAk,
             const float* __restrict__ B, int64_t sBk, int64_t sBn, float* __restrict__ C, int64_t ldc,
             int accumulate, int64_t kchunk, int use_atomic) {
    __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * SG_BM, n0 = (int64_t)blockIdx.x * SG_BN;
    const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
    const int64_t kend = (kbeg + kchunk < K) ? kbeg + kchunk : K;
    float acc[SG_TM][SG_TN];
#pragma unroll
    for (int i = 0; i < SG_TM; ++i)
#pragma unroll
        for (int j = 0; j < SG_TN; ++j) acc[i][j] = 0.f;

    const bool a_kfast = (sAk == 1);
    const bool b_nfast = (sBn == 1);
    for (int64_t k0 = kbeg; k0 < kend; k0 += SG_BK) {
#pragma unroll
        for (int i = 0; i < (SG_BM * SG_BK) / SG_THREADS; ++i) {
            int e = tid + i * SG_THREADS;
            int kk, mm;
            if (a_kfast) { kk = e % SG_BK; mm = e / SG_BK; } else { mm = e % SG_BM; kk = e / SG_BM; }
            int64_t gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < kend) ? A[gm * sAm + gk * sAk] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < (SG_BN * SG_BK) / SG_THREADS; ++i) {
            int e = tid + i * SG_THREADS;
            int kk, nn;
            if (b_nfast) { nn = e % SG_BN; kk = e / SG_BN; } else { kk = e % SG_BK; nn = e / SG_BK; }
            int64_t gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < kend) ? B[gk * sBk + gn * sBn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            float a[SG_TM], b[SG_TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < SG_TM; ++i)
#pragma unroll
                for (int j = 0; j < SG_TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < SG_TM; ++i) {
        int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < SG_TN; ++j) {
            int64_t gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= N) continue;
            float* c = C + gm * ldc + gn;
            if (use_atomic) atomicAdd(c, acc[i][j]);
            else if (accumulate) *c += acc[i][j];
            else *c = acc[i][j];
        }
    }
}

}  // namespace kprn
