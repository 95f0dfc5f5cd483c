// umma_probe.cu -- hardware probe for the tcgen05 operand forms the training kernels rely on (run on the GPU box):
//   1. MN-major ("transposed") no-swizzle operands: the bytes [c/8][r 128][8] that serve as a K-major operand with
//      (row = r, k = c) read as an MN-major operand with (mn = c, k = r); which of LBO / SBO carries which stride;
//   2. MN-major SWIZZLE_128B B operand as TMA writes it (rows of 128 B = 64 bf16, 16-byte chunks XORed with row & 7).
// Prints max |D - ref| for each candidate encoding; the kernels use the one that reads 0.
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../kprn_b200/csrc/ptx.cuh"
using namespace kprn;
typedef __nv_bfloat16 bf16;

struct Args { const bf16* A; const bf16* B; float* D; int a_mode, b_mode; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int N; int nk16; uint32_t a_k16, b_k16; };
// modes: 0 = K-major no-swizzle, 1 = MN-major no-swizzle, 2 = MN-major SWIZZLE_128B

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, int sw128) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    if (sw128) d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                  // 32 KB
    uint8_t* sB = smem + 32768;          // 64 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 98304);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::fence_mbar_init(); }
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(slot), 256); ptx::tmem_relinquish(); }
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(a.A)[i];
    for (int i = threadIdx.x; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(a.B)[i];
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = ptx::make_idesc_f16(1, 128, a.N);
        if (a.a_mode) idesc |= 1u << 15;
        if (a.b_mode) idesc |= 1u << 16;
        for (int k = 0; k < a.nk16; ++k) {
            uint64_t da = mkdesc(ptx::smem_u32(sA) + k * a.a_k16, a.a_lbo, a.a_sbo, a.a_mode == 2);
            uint64_t db = mkdesc(ptx::smem_u32(sB) + k * a.b_k16, a.b_lbo, a.b_sbo, a.b_mode == 2);
            ptx::umma_f16(tmem, da, db, idesc, k != 0);
        }
        ptx::umma_commit(ptx::smem_u32(bar));
    }
    ptx::mbar_wait(ptx::smem_u32(bar), 0);
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < a.N; c0 += 32) {
        float v[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tc_wait_ld();
        for (int i = 0; i < 32; ++i) a.D[(size_t)threadIdx.x * a.N + c0 + i] = v[i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

static float frand() { return (float)(rand() % 17 - 8) / 8.f; }   // exactly representable in bf16

int main() {
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    // logical matrices: X[r 128][c 128] (A source), Y[r 128][c 256] (B source)
    std::vector<float> X(128 * 128), Y(128 * 256);
    for (auto& v : X) v = frand();
    for (auto& v : Y) v = frand();
    // interleaved no-swizzle images: elem (r, c) at ((c/8)*128 + r)*8 + c%8
    std::vector<bf16> Xi(128 * 128), Yi(128 * 256);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 128; ++c) Xi[((c / 8) * 128 + r) * 8 + c % 8] = __float2bfloat16(X[r * 128 + c]);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 256; ++c) Yi[((c / 8) * 128 + r) * 8 + c % 8] = __float2bfloat16(Y[r * 256 + c]);
    // SW128 image of Y: panels of 64 columns; panel p: [row 128][128 B], 16-byte chunk j of row r stored at chunk j ^ (r & 7)
    std::vector<bf16> Ys(128 * 256);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 256; ++c) {
        int p = c / 64, cc = c % 64, j = cc / 8, e = cc % 8;
        Ys[(size_t)p * 128 * 64 + r * 64 + ((j ^ (r & 7)) * 8) + e] = __float2bfloat16(Y[r * 256 + c]);
    }
    bf16 *dA, *dB; float* dD;
    cudaMalloc(&dA, 32768); cudaMalloc(&dB, 65536); cudaMalloc(&dD, 128 * 256 * 4);
    std::vector<float> D(128 * 256);
    auto run = [&](const char* name, const std::vector<bf16>& Ah, const std::vector<bf16>& Bh, Args a, auto ref) {
        cudaMemcpy(dA, Ah.data(), 32768, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bh.data(), 65536, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xFF, 128 * 256 * 4);
        a.A = dA; a.B = dB; a.D = dD;
        probe_kernel<<<1, 128, 100 * 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-58s CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(D.data(), dD, 128 * a.N * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < a.N; ++n) { double d = fabs((double)D[m * a.N + n] - ref(m, n)); if (!(d <= worst)) worst = d; }
        printf("%-58s max|err| = %g %s\n", name, worst, worst == 0 ? "OK" : "");
    };
    // 0. baseline: K-major both.  D[m][n] = sum_c X[m][c] Y[n][c], K = 128 (c), N = 128 rows of Y
    { Args a = {}; a.a_mode = 0; a.b_mode = 0; a.a_lbo = 2048; a.a_sbo = 128; a.b_lbo = 2048; a.b_sbo = 128; a.N = 128; a.nk16 = 8; a.a_k16 = 4096; a.b_k16 = 4096;
      run("K-major A, K-major B (baseline)", Xi, Yi, a, [&](int m, int n) { double s = 0; for (int c = 0; c < 128; ++c) s += X[m * 128 + c] * Y[n * 256 + c]; return s; }); }
    // 1. A MN-major: D[m][n] = sum_r X[r][m] Y[n][r]?  B K-major needs k = c of Y: use k = r for A and k = c for B -> D[m][n] = sum_k X[k][m] * Y[n][k]
    for (int v = 0; v < 2; ++v) {
        Args a = {}; a.a_mode = 1; a.b_mode = 0; a.a_lbo = v ? 2048 : 128; a.a_sbo = v ? 128 : 2048; a.b_lbo = 2048; a.b_sbo = 128; a.N = 128; a.nk16 = 8;
        a.a_k16 = 256; a.b_k16 = 4096;      // A: 16 k-rows further = 256 B; B: 2 chunks of 8 columns = 4096 B
        run(v ? "MN-major A (LBO=2048 mn, SBO=128 k), K-major B" : "MN-major A (LBO=128 k, SBO=2048 mn), K-major B", Xi, Yi, a,
            [&](int m, int n) { double s = 0; for (int k = 0; k < 128; ++k) s += X[k * 128 + m] * Y[n * 256 + k]; return s; });
    }
    // 2. both MN-major: D[m][n] = sum_r X[r][m] Y[r][n], N = 256
    for (int v = 0; v < 2; ++v) {
        Args a = {}; a.a_mode = 1; a.b_mode = 1; a.a_lbo = v ? 2048 : 128; a.a_sbo = v ? 128 : 2048; a.b_lbo = a.a_lbo; a.b_sbo = a.a_sbo; a.N = 256; a.nk16 = 8;
        a.a_k16 = 256; a.b_k16 = 256;
        run(v ? "MN-major A and B (LBO=2048 mn, SBO=128 k), N=256" : "MN-major A and B (LBO=128 k, SBO=2048 mn), N=256", Xi, Yi, a,
            [&](int m, int n) { double s = 0; for (int k = 0; k < 128; ++k) s += X[k * 128 + m] * Y[k * 256 + n]; return s; });
    }
    // 3. A MN-major no-swizzle, B MN-major SWIZZLE_128B (TMA image), N = 256 (4 panels of 16 KB)
    for (int v = 0; v < 2; ++v) {
        Args a = {}; a.a_mode = 1; a.b_mode = 2; a.a_lbo = 128; a.a_sbo = 2048; a.b_lbo = v ? 1024 : 16384; a.b_sbo = v ? 16384 : 1024; a.N = 256; a.nk16 = 8;
        a.a_k16 = 256; a.b_k16 = 2048;      // 16 rows of 128 B
        run(v ? "MN-major A, SW128 MN-major B (LBO=1024 k, SBO=16384 panel)" : "MN-major A, SW128 MN-major B (LBO=16384 panel, SBO=1024 k)", Xi, Ys, a,
            [&](int m, int n) { double s = 0; for (int k = 0; k < 128; ++k) s += X[k * 128 + m] * Y[k * 256 + n]; return s; });
    }
    // 3b. same with A taken with the other no-swizzle assignment (in case 1 showed that one)
    for (int v = 0; v < 2; ++v) {
        Args a = {}; a.a_mode = 1; a.b_mode = 2; a.a_lbo = 2048; a.a_sbo = 128; a.b_lbo = v ? 1024 : 16384; a.b_sbo = v ? 16384 : 1024; a.N = 256; a.nk16 = 8;
        a.a_k16 = 256; a.b_k16 = 2048;
        run(v ? "MN-major A', SW128 MN-major B (LBO=1024 k, SBO=16384 panel)" : "MN-major A', SW128 MN-major B (LBO=16384 panel, SBO=1024 k)", Xi, Ys, a,
            [&](int m, int n) { double s = 0; for (int k = 0; k < 128; ++k) s += X[k * 128 + m] * Y[k * 256 + n]; return s; });
    }
    // 4. how the tensor core adds a K=16 block to the fp32 accumulator: acc = 1.0, then seven blocks of +0.75 ulp(1.0) each.
    //    round-to-nearest per block -> 1 + 7 ulp;  exact / wide accumulator -> 1 + 5 ulp (5.25 rounded);  truncation -> 1.0
    {
        std::vector<bf16> Aq(128 * 128, __float2bfloat16(0.f)), Bq(128 * 256, __float2bfloat16(0.f));
        auto seti = [&](std::vector<bf16>& v, int r, int c, float x) { v[((c / 8) * 128 + r) * 8 + c % 8] = __float2bfloat16(x); };
        for (int r = 0; r < 128; ++r) {
            seti(Aq, r, 0, 1.0f); seti(Bq, r, 0, 1.0f);
            for (int kb = 1; kb < 8; ++kb) { seti(Aq, r, kb * 16, 1.5f); seti(Bq, r, kb * 16, 5.9604644775390625e-08f); }     // 2^-24
        }
        Args a = {}; a.a_mode = 0; a.b_mode = 0; a.a_lbo = 2048; a.a_sbo = 128; a.b_lbo = 2048; a.b_sbo = 128; a.N = 128; a.nk16 = 8; a.a_k16 = 4096; a.b_k16 = 4096;
        cudaMemcpy(dA, Aq.data(), 32768, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bq.data(), 65536, cudaMemcpyHostToDevice);
        a.A = dA; a.B = dB; a.D = dD;
        probe_kernel<<<1, 128, 100 * 1024>>>(a);
        cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        const double ulp = 1.1920928955078125e-07;
        printf("fp32 accumulation of 1.0 + 7 x 0.75 ulp: (D - 1) / ulp = %.3f   [7 = round-to-nearest per block, 5 = exact sum rounded once, 0 = truncation]\n",
               ((double)D[0] - 1.0) / ulp);
    }
    return 0;
}
