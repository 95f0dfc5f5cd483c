// tc_state.cuh -- constants and per-ctx state shared by the tensor-core translation units (tc_path.cu: scoring forward;
// tc_train.cu: training forward caches, BPTT and weight-gradient kernels).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace kprn {

using bf16 = __nv_bfloat16;
constexpr int FMT_F16 = 0, FMT_BF16 = 1;       // operand split formats (see "operand split formats" below)

// wait on an mbarrier and, when debugging, charge the waited cycles to counter `slot`
#define TWAIT(slot, bar, par)                                              \
    do {                                                                  \
        if (a.dbg) { long long _t0 = clock64(); ptx::mbar_wait((bar), (par)); dbg_acc[slot] += clock64() - _t0; } \
        else ptx::mbar_wait((bar), (par));                                \
    } while (0)

constexpr int TILE_M = 128;           // paths per CTA tile (UMMA M)
constexpr int B_SLAB_ELEMS = 8 * 128 * 8;   // weight slab: 64 k x 128 gate columns (16 KB), both kernels
constexpr int PROJ_RING = 5;
constexpr int RECUR_RING = 6;          // 16 KB weight slabs in flight (1-CTA recurrence)
constexpr int RECUR_RING_PAIR = 8;     // 8 KB half-slabs in flight per CTA (CTA-pair recurrence)
constexpr int RECUR_BAR_BYTES = 384;   // barrier block at the end of the recurrence's shared memory
constexpr int MAX_NS = 6;             // k-slices of the proj A tile: KE/64 entity slices + NOH one-hot slices
constexpr int NUM_THREADS = 384;      // proj: warps 0 TMA producer, 1 MMA issuer, 2-3 gather/alloc, 4-11 epilogue
constexpr int RECUR_THREADS = 608;    // recur: warps 0 TMA producer, 1 MMA issuer, 2 TMEM alloc + h loader, 3-18 epilogue (4 groups); 19 warps -> 104 registers per thread

struct TcTrain;   // tc_train.cu

struct TcState {
    TcTrain* tt = nullptr;             // training-side buffers and packed operands
    uint64_t packed_version = 0;       // ctx->params_version the packed operands (entity tables included) were built from
    uint64_t packed_small = 0;         // ctx->small_version the packed WEIGHTS were built from (a sparse training step leaves the
                                       // entity tables current row by row and only bumps small_version)
    int Hp = 0, KE = 0, NCH = 0, KSh = 0, KSe = 0, G4 = 0, NOH = 0, NS = 0;
    bf16* E_hi = nullptr; bf16* E_lo = nullptr;
    bf16* WxE = nullptr; bf16* Wh = nullptr;
    bf16* WxE2 = nullptr;              // projection weights of entity columns 256.. (d_e > 256)
    int KE0 = 0, KSe1 = 0;             // entity columns / k-slices of the first and second projection launch
    DBuf wide_h0, wide_h1, wide_c;     // per-step recurrence (H > 256): ping-pong h (operand layout) and c
    float* Ptr = nullptr;
    DBuf xw, hT, hscratch, dbg, part4;
    DedupBufs dd;              // step-input deduplication
    int fmt = FMT_F16;         // operand split format (KPRN_TC_FP16=0 selects bf16 halves)
    float* scales = nullptr;   // [SC_COUNT] device floats, see the enum above
    bool dedup = true;         // project each distinct (type, entity, relation) triple once (KPRN_TC_DEDUP=0: every step)
    bool debug = false;
    bool have_tmaps = false;
    CUtensorMap tm_hi, tm_lo;  // 2-D tensor maps over the pre-split entity tables (TMA tile::gather4)
    bool recur_pairs = true;   // cta_group::2 recurrence (KPRN_TC_RECUR_PAIRS=0 falls back to the 1-CTA form)
    int num_sms = 148;
    bool attrs_set = false;
};


static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
// hidden size padded to an instantiated recur_kernel<HP> (padded units have zero weights and stay exactly 0)
static inline int pad_hidden(int H) { return H <= 64 ? 64 : (H <= 128 ? 128 : (H <= 256 ? 256 : 512)); }

// permuted gate column g' = chunk*128 + gate*32 + uu  <->  reference row gate*H + chunk*32 + uu
__host__ __device__ __forceinline__ int perm_to_row(int gp, int H, int& u) {
    int j = gp >> 7, q = (gp >> 5) & 3, uu = gp & 31;
    u = j * 32 + uu;
    return q * H + u;
}
__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---- operand split formats -------------------------------------------------------------------------------------------
// Every fp32 operand of a tensor-core product is split v = hi + lo into two 16-bit halves and a product is formed as
// hi*hi + lo*hi + hi*lo (three tcgen05.mma kind::f16 passes, fp32 accumulation in TMEM).
//   FMT_BF16: hi, lo = bf16 of v.  No range management, ~2^-18 relative error per product (8 + 8 mantissa bits).
//   FMT_F16 : hi, lo = fp16 of v * s, s a power of two per tensor chosen from the tensor's magnitude so that the largest
//             element lands near 2^11 (32x headroom below 65504; conversions saturate).  11 + 11 mantissa bits: ~2^-22 per
//             product; elements more than 2^14 below the tensor's maximum lose low-order bits of `lo` to the fp16 subnormal
//             range, an absolute error of 2^-25 of the maximum.  The accumulator is descaled by 1/(s_a s_b), exact.
// The format value is the a_format / b_format field of the instruction descriptor.  Buffers are declared bf16* whatever they hold.

__device__ __forceinline__ uint32_t split16_pair(float v0, float v1, int fmt, uint32_t& lo) {       // returns hi pair (v0 low half)
    if (fmt == FMT_BF16) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xFFFF0000u));
        lo = *reinterpret_cast<const uint32_t*>(&l2);
        return hb;
    }
    v0 = fminf(fmaxf(v0, -65504.f), 65504.f); v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
    const __half2 h2 = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
    return *reinterpret_cast<const uint32_t*>(&h2);
}
__device__ __forceinline__ void split16(float v, int fmt, bf16& hi, bf16& lo) {
    uint32_t l;
    const uint32_t h = split16_pair(v, 0.f, fmt, l);
    *reinterpret_cast<uint16_t*>(&hi) = (uint16_t)(h & 0xFFFFu);
    *reinterpret_cast<uint16_t*>(&lo) = (uint16_t)(l & 0xFFFFu);
}
// 8 consecutive values -> two 16-byte words (hi, lo)
__device__ __forceinline__ void split16x8(const float* v, int fmt, uint4& hi, uint4& lo) {
    uint32_t wh[4], wl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wh[i] = split16_pair(v[2 * i], v[2 * i + 1], fmt, wl[i]);
    hi = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    lo = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}
// the 16-bit pattern of the value `v` (a power of two) in the format: the one-hot operand slices
__device__ __forceinline__ uint32_t bits16(float v, int fmt) {
    if (fmt == FMT_BF16) { const bf16 b = __float2bfloat16_rn(v); return *reinterpret_cast<const uint16_t*>(&b); }
    const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h);
}

// per-tensor scales (device floats, written by tc_scales_kernel when the weights are packed; all 1 in FMT_BF16)
enum {
    SC_E = 0,         // entity table (E_hi / E_lo)
    SC_WXE,           // entity columns of Wx (gate-scaled), forward projection B operand
    SC_P,             // hoisted (type,relation) table (gate-scaled)
    SC_OH,            // the value of the one-hot slice in the projection: SC_E * SC_WXE / SC_P
    SC_INV_PROJ,      // 1 / (SC_E * SC_WXE): descale of the projection accumulator
    SC_WH,            // recurrent weights (gate-scaled)
    SC_H,             // h (|h| < 1): 2^13
    SC_INV_RECUR,     // 1 / (SC_H * SC_WH)
    SC_WT0,           // backward: Wh (unscaled values), B of dh = dG Wh
    SC_WT1,           // backward: Wx entity columns (unscaled values), B of dX = dG Wx_e
    SC_EMAX,          // max |E| seen by the last full split of the table (input of SC_E)
    SC_COUNT = 16
};

// wide hidden sizes (tc_wide.cu): one launch per step, state through HBM
struct RecurWide {
    const float* XW; const bf16* W; bf16* h0; bf16* h1; float* c; int64_t N; int T; int H; int Hp; int num_sms;
    const uint32_t* pos; const uint32_t* slot_of; const float* wo; float* part4; int64_t npad; float* hT;
    int fmt; const float* sc;
};
int wide_recurrence(kprn_ctx* ctx, const RecurWide& w);
int pack_wh_wide(kprn_ctx* ctx, const float* Wh, int H, int Hp, bf16* out, int fmt, const float* sc);

// training side (tc_train.cu)
int tc_train_pack(kprn_ctx* ctx);
int tc_train_buffers(kprn_ctx* ctx, int64_t N, float** cache, bf16** hcache);
void tc_train_destroy(kprn_ctx* ctx);

}  // namespace kprn
