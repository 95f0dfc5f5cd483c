// tc_state.cuh -- constants and per-ctx state shared by the tensor-core translation units (tc_path.cu: scoring forward;
// tc_train.cu: training forward caches, BPTT and weight-gradient kernels).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace kprn {

using bf16 = __nv_bfloat16;

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
constexpr int RECUR_THREADS = 608;    // recur: warps 0 TMA producer, 1 MMA issuer, 2 TMEM alloc + h loader, 3-18 epilogue (4 groups); 96 registers per thread
                                      // (setmaxnreg measured: the pool is the launch allocation, 640 x 96: 104 for the epilogue does not remove its spills)

struct TcTrain;   // tc_train.cu

// leading-token reuse of the scoring recurrence (tc_prefix.cu)
struct PrefixState {
    DBuf token, hist, p, ps, perm, ids, pre, chain_u, chain_c;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_tok = nullptr, ev_chain = nullptr;
};
struct PrefixOut { const int32_t* ids; const int32_t* perm; const uint8_t* ps; const float* chain_u; const float* chain_c; };

struct TcState {
    TcTrain* tt = nullptr;             // training-side buffers and packed operands
    PrefixState* px = nullptr;
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
// The recurrent kernel evaluates the gates as sigmoid(x) = 1/(1+2^(-x log2 e)) and tanh(x) = 1 - 2/(1+2^(2x log2 e)) with
// ex2.approx; the factor (-log2 e for i,f,o and +2 log2 e for g) is folded into the packed weights, the hoisted table and
// hence XW and the accumulators, so the cell math needs no multiply in front of the four gate exponentials.
__device__ __forceinline__ float gate_scale(int gp) {
    const int q = (gp >> 5) & 3;
    return q == 1 ? 2.8853900817779268f : -1.4426950408889634f;
}
__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// wide hidden sizes (tc_wide.cu): one launch per step, state through HBM
struct RecurWide {
    const float* XW; const bf16* W; bf16* h0; bf16* h1; float* c; int64_t N; int T; int H; int Hp; int num_sms;
    const uint32_t* pos; const uint32_t* slot_of; const float* wo; float* part4; int64_t npad; float* hT;
};
int wide_recurrence(kprn_ctx* ctx, const RecurWide& w);
int pack_wh_wide(kprn_ctx* ctx, const float* Wh, int H, int Hp, bf16* out);

// leading-token reuse (tc_prefix.cu)
bool prefix_enabled(const kprn_ctx* ctx, int64_t N);
int prefix_build(kprn_ctx* ctx, const int32_t* d_ids, int64_t N, PrefixOut* out);
int prefix_join(kprn_ctx* ctx);
void prefix_destroy(kprn_ctx* ctx);

// training side (tc_train.cu)
int tc_train_pack(kprn_ctx* ctx);
int tc_train_buffers(kprn_ctx* ctx, int64_t N, float** cache, bf16** hcache);
void tc_train_destroy(kprn_ctx* ctx);

}  // namespace kprn
