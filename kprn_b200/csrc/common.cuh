// common.cuh -- ctx layout and helpers shared by the kprn_b200 translation units.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/kprn_b200.h"

namespace kprn {

struct Layout {
    int64_t E_t, E_e, E_r, Wx, bx, Wh, bh, W1, b1, Wo, bo, total;
    int32_t G, D;
};

// Growable device buffer.
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct TcState;     // tensor-core path state (tc_path.cu)
struct TrainState;  // sparse training state (train.cu)

// Slot map of one batch (tc_path.cu dedup_build): pos[o] = hash-table position of the key read at index o (0xFFFFFFFF = none),
// slot_of[pos] = slot of that key (slots numbered by first occurrence, blocks of 1024 indices placed atomically), ids = the
// key of every slot (mode 0: the (type, entity, relation) triple, 3 ints; modes 1/2: one int), count[0] = number of slots.
struct DedupBufs {
    DBuf table, slot_of, pos, ids, count;
    int64_t n_order = 0;
};

}  // namespace kprn

struct kprn_ctx {
    kprn_config cfg;
    kprn::Layout lay;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    float* d_params = nullptr;
    float* d_grads = nullptr;
    float* d_m = nullptr;
    float* d_v = nullptr;
    int64_t adam_t = 0;
    uint64_t params_version = 1;   // bumped on every parameter mutation (tc path repacks weights)
    // last-forward state
    int64_t N = 0, S = 0;
    bool have_fwd = false, have_train_cache = false, have_loss_grad = false;
    bool ids_on_host_copy = false;
    bool cls_only = false;          // last forward materialised only the class column (scores-only fast path)
    const float* head_part4 = nullptr;   // scores-only tensor-core forward: [4][head_npad] partial head dot products (else null)
    int64_t head_npad = 0;
    bool tc_train_fwd = false;      // the last training forward ran on the fused tensor-core kernels (backward: tc_backward)
    bool gemm_on_tc = false;        // scoring forward of a shape the fused tensor-core kernels do not cover: products via tc_gemm
    const int32_t* d_ids = nullptr;      // ids of the last forward (ctx-owned copy or caller's device ptr)
    const int32_t* d_seg = nullptr;
    kprn::DBuf ids_buf, seg_buf, path_logits, seg_logits, scores, dscores, dseg_logits, labels, tmp_small;
    kprn::DBuf act, cst, hst, z1;                  // training caches [N,T,G] [N,T,H] [N,T,H] [N,H]
    kprn::DBuf ws_x, ws_xw, ws_h0, ws_h1, ws_c, ws_dg, ws_dx, ws_dh, ws_dc, ws_ds, ws_dz;
    kprn::DBuf pair_buf;
    int32_t* d_err = nullptr;           // device error flag (bad ids)
    // pinned staging
    void* h_stage = nullptr;
    size_t h_stage_cap = 0;
    int64_t launches = 0;
    bool profiling = false;
    float prof_ms[3] = {0.f, 0.f, 0.f};
    int64_t prof_launches[3] = {0, 0, 0};
    struct ProfSpan { cudaEvent_t e0, e1; int slot; };
    std::vector<ProfSpan> prof_pending;
    kprn::TcState* tc = nullptr;
    kprn::TrainState* tr = nullptr;
    uint64_t small_version = 1;    // bumped when only the SMALL dense parameters (everything but the entity table) changed
    // multi-GPU (comm.cu): one process per GPU, communicator owned by the ctx
    void* comm = nullptr;
    int rank = 0, world = 1;
    int64_t comm_bytes = 0, comm_calls = 0;
    kprn::DBuf tmp_comm;
    kprn::DBuf ws_part;                 // per-block partial sums of the column-sum / scalar-sum reductions (summed in block order: deterministic)
    kprn::DBuf tmp_pool;                // (m, z) partials of the path-sharded pool
    // sparse training step (train.cu): while set, the backward scatters entity-row gradients into the compact buffer
    // sparse_gE[slot_of[pos[o]]] (o = read-order index of the step) instead of the dense table in d_grads
    float* sparse_gE = nullptr;
    const uint32_t* sparse_pos = nullptr;
    const uint32_t* sparse_slot_of = nullptr;
    bool in_train_step = false;
    // a sparse training step without gradient clipping lets the entity-row optimizer run on the backward's side stream, beside
    // the weight-gradient kernel: tc_backward then leaves the join of that stream to the caller (eg_stream / eg_event set)
    bool defer_eg_join = false;
    cudaStream_t eg_stream = nullptr;
    cudaEvent_t eg_event = nullptr;
    double* last_loss = nullptr;    // device scalar of the last training step's loss
    int opt_dedup = -1;                 // kprn_set_option("dedup"): -1 = KPRN_TC_DEDUP / default (on), 0 / 1 = forced
    int opt_prefix = -1;                // kprn_set_option("prefix"): leading-token reuse in the scoring recurrence; -1 = KPRN_TC_PREFIX / default (on)
    int opt_path_sharded = 0;           // kprn_set_option("path_sharded", 1): the PATHS of every (user,item) segment are spread over the ranks
                                        // (every rank passes all S segments, some possibly empty); the pooled logits are combined across
                                        // ranks as log-sum-exp partials (m, z) right after the local pool -- SURVEY 8e "logit allreduce"
};

namespace kprn {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define KPRN_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return kprn::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define KPRN_TRY(expr)              \
    do {                            \
        int _rc = (expr);           \
        if (_rc != KPRN_OK) return _rc; \
    } while (0)

#define KPRN_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                        \
        (ctx)->launches++;                                                      \
        cudaError_t _e = cudaGetLastError();                                    \
        if (_e != cudaSuccess) return kprn::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

int ensure(DBuf& b, size_t bytes);
int ensure_stage(kprn_ctx* ctx, size_t bytes);
// profiling spans (no-ops unless ctx->profiling)
int prof_begin(kprn_ctx* ctx, cudaEvent_t* e0);
int prof_end(kprn_ctx* ctx, cudaEvent_t e0, int slot);

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- fp32 CUDA-core path (kernels_simple.cu) ----
int simple_forward(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, bool train);
int simple_backward(kprn_ctx* ctx, float scale);
int pool_forward(kprn_ctx* ctx, const int32_t* d_seg, int64_t S, float* d_out_scores, float* d_out_logits);
int head_forward(kprn_ctx* ctx, const float* d_hT, int64_t ldh, int64_t n0, int64_t nb, float* z1_opt);
int head_forward_cls(kprn_ctx* ctx, const float* d_hT, int64_t ldh, int64_t n0, int64_t nb);

// generic fp32 GEMM: C[m,n] (=|+=) alpha * sum_k A(m,k) B(k,n); element strides given explicitly.
int sgemm(kprn_ctx* ctx, int64_t M, int64_t N, int64_t K, const float* A, int64_t sAm, int64_t sAk,
          const float* B, int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, int splitk = 1);

// generic-stride GEMM on tensor cores (tc_gemm.cu); *handled == false -> caller falls back to the FFMA kernel
int tc_gemm(kprn_ctx* ctx, int64_t M, int64_t N, int64_t K, const float* A, int64_t sAm, int64_t sAk, const float* B,
            int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, int splitk, bool* handled);

// ---- tensor-core path (tc_path.cu) ----
bool tc_supported(const kprn_ctx* ctx);
bool tc_validates_ids(const kprn_ctx* ctx, int64_t N);   // true when tc_forward(train = false) will flag out-of-range ids itself
int tc_forward(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, bool train);
void tc_destroy(kprn_ctx* ctx);
bool tc_train_supported(const kprn_ctx* ctx);             // training forward + backward on the fused tensor-core kernels (tc_train.cu)
int tc_backward(kprn_ctx* ctx, float scale);
// pool + head backward for the live class column on the fp32 kernels: leaves ds_c[N] in ctx->ws_ds, adds gbo[c] and gWo[c,:]
int head_backward_cls(kprn_ctx* ctx, float scale, const float* hT, int64_t ldh);
int dedup_build(kprn_ctx* ctx, DedupBufs& b, int mode, const int32_t* ids, int64_t N, int T, int64_t n_keys);
void dedup_free(DedupBufs& b);

// ---- multi-GPU (comm.cu) ----
int comm_allreduce_f32(kprn_ctx* ctx, float* buf, int64_t n, bool max_op);
int comm_allreduce_f32_2(kprn_ctx* ctx, float* a, int64_t na, float* b, int64_t nb);
int comm_allreduce_i64(kprn_ctx* ctx, int64_t* buf, int64_t n, bool max_op);
int comm_allgather_bytes(kprn_ctx* ctx, const void* send, void* recv, int64_t bytes);
int comm_broadcast_f32(kprn_ctx* ctx, float* buf, int64_t n, int root);
int pool_allreduce_lse(kprn_ctx* ctx, const int32_t* d_seg, int64_t S, float* d_out_scores, float* d_out_logits);

// ---- pieces of the API layer the training step composes (api.cu) ----
int forward_device_impl(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg, int64_t N, int64_t S, bool train, bool want_logits,
                        float* d_out_scores, float* d_out_logits);
int loss_bce_device(kprn_ctx* ctx, const float* d_labels, int64_t S, double** d_loss_out);
int loss_bpr_device(kprn_ctx* ctx, const int32_t* d_pos, const int32_t* d_neg, int64_t n_pairs, int on_logits, double** d_loss_out);
int backward_device(kprn_ctx* ctx, float scale);
int host_to_device(kprn_ctx* ctx, void* dst, const void* src, size_t bytes, size_t stage_off);
void tc_packed_entity(kprn_ctx* ctx, __nv_bfloat16** hi, __nv_bfloat16** lo, int* KE);   // valid packed bf16 entity tables, or nulls

// ---- sparse training step (train.cu) ----
int train_note_dense_step(kprn_ctx* ctx, float step);
int train_materialize_grads(kprn_ctx* ctx);
int train_flush(kprn_ctx* ctx);          // brings every entity row up to date (lazy dense-equivalent Adam); no-op when nothing is pending
void train_destroy(kprn_ctx* ctx);

}  // namespace kprn
