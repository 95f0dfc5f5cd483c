/*
 * kprn_b200.h -- C ABI of the B200-native KPRN path-scoring engine (libkprn_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of eBay/KPRN: what a LuaJIT-FFI shim
 * (kprn_b200/lua/kprn_b200.lua, same declarations via ffi.cdef) binds instead of the
 * reference's Torch7 nn graph.  Citations are relative to
 * /root/reference/release/songPathRnn/ ("SP/").
 *
 * Conventions: C99, no C++ types, no callbacks.  Every function returns 0 on success or a
 * non-zero KPRN_E* code; kprn_last_error() gives the thread-local message (the Lua shim
 * turns it into error(), Torch7's convention).  Host pointers are owned by the caller and
 * are not retained after the call returns.  Device memory is owned by the ctx.  A ctx is
 * bound to one CUDA device and is not re-entrant; one ctx per host thread / Lua state.
 * One process per GPU; multi-GPU runs go through the ctx's NCCL communicator (kprn_comm_*).
 *
 * There is NO CPU fallback: if no CUDA device is usable every call fails with KPRN_ECUDA.
 */
#ifndef KPRN_B200_H
#define KPRN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPRN_OK 0
#define KPRN_EINVAL 1   /* bad argument / id out of range / shape mismatch */
#define KPRN_ECUDA 2    /* CUDA runtime or launch failure, or no device    */
#define KPRN_ESTATE 3   /* call order violated (e.g. backward before forward) */
#define KPRN_ENOMEM 4

/* cell: SP/model/OneModel.lua:235-254 */
#define KPRN_CELL_LSTM 0 /* nn.FastLSTM(D,H), -rnnType lstm */
#define KPRN_CELL_RNN 1  /* ReLU Elman RNN + MaskZero, -rnnType rnn (shipped config.sh:33) */
/* pool: SP/model/OneModel.lua:284-293, SP/eval/test_from_checkpoint.lua:70-79 */
#define KPRN_POOL_LSE 0  /* topK=2: nn.LogSumExp(2) (SP/model/module/LogSumExp.lua) */
#define KPRN_POOL_MAX 1  /* topK=0: nn.Max(2) */
#define KPRN_POOL_TOPK 2 /* topK=1: nn.TopK(k,2) + nn.Mean(2) (SP/model/module/TopK.lua) */
/* precision of the gate GEMMs */
#define KPRN_PREC_FP32 0     /* CUDA-core FFMA, fp32 everywhere */
#define KPRN_PREC_TC_SPLIT 1 /* tcgen05 kind::f16, 3-pass hi/lo split (parity mode on tensor cores) */
#define KPRN_PREC_TC_FAST 2  /* tcgen05 kind::f16, single pass (NOT parity-certified) */
#define KPRN_PREC_TC_ALL 3   /* like TC_SPLIT, and the training products (forward with caches, BPTT, weight gradients) also run
                                on tcgen05 with the bf16x3 split: scores/loss within 1e-4, gradients within 5e-4 of fp64
                                (TC_SPLIT keeps training on fp32 CUDA cores: gradients within 1e-4) */
/* tables for kprn_zero_row */
#define KPRN_TABLE_TYPE 0
#define KPRN_TABLE_ENTITY 1
#define KPRN_TABLE_RELATION 2

typedef struct kprn_ctx kprn_ctx;

/* Model + engine configuration.  Replaces the torch.CmdLine flags of SP/model/OneModel.lua:27-87
 * that shape the hot path (-relationEmbeddingDim, -entityTypeEmbeddingDim, -entityEmbeddingDim,
 * -rnnHidSize, -relationVocabSize, -entityVocabSize, -entityTypeVocabSize, -rnnType, -topK, -k)
 * and the hard-coded labelDimension=46 (OneModel.lua:119) / classId (MyOptimizer.lua:126). */
typedef struct kprn_config {
    int32_t T;            /* steps per path (movie_data_format.sh:2 "-m 6") */
    int32_t d_t, d_e, d_r;/* embedding widths: type, entity, relation; D = d_t+d_e+d_r */
    int32_t H;            /* rnnHidSize */
    int32_t L;            /* labelDimension (46 in the reference) */
    int64_t V_t, V_e, V_r;/* vocabulary sizes (rows of the three nn.LookupTable) */
    int32_t cell;         /* KPRN_CELL_* */
    int32_t head_layers;  /* 1 = nn.Linear(H,L) (OneModel.lua:275); 2 = Linear(H,H)-ReLU-Linear(H,L) */
    int32_t pool;         /* KPRN_POOL_* */
    int32_t topk;         /* k of KPRN_POOL_TOPK */
    float inv_gamma;      /* LSE temperature, 1.0 = reference */
    int32_t class_id;     /* 1-based column selected by nn.Select(2,classId) */
    int32_t precision;    /* KPRN_PREC_* */
    int32_t device;       /* CUDA device ordinal */
    int32_t reserved;
} kprn_config;

const char* kprn_last_error(void);
const char* kprn_version(void);

/* ctx lifetime. kprn_create allocates parameters (zero-filled), gradients and optimizer state on
 * cfg->device.  Replaces the graph construction of SP/model/OneModel.lua:203-294. */
int kprn_create(const kprn_config* cfg, kprn_ctx** out);
int kprn_destroy(kprn_ctx* ctx);
/* run subsequent work on an existing cudaStream_t (void* to keep the header CUDA-free); NULL = ctx's own stream */
int kprn_set_stream(kprn_ctx* ctx, void* cuda_stream);
int kprn_sync(kprn_ctx* ctx);

/* Flat parameter vector, order E_type,E_entity,E_relation,Wx,bx,Wh,[bh],[W1,b1],Wo,bo -- the DFS order
 * of nn.Module:getParameters() that MyOptimizer.lua:42 flattens.  offsets[11] in that order
 * (E_t,E_e,E_r,Wx,bx,Wh,bh,W1,b1,Wo,bo), -1 when absent. */
int kprn_param_count(kprn_ctx* ctx, int64_t* n);
int kprn_param_offsets(kprn_ctx* ctx, int64_t* offsets11);
int kprn_set_params(kprn_ctx* ctx, const float* host, int64_t n);       /* checkpoints / tests (torch.load) */
int kprn_set_params_f64(kprn_ctx* ctx, const double* host, int64_t n);  /* Torch7 CPU checkpoints are DoubleTensor */
int kprn_get_params(kprn_ctx* ctx, float* host, int64_t n);             /* saver hook OneModel.lua:393-401 */
int kprn_get_grads(kprn_ctx* ctx, float* host, int64_t n);
int kprn_init_uniform(kprn_ctx* ctx, float scale, uint64_t seed);       /* param:uniform(-s,s), OneModel.lua:306-309 */
/* raw device views of the flat vectors (an external optimizer / collective working on the device; replaces the flat
 * CudaTensor views of getParameters(), MyOptimizer.lua:42).  kprn_params_device marks the parameters changed once, at the
 * call; a caller that KEEPS the pointer and writes through it later must call kprn_params_dirty(ctx) after every write so
 * that the tensor-core forward re-packs its bf16 operand copies (otherwise it scores with the previous values). */
int kprn_params_device(kprn_ctx* ctx, float** dptr, int64_t* n);
int kprn_params_dirty(kprn_ctx* ctx);
int kprn_grads_device(kprn_ctx* ctx, float** dptr, int64_t* n);

/* params[1][idx]:zero() of MyOptimizer:zeroPadTokens (MyOptimizer.lua:74-93); row is 1-based. */
int kprn_zero_row(kprn_ctx* ctx, int table, int64_t row_1based);

/*
 * model:forward(inputs) of MyOptimizer.lua:189 / test_from_checkpoint.lua:109 for the fused graph
 * MapReduce(predictor_net, reducer) -> Sigmoid -> Select(2,classId)   (OneModel.lua:294).
 *   ids         [N,T,3] int32, 1-based, columns (type, entity, relation)   (Batcher.lua:43-54 data)
 *   seg_offsets [S+1] CSR offsets over consecutive paths; NULL + paths_per_seg>0 = dense bucket [S,P]
 *   out_scores  [S]   sigmoid(pool)[:,class_id]          (may be NULL)
 *   out_logits  [S,L] pooled logits before the sigmoid   (may be NULL)
 * train != 0 keeps the activations needed by kprn_backward.
 * Host-pointer variant: copies in/out through pinned staging inside the timed call.
 */
int kprn_forward(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S,
                 int32_t paths_per_seg, int32_t train, float* out_scores, float* out_logits);
/* same, reference tensor layout: data[B,P,T,3] DoubleTensor of 1-based ids (SP/util/Util.lua:121-201) */
int kprn_forward_f64ids(kprn_ctx* ctx, const double* ids_f64, int64_t B, int64_t P, int32_t train,
                        float* out_scores, float* out_logits);
/* device-pointer variant: everything already resident in HBM, fully asynchronous on the ctx stream */
int kprn_forward_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg_offsets, int64_t N, int64_t S,
                        int32_t train, float* d_out_scores, float* d_out_logits);
/* Page-locked host buffers for the caller's batch tensors (optional).  kprn_forward* detect pinned host pointers and DMA
 * from them directly; pageable pointers are staged through an internal pinned buffer first (one extra host memcpy).
 * Replaces BatcherFileList:preallocateTensorToGPU / populateGPUTensor (SP/model/batcher/BatcherFileList.lua:53-96). */
int kprn_pinned_alloc(int64_t bytes, void** host_ptr);
int kprn_pinned_free(void* host_ptr);

/* per-path head outputs s[N,L] of the last forward (MapReduce's "mapped", MapReduce.lua:33) -- tests/diagnostics */
int kprn_get_path_logits(kprn_ctx* ctx, float* host, int64_t n_floats);

/* criterion:forward + criterion:backward (MyOptimizer.lua:193-194) on the scores of the last forward;
 * caches dL/dscore for kprn_backward.  BCE = nn.BCECriterion (OneModel.lua:300), labels[S] in {0,1}. */
int kprn_loss_bce(kprn_ctx* ctx, const float* labels, int64_t S, float* loss);
/* nn.BPRLoss (SP/model/criterion/BPRLoss.lua:18-32) over pairs of segments of the last forward:
 * input = { scores[pos_seg[i]], scores[neg_seg[i]] } (0-based segment indices), reference's gradient
 * (not divided by n).  apply_to_logits != 0 pairs the pooled class logits instead of the sigmoid scores. */
int kprn_loss_bpr(kprn_ctx* ctx, const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs,
                  int32_t apply_to_logits, float* loss);

/* model:backward(inputs, df_do) (MyOptimizer.lua:195): BPTT through pool, head, cell and the three
 * lookup tables; ACCUMULATES into the gradient buffer (accGradParameters), times `scale`.
 * grad_scores [S] = dL/dscore from the caller, or NULL to use the cached loss gradient. */
int kprn_backward(kprn_ctx* ctx, const float* grad_scores, float scale);
int kprn_zero_grad(kprn_ctx* ctx);                      /* model:zeroGradParameters(), MyOptimizer.lua:186 */
int kprn_grad_norm(kprn_ctx* ctx, float* norm);         /* gradParameters:norm(), MyOptimizer.lua:198 */
int kprn_grad_scale(kprn_ctx* ctx, float s);            /* gradParameters:mul(), MyOptimizer.lua:200 */
int kprn_grad_add_params(kprn_ctx* ctx, float l2);      /* gradParameters:add(l2, parameters), :203/:212 */
/* optim.adam on the flat vector (OneModel.lua:340-355, MyOptimizer.lua:218); dense like the reference. */
int kprn_adam_step(kprn_ctx* ctx, float lr, float beta1, float beta2, float eps);
/* one whole MyOptimizer:trainBatch (MyOptimizer.lua:177-221) with BCE: zero pad rows -> zero grad ->
 * forward -> BCE -> backward -> [clip + l2] -> Adam -> zero pad rows.  pad_rows[3] 1-based, 0 = none.
 * (= kprn_train_step with default options: row-sparse evaluation, same results.) */
int kprn_train_batch_bce(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S,
                         const float* labels, const int64_t* pad_rows3, int32_t regularize, int32_t use_clip,
                         float clip_norm, float l2, float lr, float beta1, float beta2, float eps, float* loss);

/* nn.BPRLoss flavour of the same step (SP/model/criterion/BPRLoss.lua:18-32 on pooled scores; BASELINE configs[3]):
 * pos_seg/neg_seg[n_pairs] are 0-based segment indices of the batch, paired element-wise. */
int kprn_train_batch_bpr(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S,
                         const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs, int32_t apply_to_logits,
                         const int64_t* pad_rows3, int32_t regularize, int32_t use_clip, float clip_norm, float l2, float lr,
                         float beta1, float beta2, float eps, float* loss);

/* The general form of MyOptimizer:trainBatch (MyOptimizer.lua:177-221).  By default the 2.85 M-row entity table is handled
 * row-sparsely WITHOUT changing the result: the entity gradient is kept per distinct row of the batch, and optim.adam's
 * dense update (rows without gradient still decay and move) is evaluated lazily -- replayed, with the same arithmetic, the
 * next time a row is read, or by any call that exposes the parameters (scoring forward, kprn_get_params, ...).
 * dense = 1 runs the reference's flow literally (dense gradient, dense Adam); it is also what l2 != 0 falls back to. */
typedef struct kprn_train_opts {
    int32_t loss;           /* 0 = nn.BCECriterion on labels[S]; 1 = nn.BPRLoss on (pos_seg, neg_seg) pairs */
    int32_t bpr_on_logits;  /* BPR on the pooled class logits instead of the sigmoid scores */
    int32_t regularize;     /* MyOptimizer.lua:196: clip and l2 apply only when set (shipped config: 0) */
    int32_t use_clip;
    float clip_norm, l2;
    float lr, beta1, beta2, eps;   /* optim.adam (OneModel.lua:340-355) */
    float grad_scale;       /* multiplies dL/dscore; S_local / S_global when this batch is one rank's shard of a step (0 = 1) */
    int32_t dense;          /* 1 = reference-literal dense flow */
    int32_t reserved[4];
} kprn_train_opts;
/* Optional hint from a host with a look-ahead batcher (BatcherFileList.lua:53-96 stages the next bucket while one trains): the
 * device-resident ids [N_next, T, 3] of the batch the NEXT kprn_train_step* call will get.  The step running in between brings
 * that batch's entity rows up to date at its tail, beside its weight-gradient kernel, instead of at the head of the next step.
 * Purely a scheduling hint: a wrong or missing hint changes no result.  The buffer must stay valid until that step returns. */
int kprn_train_prefetch_device(kprn_ctx* ctx, const int32_t* d_ids_next, int64_t N_next);
int kprn_train_step(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S, const float* labels,
                    const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs, const int64_t* pad_rows3,
                    const kprn_train_opts* opts, float* loss);
/* the same with the batch already resident in HBM; asynchronous on the ctx stream when single-GPU (with a communicator the
 * row exchange needs one host round trip).  The loss stays on the device until kprn_last_loss. */
int kprn_train_step_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg_offsets, int64_t N, int64_t S,
                           const float* d_labels, const int32_t* d_pos_seg, const int32_t* d_neg_seg, int64_t n_pairs,
                           const int64_t* pad_rows3, const kprn_train_opts* opts);
int kprn_last_loss(kprn_ctx* ctx, float* loss);
/* distinct entity rows of the last training batch on this rank / in the union over ranks / rows moved by the all-gather */
int kprn_train_info(kprn_ctx* ctx, int64_t* rows_local, int64_t* rows_union, int64_t* rows_gathered);

/*
 * Multi-GPU: one process per GPU, an NCCL communicator owned by the ctx (the reference is single-device,
 * OneModel.lua:173-179).  Rank 0 calls kprn_comm_unique_id and ships the 128 bytes to the other processes (a file, a
 * socket, torch.distributed ...); every rank then calls kprn_comm_init.  NCCL is bound at run time (libnccl.so.2).
 * With a communicator, kprn_train_step* all-reduces the small tensors' gradients (~4 MB) and all-gathers the entity-row
 * gradients (row ids + rows); every rank reduces each row's contributions in rank order and applies the same update, so
 * replicas stay bit-identical.  Scoring shards by (user,item) segment and needs no collective.
 */
int kprn_comm_unique_id(void* id128);
int kprn_comm_init(kprn_ctx* ctx, int32_t rank, int32_t world, const void* id128);
int kprn_comm_destroy(kprn_ctx* ctx);
int kprn_comm_info(kprn_ctx* ctx, int32_t* rank, int32_t* world, int64_t* bytes_moved, int64_t* collective_calls, int32_t reset);
int kprn_broadcast_params(kprn_ctx* ctx, int32_t root);                 /* replicas start from rank `root`'s parameters */
int kprn_allreduce_grads(kprn_ctx* ctx, int32_t average);               /* dense flat gradient, one ncclAllReduce (reference-style flow) */
int kprn_allreduce_host_f64(kprn_ctx* ctx, double* vals, int64_t n, int32_t op_max);   /* loss / counters: n <= 64 host doubles, sum or max */

/* engine knobs (no reference counterpart; they never change results).  Keys:
 *   "dedup"  1/0: project each distinct (type, entity, relation) step input once / every step (-1: KPRN_TC_DEDUP env, default on)
 *   "prefix" 1/0: scoring recurrence steps through the shared leading (pad) tokens of the batch once / per path
 *                 (-1: KPRN_TC_PREFIX env, default on for batches of >= 4096 paths; 1 forces it for any size)
 * and one that changes what a multi-GPU call means (BASELINE configs[2] "logit allreduce"):
 *   "path_sharded" 1/0: with a communicator, the PATHS of each (user,item) segment are spread over the ranks: every rank passes
 *                 all S segments (its share of each, possibly empty) and the forward combines the pooled logits across ranks as
 *                 log-sum-exp partials (two small all-reduces), so scores / losses / gradients are those of the whole batch;
 *                 training then uses grad_scale = 1.  Default 0: ranks own whole segments, no data-path collective. */
int kprn_set_option(kprn_ctx* ctx, const char* key, int64_t value);

/* counters for bench.py: kernels launched by this library since the last reset */
int kprn_launch_count(kprn_ctx* ctx, int64_t* n, int32_t reset);
/* device time (ms) measured with CUDA events on the ctx stream while profiling is enabled (kprn_profile(ctx,1)),
 * accumulated since the last reset: slot 0 = whole gate phase of a forward (gathers + LSTM + head), slot 1 = the
 * x-projection kernel ("proj"), slot 2 = the recurrent kernel ("recur"); n3 = number of launches per slot.
 * Slots 1-2 are zero on the fp32 CUDA-core engine. */
int kprn_profile(kprn_ctx* ctx, int32_t enable);
int kprn_profile_read(kprn_ctx* ctx, float* ms3, int64_t* n3, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* KPRN_B200_H */
