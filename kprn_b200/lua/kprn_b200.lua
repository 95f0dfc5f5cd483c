--[[
kprn_b200.lua -- LuaJIT-FFI shim: the reference-side binding of libkprn_b200.so.

Defines nn.KPRNPathScorer, a single nn.Module that replaces
    nn.Sequential():add(nn.MapReduce(predictor_net, reducer)):add(nn.Sigmoid())   (model/OneModel.lua:294)
plus the nn.Select(2,classId) that MyOptimizer.lua:126 / test_from_checkpoint.lua:82 append, and
kprn.Optimizer, a drop-in for MyOptimizer (model/optimizer/MyOptimizer.lua) whose trainBatch runs the whole
step (pad-row zeroing, forward, BCE, backward, clip/L2, Adam) on the device.

No cutorch / cunn: host code stays Lua(JIT), all device work goes through the C ABI in include/kprn_b200.h.

STATUS: this image has no Lua/LuaJIT/Torch7 (SURVEY.md App. C), so this file has NOT been executed here.
The identical C ABI is exercised through ctypes by kprn_b200/engine.py and tests/.
]]
local ffi = require 'ffi'
require 'torch'
require 'nn'

ffi.cdef[[
typedef struct kprn_ctx kprn_ctx;
typedef struct kprn_config {
    int32_t T; int32_t d_t, d_e, d_r; int32_t H; int32_t L;
    int64_t V_t, V_e, V_r;
    int32_t cell; int32_t head_layers; int32_t pool; int32_t topk; float inv_gamma;
    int32_t class_id; int32_t precision; int32_t device; int32_t reserved;
} kprn_config;
const char* kprn_last_error(void);
const char* kprn_version(void);
int kprn_create(const kprn_config* cfg, kprn_ctx** out);
int kprn_destroy(kprn_ctx* ctx);
int kprn_set_stream(kprn_ctx* ctx, void* cuda_stream);
int kprn_sync(kprn_ctx* ctx);
int kprn_param_count(kprn_ctx* ctx, int64_t* n);
int kprn_param_offsets(kprn_ctx* ctx, int64_t* offsets11);
int kprn_set_params(kprn_ctx* ctx, const float* host, int64_t n);
int kprn_set_params_f64(kprn_ctx* ctx, const double* host, int64_t n);
int kprn_get_params(kprn_ctx* ctx, float* host, int64_t n);
int kprn_get_grads(kprn_ctx* ctx, float* host, int64_t n);
int kprn_init_uniform(kprn_ctx* ctx, float scale, uint64_t seed);
int kprn_params_device(kprn_ctx* ctx, float** dptr, int64_t* n);
int kprn_params_dirty(kprn_ctx* ctx);
int kprn_grads_device(kprn_ctx* ctx, float** dptr, int64_t* n);
int kprn_zero_row(kprn_ctx* ctx, int table, int64_t row_1based);
int kprn_forward(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S,
                 int32_t paths_per_seg, int32_t train, float* out_scores, float* out_logits);
int kprn_forward_f64ids(kprn_ctx* ctx, const double* ids_f64, int64_t B, int64_t P, int32_t train,
                        float* out_scores, float* out_logits);
int kprn_forward_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg_offsets, int64_t N, int64_t S,
                        int32_t train, float* d_out_scores, float* d_out_logits);
int kprn_pinned_alloc(int64_t bytes, void** host_ptr);
int kprn_pinned_free(void* host_ptr);
int kprn_get_path_logits(kprn_ctx* ctx, float* host, int64_t n_floats);
int kprn_loss_bce(kprn_ctx* ctx, const float* labels, int64_t S, float* loss);
int kprn_loss_bpr(kprn_ctx* ctx, const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs,
                  int32_t apply_to_logits, float* loss);
int kprn_backward(kprn_ctx* ctx, const float* grad_scores, float scale);
int kprn_zero_grad(kprn_ctx* ctx);
int kprn_grad_norm(kprn_ctx* ctx, float* norm);
int kprn_grad_scale(kprn_ctx* ctx, float s);
int kprn_grad_add_params(kprn_ctx* ctx, float l2);
int kprn_adam_step(kprn_ctx* ctx, float lr, float beta1, float beta2, float eps);
int kprn_train_batch_bce(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S,
                         const float* labels, const int64_t* pad_rows3, int32_t regularize, int32_t use_clip,
                         float clip_norm, float l2, float lr, float beta1, float beta2, float eps, float* loss);
int kprn_train_batch_bpr(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S,
                         const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs, int32_t apply_to_logits,
                         const int64_t* pad_rows3, int32_t regularize, int32_t use_clip, float clip_norm, float l2, float lr,
                         float beta1, float beta2, float eps, float* loss);
typedef struct kprn_train_opts {
    int32_t loss; int32_t bpr_on_logits; int32_t regularize; int32_t use_clip; float clip_norm, l2;
    float lr, beta1, beta2, eps; float grad_scale; int32_t dense; int32_t reserved[4];
} kprn_train_opts;
int kprn_train_step(kprn_ctx* ctx, const int32_t* ids, const int32_t* seg_offsets, int64_t N, int64_t S, const float* labels,
                    const int32_t* pos_seg, const int32_t* neg_seg, int64_t n_pairs, const int64_t* pad_rows3,
                    const kprn_train_opts* opts, float* loss);
int kprn_train_step_device(kprn_ctx* ctx, const int32_t* d_ids, const int32_t* d_seg_offsets, int64_t N, int64_t S,
                           const float* d_labels, const int32_t* d_pos_seg, const int32_t* d_neg_seg, int64_t n_pairs,
                           const int64_t* pad_rows3, const kprn_train_opts* opts);
int kprn_train_prefetch_device(kprn_ctx* ctx, const int32_t* d_ids_next, int64_t N_next);
int kprn_last_loss(kprn_ctx* ctx, float* loss);
int kprn_train_info(kprn_ctx* ctx, int64_t* rows_local, int64_t* rows_union, int64_t* rows_gathered);
int kprn_comm_unique_id(void* id128);
int kprn_comm_init(kprn_ctx* ctx, int32_t rank, int32_t world, const void* id128);
int kprn_comm_destroy(kprn_ctx* ctx);
int kprn_comm_info(kprn_ctx* ctx, int32_t* rank, int32_t* world, int64_t* bytes_moved, int64_t* collective_calls, int32_t reset);
int kprn_broadcast_params(kprn_ctx* ctx, int32_t root);
int kprn_allreduce_grads(kprn_ctx* ctx, int32_t average);
int kprn_allreduce_host_f64(kprn_ctx* ctx, double* vals, int64_t n, int32_t op_max);
int kprn_set_option(kprn_ctx* ctx, const char* key, int64_t value);
int kprn_launch_count(kprn_ctx* ctx, int64_t* n, int32_t reset);
int kprn_profile(kprn_ctx* ctx, int32_t enable);
int kprn_profile_read(kprn_ctx* ctx, float* ms3, int64_t* n3, int32_t reset);
]]

local C = ffi.load(os.getenv('KPRN_B200_LIB') or 'kprn_b200')

local function check(rc)
   if rc ~= 0 then error('kprn_b200: ' .. ffi.string(C.kprn_last_error()), 2) end
end

kprn = kprn or {}

------------------------------------------------------------------------------------------------
-- nn.KPRNPathScorer
------------------------------------------------------------------------------------------------
local Scorer, parent = torch.class('nn.KPRNPathScorer', 'nn.Module')

-- opt mirrors the OneModel.lua flags that shape the path (OneModel.lua:27-87):
--   relationVocabSize, relationEmbeddingDim, entityVocabSize, entityEmbeddingDim, entityTypeVocabSize,
--   entityTypeEmbeddingDim, rnnHidSize, rnnType ('lstm'|'rnn'), topK (0 max, 1 topk-mean, 2 LogSumExp), k,
--   labelDimension (46), classId (1), maxPathLength (6), precision, gpuid
function Scorer:__init(opt)
   parent.__init(self)
   local cfg = ffi.new('kprn_config')
   cfg.T = opt.maxPathLength or 6
   cfg.d_t, cfg.d_e, cfg.d_r = opt.entityTypeEmbeddingDim, opt.entityEmbeddingDim, opt.relationEmbeddingDim
   cfg.H = opt.rnnHidSize
   cfg.L = opt.labelDimension or 46
   cfg.V_t, cfg.V_e, cfg.V_r = opt.entityTypeVocabSize, opt.entityVocabSize, opt.relationVocabSize
   cfg.cell = (opt.rnnType == 'lstm') and 0 or 1
   cfg.head_layers = opt.headLayers or 1
   local topK = opt.topK or 2
   cfg.pool = (topK == 2) and 0 or ((topK == 0) and 1 or 2)
   cfg.topk = opt.k or 5
   cfg.inv_gamma = 1.0
   cfg.class_id = opt.classId or 1
   cfg.precision = opt.precision or 1
   cfg.device = opt.gpuid or 0
   self.opt = opt
   self.cfgT = cfg.T
   local pctx = ffi.new('kprn_ctx*[1]')
   check(C.kprn_create(cfg, pctx))
   self.ctx = ffi.gc(pctx[0], C.kprn_destroy)
   local n = ffi.new('int64_t[1]')
   check(C.kprn_param_count(self.ctx, n))
   self.nParams = tonumber(n[0])
   self.output = torch.FloatTensor()
   self.gradInput = torch.Tensor()          -- ids are not differentiable (MapReduce.lua:83 expects an empty tensor)
   self.train = true
   self._gradCached = nil
   self._didBackward = false
   -- thin lookup-table proxies so MyOptimizer:zeroPadTokens-style code keeps working (MyOptimizer.lua:74-93)
   local function proxy(tableId)
      local owner = self
      local p = {}
      function p:zeroRow(idx) check(C.kprn_zero_row(owner.ctx, tableId, idx)) end
      function p:parameters()
         local w = setmetatable({}, {__index = function(_, idx) return {zero = function() p:zeroRow(idx) end} end})
         return {w}, {}
      end
      return p
   end
   self.typeLookup, self.entityLookup, self.relationLookup = proxy(0), proxy(1), proxy(2)
end

function Scorer:reset(stdv)
   check(C.kprn_init_uniform(self.ctx, stdv or 0.1, 12345))   -- param:uniform(-0.1,0.1), OneModel.lua:306-309
end

-- input: DoubleTensor [B,P,T,3] of 1-based ids (Batcher.lua:43-54) -> output FloatTensor [B] = sigmoid(pool)[:,classId]
function Scorer:updateOutput(input)
   assert(input:dim() == 4 and input:size(4) == 3, 'expected [B,P,T,3] ids')
   local B, P = input:size(1), input:size(2)
   input = input:contiguous()
   self.output:resize(B)
   if input:type() == 'torch.DoubleTensor' then
      check(C.kprn_forward_f64ids(self.ctx, input:data(), B, P, self.train and 1 or 0, self.output:data(), nil))
   else
      local ids = input:int():contiguous()
      check(C.kprn_forward(self.ctx, ids:data(), nil, B * P, B, P, self.train and 1 or 0, self.output:data(), nil))
   end
   self._gradCached = nil
   self._didBackward = false
   return self.output
end

-- updateGradInput / accGradParameters may be invoked separately (MapReduce.lua:58-71): cache the upstream
-- gradient in the former, run BPTT once per forward in the latter.
function Scorer:updateGradInput(input, gradOutput)
   -- gradOutput == nil: the criterion ran on the device (model:bce / model:bpr) and its gradient is cached there
   self._gradCached = gradOutput and gradOutput:float():contiguous() or nil
   return self.gradInput
end

function Scorer:accGradParameters(input, gradOutput, scale)
   if self._didBackward then return end
   local g = (gradOutput and gradOutput:float():contiguous()) or self._gradCached
   -- nil -> NULL: kprn_backward then uses the loss gradient cached by kprn_loss_bce / kprn_loss_bpr (as nn.py does)
   check(C.kprn_backward(self.ctx, g and g:data() or nil, scale or 1))
   self._didBackward = true
end

function Scorer:backward(input, gradOutput, scale)
   self:updateGradInput(input, gradOutput)
   self:accGradParameters(input, gradOutput, scale)
   return self.gradInput
end

function Scorer:zeroGradParameters() check(C.kprn_zero_grad(self.ctx)) end

-- Flat parameter / gradient vectors.  In Torch7 getParameters() returns tensors that ALIAS the module's storage, and
-- MyOptimizer relies on that (optim.adam mutates `parameters` in place, MyOptimizer.lua:42,218).  The parameters of this
-- module live in HBM, so a host tensor cannot alias them: an unmodified MyOptimizer would train a stale copy without any
-- error.  getParameters() therefore REFUSES unless the caller states that it wants host mirrors and will push them back:
--     local p, g = model:getParameters{mirror = true} ... model:pushParameters(p)
-- The supported training route is kprn.Optimizer (below), whose trainBatch runs the whole step on the device.
function Scorer:getParameters(how)
   if not (type(how) == 'table' and how.mirror) then
      error('nn.KPRNPathScorer:getParameters(): parameters live on the device and cannot alias a host tensor; an in-place '
            .. 'host optimizer (MyOptimizer + optim.adam) would train a stale copy. Use kprn.Optimizer, or ask for mirrors '
            .. 'explicitly with getParameters{mirror=true} and call pushParameters() after every host-side update.', 2)
   end
   self._flat = self._flat or torch.FloatTensor(self.nParams)
   self._flatGrad = self._flatGrad or torch.FloatTensor(self.nParams)
   check(C.kprn_get_params(self.ctx, self._flat:data(), self.nParams))
   check(C.kprn_get_grads(self.ctx, self._flatGrad:data(), self.nParams))
   return self._flat, self._flatGrad
end
function Scorer:parameters()
   local p, g = self:getParameters{mirror = true}
   return {p}, {g}
end
function Scorer:pushParameters(flat)          -- after a host-side optimizer touched the mirror
   flat = flat or self._flat
   if flat:type() == 'torch.DoubleTensor' then check(C.kprn_set_params_f64(self.ctx, flat:data(), self.nParams))
   else check(C.kprn_set_params(self.ctx, flat:float():data(), self.nParams)) end
end
function Scorer:pullGradients()
   check(C.kprn_get_grads(self.ctx, self._flatGrad:data(), self.nParams))
   return self._flatGrad
end

-- engine knobs (include/kprn_b200.h, kprn_set_option): "dedup", "prefix" never change results; "pathSharded" (with a communicator)
-- spreads the paths of every (user,item) pair over the ranks and combines the pooled logits as log-sum-exp partials
function Scorer:setOption(key, value)
   if key == 'pathSharded' then key = 'path_sharded' end
   check(C.kprn_set_option(self.ctx, key, value and (value == true and 1 or value) or 0))
   return self
end
function Scorer:training() self.train = true; return self end
function Scorer:evaluate() self.train = false; return self end
function Scorer:type() return self end
function Scorer:cuda() return self end
function Scorer:double() return self end
function Scorer:float() return self end
function Scorer:clearState() self.output:set(); self._gradCached = nil; return self end
function Scorer:__tostring__()
   return string.format('nn.KPRNPathScorer(%s, %d params)', ffi.string(C.kprn_version()), self.nParams)
end
-- torch.save / torch.load of the module object (saver hook OneModel.lua:393-401): parameters travel as a FloatTensor
function Scorer:write(file)
   local flat = self:getParameters{mirror = true}
   file:writeObject({opt = self.opt, flat = flat})
end
function Scorer:read(file)
   local t = file:readObject()
   self:__init(t.opt)
   self:pushParameters(t.flat)
end

------------------------------------------------------------------------------------------------
-- criteria on the cached scores (MyOptimizer.lua:193-194)
------------------------------------------------------------------------------------------------
function Scorer:bce(labels)                       -- nn.BCECriterion, OneModel.lua:300
   local y = labels:float():contiguous()
   local loss = ffi.new('float[1]')
   check(C.kprn_loss_bce(self.ctx, y:data(), y:nElement(), loss))
   return loss[0]
end
function Scorer:bpr(posSeg, negSeg, onLogits)     -- nn.BPRLoss, criterion/BPRLoss.lua:18-32 (0-based segment ids)
   local a, b = posSeg:int():contiguous(), negSeg:int():contiguous()
   local loss = ffi.new('float[1]')
   check(C.kprn_loss_bpr(self.ctx, a:data(), b:data(), a:nElement(), onLogits and 1 or 0, loss))
   return loss[0]
end

------------------------------------------------------------------------------------------------
-- kprn.Optimizer: drop-in for MyOptimizer (same constructor fields that matter, same train(batcher) loop)
------------------------------------------------------------------------------------------------
local Opt = torch.class('kprn.Optimizer')

function Opt:__init(model, optConfig, optInfo)
   self.model = model                                  -- nn.KPRNPathScorer
   self.lr = optConfig.learningRate or 1e-3            -- OneModel.lua:340-355
   self.beta1, self.beta2 = optConfig.beta1 or 0.9, optConfig.beta2 or 0.999
   self.epsilon = optConfig.epsilon or 1e-8
   self.regularize = optInfo.regularize or 0
   self.useGradClip = optInfo.useGradClip and 1 or 0
   self.gradClipNorm = optInfo.gradClipNorm or 5
   self.l2 = optInfo.l2 or 0
   self.pads = ffi.new('int64_t[3]', optInfo.entityTypePadToken or 0, optInfo.entityPadToken or 0,
                       optInfo.relationPadToken or 0)      -- OneModel.lua:374-376
   self.numEpochs = optInfo.numEpochs or 1
end

function Opt:_opts(loss)
   local o = ffi.new('kprn_train_opts')
   o.loss = loss
   o.bpr_on_logits = self.bprOnLogits and 1 or 0
   o.regularize, o.use_clip, o.clip_norm, o.l2 = self.regularize, self.useGradClip, self.gradClipNorm, self.l2
   o.lr, o.beta1, o.beta2, o.eps = self.lr, self.beta1, self.beta2, self.epsilon
   o.grad_scale = self.gradScale or 1      -- S_local / S_global when this process holds one shard of the batch
   o.dense = self.dense and 1 or 0         -- 1: the reference's dense flow literally (dense gradient, dense optim.adam)
   return o
end

local function denseSegments(B, P)
   local seg = torch.IntTensor(B + 1)
   for i = 0, B do seg[i + 1] = i * P end
   return seg
end

-- MyOptimizer:trainBatch (MyOptimizer.lua:177-221) entirely on the device, nn.BCECriterion (OneModel.lua:300)
function Opt:trainBatch(inputs, targets)
   local B, P = inputs:size(1), inputs:size(2)
   local ids = inputs:int():contiguous()
   local seg = denseSegments(B, P)
   local y = targets:float():contiguous()
   local loss = ffi.new('float[1]')
   check(C.kprn_train_step(self.model.ctx, ids:data(), seg:data(), B * P, B, y:data(), nil, nil, 0, self.pads, self:_opts(0), loss))
   return loss[0]
end

-- the same step with nn.BPRLoss (criterion/BPRLoss.lua:18-32): posSeg/negSeg are 0-based bucket indices, paired element-wise
function Opt:trainBatchBPR(inputs, posSeg, negSeg)
   local B, P = inputs:size(1), inputs:size(2)
   local ids = inputs:int():contiguous()
   local seg = denseSegments(B, P)
   local a, b = posSeg:int():contiguous(), negSeg:int():contiguous()
   local loss = ffi.new('float[1]')
   check(C.kprn_train_step(self.model.ctx, ids:data(), seg:data(), B * P, B, nil, a:data(), b:data(), a:nElement(), self.pads,
                           self:_opts(1), loss))
   return loss[0]
end

-- multi-GPU: one Lua process per GPU.  Rank 0 writes the 128-byte NCCL id to `path`, the others read it.
function Opt:initComm(rank, world, path)
   local id = ffi.new('char[128]')
   if rank == 0 then
      check(C.kprn_comm_unique_id(id))
      local f = assert(io.open(path .. '.tmp', 'wb')); f:write(ffi.string(id, 128)); f:close()
      os.rename(path .. '.tmp', path)
   else
      local f
      repeat f = io.open(path, 'rb') until f
      ffi.copy(id, f:read(128), 128); f:close()
   end
   check(C.kprn_comm_init(self.model.ctx, rank, world, id))
   check(C.kprn_broadcast_params(self.model.ctx, 0))
end

-- MyOptimizer:train (MyOptimizer.lua:95-172): epoch loop over a BatcherFileList
function Opt:train(trainBatcher)
   self.model:training()
   for epoch = 1, self.numEpochs do
      local total, nb = 0, 0
      while true do
         local targets, inputs = trainBatcher:getBatch()
         if targets == nil then break end
         total = total + self:trainBatch(inputs, targets)
         nb = nb + 1
      end
      print(string.format('Iter: %d\navg loss in epoch = %f', epoch, total / math.max(nb, 1)))
      trainBatcher:reset()
   end
end

return kprn
