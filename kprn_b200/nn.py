"""Host-side mirror of the reference's Torch7 operator surface for the path (same names, argument meaning, error
behaviour), over the C ABI.  This is the Python twin of kprn_b200/lua/kprn_b200.lua, so that the parity tests read
like the reference's own driver code:

    model = KPRNPathScorer(opt)                     # nn.MapReduce(predictor_net, reducer) -> Sigmoid -> Select(2,classId)
    output = model:forward(inputs)                  # MyOptimizer.lua:189, test_from_checkpoint.lua:109
    err = criterion:forward(output, targets); df_do = criterion:backward(output, targets)     # MyOptimizer.lua:193-194
    model:backward(inputs, df_do)                   # MyOptimizer.lua:195

Paths relative to /root/reference/release/songPathRnn/.
"""
import numpy as np

from . import engine as _e


class KPRNPathScorer:
    """nn.Module surface (torch/nn Module.lua): forward/updateOutput, backward = updateGradInput + accGradParameters,
    zeroGradParameters, parameters, getParameters, training/evaluate, type/cuda/double/float, clearState."""

    # flag names of model/OneModel.lua:27-87
    def __init__(self, entityTypeVocabSize, entityTypeEmbeddingDim, entityVocabSize, entityEmbeddingDim,
                 relationVocabSize, relationEmbeddingDim, rnnHidSize, rnnType="lstm", topK=2, k=5,
                 labelDimension=46, classId=1, maxPathLength=6, headLayers=1, precision=_e.PREC_TC_SPLIT, gpuid=0):
        pool = {2: _e.POOL_LSE, 0: _e.POOL_MAX, 1: _e.POOL_TOPK}[topK]      # OneModel.lua:284-293
        cell = _e.CELL_LSTM if rnnType == "lstm" else _e.CELL_RNN           # OneModel.lua:235-254
        if cell == _e.CELL_RNN and precision != _e.PREC_FP32:
            precision = _e.PREC_FP32                                         # the tensor-core path covers the LSTM cell
        self.engine = _e.Engine(T=maxPathLength, d_t=entityTypeEmbeddingDim, d_e=entityEmbeddingDim,
                                d_r=relationEmbeddingDim, H=rnnHidSize, L=labelDimension, V_t=entityTypeVocabSize,
                                V_e=entityVocabSize, V_r=relationVocabSize, cell=cell, head_layers=headLayers,
                                pool=pool, topk=k, class_id=classId, precision=precision, device=gpuid)
        self.train = True
        self.output = None
        self.gradInput = np.empty(0)          # ids are not differentiable (MapReduce.lua:83 expects an empty tensor)
        self._grad_cached = None
        self._did_backward = False
        # thin proxies for MyOptimizer:zeroPadTokens (MyOptimizer.lua:74-93): lookup:parameters()[1][idx]:zero()
        self.typeLookup = _LookupProxy(self.engine, _e.TABLE_TYPE)
        self.entityLookup = _LookupProxy(self.engine, _e.TABLE_ENTITY)
        self.relationLookup = _LookupProxy(self.engine, _e.TABLE_RELATION)

    # -- forward ---------------------------------------------------------------------------------
    def updateOutput(self, input):
        """input: [B,P,T,3] tensor of 1-based ids (float64 like Batcher.lua:43-54, or any integer dtype)."""
        a = np.asarray(input)
        if a.ndim != 4 or a.shape[3] != 3:
            raise ValueError("expected [B,P,T,3] ids")
        if a.dtype == np.float64:
            self.output = self.engine.forward_f64ids(a, train=self.train)
        else:
            B, P = a.shape[0], a.shape[1]
            self.output = self.engine.forward(a.reshape(B * P, a.shape[2], 3), None, paths_per_seg=P, train=self.train)
        self._grad_cached = None
        self._did_backward = False
        return self.output

    forward = updateOutput

    def forward_csr(self, ids, seg):
        """Variable paths-per-pair batches (CSR), the generalisation of the reference's dense buckets."""
        self.output = self.engine.forward(ids, seg, train=self.train)
        self._grad_cached = None
        self._did_backward = False
        return self.output

    # -- backward: the two halves may be called separately (MapReduce.lua:58-71) -------------------
    def updateGradInput(self, input, gradOutput):
        self._grad_cached = None if gradOutput is None else np.ascontiguousarray(gradOutput, dtype=np.float32)
        return self.gradInput

    def accGradParameters(self, input, gradOutput=None, scale=1.0):
        if self._did_backward:
            return
        g = self._grad_cached if gradOutput is None else np.ascontiguousarray(gradOutput, dtype=np.float32)
        self.engine.backward(g, scale)      # g is None -> the loss gradient cached on the device by the criterion
        self._did_backward = True

    def backward(self, input, gradOutput, scale=1.0):
        self.updateGradInput(input, gradOutput)
        self.accGradParameters(input, gradOutput, scale)
        return self.gradInput

    def zeroGradParameters(self):
        self.engine.zero_grad()

    # -- parameters --------------------------------------------------------------------------------
    def getParameters(self, mirror=False):
        """Flat (parameters, gradParameters) in nn.Module:getParameters() order (MyOptimizer.lua:42).  Torch7 returns tensors
        that ALIAS the module storage and MyOptimizer relies on it (optim.adam mutates them in place, MyOptimizer.lua:218); these
        parameters live in HBM, so host arrays cannot alias them and an unmodified host optimizer would train a stale copy.
        Like the Lua shim, the call refuses unless the caller asks for mirrors explicitly (and pushes them back with
        setParameters); the supported training route is MyOptimizer below (the whole step on the device)."""
        if not mirror:
            raise RuntimeError("KPRNPathScorer.getParameters(): parameters live on the device and cannot alias a host tensor; "
                               "use nn.MyOptimizer (device-side trainBatch) or getParameters(mirror=True) + setParameters()")
        return self.engine.get_params(), self.engine.get_grads()

    def parameters(self):
        p, g = self.getParameters(mirror=True)
        return [p], [g]

    def setParameters(self, flat):
        self.engine.set_params(flat)

    def reset(self, stdv=0.1, seed=12345):
        self.engine.init_uniform(stdv, seed)      # param:uniform(-0.1, 0.1), OneModel.lua:306-309

    def setOption(self, key, value):
        """Engine knobs (kprn_set_option): "dedup", "prefix" (never change results), "pathSharded" (multi-GPU pooling mode)."""
        self.engine.set_option("path_sharded" if key == "pathSharded" else key, int(value))
        return self

    def training(self):
        self.train = True
        return self

    def evaluate(self):
        self.train = False
        return self

    def type(self, *_):
        return self

    cuda = double = float = type

    def clearState(self):
        self.output = None
        self._grad_cached = None
        return self

    def __repr__(self):
        return "nn.KPRNPathScorer(%d params)" % self.engine.n_params


class _LookupProxy:
    def __init__(self, eng, table):
        self._eng, self._table = eng, table

    def zeroRow(self, idx_1based):
        self._eng.zero_row(self._table, idx_1based)


class BCECriterion:
    """nn.BCECriterion (OneModel.lua:300) evaluated on the scores cached by the last forward."""

    def __init__(self, model):
        self.model = model
        self.output = None

    def forward(self, output, targets):
        self.output = self.model.engine.loss_bce(targets)
        return self.output

    def backward(self, output, targets):
        # the device already holds dL/dscore chained to the pooled logit; None tells model.backward to use it
        return None


class BPRLoss:
    """nn.BPRLoss (model/criterion/BPRLoss.lua:18-32); input = {pos, neg} as 0-based segment indices of the last forward."""

    def __init__(self, model, on_logits=False):
        self.model, self.on_logits = model, on_logits

    def forward(self, pos_seg, neg_seg):
        return self.model.engine.loss_bpr(pos_seg, neg_seg, self.on_logits)


class MyOptimizer:
    """model/optimizer/MyOptimizer.lua mirror: trainBatch = zeroPadTokens -> fEval(forward, criterion, backward,
    clip + L2) -> optim.adam -> zeroPadTokens (MyOptimizer.lua:177-221), entirely on the device."""

    def __init__(self, model, learningRate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, regularize=0, useGradClip=False,
                 gradClipNorm=5.0, l2=0.0, entityTypePadToken=0, entityPadToken=0, relationPadToken=0, numEpochs=1):
        self.model = model
        self.opt = dict(lr=learningRate, b1=beta1, b2=beta2, eps=epsilon)
        self.regularize, self.useGradClip, self.gradClipNorm, self.l2 = regularize, useGradClip, gradClipNorm, l2
        self.pads = (entityTypePadToken, entityPadToken, relationPadToken)       # OneModel.lua:374-376
        self.numEpochs = numEpochs
        self.totalError = 0.0

    def trainBatch(self, inputs, targets):
        a = np.asarray(inputs)
        B, P = a.shape[0], a.shape[1]
        ids = a.reshape(B * P, a.shape[2], 3).astype(np.int32)
        seg = np.arange(0, B * P + 1, P, dtype=np.int32)
        return self.trainBatchCSR(ids, seg, targets)

    def trainBatchCSR(self, ids, seg, targets):
        err = self.model.engine.train_batch_bce(ids, seg, targets, pad_rows=self.pads, regularize=bool(self.regularize),
                                                use_clip=self.useGradClip, clip_norm=self.gradClipNorm, l2=self.l2, **self.opt)
        self.totalError += err
        return err

    def trainBatchBPR(self, ids, seg, pos_seg, neg_seg, on_logits=False):
        """The same step with nn.BPRLoss on (pos, neg) pairs of segments (criterion/BPRLoss.lua:18-32; BASELINE configs[3])."""
        err = self.model.engine.train_batch_bpr(ids, seg, pos_seg, neg_seg, on_logits=on_logits, pad_rows=self.pads,
                                                regularize=bool(self.regularize), use_clip=self.useGradClip,
                                                clip_norm=self.gradClipNorm, l2=self.l2, **self.opt)
        self.totalError += err
        return err

    def train(self, batcher):
        """MyOptimizer:train (MyOptimizer.lua:95-172): epochs over a BatcherFileList-like object."""
        self.model.training()
        history = []
        for _ in range(self.numEpochs):
            self.totalError, nb = 0.0, 0
            while True:
                batch = batcher.getBatch()
                if batch is None:
                    break
                labels, data = batch[0], batch[1]
                self.trainBatch(data, labels)
                nb += 1
            history.append(self.totalError / max(nb, 1))
            batcher.reset()
        return history
