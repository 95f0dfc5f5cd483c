"""ctypes binding of libkprn_b200.so -- the same C ABI (include/kprn_b200.h) the LuaJIT-FFI shim binds.

There is no CPU fallback here: if the CUDA extension is missing or no device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KPRN_B200_LIB") or os.path.join(_HERE, "libkprn_b200.so")   # same override as the Lua shim

CELL_LSTM, CELL_RNN = 0, 1
POOL_LSE, POOL_MAX, POOL_TOPK = 0, 1, 2
PREC_FP32, PREC_TC_SPLIT, PREC_TC_FAST, PREC_TC_ALL = 0, 1, 2, 3
TABLE_TYPE, TABLE_ENTITY, TABLE_RELATION = 0, 1, 2

PARAM_NAMES = ["E_t", "E_e", "E_r", "Wx", "bx", "Wh", "bh", "W1", "b1", "Wo", "bo"]


class KprnConfig(C.Structure):
    _fields_ = [("T", C.c_int32), ("d_t", C.c_int32), ("d_e", C.c_int32), ("d_r", C.c_int32),
                ("H", C.c_int32), ("L", C.c_int32),
                ("V_t", C.c_int64), ("V_e", C.c_int64), ("V_r", C.c_int64),
                ("cell", C.c_int32), ("head_layers", C.c_int32), ("pool", C.c_int32), ("topk", C.c_int32),
                ("inv_gamma", C.c_float), ("class_id", C.c_int32), ("precision", C.c_int32),
                ("device", C.c_int32), ("reserved", C.c_int32)]


class KprnTrainOpts(C.Structure):
    _fields_ = [("loss", C.c_int32), ("bpr_on_logits", C.c_int32), ("regularize", C.c_int32), ("use_clip", C.c_int32),
                ("clip_norm", C.c_float), ("l2", C.c_float), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("grad_scale", C.c_float), ("dense", C.c_int32), ("reserved", C.c_int32 * 4)]


class KprnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kprn_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None

_I64P = C.POINTER(C.c_int64)
_F32P = C.POINTER(C.c_float)
_I32P = C.POINTER(C.c_int32)
_F64P = C.POINTER(C.c_double)

# every entry point declared in include/kprn_b200.h: name -> argtypes (restype is int unless noted)
ABI = {
    "kprn_create": [C.POINTER(KprnConfig), C.POINTER(C.c_void_p)],
    "kprn_destroy": [C.c_void_p],
    "kprn_set_stream": [C.c_void_p, C.c_void_p],
    "kprn_sync": [C.c_void_p],
    "kprn_param_count": [C.c_void_p, _I64P],
    "kprn_param_offsets": [C.c_void_p, _I64P],
    "kprn_set_params": [C.c_void_p, _F32P, C.c_int64],
    "kprn_set_params_f64": [C.c_void_p, _F64P, C.c_int64],
    "kprn_get_params": [C.c_void_p, _F32P, C.c_int64],
    "kprn_get_grads": [C.c_void_p, _F32P, C.c_int64],
    "kprn_init_uniform": [C.c_void_p, C.c_float, C.c_uint64],
    "kprn_params_device": [C.c_void_p, C.POINTER(C.c_void_p), _I64P],
    "kprn_params_dirty": [C.c_void_p],
    "kprn_grads_device": [C.c_void_p, C.POINTER(C.c_void_p), _I64P],
    "kprn_zero_row": [C.c_void_p, C.c_int, C.c_int64],
    "kprn_forward": [C.c_void_p, _I32P, _I32P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _F32P, _F32P],
    "kprn_forward_f64ids": [C.c_void_p, _F64P, C.c_int64, C.c_int64, C.c_int32, _F32P, _F32P],
    "kprn_forward_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p],
    "kprn_pinned_alloc": [C.c_int64, C.POINTER(C.c_void_p)],
    "kprn_pinned_free": [C.c_void_p],
    "kprn_get_path_logits": [C.c_void_p, _F32P, C.c_int64],
    "kprn_loss_bce": [C.c_void_p, _F32P, C.c_int64, _F32P],
    "kprn_loss_bpr": [C.c_void_p, _I32P, _I32P, C.c_int64, C.c_int32, _F32P],
    "kprn_backward": [C.c_void_p, _F32P, C.c_float],
    "kprn_zero_grad": [C.c_void_p],
    "kprn_grad_norm": [C.c_void_p, _F32P],
    "kprn_grad_scale": [C.c_void_p, C.c_float],
    "kprn_grad_add_params": [C.c_void_p, C.c_float],
    "kprn_adam_step": [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float],
    "kprn_train_batch_bce": [C.c_void_p, _I32P, _I32P, C.c_int64, C.c_int64, _F32P, _I64P, C.c_int32, C.c_int32,
                             C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _F32P],
    "kprn_train_batch_bpr": [C.c_void_p, _I32P, _I32P, C.c_int64, C.c_int64, _I32P, _I32P, C.c_int64, C.c_int32, _I64P, C.c_int32,
                             C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _F32P],
    "kprn_train_step": [C.c_void_p, _I32P, _I32P, C.c_int64, C.c_int64, _F32P, _I32P, _I32P, C.c_int64, _I64P,
                        C.POINTER(KprnTrainOpts), _F32P],
    "kprn_train_step_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int64, _I64P, C.POINTER(KprnTrainOpts)],
    "kprn_train_prefetch_device": [C.c_void_p, C.c_void_p, C.c_int64],
    "kprn_last_loss": [C.c_void_p, _F32P],
    "kprn_train_info": [C.c_void_p, _I64P, _I64P, _I64P],
    "kprn_comm_unique_id": [C.c_void_p],
    "kprn_comm_init": [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p],
    "kprn_comm_destroy": [C.c_void_p],
    "kprn_comm_info": [C.c_void_p, _I32P, _I32P, _I64P, _I64P, C.c_int32],
    "kprn_broadcast_params": [C.c_void_p, C.c_int32],
    "kprn_allreduce_grads": [C.c_void_p, C.c_int32],
    "kprn_allreduce_host_f64": [C.c_void_p, _F64P, C.c_int64, C.c_int32],
    "kprn_set_option": [C.c_void_p, C.c_char_p, C.c_int64],
    "kprn_launch_count": [C.c_void_p, _I64P, C.c_int32],
    "kprn_profile": [C.c_void_p, C.c_int32],
    "kprn_profile_read": [C.c_void_p, _F32P, _I64P, C.c_int32],
}


def load_library():
    """Loads the C-ABI library and declares every symbol of include/kprn_b200.h. Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("kprn_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in ABI.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.kprn_last_error.restype = C.c_char_p
    lib.kprn_last_error.argtypes = []
    lib.kprn_version.restype = C.c_char_p
    lib.kprn_version.argtypes = []
    _lib = lib
    return lib


def _chk(rc):
    if rc != 0:
        raise KprnError(rc, load_library().kprn_last_error().decode("utf-8", "replace"))


def _f32p(a):
    return None if a is None else a.ctypes.data_as(_F32P)


def _i32p(a):
    return None if a is None else a.ctypes.data_as(_I32P)


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory from kprn_pinned_alloc (freed when the array is collected)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    _chk(lib.kprn_pinned_alloc(max(n * dtype.itemsize, 1), C.byref(p)))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    import weakref
    weakref.finalize(buf, lib.kprn_pinned_free, C.c_void_p(p.value))
    return arr


class Engine:
    """One kprn_ctx on one CUDA device."""

    def __init__(self, T=6, d_t=256, d_e=256, d_r=256, H=256, L=46, V_t=6, V_e=2851220, V_r=9,
                 cell=CELL_LSTM, head_layers=1, pool=POOL_LSE, topk=5, inv_gamma=1.0, class_id=1,
                 precision=PREC_FP32, device=0):
        self.lib = load_library()
        self.cfg = KprnConfig(T, d_t, d_e, d_r, H, L, V_t, V_e, V_r, cell, head_layers, pool, topk,
                              inv_gamma, class_id, precision, device, 0)
        h = C.c_void_p()
        _chk(self.lib.kprn_create(C.byref(self.cfg), C.byref(h)))
        self.h = h
        n = C.c_int64()
        _chk(self.lib.kprn_param_count(self.h, C.byref(n)))
        self.n_params = int(n.value)
        off = (C.c_int64 * 11)()
        _chk(self.lib.kprn_param_offsets(self.h, off))
        self.offsets = dict(zip(PARAM_NAMES, [int(o) for o in off]))
        self.D = d_t + d_e + d_r
        self.G = (4 if cell == CELL_LSTM else 1) * H
        self.last_S = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.kprn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters -------------------------------------------------------------------------
    def set_params(self, flat):
        flat = np.asarray(flat)
        if flat.dtype == np.float64:
            flat = np.ascontiguousarray(flat)
            _chk(self.lib.kprn_set_params_f64(self.h, flat.ctypes.data_as(_F64P), flat.size))
        else:
            flat = np.ascontiguousarray(flat, dtype=np.float32)
            _chk(self.lib.kprn_set_params(self.h, _f32p(flat), flat.size))

    def get_params(self):
        out = np.empty(self.n_params, dtype=np.float32)
        _chk(self.lib.kprn_get_params(self.h, _f32p(out), out.size))
        return out

    def get_grads(self):
        out = np.empty(self.n_params, dtype=np.float32)
        _chk(self.lib.kprn_get_grads(self.h, _f32p(out), out.size))
        return out

    def init_uniform(self, scale=0.1, seed=12345):
        _chk(self.lib.kprn_init_uniform(self.h, scale, seed))

    def zero_row(self, table, row_1based):
        _chk(self.lib.kprn_zero_row(self.h, int(table), int(row_1based)))

    def params_device(self):
        p, n = C.c_void_p(), C.c_int64()
        _chk(self.lib.kprn_params_device(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def params_dirty(self):
        """After writing through a pointer kept from params_device(): the tensor-core forward must re-pack its operands."""
        _chk(self.lib.kprn_params_dirty(self.h))

    def grads_device(self):
        p, n = C.c_void_p(), C.c_int64()
        _chk(self.lib.kprn_grads_device(self.h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, ids, seg=None, paths_per_seg=0, train=False, want_logits=False):
        """ids [N,T,3] int32 1-based (host); seg [S+1] CSR or None + paths_per_seg. Returns scores[S] (, logits[S,L])."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        N = ids.shape[0]
        if ids.shape != (N, self.cfg.T, 3):
            raise KprnError(1, "ids must be [N,%d,3]" % self.cfg.T)
        if seg is not None:
            seg = np.ascontiguousarray(seg, dtype=np.int32)
            S = seg.shape[0] - 1
        else:
            S = N // paths_per_seg if paths_per_seg > 0 else 0
        scores = np.empty(max(S, 0), dtype=np.float32)
        logits = np.empty((max(S, 0), self.cfg.L), dtype=np.float32) if want_logits else None
        _chk(self.lib.kprn_forward(self.h, _i32p(ids), _i32p(seg), N, S, paths_per_seg, int(train),
                                   _f32p(scores), _f32p(logits)))
        self.last_S = S
        return (scores, logits) if want_logits else scores

    def forward_f64ids(self, data, train=False, want_logits=False):
        """data [B,P,T,3] float64 of 1-based ids -- the reference's Batcher tensor (Batcher.lua:43-54)."""
        data = np.ascontiguousarray(data, dtype=np.float64)
        B, P = data.shape[0], data.shape[1]
        scores = np.empty(B, dtype=np.float32)
        logits = np.empty((B, self.cfg.L), dtype=np.float32) if want_logits else None
        _chk(self.lib.kprn_forward_f64ids(self.h, data.ctypes.data_as(_F64P), B, P, int(train), _f32p(scores), _f32p(logits)))
        self.last_S = B
        return (scores, logits) if want_logits else scores

    def forward_device(self, d_ids, d_seg, N, S, train=False, d_scores=0, d_logits=0):
        """Raw device pointers (ints); asynchronous on the ctx stream."""
        _chk(self.lib.kprn_forward_device(self.h, C.c_void_p(d_ids), C.c_void_p(d_seg), N, S, int(train),
                                          C.c_void_p(d_scores) if d_scores else None,
                                          C.c_void_p(d_logits) if d_logits else None))
        self.last_S = S

    def path_logits(self, N):
        out = np.empty((N, self.cfg.L), dtype=np.float32)
        _chk(self.lib.kprn_get_path_logits(self.h, _f32p(out), out.size))
        return out

    # ---- loss / backward / optimizer ----------------------------------------------------------
    def loss_bce(self, labels):
        labels = np.ascontiguousarray(labels, dtype=np.float32)
        loss = C.c_float()
        _chk(self.lib.kprn_loss_bce(self.h, _f32p(labels), labels.size, C.byref(loss)))
        return float(loss.value)

    def loss_bpr(self, pos_seg, neg_seg, on_logits=False):
        pos_seg = np.ascontiguousarray(pos_seg, dtype=np.int32)
        neg_seg = np.ascontiguousarray(neg_seg, dtype=np.int32)
        loss = C.c_float()
        _chk(self.lib.kprn_loss_bpr(self.h, _i32p(pos_seg), _i32p(neg_seg), pos_seg.size, int(on_logits), C.byref(loss)))
        return float(loss.value)

    def backward(self, grad_scores=None, scale=1.0):
        g = None if grad_scores is None else np.ascontiguousarray(grad_scores, dtype=np.float32)
        _chk(self.lib.kprn_backward(self.h, _f32p(g), scale))

    def zero_grad(self):
        _chk(self.lib.kprn_zero_grad(self.h))

    def grad_norm(self):
        v = C.c_float()
        _chk(self.lib.kprn_grad_norm(self.h, C.byref(v)))
        return float(v.value)

    def grad_scale(self, s):
        _chk(self.lib.kprn_grad_scale(self.h, s))

    def grad_add_params(self, l2):
        _chk(self.lib.kprn_grad_add_params(self.h, l2))

    def adam_step(self, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        _chk(self.lib.kprn_adam_step(self.h, lr, b1, b2, eps))

    def train_batch_bce(self, ids, seg, labels, pad_rows=(0, 0, 0), regularize=False, use_clip=False, clip_norm=5.0,
                        l2=0.0, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        labels = np.ascontiguousarray(labels, dtype=np.float32)
        pads = (C.c_int64 * 3)(*[int(x) for x in pad_rows])
        loss = C.c_float()
        _chk(self.lib.kprn_train_batch_bce(self.h, _i32p(ids), _i32p(seg), ids.shape[0], seg.shape[0] - 1, _f32p(labels),
                                           pads, int(regularize), int(use_clip), clip_norm, l2, lr, b1, b2, eps, C.byref(loss)))
        return float(loss.value)

    def train_opts(self, loss="bce", bpr_on_logits=False, regularize=False, use_clip=False, clip_norm=5.0, l2=0.0, lr=1e-3,
                   b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0, dense=False):
        o = KprnTrainOpts()
        o.loss = 0 if loss == "bce" else 1
        o.bpr_on_logits, o.regularize, o.use_clip = int(bpr_on_logits), int(regularize), int(use_clip)
        o.clip_norm, o.l2, o.lr, o.beta1, o.beta2, o.eps = clip_norm, l2, lr, b1, b2, eps
        o.grad_scale, o.dense = grad_scale, int(dense)
        return o

    def train_step(self, ids, seg, labels=None, pos_seg=None, neg_seg=None, pad_rows=(0, 0, 0), **kw):
        """One MyOptimizer:trainBatch (host buffers). BCE with labels[S], or BPR with (pos_seg, neg_seg) 0-based segment pairs."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        if labels is None and pos_seg is not None:
            kw.setdefault("loss", "bpr")
        o = self.train_opts(**kw)
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.float32)
        ps = None if pos_seg is None else np.ascontiguousarray(pos_seg, dtype=np.int32)
        ns = None if neg_seg is None else np.ascontiguousarray(neg_seg, dtype=np.int32)
        pads = (C.c_int64 * 3)(*[int(x) for x in pad_rows])
        loss = C.c_float()
        _chk(self.lib.kprn_train_step(self.h, _i32p(ids), _i32p(seg), ids.shape[0], seg.shape[0] - 1, _f32p(lab), _i32p(ps), _i32p(ns),
                                      0 if ps is None else ps.size, pads, C.byref(o), C.byref(loss)))
        return float(loss.value)

    def train_step_device(self, d_ids, d_seg, N, S, d_labels=0, d_pos=0, d_neg=0, n_pairs=0, pad_rows=(0, 0, 0), opts=None):
        """Raw device pointers (ints); the loss stays on the device (last_loss())."""
        pads = (C.c_int64 * 3)(*[int(x) for x in pad_rows])
        o = opts if opts is not None else self.train_opts()
        _chk(self.lib.kprn_train_step_device(self.h, C.c_void_p(d_ids), C.c_void_p(d_seg), N, S, C.c_void_p(d_labels) if d_labels else None,
                                             C.c_void_p(d_pos) if d_pos else None, C.c_void_p(d_neg) if d_neg else None, n_pairs, pads,
                                             C.byref(o)))

    def train_prefetch_device(self, d_ids_next, N_next):
        """Scheduling hint (kprn_train_prefetch_device): device pointer of the ids the NEXT train_step* call will get."""
        _chk(self.lib.kprn_train_prefetch_device(self.h, C.c_void_p(d_ids_next) if d_ids_next else None, int(N_next)))

    def last_loss(self):
        v = C.c_float()
        _chk(self.lib.kprn_last_loss(self.h, C.byref(v)))
        return float(v.value)

    def train_info(self):
        a, b, c_ = C.c_int64(), C.c_int64(), C.c_int64()
        _chk(self.lib.kprn_train_info(self.h, C.byref(a), C.byref(b), C.byref(c_)))
        return {"rows_local": int(a.value), "rows_union": int(b.value), "rows_gathered": int(c_.value)}

    def train_batch_bpr(self, ids, seg, pos_seg, neg_seg, on_logits=False, pad_rows=(0, 0, 0), regularize=False, use_clip=False,
                        clip_norm=5.0, l2=0.0, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        ps = np.ascontiguousarray(pos_seg, dtype=np.int32)
        ns = np.ascontiguousarray(neg_seg, dtype=np.int32)
        pads = (C.c_int64 * 3)(*[int(x) for x in pad_rows])
        loss = C.c_float()
        _chk(self.lib.kprn_train_batch_bpr(self.h, _i32p(ids), _i32p(seg), ids.shape[0], seg.shape[0] - 1, _i32p(ps), _i32p(ns), ps.size,
                                           int(on_logits), pads, int(regularize), int(use_clip), clip_norm, l2, lr, b1, b2, eps, C.byref(loss)))
        return float(loss.value)

    # ---- multi-GPU (one process per GPU; NCCL communicator owned by the ctx) --------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        _chk(load_library().kprn_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, world, id128):
        buf = (C.c_char * 128).from_buffer_copy(id128)
        _chk(self.lib.kprn_comm_init(self.h, rank, world, buf))

    def comm_destroy(self):
        _chk(self.lib.kprn_comm_destroy(self.h))

    def comm_info(self, reset=False):
        r, w, b, c_ = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        _chk(self.lib.kprn_comm_info(self.h, C.byref(r), C.byref(w), C.byref(b), C.byref(c_), int(reset)))
        return {"rank": r.value, "world": w.value, "bytes": int(b.value), "calls": int(c_.value)}

    def broadcast_params(self, root=0):
        _chk(self.lib.kprn_broadcast_params(self.h, root))

    def allreduce_grads(self, average=False):
        _chk(self.lib.kprn_allreduce_grads(self.h, int(average)))

    def allreduce_host(self, vals, op_max=False):
        v = np.ascontiguousarray(vals, dtype=np.float64).copy()
        _chk(self.lib.kprn_allreduce_host_f64(self.h, v.ctypes.data_as(_F64P), v.size, int(op_max)))
        return v

    # ---- plumbing -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        _chk(self.lib.kprn_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def sync(self):
        _chk(self.lib.kprn_sync(self.h))

    def set_option(self, key, value):
        _chk(self.lib.kprn_set_option(self.h, key.encode(), int(value)))

    def launch_count(self, reset=False):
        n = C.c_int64()
        _chk(self.lib.kprn_launch_count(self.h, C.byref(n), int(reset)))
        return int(n.value)

    def profile(self, enable=True):
        _chk(self.lib.kprn_profile(self.h, int(enable)))

    def profile_read(self, reset=True):
        """Returns ([ms_gate_phase, ms_proj, ms_recur], [n, n, n]) accumulated while profiling was enabled."""
        ms, n = (C.c_float * 3)(), (C.c_int64 * 3)()
        _chk(self.lib.kprn_profile_read(self.h, ms, n, int(reset)))
        return [float(x) for x in ms], [int(x) for x in n]
