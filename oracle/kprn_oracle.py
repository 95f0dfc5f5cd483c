"""ctypes wrapper around oracle/libkprn_oracle.so (the fp64 CPU restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Never imported from kprn_b200/.
See kprn_oracle.c for the reference citations; PARITY UNPINNED (no reference goldens exist).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkprn_oracle.so")


class OCfg(C.Structure):
    _fields_ = [("T", C.c_int32), ("d_t", C.c_int32), ("d_e", C.c_int32), ("d_r", C.c_int32),
                ("H", C.c_int32), ("L", C.c_int32),
                ("V_t", C.c_int64), ("V_e", C.c_int64), ("V_r", C.c_int64),
                ("cell", C.c_int32), ("head_layers", C.c_int32), ("pool", C.c_int32),
                ("topk", C.c_int32), ("inv_gamma", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "kprn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkprn_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.okprn_param_count.restype = C.c_int64
        _lib.okprn_bce.restype = C.c_double
        _lib.okprn_bpr.restype = C.c_double
    return _lib


def _p(a, t=C.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t))


PARAM_NAMES = ["E_t", "E_e", "E_r", "Wx", "bx", "Wh", "bh", "W1", "b1", "Wo", "bo"]


class Oracle:
    """fp64 oracle for one model configuration (dims follow SURVEY.md section 8 notation)."""

    def __init__(self, T=6, d_t=4, d_e=8, d_r=4, H=8, L=46, V_t=6, V_e=50, V_r=9,
                 cell=0, head_layers=1, pool=0, topk=5, inv_gamma=1.0):
        self.cfg = OCfg(T, d_t, d_e, d_r, H, L, V_t, V_e, V_r, cell, head_layers, pool, topk, inv_gamma)
        self.n_params = int(lib().okprn_param_count(C.byref(self.cfg)))
        off = (C.c_int64 * 11)()
        lib().okprn_param_offsets(C.byref(self.cfg), off)
        self.offsets = dict(zip(PARAM_NAMES, [int(o) for o in off]))
        self.D = d_t + d_e + d_r
        self.G = (4 if cell == 0 else 1) * H

    def shapes(self):
        c = self.cfg
        s = {"E_t": (c.V_t, c.d_t), "E_e": (c.V_e, c.d_e), "E_r": (c.V_r, c.d_r),
             "Wx": (self.G, self.D), "bx": (self.G,), "Wh": (self.G, c.H), "Wo": (c.L, c.H), "bo": (c.L,)}
        if c.cell == 1:
            s["bh"] = (c.H,)
        if c.head_layers == 2:
            s["W1"] = (c.H, c.H)
            s["b1"] = (c.H,)
        return s

    def view(self, flat, name):
        shp = self.shapes()[name]
        o = self.offsets[name]
        return flat[o:o + int(np.prod(shp))].reshape(shp)

    def init_params(self, seed=12345, scale=0.1):
        """U(-0.1, 0.1) init (model/OneModel.lua:306-309); PCG64 stream, seed = OneModel.lua:115."""
        rng = np.random.Generator(np.random.PCG64(seed))
        return rng.uniform(-scale, scale, self.n_params)

    def zero_row(self, params, table, row_1based):
        assert lib().okprn_zero_row(C.byref(self.cfg), _p(params), int(table), C.c_int64(row_1based)) == 0

    def forward(self, params, ids, seg, class_id=1, cache=False):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64)
        c = self.cfg
        N, S = ids.shape[0], seg.shape[0] - 1
        assert ids.shape == (N, c.T, 3) and seg[-1] == N
        out = {"path_logits": np.zeros((N, c.L)), "seg_logits": np.zeros((S, c.L)), "scores": np.zeros(S)}
        act = cst = hst = z1 = None
        if cache:
            act = np.zeros((N, c.T, self.G))
            cst = np.zeros((N, c.T, c.H))
            hst = np.zeros((N, c.T, c.H))
            z1 = np.zeros((N, c.H))
            out.update(act=act, cst=cst, hst=hst, z1=z1)
        rc = lib().okprn_forward_paths(C.byref(c), _p(params), _p(ids, C.c_int32), C.c_int64(N),
                                       _p(out["path_logits"]), _p(act), _p(cst), _p(hst), _p(z1))
        if rc != 0:
            raise ValueError("oracle: id out of range")
        lib().okprn_pool(C.byref(c), _p(out["path_logits"]), _p(seg, C.c_int32), C.c_int64(S),
                         int(class_id), _p(out["seg_logits"]), _p(out["scores"]))
        return out

    @staticmethod
    def bce(scores, labels):
        scores = np.ascontiguousarray(scores, dtype=np.float64)
        labels = np.ascontiguousarray(labels, dtype=np.float64)
        d = np.zeros_like(scores)
        loss = lib().okprn_bce(_p(scores), _p(labels), C.c_int64(scores.size), _p(d))
        return float(loss), d

    @staticmethod
    def bpr(pos, neg):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        neg = np.ascontiguousarray(neg, dtype=np.float64)
        dp, dn = np.zeros_like(pos), np.zeros_like(neg)
        loss = lib().okprn_bpr(_p(pos), _p(neg), C.c_int64(pos.size), _p(dp), _p(dn))
        return float(loss), dp, dn

    def backward(self, params, ids, seg, fwd, dscores, class_id=1, scale=1.0, grads=None):
        """dscores = dL/dscore[S]; returns flat grads (accumulated into `grads` if given)."""
        c = self.cfg
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        N, S = ids.shape[0], seg.shape[0] - 1
        dsl = np.zeros((S, c.L))
        lib().okprn_score_backward(C.byref(c), _p(fwd["scores"]), _p(np.ascontiguousarray(dscores, dtype=np.float64)),
                                   C.c_int64(S), int(class_id), _p(dsl))
        return self.backward_logits(params, ids, seg, fwd, dsl, scale, grads)

    def backward_logits(self, params, ids, seg, fwd, dseg_logits, scale=1.0, grads=None):
        c = self.cfg
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        N, S = ids.shape[0], seg.shape[0] - 1
        if grads is None:
            grads = np.zeros(self.n_params)
        dsl = np.ascontiguousarray(dseg_logits, dtype=np.float64)
        rc = lib().okprn_backward(C.byref(c), _p(params), _p(ids, C.c_int32), _p(seg, C.c_int32),
                                  C.c_int64(N), C.c_int64(S), _p(fwd["path_logits"]), _p(fwd["seg_logits"]),
                                  _p(dsl), _p(fwd["act"]), _p(fwd["cst"]), _p(fwd["hst"]), _p(fwd["z1"]),
                                  C.c_double(scale), _p(grads))
        assert rc == 0
        return grads

    @staticmethod
    def clip_l2(g, theta, use_clip, clipnorm, l2):
        lib().okprn_clip_l2(_p(g), _p(theta), C.c_int64(g.size), int(use_clip), C.c_double(clipnorm), C.c_double(l2))

    @staticmethod
    def adam(x, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        lib().okprn_adam(_p(x), _p(g), _p(m), _p(v), C.c_int64(x.size), C.c_int64(t),
                         C.c_double(lr), C.c_double(b1), C.c_double(b2), C.c_double(eps))

    @staticmethod
    def num_threads():
        return int(lib().okprn_num_threads())

    @staticmethod
    def set_num_threads(n):
        """OpenMP threads of the C restatement (torchrun exports OMP_NUM_THREADS=1; bench.py's CPU arm asks for all host cores)."""
        lib().okprn_set_num_threads(int(n))
