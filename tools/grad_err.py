"""Prints per-tensor gradient errors (vs the fp64 oracle) of the fp32 and tensor-core training paths."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_case, engine_for

for kw in (dict(N=1500, seed=32, d_t=16, d_e=128, d_r=16, H=128, V_e=2000), dict(N=333, seed=33, d_t=50, d_e=100, d_r=50, H=250, V_e=3000), dict(N=900, seed=34, d_t=32, d_e=256, d_r=32, H=256, V_e=5000)):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg, cache=True)
    loss_ref, dsc = o.bce(ref["scores"], labels)
    g_ref = o.backward(params, ids, seg, ref, dsc)
    for prec in (0, 3):
        e = engine_for(o, precision=prec)
        e.set_params(params); e.zero_grad()
        e.forward(ids, seg, train=True); e.loss_bce(labels); e.backward()
        g = e.get_grads()
        out = []
        for name in o.shapes():
            a, r = o.view(g, name).astype(np.float64), o.view(g_ref, name)
            rms = np.sqrt(np.mean(r * r))
            err = np.abs(a - r) / np.maximum(np.abs(r), rms)
            out.append("%s %.1e" % (name, err.max()))
        print("H=%d prec %d:" % (kw["H"], prec), "  ".join(out), flush=True)
        e.close()
