"""Bit-reproducibility check of the training step (and of the next-batch hint): runs the same 6 steps several times, prints the
number of parameters that differ between runs (expected: 0).  usage: python tools/repro_check.py  (one B200)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import engine_for, make_case
from tests.test_gpu_train import _batches
V_e = 4000
o, params, _, _, _ = make_case(N=64, seed=90, d_t=8, d_e=128, d_r=8, H=128, V_e=V_e)
c = o.cfg
pads = (c.V_t, c.V_e, c.V_r)
batches = _batches(o, 3, 700, 900, V_e)
dev = [(torch.from_numpy(i).cuda(), torch.from_numpy(s).cuda(), torch.from_numpy(l.astype(np.float32)).cuda(), len(s) - 1) for i, s, l in batches]
out = {}
for mode in ("none", "none2", "none3", "right", "wrong"):
    e = engine_for(o, precision=3)
    e.set_params(params)
    for t in range(6):
        ids_t, seg_t, lab_t, S = dev[t % 3]
        if mode == "right":
            e.train_prefetch_device(dev[(t + 1) % 3][0].data_ptr(), 700)
        elif mode == "wrong":
            e.train_prefetch_device(dev[(t + 2) % 3][0].data_ptr(), 700)
        e.train_step_device(ids_t.data_ptr(), seg_t.data_ptr(), 700, S, d_labels=lab_t.data_ptr(), pad_rows=pads)
    out[mode] = e.get_params().copy()
    e.close()
for m in ("none2", "none3", "right", "wrong"):
    d = np.abs(out[m].astype(np.float64) - out["none"])
    print(m, "max diff", d.max(), "n diff", int((d > 0).sum()))
    for name in o.shapes():
        dd = np.abs(o.view(out[m], name).astype(np.float64) - o.view(out["none"], name))
        if dd.max() > 0:
            print("   ", name, dd.max(), int((dd > 0).sum()), "rows" if dd.ndim == 2 else "", np.unique(np.nonzero(dd)[0])[:10] if dd.ndim == 2 else "")
