"""Times one MyOptimizer:trainBatch-equivalent step (kprn_train_batch_bce: pad zero -> fwd -> BCE -> bwd -> Adam) on one
GPU for the BASELINE dims (d=256 conv. A, 65 536 paths, V_e = 2.85 M) -- BASELINE configs[3] flavour, secondary metric."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kprn_b200 import engine, synth

V_E, N, d = 2851220, 65536, 256
res = {}
for prec in [int(x) for x in (sys.argv[1:] or ["0", "3"])]:
    e = engine.Engine(d_t=d, d_e=d, d_r=d, H=d, V_e=V_E, precision=prec)
    e.init_uniform(0.1, 1)
    ids = synth.make_paths(N, V_e=V_E, seed=5)
    seg = synth.make_segments(N, seed=5)
    labels = synth.make_labels(len(seg) - 1, 5)
    pads = (6, V_E, 9)
    losses = []
    for i in range(2):
        losses.append(e.train_batch_bce(ids, seg, labels, pad_rows=pads))
    e.sync()
    t0 = time.perf_counter()
    K = 5
    for i in range(K):
        losses.append(e.train_batch_bce(ids, seg, labels, pad_rows=pads))
    e.sync()
    dt = (time.perf_counter() - t0) / K
    res[prec] = {"ms_per_step": 1e3 * dt, "paths_per_s": N / dt, "loss_first": losses[0], "loss_last": losses[-1]}
    print("precision", prec, json.dumps(res[prec]), flush=True)
    e.close()
