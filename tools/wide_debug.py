"""Debug helper: path-logit error of the wide shapes (d_e > 256 / H > 256) by feature."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_case, engine_for

for name, kw in (("halves only   d_e=320 H=128", dict(N=300, seed=16, d_t=8, d_e=320, d_r=8, H=128, V_e=800)),
                 ("wide only T=1 d_e=64  H=512", dict(N=300, seed=17, T=1, d_t=8, d_e=64, d_r=8, H=512, V_e=800)),
                 ("wide only T=2 d_e=64  H=512", dict(N=300, seed=17, T=2, d_t=8, d_e=64, d_r=8, H=512, V_e=800)),
                 ("wide only T=6 d_e=64  H=512", dict(N=300, seed=17, d_t=8, d_e=64, d_r=8, H=512, V_e=800)),
                 ("wide H=320    d_e=64", dict(N=300, seed=18, d_t=8, d_e=64, d_r=8, H=320, V_e=800)),
                 ("both          d_e=512 H=512", dict(N=300, seed=15, d_t=8, d_e=512, d_r=8, H=512, V_e=800))):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=1)
    e.set_params(params)
    scores, logits = e.forward(ids, seg, want_logits=True)
    pl = e.path_logits(ids.shape[0]).astype(np.float64)
    r = ref["path_logits"]
    err = np.abs(pl - r) / np.maximum(np.abs(r), np.sqrt(np.mean(r * r)))
    rows = np.nonzero(err.max(axis=1) > 1e-4)[0]
    s2 = e.forward(ids, seg)
    es = np.abs(s2 - ref["scores"]).max()
    print("%-30s max rel err %.2e  bad rows %d of %d (first %s)  scores-only abs err %.2e" % (name, err.max(), len(rows), len(err), rows[:8], es), flush=True)
    e.close()
