"""Debug driver for the tensor-core path: prints error stats against the oracle (run under gpurun with a timeout)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_case, engine_for  # noqa: E402


def run(kw, precision=1):
    o, params, ids, seg, labels = make_case(**kw)
    ref = o.forward(params, ids, seg, cache=True)
    e = engine_for(o, precision=precision)
    e.set_params(params)
    t0 = time.time()
    scores, logits = e.forward(ids, seg, want_logits=True)
    dt = time.time() - t0
    pl = e.path_logits(ids.shape[0])
    rp = ref["path_logits"]
    scale = np.sqrt(np.mean(rp * rp))
    err = np.abs(pl - rp)
    print("case %s prec %d: path_logits max abs err %.3e (rms ref %.3e) rel-to-rms %.3e; scores max rel %.3e; %.2fs"
          % ({k: v for k, v in kw.items() if k in ("N", "d_e", "H", "T", "d_t")}, precision, err.max(), scale,
             err.max() / scale, np.max(np.abs(scores - ref["scores"]) / ref["scores"]), dt), flush=True)
    bad = np.argwhere(err > 1e-3 * scale)
    if len(bad):
        print("   first bad entries (path, col):", bad[:5].tolist(), "of", len(bad), flush=True)
    e.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(dict(N=128, seed=1, d_t=8, d_e=64, d_r=8, H=64, V_e=500))
        run(dict(N=128, seed=1, d_t=8, d_e=64, d_r=8, H=64, V_e=500), precision=2)
    elif which == "mid":
        run(dict(N=300, seed=2, d_t=16, d_e=128, d_r=16, H=128, V_e=2000))
        run(dict(N=1000, seed=3, d_t=32, d_e=256, d_r=32, H=256, V_e=5000))
        run(dict(N=333, seed=4, d_t=50, d_e=100, d_r=50, H=250, V_e=3000))
    elif which == "big":
        run(dict(N=40000, seed=5, d_t=8, d_e=64, d_r=8, H=64, V_e=50000))
        run(dict(N=20000, seed=6, d_t=16, d_e=256, d_r=16, H=256, V_e=50000, length_mode="realistic"))
