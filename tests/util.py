"""Shared helpers for the parity tests."""
import numpy as np

from kprn_b200 import synth
from oracle.kprn_oracle import Oracle

RTOL = 1e-4   # north_star: "within 1e-4 relative (fp32)" against the fp64 CPU result


def assert_rel(actual, expected, rtol=RTOL, what=""):
    """|a-e| <= rtol * max(|e|, rms(e)): elementwise relative error, with the tensor's own RMS as the floor so that
    entries that are (near) zero by cancellation are judged against the tensor's scale."""
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, (what, a.shape, e.shape)
    assert np.isfinite(a).all(), what + ": non-finite values"
    floor = np.sqrt(np.mean(e * e)) if e.size else 0.0
    denom = np.maximum(np.abs(e), floor)
    denom = np.where(denom == 0, 1.0, denom)
    err = np.abs(a - e) / denom
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, "%s: max rel err %.3e > %.1e (at %s)" % (what, worst, rtol, np.unravel_index(err.argmax(), err.shape))
    return worst


def make_case(N, seed, T=6, d_t=8, d_e=16, d_r=8, H=16, L=46, V_e=300, seg_mode="realistic", fixed_p=4,
              length_mode="stress", V_t=6, V_r=9, **kw):
    o = Oracle(T=T, d_t=d_t, d_e=d_e, d_r=d_r, H=H, L=L, V_e=V_e, V_t=V_t, V_r=V_r, **kw)
    c = o.cfg
    ids = synth.make_paths(N, T=T, V_e=V_e, seed=seed, length_mode=length_mode, V_t=c.V_t, V_r=c.V_r)
    seg = synth.make_segments(N, seed=seed, mode=seg_mode, fixed_p=fixed_p)
    params = o.init_params(seed)
    for tbl, row in ((0, c.V_t), (1, c.V_e), (2, c.V_r)):
        o.zero_row(params, tbl, row)
    labels = synth.make_labels(len(seg) - 1, seed)
    return o, params, ids, seg, labels


def engine_for(o, precision=0, device=0):
    from kprn_b200.engine import Engine
    c = o.cfg
    return Engine(T=c.T, d_t=c.d_t, d_e=c.d_e, d_r=c.d_r, H=c.H, L=c.L, V_t=c.V_t, V_e=c.V_e, V_r=c.V_r,
                  cell=c.cell, head_layers=c.head_layers, pool=c.pool, topk=c.topk, inv_gamma=c.inv_gamma,
                  class_id=1, precision=precision, device=device)
