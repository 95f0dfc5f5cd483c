"""Leading-token reuse of the scoring recurrence (kprn_b200/csrc/tc_prefix.cu): the batch is sorted by the number of leading
copies of one token (the left pad of SP/model/batcher/Batcher.lua:43-54), tiles start from the shared state chain.  Results must
equal the oracle's -- which steps through every pad -- and the engine's own results with the option off."""
import numpy as np
import pytest

from kprn_b200 import synth
from tests.util import assert_rel, engine_for, make_case

pytestmark = pytest.mark.gpu


def _run(o, params, ids, seg, prefix, precision=1, logits=False):
    e = engine_for(o, precision=precision)
    e.set_option("prefix", prefix)
    e.set_params(params)
    if logits:
        out, _ = e.forward(ids, seg, want_logits=True)
        pl = e.path_logits(ids.shape[0])
    else:
        out, pl = e.forward(ids, seg), None
    e.close()
    return out, pl


@pytest.mark.parametrize("kw", [
    dict(N=1500, seed=71, d_t=16, d_e=128, d_r=16, H=128, V_e=2000, length_mode="realistic"),
    dict(N=2100, seed=72, d_t=32, d_e=256, d_r=32, H=256, V_e=5000, length_mode="stress"),
    dict(N=777, seed=73, d_t=50, d_e=100, d_r=50, H=250, V_e=3000, length_mode="stress"),      # ragged last tile, odd tile count
    dict(N=640, seed=74, d_t=8, d_e=64, d_r=8, H=64, V_e=500, length_mode=6),                   # no pads at all: p = 0 everywhere
    dict(N=900, seed=75, d_t=8, d_e=64, d_r=8, H=64, V_e=500, length_mode=1),                   # one real step: tiles start at T-1
])
def test_prefix_reuse_matches_oracle_and_the_plain_recurrence(kw):
    o, params, ids, seg, _ = make_case(**kw)
    ref = o.forward(params, ids, seg)
    on, pl_on = _run(o, params, ids, seg, 1, logits=True)
    off, pl_off = _run(o, params, ids, seg, 0, logits=True)
    assert_rel(on, ref["scores"], what="scores, prefix on")
    assert_rel(pl_on, ref["path_logits"], what="path logits, prefix on")
    assert_rel(on, off, rtol=1e-5, what="scores on vs off")
    assert_rel(pl_on, pl_off, rtol=1e-5, what="path logits on vs off")


def test_prefix_reuse_scores_only_path_and_shuffled_batch():
    """The scores-only entry (no logits materialised: head fused into the recurrence, results scattered through the sort's
    permutation) on a batch whose pad lengths are interleaved, so the sort really moves rows."""
    o, params, _, _, _ = make_case(N=64, seed=76, d_t=16, d_e=128, d_r=16, H=128, V_e=4000)
    ids = synth.make_paths(5000, T=o.cfg.T, V_e=4000, seed=760, length_mode="stress")
    seg = synth.make_segments(5000, seed=760)
    ref = o.forward(params, ids, seg)
    on, _ = _run(o, params, ids, seg, -1)          # default: on for N >= 4096
    off, _ = _run(o, params, ids, seg, 0)
    assert_rel(on, ref["scores"], what="scores, default options")
    assert_rel(on, off, rtol=1e-5, what="scores on vs off")


def test_prefix_reuse_with_deduplication_off():
    """Both knobs are independent: with the projection running on every step of the SORTED batch (dedup off) the recurrence
    reads its own rows of XW, and tiles still start from the shared state chain."""
    o, params, ids, seg, _ = make_case(N=1300, seed=77, d_t=16, d_e=128, d_r=16, H=128, V_e=2000, length_mode="realistic")
    ref = o.forward(params, ids, seg)
    e = engine_for(o, precision=1)
    e.set_option("prefix", 1)
    e.set_option("dedup", 0)
    e.set_params(params)
    sc, _ = e.forward(ids, seg, want_logits=True)
    pl = e.path_logits(ids.shape[0])
    e.close()
    assert_rel(sc, ref["scores"], what="scores, prefix on / dedup off")
    assert_rel(pl, ref["path_logits"], what="path logits, prefix on / dedup off")
